#!/usr/bin/env python
"""Benchmark of the unet-only hot path: U-Net-only frames/sec in bf16.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--impl native|reference|cudnn]

--config selects one of BASELINE.json's configurations (default 1, the one the metric is quoted on):

  0  unet-only pipeline, 2,000-frame 256x256 clip, batch 32 (the reference's own CPU-runnable case)
  1  256x256 GIRAFE-shaped clip, batch 512 per step per GPU                       [default]
  2  BAGLS-shaped 512(H) x 256(W) frames, native resolution, batch 256 per step per GPU
     (200,000 frames over 8 GPUs = 25,000 per rank; weak scaling)
  3  YOLO-crop+UNet stage: pre-cropped 256x256 ROIs streamed in batches of 512
  4  10^6-frame clip, STRONG scaling: the frames are sharded over the ranks, the int32 areas are
     all-gathered and the kinematic features computed, all inside the timed region

A step = one batch of synthetic gray frames through stem -> U-Net -> threshold -> per-frame area.
`value` times K steps with the clip resident in HBM; `e2e` times the same work through the public
API from pinned HOST memory (H2D of the frames and D2H of the areas inside the timed region). After
the K steps the area waveform is gathered (NCCL when N > 1) and the kinematic features are
computed, inside the timed region.

--impl reference times the reference's own per-frame CPU loop (features.py:234-238 ->
utils.py:218-241 -> models/unet.py:74-88; batch 1, fp32, all host threads): the unmodified
reference package from oracle/_ref when oracle/make_ref.py has put it there (kind "reference"),
else the oracle's restatement of the same loop (kind "port"). One CPU process on rank 0, whatever
--gpus says.
--impl cudnn is a CONTEXT arm, not the product: the reference's UNet under PyTorch eager + cuDNN
(bf16, channels_last, same batch) on the same GPU -- "the library on the same box".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import re
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "unet_only_frames_per_sec_256x256_bf16"
FEATS = (32, 64, 128, 256)
L2_NOTE = ("inputs cycle through a resident clip of {frames} frames ({mb:.0f} MB) and {act:.1f} GB of "
           "activations per step, both larger than the 126 MB L2")

CONFIGS = {
    0: dict(workload="unet-only pipeline, 2,000-frame 256x256 clip, batch 32 (BASELINE.json configs[0])",
            hgt=256, wid=256, batch=32, clip=4096, steps=62, warmup=20, seed=0),
    1: dict(workload="unet-only 256x256 GIRAFE-shaped clip, batch 512 per step per GPU, bf16 "
                     "tensor-core path (BASELINE.json configs[1])",
            hgt=256, wid=256, batch=512, clip=4096, steps=100, warmup=20, seed=0),
    2: dict(workload="unet-only BAGLS-shaped 512x256 frames at native resolution, batch 256 per step "
                     "per GPU (BASELINE.json configs[2]: 200,000 frames = 25,000 per rank on 8 GPUs)",
            hgt=512, wid=256, batch=256, clip=2048, steps=98, warmup=20, seed=2),
    3: dict(workload="YOLO-crop+UNet stage: pre-cropped 256x256 ROIs streamed in batches of 512 "
                     "(BASELINE.json configs[3])",
            hgt=256, wid=256, batch=512, clip=4096, steps=100, warmup=20, seed=3),
    4: dict(workload="area waveform + kinematic features of a 1,000,000-frame 256x256 clip, frames "
                     "sharded over the GPUs (BASELINE.json configs[4])",
            hgt=256, wid=256, batch=512, clip=4096, steps=None, warmup=20, seed=0, frames=1_000_000),
}


def module_flops(hgt: int, wid: int) -> dict[str, float]:
    """Algorithmic FLOPs per frame of every reference module on the path (2 x MACs of the
    reference formulation, /root/reference/openglottal/models/unet.py:50-72),
    keyed by the names ogl_unet_launch_name() uses."""
    def conv(cin, cout, h, w):
        return 2.0 * 9 * cin * cout * h * w

    def convt(cin, cout, h, w):      # h, w = INPUT resolution
        return 2.0 * 4 * cin * cout * h * w

    fl = {"stem": conv(1, 32, hgt, wid), "downs.0.net.3+pool": conv(32, 32, hgt, wid)}
    cin = 32
    for lvl in range(1, 4):
        f = FEATS[lvl]
        h, w = hgt >> lvl, wid >> lvl
        fl[f"downs.{lvl}.net.0"] = conv(cin, f, h, w)
        fl[f"downs.{lvl}.net.3+pool"] = conv(f, f, h, w)
        cin = f
    fl["bottleneck.net.0"] = conv(256, 512, hgt >> 4, wid >> 4)
    fl["bottleneck.net.3"] = conv(512, 512, hgt >> 4, wid >> 4)
    for k in range(4):
        lvl = 3 - k
        f = FEATS[lvl]
        h, w = hgt >> lvl, wid >> lvl
        fl[f"ups.{2 * k}(convT)"] = convt(2 * f, f, h // 2, w // 2)
        fl[f"ups.{2 * k + 1}.net.0(cat)"] = conv(2 * f, f, h, w)
        fl[f"ups.{2 * k + 1}.net.3"] = conv(f, f, h, w)
    fl["ups.7.net.3+head"] = fl.pop("ups.7.net.3") + 2.0 * 32 * hgt * wid   # 1x1 head fused in
    return fl


def launch_flops(name: str, fl: dict[str, float]) -> float:
    """A launch that computes several reference modules is named 'a+b' with module names
    (module names may themselves contain '+', e.g. 'downs.0.net.3+pool')."""
    if name in fl:
        return fl[name]
    for i, ch in enumerate(name):
        if ch == "+" and name[:i] in fl:
            return fl[name[:i]] + launch_flops(name[i + 1:], fl)
    raise KeyError(f"launch {name!r} does not name reference modules")


def layer_flops(hgt: int, wid: int) -> list[float]:
    return list(module_flops(hgt, wid).values())


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples: list[int] = []
        self.power: list[float] = []
        self.reasons: set[str] = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        try:    # board energy counter (mJ): the step runs at the power cap, so J/step is the cost
            self.e0, self.t0 = nv.nvmlDeviceGetTotalEnergyConsumption(self.h), time.perf_counter()
        except Exception:
            self.e0 = None
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.01)

    def stop(self) -> dict:
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        pw = sorted(self.power)
        out = {"sm_mhz": s[len(s) // 2], "sm_mhz_min": s[0], "sm_max_mhz": self.max_mhz,
               "power_w": round(pw[len(pw) // 2], 1) if pw else None,
               "reasons": sorted(self.reasons), "samples": len(s)}
        try:    # mean board power over the sampled window from the energy counter (power_w above is
            # NVML's own ~1 s moving average and lags a 0.2 s timed region)
            if getattr(self, "e0", None) is not None:
                e1, t1 = self.nv.nvmlDeviceGetTotalEnergyConsumption(self.h), time.perf_counter()
                out["window_s"] = round(t1 - self.t0, 4)
                out["energy_j"] = round((e1 - self.e0) * 1e-3, 3)
                out["mean_w"] = round((e1 - self.e0) * 1e-3 / max(t1 - self.t0, 1e-6), 1)
        except Exception:
            pass
        return out


def bench_state(seed: int = 0):
    """Weights for the timed runs: the synthetically trained state dict (seed 0) when its cache
    travelled with the repo, else the seeded calibrated one (same architecture; timing is identical).
    Seeds 2 and 3 stand in for openglottal_unet_bagl_50epochs.pt / openglottal_unet_cropped.pt."""
    import synthdata as synth

    cache = ROOT / "tests" / "golden" / "_cache" / "trained_seed0_s150_r128.pt"
    if seed == 0 and cache.exists():
        import torch

        return torch.load(cache, map_location="cpu", weights_only=True), "synthetically-trained(seed0)"
    return synth.calibrated_state(seed), f"calibrated-random(seed{seed})"


def synthetic_clip(n: int, seed: int, hgt: int = 256, wid: int = 256):
    """n distinct frames: a 64-frame seeded glottis cycle tiled with per-frame shifts."""
    import numpy as np
    import synthdata as synth

    base, _ = synth.glottis_clip(64, hgt, wid, seed=seed, period=16.0)
    reps = (n + 63) // 64
    out = np.concatenate([np.roll(base, shift=3 * r, axis=2) for r in range(reps)])[:n]
    return np.ascontiguousarray(out)


# --------------------------------------------------------------------------- CPU baseline
def reference_segmenter(sd):
    """(callable frame_gray -> area, kind): the reference's per-frame step features.py:236-238.
    kind "reference": the unmodified package from oracle/_ref; "port": the oracle restatement."""
    import numpy as np
    import torch

    try:
        from oracle.make_ref import import_reference

        og = import_reference()
    except Exception:
        og = None
    if og is not None:
        from openglottal.utils import unet_segment_frame   # the reference's own function

        model = og.UNet(1, 1, (32, 64, 128, 256))
        model.load_state_dict(sd)
        model.eval()
        dev = torch.device("cpu")

        def seg(frame_gray):
            mask_full = unet_segment_frame(frame_gray, model, dev)       # features.py:236
            return float(np.sum(mask_full > 0))                          # features.py:238
        return seg, "reference"
    from oracle import unet_oracle as uo

    def seg(frame_gray):
        return float(np.sum(uo.segment_frame(sd, frame_gray) > 0))
    return seg, "port"


def cpu_reference_fps(sd, frames, warmup: int, seconds: float, max_frames: int):
    """The reference's per-frame loop (features.py:234-238 -> utils.py:218-241 -> unet.py:74-88):
    batch 1, fp32, torch CPU with all host threads. Returns (fps, n, cores, kind)."""
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    seg, kind = reference_segmenter(sd)
    for f in frames[:warmup]:
        seg(f)
    t0 = time.perf_counter()
    n = 0
    while n < max_frames:
        seg(frames[n % len(frames)])
        n += 1
        if time.perf_counter() - t0 > seconds and n >= 8:
            break
    dt = time.perf_counter() - t0
    return n / dt, n, cores, kind


def cpu_batched_fps(sd, frames, seconds: float = 5.0, batch: int = 32):
    """Fairness figure beside the per-frame loop: the reference's UNet module
    (unet.py:74-88) on the host cores with a batch of `batch` frames per forward, fp32, /255 input and
    `sigmoid > 0.5` area as utils.py:235-241 has them, frames at their own size (no resize). The
    reference itself never runs this way -- its loop is batch 1 -- so it is reported next to
    `cpu_baseline.value`, never in its place. Returns (fps, frames timed) or None."""
    import numpy as np
    import torch

    try:
        from oracle.make_ref import import_reference

        og = import_reference()
    except Exception:
        og = None
    if og is None:
        return None
    model = og.UNet(1, 1, (32, 64, 128, 256))
    model.load_state_dict(sd)
    model.eval()
    x = torch.from_numpy(np.ascontiguousarray(frames[:batch])).float().div_(255.0).unsqueeze(1)
    with torch.no_grad():
        model(x[:2])
        t0 = time.perf_counter()
        n = 0
        while True:
            prob = torch.sigmoid(model(x))
            (prob > 0.5).flatten(1).sum(1)
            n += x.shape[0]
            if time.perf_counter() - t0 > seconds:
                break
    return n / (time.perf_counter() - t0), n


def cpu_features_timing(area, seconds: float = 20.0):
    """The reference's _kinematic_features (features.py:38-68) on prefixes of the job's area
    waveform. Its np.correlate(..., 'full') (features.py:55) is O(n^2): timed at growing n within
    a time budget and extrapolated quadratically to the full length (stated in the result)."""
    import numpy as np

    try:
        from oracle.make_ref import import_reference

        og = import_reference()
    except Exception:
        og = None
    if og is not None:
        from openglottal.features import _kinematic_features as ref_features
        kind = "reference"
    else:
        from oracle.features_oracle import kinematic_features

        def ref_features(a):
            return kinematic_features(a, exact_correlate=True)
        kind = "port"
    from oracle.features_oracle import kinematic_features as lag50

    out = {"kind": kind, "samples": []}
    n, spent = 25_000, 0.0
    while n <= len(area):
        t0 = time.perf_counter()
        ref_features(list(area[:n].astype(np.float64)))
        dt = time.perf_counter() - t0
        out["samples"].append({"n": n, "seconds": round(dt, 3)})
        spent += dt
        if spent + 4 * dt > seconds:
            break
        n *= 2
    last = out["samples"][-1]
    out["extrapolated_seconds_full"] = round(last["seconds"] * (len(area) / last["n"]) ** 2, 1)
    out["extrapolation"] = f"quadratic from n = {last['n']} to n = {len(area)} (np.correlate 'full')"
    t0 = time.perf_counter()
    lag50(area.astype(np.float64))
    out["lag50_port_seconds_full"] = round(time.perf_counter() - t0, 3)
    return out


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    hgt, wid = cfg["hgt"], cfg["wid"]
    sd, wname = bench_state(cfg["seed"])
    frames = synthetic_clip(64, seed=1, hgt=hgt, wid=wid)
    per_step = 8
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    seg, kind = reference_segmenter(sd)
    steps = args.steps if args.steps is not None else 4
    for i in range(args.warmup):
        for f in frames[:per_step]:
            seg(f)
    t0 = time.perf_counter()
    for s in range(steps):
        for j in range(per_step):
            seg(frames[(s * per_step + j) % len(frames)])
    dt = time.perf_counter() - t0
    fps = steps * per_step / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "strong" if args.config == 4 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"] + " -- the reference's own per-frame loop: batch 1, fp32, "
                               "CPU (frames other than 256x256 are squashed to 256x256 by utils.py:234)",
                   "config_index": args.config, "frames_per_step": per_step, "height": hgt, "width": wid,
                   "weights": wname,
                   "processes": "ONE CPU process on rank 0 with all host threads, whatever --gpus "
                                "says: at N > 1 a GPU/CPU ratio divides N GPUs by one host"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"{steps * per_step} frames of the synthetic {hgt}x{wid} clip, "
                                   f"{per_step} per step, torch CPU fp32 batch-1 loop "
                                   + ("(unmodified reference package, oracle/_ref)" if kind == "reference"
                                      else "(oracle restatement)")},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- library context arm
def run_cudnn(args) -> None:
    """The reference's UNet under PyTorch eager + cuDNN on the same GPU (bf16, channels_last, the
    config's batch), followed by the same threshold and per-frame count in eager torch. Context
    for the native number -- none of this repo's kernels run here."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    hgt, wid, batch = cfg["hgt"], cfg["wid"], cfg["batch"]
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    sd, wname = bench_state(cfg["seed"])
    try:
        from oracle.make_ref import import_reference

        og = import_reference()
    except Exception:
        og = None
    if og is None:
        print(json.dumps({"impl": "cudnn", "unavailable": "oracle/_ref (the reference package) is not built"}))
        return
    model = og.UNet(1, 1, (32, 64, 128, 256))
    model.load_state_dict(sd)
    model = model.to(dev).eval().to(torch.bfloat16).to(memory_format=torch.channels_last)
    torch.backends.cudnn.benchmark = True
    clip = torch.from_numpy(synthetic_clip(cfg["clip"], seed=1, hgt=hgt, wid=wid)).to(dev)
    nb = cfg["clip"] // batch
    steps = args.steps if args.steps is not None else (cfg["steps"] or 100)

    @torch.no_grad()
    def step(i):
        lo = (i % nb) * batch
        x = (clip[lo:lo + batch].to(torch.bfloat16) / 255.0).unsqueeze(1).contiguous(memory_format=torch.channels_last)
        z = model(x)
        return (z[:, 0] > 0).flatten(1).sum(1, dtype=torch.int32)

    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(dev.index)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        a = step(i)
    e1.record()
    torch.cuda.synchronize(dev)
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    fps = steps * batch / (ms * 1e-3)
    fl = sum(layer_flops(hgt, wid))
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    print(json.dumps({
        "impl": "cudnn", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": 1, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": cfg["workload"] + " -- CONTEXT: the reference's UNet module under PyTorch "
                               f"{torch.__version__} eager + cuDNN {torch.backends.cudnn.version()}, "
                               "bf16, channels_last, cudnn.benchmark",
                   "config_index": args.config, "frames_per_step_per_gpu": batch, "height": hgt,
                   "width": wid, "weights": wname},
        "pct_of_tc_roofline": 100.0 * fps * fl / (peak_tf * 1e12), "clocks": clocks,
        "last_areas": a[:4].tolist(),
    }))


# --------------------------------------------------------------------------- GPU arm
def read_launch_times(lib, model, hgt, wid, batch):
    """Per-launch times of the last min(K, 16) profiled forwards (CUDA events on the launching
    stream, recorded by ogl_unet_forward between its launches) -> the roofline object."""
    import numpy as np
    from openglottal_b200 import _native

    buf = (C.c_float * 64)()
    cnt = C.c_int(0)
    _native.check(lib.ogl_unet_layer_times(model._handle, buf, 64, C.byref(cnt)))
    nl = cnt.value
    layer_ms = np.array(buf[:nl])
    names = [lib.ogl_unet_launch_name(model._handle, i).decode() for i in range(nl)]
    mfl = module_flops(hgt, wid)
    fl = np.array([launch_flops(n_, mfl) for n_ in names]) * batch
    assert abs(fl.sum() / batch - sum(mfl.values())) < 1.0, "launch names do not cover the path"
    is_tc = np.array([n_ != "stem" for n_ in names])     # every launch but the CUDA-core stem
    n_tc = int(is_tc.sum())
    tc_ms = float(layer_ms[is_tc].sum())
    tc_flops = float(fl[is_tc].sum())
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    which = "fallback"
    if pk.exists():
        peaks = json.loads(pk.read_text())
        which = "measured"
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))   # kernels timed inside a long step
    achieved_tf = tc_flops / (tc_ms * 1e-3) / 1e12
    # DRAM bytes per launch (read + write) from the newest committed `ncu --set full` capture
    # (scripts/ncu_summary.py), scaled to this step's pixel count
    traffic = None
    captures = sorted((ROOT / "profiles").glob("traffic_*.json"),
                      key=lambda q: [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", q.name)])
    if captures:
        cap = json.loads(captures[-1].read_text())
        per_frame = cap.get("dram_bytes_per_frame_tc_launches")      # of a 256x256 frame
        if per_frame is not None:
            traffic = per_frame * batch * (hgt * wid) / (256 * 256) / n_tc
    roofline = {
        "bound": "tensor",
        "kernel": f"conv_tc_kernel + upcat_tc_kernel + s2d_tc_kernel ({n_tc} tcgen05 launches/step: "
                  "every conv3x3, ConvTranspose2d and the head; the Cin=1 stem is a K=16 GEMM inside "
                  "the first of them when the frames are u8)",
        "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
        "peak_source": f"{which} bf16_tflops_sustained", "traffic": traffic,
        "flops_per_launch_avg": tc_flops / n_tc, "ms_per_launch_avg": tc_ms / n_tc,
        "tc_share_of_step": tc_ms / float(layer_ms.sum()),
    }
    layers = [{"layer": n_, "ms": float(m), "tflops": float(f / (m * 1e-3) / 1e12) if m > 0 else None}
              for n_, m, f in zip(names, layer_ms, fl)]
    return roofline, layers, nl, peak_tf


def run_native(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    import openglottal_b200 as ogl
    from openglottal_b200 import _native, sharding

    cfg = CONFIGS[args.config]
    hgt, wid, batch, clip_frames = cfg["hgt"], cfg["wid"], cfg["batch"], cfg["clip"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: openglottal_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sd, wname = bench_state(cfg["seed"])
    model = ogl.UNet().to(dev)
    model.load_state_dict(sd)
    model.eval()
    model.max_batch = max(batch, 512)
    if args.no_graph:
        model.use_graphs = False
    lib = _native.load()

    clip_host = torch.from_numpy(synthetic_clip(clip_frames, seed=1 + rank, hgt=hgt, wid=wid)).pin_memory()
    clip_dev = clip_host.to(dev)
    nb = clip_frames // batch
    strong = args.config == 4
    if strong:
        total_frames = args.frames or cfg["frames"]
        lo_f, hi_f = sharding.shard_range(total_frames, rank, world)
        local_frames = hi_f - lo_f
        steps = (local_frames + batch - 1) // batch          # the last step of a rank may be short
    else:
        steps = args.steps if args.steps is not None else cfg["steps"]
        total_frames = world * steps * batch
        local_frames = steps * batch
    warmup = args.warmup
    area_all = torch.zeros(local_frames, dtype=torch.int32, device=dev)

    def step_dev(i: int, out: torch.Tensor | None):
        lo = (i % nb) * batch
        m = min(batch, local_frames - i * batch) if out is not None else batch
        _, _, a = model.run(clip_dev[lo:lo + m], want_mask=True)
        if out is not None:
            out[i * batch:i * batch + m] = a

    def finish(local_area: torch.Tensor):
        full = sharding.gather_area(local_area, total_frames)
        return ogl.kinematic_features_device(full), full

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up
    for i in range(warmup):
        step_dev(i, None)
    finish(area_all)
    barrier()

    # ---- device-resident timed region (value). Events between the launches are recorded INSIDE
    # this region (ogl_unet_set_profiling keeps the last 16 forwards), so the per-launch times
    # behind `roofline` come from the same steps, at the same clocks, as `value`. A forward replayed
    # as a CUDA graph (small batches) records no events: its launches are timed in a second, eager
    # loop of the same steps right after the timed region (stated in roofline.timed_in).
    graphed = model.use_graphs and batch <= model.graph_max_batch
    sampler = ClockSampler(local_rank)
    sampler.start()
    if not graphed:
        _native.check(lib.ogl_unet_set_profiling(model._handle, 1))
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    barrier()
    e0.record()
    for i in range(steps):
        step_dev(i, area_all)
    e1.record()
    feats, area_full = finish(area_all)
    e2.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e2)
    feature_ms = e1.elapsed_time(e2)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = total_frames / (ms * 1e-3)

    timed_in = "the timed steps"
    if graphed:
        use = model.use_graphs
        model.use_graphs = False
        _native.check(lib.ogl_unet_set_profiling(model._handle, 1))
        for i in range(min(steps, 32)):
            step_dev(i, None)
        torch.cuda.synchronize(dev)
        model.use_graphs = use
        timed_in = "an eager loop of the same steps after the timed region (the timed steps replay CUDA graphs)"
    roofline, layers, nl, peak_tf = read_launch_times(lib, model, hgt, wid, batch)
    roofline["timed_in"] = timed_in
    _native.check(lib.ogl_unet_set_profiling(model._handle, 0))

    # ---- end-to-end from pinned host memory through the public API: segment_clip streams the
    # batches (double-buffered H2D on a copy stream), masks are produced as in the value leg, the
    # areas are read back to the host; then the gather and the feature step.
    e2e_frames = local_frames
    host_frames = min(e2e_frames, 16384)
    reps = (host_frames + clip_frames - 1) // clip_frames
    host = (clip_host if reps == 1 else clip_host.repeat(reps, 1, 1).pin_memory())[:host_frames]
    ogl.segment_clip(host[:min(host_frames, 2 * batch)], model, batch=batch, want_masks=True)[0].cpu()   # warm
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record()
    parts, done = [], 0
    while done < e2e_frames:
        m = min(host_frames, e2e_frames - done)
        a, _ = ogl.segment_clip(host[:m], model, batch=batch, want_masks=True)
        parts.append(a.cpu())
        done += m
    area_host = torch.cat(parts)
    feats_e2e = ogl.kinematic_features_device(
        sharding.gather_area(area_host.to(dev, non_blocking=True), total_frames))
    t1.record()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - w0
    ms_e2e = max(t0.elapsed_time(t1), wall * 1e3)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    e2e_value = total_frames / (ms_e2e * 1e-3)
    k = min(batch, local_frames)
    same = bool(torch.equal(area_host[:k].to(dev), area_all[:k])) if rank == 0 else True

    act_gb = lib.ogl_unet_workspace_bytes(model._handle, batch, hgt, wid, 0) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {
            "workload": cfg["workload"], "config_index": args.config,
            "frames_per_step_per_gpu": batch, "height": hgt, "width": wid,
            "total_frames": total_frames,
            "weights": wname, "parallelism": f"frame-range shards x{world}, area all-gather",
            "l2": L2_NOTE.format(frames=clip_frames, mb=clip_frames * hgt * wid / 1e6, act=act_gb),
            "cuda_graphs": bool(graphed),
        },
        "e2e": {"value": e2e_value, "unit": "frames/s",
                "h2d_bytes_per_step": batch * hgt * wid, "d2h_bytes_per_step": batch * 4,
                "ms_total": ms_e2e, "matches_device_run": same,
                "note": "same masks + areas as the value leg; equal to `value` within box noise when "
                        "the copies hide behind the compute"},
        "gpu_launches": steps * (nl + 1) + 70,
        "feature_step_ms": feature_ms,
        "roofline": roofline,
        "clocks": clocks,
        "pct_of_tc_roofline": 100.0 * (value / world) * sum(layer_flops(hgt, wid)) / (peak_tf * 1e12),
        "features": {k_: (None if v is None else float(v)) for k_, v in (feats or {}).items()
                     if not k_.startswith("_")} if feats else None,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        secs = args.cpu_seconds
        fps, n, cores, kind = cpu_reference_fps(sd, synthetic_clip(64, seed=1, hgt=hgt, wid=wid),
                                                warmup=3, seconds=secs, max_frames=2000)
        line["cpu_baseline"] = {
            "value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
            "sample": f"{n} frames of the same synthetic {hgt}x{wid} clip through the reference's "
                      "per-frame loop (torch CPU fp32, batch 1; "
                      + ("the unmodified reference package from oracle/_ref)" if kind == "reference"
                         else "restated in oracle/)")}
        b32 = cpu_batched_fps(sd, synthetic_clip(32, seed=1, hgt=hgt, wid=wid), seconds=5.0)
        if b32 is not None:
            line["cpu_baseline"]["batch32"] = {
                "value": b32[0], "unit": "frames/s", "cores": cores,
                "sample": f"{b32[1]} frames, the reference's UNet module on the same host cores with 32 "
                          "frames per forward (fp32): what batching alone buys the CPU; the "
                          "reference's pipeline is the batch-1 loop above"}
        if strong:
            line["cpu_baseline"]["features"] = cpu_features_timing(area_full.cpu().numpy(), seconds=20.0)
            line["cpu_baseline"]["extrapolated_job_seconds"] = round(total_frames / fps, 1)
    if rank == 0:
        if args.layers_out:
            Path(args.layers_out).parent.mkdir(parents=True, exist_ok=True)
            Path(args.layers_out).write_text(json.dumps(layers, indent=1))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    # defaults (config 1): 20 + 100 steps = 1.3 s of device time. The board runs this path at its
    # power cap and needs a few hundred ms to settle its clocks: 3 + 20 steps measure a boost
    # transient (10.8-11.0 ms/step) rather than the sustained rate of a 100 000-frame clip.
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=None, help="config 4: frames of the whole job (10^6)")
    ap.add_argument("--impl", default="native", choices=["native", "reference", "cudnn"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="never replay forwards as CUDA graphs")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--layers-out", default=None,
                    help="write the per-launch times of this run to this JSON file (never written "
                         "by default: a run under ncu must not overwrite a real run's file)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "cudnn":
        run_cudnn(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
