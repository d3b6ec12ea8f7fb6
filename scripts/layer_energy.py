"""Energy and time of every launch of the bf16 forward, from the board's energy counter.

The forward runs under the board's power cap (clocks.reasons = sw_power_cap in every bench line),
so a step costs (energy per step) / (power cap): what is left to gain is energy, not pipe time.
This script attributes the energy: with ogl_unet_set_repeat one launch is enqueued R times per
forward; (E(R) - E(1)) / (R - 1) is its energy, the same difference of the loop times its duration
at the clocks of that mix. torch.matmul (8192^3 bf16) and a device-to-device copy are measured
the same way as yardsticks (pJ/FLOP of a dense GEMM, pJ/byte of HBM traffic).

Usage: layer_energy.py [batch] [seconds per measurement] [repeat] [launch,launch,...]
-> one JSON document on stdout (all launches when no list is given; experiment switches such as
OGL_FUSE_STEM come from the environment)
"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import pynvml
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import openglottal_b200 as ogl  # noqa: E402
from openglottal_b200 import _native  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
R = int(sys.argv[3]) if len(sys.argv) > 3 else 9
only = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else None

pynvml.nvmlInit()
nv = pynvml.nvmlDeviceGetHandleByIndex(0)


def energy_mj() -> int:
    return pynvml.nvmlDeviceGetTotalEnergyConsumption(nv)


def measure(fn, seconds):
    """(J per call, ms per call, mean W, SM MHz at the end) of fn over about `seconds`."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    one = max(time.perf_counter() - t0, 1e-5)
    n = max(4, int(seconds / one))
    for _ in range(n // 4):          # bring the board to the steady clocks of this mix
        fn()
    torch.cuda.synchronize()
    e0, t0 = energy_mj(), time.perf_counter()
    for _ in range(n):
        fn()
    mhz = pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM)
    torch.cuda.synchronize()
    e1, t1 = energy_mj(), time.perf_counter()
    joule = (e1 - e0) * 1e-3
    return joule / n, (t1 - t0) / n * 1e3, joule / (t1 - t0), mhz


sd, _ = bench.bench_state()
model = ogl.UNet().to("cuda")
model.load_state_dict(sd)
model.eval()
model.max_batch = batch
lib = _native.load()
g = torch.Generator().manual_seed(0)
frames = torch.randint(0, 256, (4 * batch, 256, 256), dtype=torch.uint8, generator=g).cuda()
# realistic activations: the synthetic glottis clip, not noise (toggle rates drive the MMA power)
clip = torch.from_numpy(bench.synthetic_clip(4 * batch, seed=1)).cuda()
state = {"i": 0, "area": True}


def forward(src=clip):
    i = state["i"] = (state["i"] + 1) % 4
    model.run(src[i * batch:(i + 1) * batch], want_area=state["area"])


forward()
torch.cuda.synchronize()
names = [lib.ogl_unet_launch_name(model._handle, i).decode() for i in range(lib.ogl_unet_launch_count(model._handle))]
mfl = bench.module_flops(256, 256)
flops = [bench.launch_flops(n_, mfl) * batch for n_ in names]

time.sleep(1.0)
e0, t0 = energy_mj(), time.perf_counter()
time.sleep(1.0)
idle_w = (energy_mj() - e0) * 1e-3 / (time.perf_counter() - t0)

base_j, base_ms, base_w, base_mhz = measure(forward, 2 * seconds)
noise_j, noise_ms, noise_w, noise_mhz = measure(lambda: forward(frames), 2 * seconds)
rows = []
for i, name in enumerate(names):
    if only is not None and i not in only:
        continue
    _native.check(lib.ogl_unet_set_repeat(model._handle, i, R))
    state["area"] = "head" not in name     # a repeated head launch must not accumulate areas
    j, ms, w, mhz = measure(forward, seconds)
    state["area"] = True
    lj, lms = (j - base_j) / (R - 1), (ms - base_ms) / (R - 1)
    rows.append({"layer": name, "joule": lj, "ms": lms, "watt": lj / (lms * 1e-3) if lms > 0 else None,
                 "sm_mhz": mhz, "tflop": flops[i] / 1e12,
                 "pj_per_flop": lj / flops[i] * 1e12 if flops[i] else None})
_native.check(lib.ogl_unet_set_repeat(model._handle, -1, 1))

# yardsticks
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
c = torch.empty(8192, 8192, device="cuda", dtype=torch.bfloat16)
gj, gms, gw, gmhz = measure(lambda: torch.matmul(a, b, out=c), seconds)
ar = torch.relu(a)          # half of the entries zero, like post-ReLU activations
rj, rms, rw, rmhz = measure(lambda: torch.matmul(ar, b, out=c), seconds)
src = torch.empty(1 << 30, device="cuda", dtype=torch.uint8)
dst = torch.empty_like(src)
cj, cms, cw, cmhz = measure(lambda: dst.copy_(src), seconds)
gemm_flop = 2 * 8192 ** 3
print(json.dumps({
    "batch": batch, "repeat": R, "idle_watt": idle_w,
    "forward": {"joule": base_j, "ms": base_ms, "watt": base_w, "sm_mhz": base_mhz,
                "pj_per_flop": base_j / sum(flops) * 1e12, "data": "synthetic glottis clip"},
    "forward_noise_frames": {"joule": noise_j, "ms": noise_ms, "watt": noise_w, "sm_mhz": noise_mhz},
    "layers": rows,
    "sum_of_layers": {"joule": sum(r["joule"] for r in rows), "ms": sum(r["ms"] for r in rows)},
    "yardsticks": {
        "matmul_8192_bf16": {"joule": gj, "ms": gms, "watt": gw, "sm_mhz": gmhz,
                             "tflops": gemm_flop / gms / 1e9, "pj_per_flop": gj / gemm_flop * 1e12},
        "matmul_8192_bf16_relu_a": {"joule": rj, "ms": rms, "watt": rw, "sm_mhz": rmhz,
                                    "tflops": gemm_flop / rms / 1e9, "pj_per_flop": rj / gemm_flop * 1e12},
        "copy_1gib": {"joule": cj, "ms": cms, "watt": cw, "sm_mhz": cmhz,
                      "gbs": 2 * (1 << 30) / cms / 1e6, "pj_per_byte": cj / (2 * (1 << 30)) * 1e12},
    },
}, indent=1))
