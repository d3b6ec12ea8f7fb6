"""CPU: host-side logic -- C-ABI exports, state-dict tree, sharding (incl. world_size-2 gloo),
loud failures without a GPU. No compute calls into the library here."""
import ctypes as C
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(lib):
    from openglottal_b200 import _native

    header = (ROOT / "include" / "openglottal_b200.h").read_text()
    declared = set(re.findall(r"\b(ogl_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ogl_version() == 100


def test_struct_layout_matches_header():
    from openglottal_b200 import _native

    ptr = C.sizeof(C.c_void_p)
    assert C.sizeof(_native.ConvBN) == 5 * ptr
    assert C.sizeof(_native.ConvT) == 2 * ptr
    # 18 conv+bn, 4 convT, head w/b, eps (padded)
    assert C.sizeof(_native.UNetState) == (18 * 5 + 4 * 2 + 2) * ptr + ptr


def test_module_tree_matches_reference_state_dict(calibrated_sd):
    import openglottal_b200 as ogl

    m = ogl.UNet()
    sd = m.state_dict()
    assert len(sd) == 118
    assert sum(p.numel() for p in m.parameters()) == 7_762_465
    assert set(sd) == set(calibrated_sd)
    for k, v in calibrated_sd.items():
        assert sd[k].shape == v.shape, k
    m.load_state_dict(calibrated_sd, strict=True)
    keys = list(sd)
    assert keys[0] == "downs.0.net.0.weight" and keys[-1] == "head.bias"
    assert keys.index("ups.0.weight") < keys.index("bottleneck.net.0.weight") < keys.index("head.weight")
    assert sd["ups.0.weight"].shape == (512, 256, 2, 2)
    assert sd["ups.7.net.0.weight"].shape == (32, 64, 3, 3)


def test_no_cpu_fallback(lib):
    import openglottal_b200 as ogl

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.ogl_unet_create(C.byref(h), 0) != 0
    assert b"no CUDA device" in lib.ogl_last_error()
    m = ogl.UNet().eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 1, 32, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        ogl._kinematic_features([1.0, 2.0, 3.0])
    # reference-identical edge cases are decided before any device work
    assert ogl._kinematic_features([0.0] * 4) is None
    with pytest.raises(ValueError):
        ogl._kinematic_features([3.0])


def test_product_does_not_import_oracle():
    pkg = ROOT / "openglottal_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.h")):
        text = p.read_text()
        assert "import oracle" not in text and "from oracle" not in text, p


def test_shard_ranges_cover_and_are_contiguous():
    from openglottal_b200 import sharding

    for n in (0, 1, 7, 8, 9, 2000, 100_000, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            edges = [sharding.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and a <= b and c <= d
            assert max(b - a for a, b in edges) == sharding.shard_size(n, world) or n == 0
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["OGL_ROOT"])
from openglottal_b200 import sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["OGL_PORT"],
                        rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = sharding.rank_world()
for n in (9, 10, 1, 1001):
    full = (torch.arange(n, dtype=torch.int32) * 7 + 3)
    lo, hi = sharding.shard_range(n, rank, world)
    got = sharding.gather_area(full[lo:hi].clone(), n)
    assert torch.equal(got, full), (n, rank, got[:5])
# ranks agree on a branch before a collective (extract_features_unet's decode fallback)
assert sharding.agree_any(rank == world - 1, "cpu") is True
assert sharding.agree_any(False, "cpu") is False
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


@pytest.mark.parametrize("world", [2, 3])
def test_gather_area_gloo(world, tmp_path):
    port = 29500 + (os.getpid() % 2000) + world
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), OGL_PORT=str(port),
                   OGL_ROOT=str(ROOT), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, "-c", _GLOO_WORKER], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out.decode()
        assert b"ok" in out


def test_cli_rejects_out_of_scope_requests(capsys):
    from openglottal_b200 import cli

    with pytest.raises(SystemExit):
        cli.main(["run", "x.avi", "--pipeline", "vft", "--unet-weights", "w.pt"])
    with pytest.raises(SystemExit):
        cli.main(["run", "x.avi", "--pipeline", "unet-only"])          # weights required
    with pytest.raises(SystemExit):
        cli.main(["run", "x.avi", "--unet-weights", "w.pt", "--device", "cpu"])


def test_dice_matches_reference_definition():
    from openglottal_b200 import dice

    a = np.zeros((4, 4), np.uint8)
    assert dice(a, a) == 1.0
    b = a.copy(); b[0, :2] = 255
    c = a.copy(); c[0, 1:3] = 255
    assert dice(b, c) == pytest.approx(0.5)


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU loop restated in oracle/) runs without a
    GPU and prints one JSON line with the keys the driver reads."""
    import json

    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["metric"] == "unet_only_frames_per_sec_256x256_bf16" and line["higher_is_better"] is True
    # "reference": the unmodified package vendored into oracle/_ref by oracle/make_ref.py (what the
    # build container and the GPU box have); "port": the oracle restatement when it is absent
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["config_index"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"]


def test_batched_cpu_figure_of_the_bench(calibrated_sd):
    """bench.py's cpu_baseline.batch32 leg (a batched fp32 CPU forward beside the
    reference's batch-1 loop): runs the reference module from oracle/_ref and returns (fps, frames);
    None when the reference package is not vendored."""
    sys.path.insert(0, str(ROOT))
    import bench
    from oracle import synth

    frames, _ = synth.glottis_clip(2, 64, 64, seed=5)
    got = bench.cpu_batched_fps(calibrated_sd, frames, seconds=0.01, batch=2)
    if (ROOT / "oracle" / "_ref" / "openglottal").is_dir():
        assert got is not None and got[0] > 0 and got[1] >= 2 and got[1] % 2 == 0
    else:
        assert got is None
    # the algorithmic FLOPs the roofline is credited with: SURVEY App. A (reference formulation)
    assert sum(bench.module_flops(256, 256).values()) == 24_066_916_352
    assert sum(bench.module_flops(512, 256).values()) == 48_133_832_704


@pytest.mark.parametrize("fourcc", ["MJPG", "FFV1"])
def test_parallel_decode_equals_the_sequential_loop(tmp_path, fourcc, write_clip):
    """Frame ranges decoded by independent VideoCapture instances (intra-only codecs: frame-exact
    seeks) are the frames of the reference's sequential loop (utils.py:43-54), for worker counts
    that do and do not divide the frame count; short clips and other codecs take that loop."""
    import numpy as np

    from openglottal_b200 import utils

    clip = tmp_path / f"clip_{fourcc}.avi"
    write_clip(clip, fourcc, 103)
    want = utils.load_frames_bgr(str(clip))
    assert len(want) == 103
    info = utils.video_info(str(clip))
    assert info["frames"] == 103 and (info["height"], info["width"]) == (64, 48)
    for workers in (2, 3, 8):
        got = utils.load_frames_bgr_parallel(str(clip), workers=workers, min_frames=16)
        assert len(got) == len(want)
        assert all(np.array_equal(a, b) for a, b in zip(got, want)), workers
    # ranges decoded in place, out of order, by one decoder (the streaming path's access pattern)
    dec = utils.RangeDecoder(str(clip))
    buf = np.empty((20, 64, 48, 3), np.uint8)
    for start in (60, 5, 83):
        assert dec.read_into(start, start + 20, buf) == 20
        assert all(np.array_equal(buf[i], want[start + i]) for i in range(20))
    assert dec.read_into(95, 115, buf) == 8        # past the end: a short count, not an error
    dec.release()
    assert not utils.parallel_decodable({"frames": 5000, "height": 64, "width": 48, "fourcc": "avc1"}, 8)
    assert not utils.parallel_decodable({"frames": 0, "height": 64, "width": 48, "fourcc": "MJPG"}, 8)
