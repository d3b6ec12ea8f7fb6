"""AVI -> features wall time with a per-stage split (after the GPU
path, video decode is the end-to-end bottleneck). Writes a synthetic 256x256 clip, then times
  * the reference's sequential cv2.VideoCapture loop (utils.py:43-54) on a prefix,
  * decode_gray_clip (threads decoding frame ranges into pinned chunks, H2D + BGR->gray on the GPU),
  * the U-Net + area + kinematic features on the resident gray clip,
  * extract_features_unet(path) end to end,
and reports which hardware decoders the box offers (libnvcuvid / libnvjpeg).
Usage: python scripts/ingest_bench.py [frames=20000] [fourcc=MJPG] [workers=auto]"""
import ctypes.util
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import cv2
import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import openglottal_b200 as ogl  # noqa: E402
from openglottal_b200 import utils  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
fourcc = sys.argv[2] if len(sys.argv) > 2 else "MJPG"
workers = int(sys.argv[3]) if len(sys.argv) > 3 else None
dev = torch.device("cuda:0")

sd, _ = bench.bench_state()
model = ogl.UNet().to(dev)
model.load_state_dict(sd)
model.eval()

tmp = Path(tempfile.mkdtemp(prefix="ogl_ingest_"))
clip = tmp / f"clip_{fourcc}.avi"
base = bench.synthetic_clip(2000, seed=5)
t0 = time.perf_counter()
wr = cv2.VideoWriter(str(clip), cv2.VideoWriter_fourcc(*fourcc), 25.0, (256, 256))
for i in range(n):
    wr.write(cv2.cvtColor(base[i % 2000], cv2.COLOR_GRAY2BGR))
wr.release()
write_s = time.perf_counter() - t0
size = clip.stat().st_size

# reference loop on a prefix (it is O(n) and slow: time 4000 frames)
cap = cv2.VideoCapture(str(clip))
t0 = time.perf_counter()
m = 0
while m < min(n, 4000):
    ok, _f = cap.read()
    if not ok:
        break
    m += 1
seq_fps = m / (time.perf_counter() - t0)
cap.release()

ogl.decode_gray_clip(str(clip), dev, workers=workers)          # warm (page cache, pinned pools)
ogl.masks_for_clip(torch.zeros((1024, 256, 256), dtype=torch.uint8, device=dev), model)   # workspace
torch.cuda.synchronize()
tim = {}
gray = ogl.decode_gray_clip(str(clip), dev, workers=workers, timings=tim)
torch.cuda.synchronize()
t0 = time.perf_counter()
area, _ = ogl.masks_for_clip(gray, model)
feats = ogl.kinematic_features_device(area)
torch.cuda.synchronize()
seg_s = time.perf_counter() - t0
del gray
t0 = time.perf_counter()
feats_e2e = ogl.extract_features_unet(str(clip), None, model)
torch.cuda.synchronize()
e2e_s = time.perf_counter() - t0
# the same streaming loop by hand, with the decoder's own statistics and per-chunk wall times
from openglottal_b200.features import iter_gray_chunks  # noqa: E402
st, marks = {}, []
t0 = time.perf_counter()
area2 = torch.empty(n, dtype=torch.int32, device=dev)
for i0, part in iter_gray_chunks(str(clip), dev, workers=workers, stats=st):
    t1 = time.perf_counter()
    area2[i0:i0 + part.shape[0]] = ogl.masks_for_clip(part, model)[0]
    marks.append((round(t1 - t0, 4), round(time.perf_counter() - t1, 4)))
torch.cuda.synchronize()
stream_s = time.perf_counter() - t0
same = all(feats[k] == feats_e2e[k] for k in ("area_mean", "area_std", "f0", "periodicity"))

def have(lib):
    if ctypes.util.find_library(lib):
        return True
    out = subprocess.run("ldconfig -p", shell=True, capture_output=True, text=True).stdout
    return f"lib{lib}.so" in out

try:
    cores = len(os.sched_getaffinity(0))
except AttributeError:
    cores = os.cpu_count()
print(json.dumps({
    "clip": {"frames": n, "fourcc": fourcc, "bytes_per_frame": round(size / n), "write_fps": round(n / write_s)},
    "host_cores": cores, "decode_workers": tim.get("workers"), "decode_mode": tim.get("mode"),
    "reference_sequential_decode_fps": round(seq_fps),
    "decode_gray_clip": {"fps": round(n / tim["total_s"]), "seconds": round(tim["total_s"], 3),
                         "decode_wait_seconds": round(tim["decode_s"], 3),
                         "h2d_gray_hidden_seconds": round(tim["total_s"] - tim["decode_s"], 3)},
    "unet_area_features_on_resident_clip": {"fps": round(n / seg_s), "seconds": round(seg_s, 3)},
    "extract_features_unet_end_to_end": {"fps": round(n / e2e_s), "seconds": round(e2e_s, 3),
                                         "same_features_as_staged_run": bool(same)},
    "streaming_loop": {"seconds": round(stream_s, 3), "decode_wait_seconds": round(st["decode_s"], 3),
                       "chunk_yield_time_and_enqueue_seconds": marks[:6]},
    "bottleneck": "decode" if tim["total_s"] > seg_s else "unet",
    "hardware_decoders": {"libnvcuvid": have("nvcuvid"), "libnvjpeg": have("nvjpeg"),
                          "cv2_cudacodec": hasattr(cv2, "cudacodec")},
    "features": {k: feats[k] for k in ("area_mean", "f0", "periodicity")},
}))
clip.unlink()
