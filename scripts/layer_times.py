"""Per-launch times of the bf16 forward (CUDA events between launches on the launching stream),
averaged over a few steps of a 512-frame 256x256 batch. Usage: layer_times.py [batch] [steps] [tag]
Prints one JSON line; experiment switches come from the environment (OGL_DBG, OGL_NA, OGL_NW)."""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import openglottal_b200 as ogl  # noqa: E402
from openglottal_b200 import _native  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
tag = sys.argv[3] if len(sys.argv) > 3 else ""
hgt = int(os.environ.get("OGL_H", "256"))
wid = int(os.environ.get("OGL_W", "256"))
sd, _ = bench.bench_state()
model = ogl.UNet().to("cuda")
model.load_state_dict(sd)
model.eval()
model.max_batch = batch
model.schedule = "direct" if os.environ.get("OGL_S2D", "1") == "0" else "s2d"
lib = _native.load()
g = torch.Generator().manual_seed(0)
frames = torch.randint(0, 256, (4 * batch, hgt, wid), dtype=torch.uint8, generator=g).cuda()
for i in range(3):
    model.run(frames[(i % 4) * batch:(i % 4 + 1) * batch])
torch.cuda.synchronize()
sampler = bench.ClockSampler(0)
sampler.start()
_native.check(lib.ogl_unet_set_profiling(model._handle, 1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps * 8):
    model.run(frames[(i % 4) * batch:(i % 4 + 1) * batch])
e1.record()
torch.cuda.synchronize()
clocks = sampler.stop()
total = e0.elapsed_time(e1) / (steps * 8)
buf = (C.c_float * 64)()
cnt = C.c_int(0)
_native.check(lib.ogl_unet_layer_times(model._handle, buf, 64, C.byref(cnt)))   # last 16 steps
_native.check(lib.ogl_unet_set_profiling(model._handle, 0))
acc = np.array(buf[:cnt.value])
names = [lib.ogl_unet_launch_name(model._handle, i).decode() for i in range(cnt.value)]
print(json.dumps({"tag": tag, "env": {k: v for k, v in os.environ.items() if k.startswith("OGL_")},
                  "batch": batch, "clocks": clocks, "ms_step": total, "fps": batch / total * 1e3,
                  "layers": dict(zip(names, [round(float(x), 4) for x in acc]))}))
