"""Timeline of CTA 0 of the fused-stem launch (OGL_TRACE, s2d_tc.cu): clock64 of every hand-off of
the first 96 tiles, per role. Prints, for tiles in steady state, each event relative to the tile's
main-MMA issue and the per-tile period of every role. Usage: stem_trace.py [batch=128]"""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

path = os.path.join(tempfile.gettempdir(), "ogl_trace.bin")
os.environ["OGL_TRACE"] = path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import openglottal_b200 as ogl  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
which = sys.argv[2] if len(sys.argv) > 2 else "stem"
os.environ["OGL_TRACE_LAUNCH"] = which
sd, _ = bench.bench_state()
model = ogl.UNet().to("cuda")
model.load_state_dict(sd)
model.eval()
model.max_batch = batch
model.use_graphs = False
frames = torch.from_numpy(bench.synthetic_clip(batch, seed=1)).cuda()
for _ in range(3):
    model.run(frames)
torch.cuda.synchronize()
t = np.fromfile(path, dtype=np.uint64).reshape(8, 96, 8).astype(np.int64)
roles = ["build w0", "drain g0", "stem issuer", "main issuer 0", "main issuer 1", "epilogue g0",
         "epilogue g1", "u8 TMA"]
ev = {0: ["u8_full", "sa_empty", "done"],
      1: ["a_empty", "sd_full A", "ld A", "-", "sd_full B", "ld B", "stored"],
      2: ["sa_full", "slots k0-3", "issued k0-3", "slots k4-5", "issued k4-5"],
      3: ["acc_empty", "a_full s0", "issued s0", "a_full s1", "issued s1"],
      4: ["acc_empty", "a_full s0", "issued s0", "a_full s1", "issued s1"],
      5: ["acc_full", "done"], 6: ["acc_full", "done"], 7: ["u8_empty"]}
if which != "stem":     # TMA producer instead of the stem roles; up to three stages per tile
    roles[7] = "producer"
    ev[7] = ["slot s0", "slot s1", "slot s2"]
    for r in (3, 4):
        ev[r] = ["acc_empty", "a_full s0", "issued s0", "a_full s1", "issued s1", "-", "-", "-"]
t0 = t[t > 0].min()
print("per-tile period (cycles), tiles 20..80, by role (first event of each tile):")
for r, name in enumerate(roles):
    x = t[r, :, 0]
    idx = [i for i in range(20, 80) if x[i] > 0 and x[i - 1] > 0]
    if len(idx) > 2:
        # issuers / epilogue groups see every other tile: their index is the tile number `li`
        d = np.diff(x[x > 0])
        print(f"  {name:14s} events {int((x > 0).sum()):3d}  median delta between consecutive recorded tiles {int(np.median(d))}")
print("timeline of tiles 40..45 (cycles since the first event):")
for tile in range(40, 46):
    print(f" tile {tile}")
    for r, name in enumerate(roles):
        vals = [(ev[r][e], int(t[r, tile, e] - t0)) for e in range(len(ev[r])) if t[r, tile, e] > 0]
        if vals:
            print(f"   {name:14s} " + "  ".join(f"{k}={v}" for k, v in vals))
