"""Tiny forward with the 16-warp tensor-core stem (fuse_stem = 3) against the fp32 CUDA-core stem."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import openglottal_b200 as ogl  # noqa: E402

sd, _ = bench.bench_state()
model = ogl.UNet().to("cuda")
model.load_state_dict(sd)
model.eval()
frames = torch.randint(0, 256, (4, 256, 256), dtype=torch.uint8).cuda()
model.fuse_stem = 1
ref = model.run(frames, want_logits=True)
torch.cuda.synchronize()
model.fuse_stem = 3
got = model.run(frames, want_logits=True)
torch.cuda.synchronize()
print("max|dz|", (ref[0] - got[0]).abs().max().item(), "mask diffs", (ref[1] != got[1]).sum().item())
