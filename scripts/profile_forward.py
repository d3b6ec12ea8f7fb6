"""One warm-up forward and one measured forward of a 512-frame 256x256 batch (22 launches each):
the command ncu wraps for the launch list and the --set full capture."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import openglottal_b200 as ogl  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sd, _ = bench.bench_state()
model = ogl.UNet().to("cuda")
model.load_state_dict(sd)
model.eval()
frames = torch.from_numpy(bench.synthetic_clip(batch, seed=1)).cuda()
for _ in range(2):
    _, mask, area = model.run(frames)
torch.cuda.synchronize()
print("ok", int(area.sum()))
