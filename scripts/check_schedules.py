"""Small forwards through every kernel schedule (space-to-depth / direct, one CTA / CTA pairs,
fused / separate stem). Prints the areas; exits non-zero when the schedules that share their
arithmetic disagree. (compute-sanitizer is closed on this pool, so this is the cheap stand-in.)"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import openglottal_b200 as ogl  # noqa: E402
import synthdata  # noqa: E402

sd = synthdata.calibrated_state(0)
model = ogl.UNet().to("cuda")
model.load_state_dict(sd)
model.eval()
ok = True
for shape in ((3, 64, 96), (2, 48, 80), (5, 32, 32)):
    frames = torch.from_numpy(synthdata.glottis_clip(*shape, seed=5)[0]).cuda()
    ref = None
    for schedule, pairs, fuse in (("s2d", 1, True), ("s2d", 3, True), ("s2d", 3, False), ("direct", 3, False)):
        model.schedule, model.cta_pairs, model.fuse_stem = schedule, pairs, fuse
        _, mask, area = model.run(frames)
        torch.cuda.synchronize()
        if schedule == "s2d":
            if ref is None:
                ref = (mask.clone(), area.clone())
            elif not (torch.equal(mask, ref[0]) and torch.equal(area, ref[1])):
                ok = False
        print(shape, schedule, pairs, fuse, area.tolist())
wave = torch.arange(1000, dtype=torch.int32, device="cuda") % 37
print(ogl.kinematic_features_device(wave)["f0"])
sys.exit(0 if ok else 1)
