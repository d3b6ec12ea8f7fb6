"""extract_features_unet on a video FILE under torch.distributed (one process per GPU): every rank
decodes and segments only its frame range, the areas are all-gathered. Checks the result against
the single-process staged run and reports the file -> features rate.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
        --master-port 29531 scripts/dist_extract_check.py [frames=40000]"""
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import cv2
import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import openglottal_b200 as ogl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
sd, _ = bench.bench_state()
model = ogl.UNet().to(dev)
model.load_state_dict(sd)
model.eval()

clip = Path(tempfile.gettempdir()) / f"ogl_dist_clip_{n}.avi"
if rank == 0:
    base = bench.synthetic_clip(2000, seed=5)
    wr = cv2.VideoWriter(str(clip), cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (256, 256))
    for i in range(n):
        wr.write(cv2.cvtColor(base[i % 2000], cv2.COLOR_GRAY2BGR))
    wr.release()
if world > 1:
    dist.barrier()
ogl.extract_features_unet(str(clip), None, model)                 # warm (page cache, workspace, NCCL)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
got = ogl.extract_features_unet(str(clip), None, model)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
secs = time.perf_counter() - t0
# single-process staged run of a prefix on every rank: same areas as the sharded run
m = min(n, 3000)
frames = ogl.load_frames_bgr(str(clip))[:m] if rank == 0 else None
ok = True
if rank == 0:
    gray = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in frames])
    want, _ = ogl.masks_for_clip(torch.from_numpy(gray).to(dev), model)
    ok = bool(np.array_equal(got["_area"][:m], want.cpu().numpy().astype(np.float64)))
    print(json.dumps({"world": world, "frames": n, "seconds": round(secs, 3), "file_to_features_fps": round(n / secs),
                      "areas_equal_single_process_prefix": ok, "len_area": int(len(got["_area"])),
                      "f0": got["f0"], "decode_workers_per_rank": max(1, ogl.utils.decode_workers() // world)}))
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
