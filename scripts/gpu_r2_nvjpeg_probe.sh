# nvJPEG feasibility: decode rate of an OpenCV-written 256x256 MJPG clip per backend, difference to cv2's frames
python - <<'PY'
import sys, cv2, numpy as np
sys.path.insert(0, '.')
import bench
base = bench.synthetic_clip(2000, seed=5)
wr = cv2.VideoWriter('/tmp/probe.avi', cv2.VideoWriter_fourcc(*'MJPG'), 25.0, (256, 256))
for i in range(16384):
    wr.write(cv2.cvtColor(base[i % 2000], cv2.COLOR_GRAY2BGR))
wr.release()
cap = cv2.VideoCapture('/tmp/probe.avi'); ok, f = cap.read(); f.tofile('/tmp/probe_f0.raw'); print('cv2 frame0', f.shape, f.mean())
PY
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
timeout 300 scripts/microbench/bin/nvjpeg_probe /tmp/probe.avi /tmp/probe_f0.raw
