"""Host-side decode scaling: frame ranges of an MJPG clip decoded by T threads of one process against
P forked processes writing into one shared-memory array. Usage: decode_scaling.py [frames=20000]"""
import multiprocessing as mp
import os
import sys
import tempfile
import threading
import time
from multiprocessing import shared_memory
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from openglottal_b200.utils import RangeDecoder  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
clip = Path(tempfile.gettempdir()) / f"ogl_scaling_{n}.avi"
if not clip.exists():
    rng = np.random.default_rng(0)
    base = (rng.integers(0, 255, (64, 256, 256), dtype=np.uint8) // 2 + 60)
    wr = cv2.VideoWriter(str(clip), cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (256, 256))
    for i in range(n):
        wr.write(cv2.cvtColor(cv2.GaussianBlur(base[i % 64], (9, 9), 3), cv2.COLOR_GRAY2BGR))
    wr.release()
shm = shared_memory.SharedMemory(create=True, size=n * 256 * 256 * 3)
arr = np.ndarray((n, 256, 256, 3), np.uint8, buffer=shm.buf)
arr[:] = 0   # fault the pages in


def work(lo, hi):
    d = RangeDecoder(str(clip))
    got = d.read_into(lo, hi, arr[lo:hi])
    d.release()
    return got


def child(name, lo, hi, q):
    s = shared_memory.SharedMemory(name=name)
    a = np.ndarray((n, 256, 256, 3), np.uint8, buffer=s.buf)
    d = RangeDecoder(str(clip))
    q.put(d.read_into(lo, hi, a[lo:hi]))
    d.release()
    s.close()


cores = len(os.sched_getaffinity(0))
for w in sorted({4, 8, 12, 16, cores // 2, cores}):
    if w > cores or w < 1:
        continue
    cut = [n * i // w for i in range(w + 1)]
    th = [threading.Thread(target=work, args=(cut[i], cut[i + 1])) for i in range(w)]
    t0 = time.perf_counter()
    [t.start() for t in th]
    [t.join() for t in th]
    t_thr = time.perf_counter() - t0
    ref = arr[::997].copy()
    arr[:] = 0
    q = mp.get_context("fork").Queue()
    ps = [mp.get_context("fork").Process(target=child, args=(shm.name, cut[i], cut[i + 1], q)) for i in range(w)]
    t0 = time.perf_counter()
    [p.start() for p in ps]
    done = sum(q.get() for _ in ps)
    [p.join() for p in ps]
    t_proc = time.perf_counter() - t0
    print(f"workers {w:3d}: threads {n / t_thr:8.0f} fps   forked processes {n / t_proc:8.0f} fps "
          f"(incl. fork)   same frames {bool(np.array_equal(ref, arr[::997]))} decoded {done}")
shm.close()
shm.unlink()
