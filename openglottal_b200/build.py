"""Builds libopenglottal_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libopenglottal_b200.so"
SOURCES = ["api.cu", "conv_tc.cu", "conv_tc_f16.cu", "s2d_tc.cu", "s2d_tc_f16.cu", "upcat_tc.cu",
           "upcat_tc_f16.cu", "stem_f32.cu", "features.cu", "frame_ops.cu", "resize.cu"]
# units that #include another source (the f16 twins)
INCLUDES = {"conv_tc_f16.cu": ["conv_tc.cu"], "s2d_tc_f16.cu": ["s2d_tc.cu"],
            "upcat_tc_f16.cu": ["upcat_tc.cu"]}
HEADERS = ["internal.h", "ptx.cuh", "../../include/openglottal_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
OBJ_DIR = LIB_DIR / "obj"


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libopenglottal_b200.so")


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    built = LIB_PATH.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [(CSRC / h).resolve() for h in HEADERS]
    return any(d.stat().st_mtime > built for d in deps)


def _compile_one(args) -> tuple[str, int, str]:
    src, obj, verbose = args
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", "-o", str(obj), str(src)]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return " ".join(cmd), res.returncode, res.stdout + res.stderr


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile the CUDA sources into ``openglottal_b200/lib/libopenglottal_b200.so``: one nvcc per
    translation unit (in parallel; only the units older than their sources or the headers), then
    one link step."""
    if not force and not is_stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    hdr_time = max((CSRC / h).resolve().stat().st_mtime for h in HEADERS)
    jobs, objs = [], []
    for s in SOURCES:
        src, obj = CSRC / s, OBJ_DIR / (Path(s).stem + ".o")
        objs.append(obj)
        newest = max([src.stat().st_mtime, hdr_time] +
                     [(CSRC / i).stat().st_mtime for i in INCLUDES.get(s, [])])
        if force or not obj.exists() or obj.stat().st_mtime < newest:
            jobs.append((src, obj, verbose))
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        for cmd, rc, log in pool.map(_compile_one, jobs):
            if rc != 0:
                raise RuntimeError(f"nvcc failed:\n{cmd}\n{log}")
            if verbose:
                print(log)
    tmp = LIB_PATH.with_suffix(".so.tmp")
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp),
           *[str(o) for o in objs]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
