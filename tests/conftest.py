import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
CACHE = GOLDEN / "_cache"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library, built on demand (nvcc cross-compiles without a GPU)."""
    from openglottal_b200 import _native, build

    build.build_library()
    return _native.load()


@pytest.fixture(scope="session")
def trained_sd():
    """Synthetically trained state dict, cached under tests/golden/_cache."""
    from oracle import synth

    return synth.trained_state(0, steps=150, size=128, batch=8, cache_dir=CACHE)


@pytest.fixture(scope="session")
def calibrated_sd():
    from oracle import synth

    return synth.calibrated_state(0)


@pytest.fixture(scope="session")
def native_model(lib, trained_sd):
    import torch
    import openglottal_b200 as ogl

    m = ogl.UNet().to("cuda")
    m.load_state_dict(trained_sd, strict=True)
    m.eval()
    return m


@pytest.fixture(scope="session")
def write_clip():
    """``write_clip(path, fourcc, n, hgt=64, wid=48)``: a synthetic glottis clip as a video file."""
    def _write(path, fourcc: str, n: int, hgt: int = 64, wid: int = 48):
        import cv2

        from oracle import synth

        frames, _ = synth.glottis_clip(n, hgt, wid, seed=11, period=9.0)
        wr = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*fourcc), 25.0, (wid, hgt))
        assert wr.isOpened()
        for f in frames:
            wr.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
        wr.release()

    return _write
