"""CUDA feature kernels vs the reference-generated known answers (tests/golden/features_kat.json)
and the NumPy oracle on random waveforms. Tolerance 1e-3 relative (north_star); f0 / None exact."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
KEYS = ("area_mean", "area_std", "area_range", "open_quotient", "periodicity", "cv")


def _close(a, b, rel=1e-3):
    return abs(a - b) <= rel * max(abs(b), 1e-9) + 1e-9


def _check(got, ref):
    if ref is None:
        assert got is None
        return
    assert got is not None
    for k in KEYS:
        assert _close(float(got[k]), float(ref[k])), (k, got[k], ref[k])
    assert (got["f0"] is None) == (ref["f0"] is None), (got["f0"], ref["f0"])
    if ref["f0"] is not None:
        assert got["f0"] == pytest.approx(ref["f0"], rel=1e-12)


def test_known_answers_from_reference(lib):
    import openglottal_b200 as ogl

    kats = json.loads((GOLDEN / "features_kat.json").read_text())
    for case in kats:
        wave = case["input"]
        if case["raises"]:
            with pytest.raises(ValueError):
                ogl._kinematic_features(wave)
            continue
        _check(ogl._kinematic_features(wave), case["output"])


@pytest.mark.parametrize("n,seed", [(2, 0), (3, 1), (50, 2), (51, 3), (500, 4), (4097, 5),
                                    (100000, 6), (1000000, 7)])
def test_random_waveforms_vs_oracle(lib, n, seed):
    import openglottal_b200 as ogl
    from oracle.features_oracle import kinematic_features

    rng = np.random.default_rng(seed)
    t = np.arange(n)
    period = rng.uniform(6, 40)
    wave = np.floor(np.maximum(0, 900 * np.sin(2 * np.pi * t / period) + rng.normal(0, 30, n)) + 50)
    area = torch.from_numpy(wave.astype(np.int32)).cuda()
    got = ogl.kinematic_features_device(area)
    ref = kinematic_features(wave)
    _check(got, ref)
    assert np.array_equal(got["_area"], wave) and got["_area"].dtype == np.float64
    # list entry point (float64 upload path when non-integral)
    if n <= 5000:
        _check(ogl._kinematic_features(list(wave + 0.25)), kinematic_features(wave + 0.25))


def test_silent_and_edge_cases(lib):
    import openglottal_b200 as ogl

    assert ogl._kinematic_features([0.0] * 10) is None
    assert ogl.kinematic_features_device(torch.zeros(64, dtype=torch.int32, device="cuda")) is None
    with pytest.raises(ValueError):
        ogl._kinematic_features([5.0])
    with pytest.raises(ValueError):
        ogl._kinematic_features([])
    const = ogl._kinematic_features([100.0] * 64)
    assert const["f0"] is None and const["periodicity"] == 0.0 and const["area_std"] == 0.0


def test_bgr_to_gray_matches_cv2(lib):
    import cv2
    import openglottal_b200 as ogl

    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 256, (5, 64, 48, 3), dtype=np.uint8)
    ref = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in bgr])
    got = ogl.bgr_to_gray(torch.from_numpy(bgr).cuda()).cpu().numpy()
    assert np.array_equal(got, ref)
