"""CPU: the oracle restatement is pinned against fixtures generated from the UNMODIFIED reference
(tests/golden/make_golden.py, run in the build container where /root/reference exists)."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

GOLDEN = Path(__file__).parent / "golden"


def test_forward_oracle_matches_reference_logits(calibrated_sd):
    from oracle import unet_oracle as uo

    g = np.load(GOLDEN / "unet_forward.npz")
    got = uo.ref_forward(calibrated_sd, uo.frames_to_input(g["frames"])).numpy()
    assert got.shape == g["logits"].shape
    # same torch ops in the same order; allow for a different CPU conv kernel selection
    assert np.abs(got - g["logits"]).max() <= 1e-4


def test_folded_fp32_matches_reference_logits(calibrated_sd):
    from oracle import unet_oracle as uo

    g = np.load(GOLDEN / "unet_forward.npz")
    got = uo.folded_forward(calibrated_sd, uo.frames_to_input(g["frames"]), bf16=False).numpy()
    assert np.abs(got - g["logits"]).max() <= 2e-4


def test_bitmodel_is_close_but_not_identical(calibrated_sd):
    from oracle import unet_oracle as uo

    g = np.load(GOLDEN / "unet_forward.npz")
    got = uo.folded_forward(calibrated_sd, uo.frames_to_input(g["frames"]), bf16=True).numpy()
    err = np.abs(got - g["logits"]).max()
    assert 1e-4 < err < 0.5


def test_segment_frame_oracle_matches_reference_masks(calibrated_sd):
    from oracle import unet_oracle as uo

    g = np.load(GOLDEN / "segment_frame.npz")
    for name, shape in (("256", (256, 256)), ("96", (96, 128))):
        frame = g["f" + name]
        ref = np.unpackbits(g["m" + name])[: shape[0] * shape[1]].reshape(shape).astype(bool)
        got = uo.segment_frame(calibrated_sd, frame) > 0
        assert got.shape == shape
        assert (got != ref).sum() <= 2, name     # borderline pixels only


def test_pipeline_oracle_matches_reference_features(calibrated_sd):
    import cv2
    from oracle import unet_oracle as uo
    from oracle.features_oracle import kinematic_features

    ref = json.loads((GOLDEN / "pipeline.json").read_text())
    cap = cv2.VideoCapture(str(GOLDEN / "pipeline_clip.avi"))
    frames = []
    while True:
        ok, frm = cap.read()
        if not ok:
            break
        frames.append(cv2.cvtColor(frm, cv2.COLOR_BGR2GRAY))
    cap.release()
    assert len(frames) == 30
    wave = uo.area_wave(calibrated_sd, frames)
    assert np.abs(np.array(wave) - np.array(ref["_area"])).max() <= 2
    feats = kinematic_features(ref["_area"])
    for k in ("area_mean", "area_std", "area_range", "open_quotient", "periodicity", "cv"):
        assert feats[k] == pytest.approx(ref[k], rel=1e-12, abs=1e-12), k
    assert feats["f0"] == ref["f0"]


def test_features_oracle_matches_reference_kats():
    from oracle.features_oracle import kinematic_features

    kats = json.loads((GOLDEN / "features_kat.json").read_text())
    assert len(kats) >= 10
    for case in kats:
        for exact in (False, True):
            if case["raises"]:
                with pytest.raises(ValueError):
                    kinematic_features(case["input"], exact_correlate=exact)
                continue
            got = kinematic_features(case["input"], exact_correlate=exact)
            ref = case["output"]
            if ref is None:
                assert got is None
                continue
            for k in ("area_mean", "area_std", "area_range", "open_quotient", "cv"):
                assert got[k] == pytest.approx(ref[k], rel=1e-13, abs=1e-13), k
            assert got["periodicity"] == pytest.approx(ref["periodicity"], rel=1e-9, abs=1e-12)
            assert got["f0"] == ref["f0"]


def test_survey_appendix_c_values():
    """Spot values quoted in SURVEY.md App. C (generated from the reference)."""
    from oracle.features_oracle import kinematic_features

    t = np.arange(500)
    f = kinematic_features(np.floor(np.maximum(0, 300 * np.sin(2 * np.pi * t / 20))))
    assert f["area_mean"] == 94.5 and f["f0"] == 0.05 and f["open_quotient"] == 0.45
    assert f["periodicity"] == pytest.approx(0.9599999999999985, rel=1e-12)
    f = kinematic_features([5.0, 7.0])
    assert f["f0"] is None and f["periodicity"] == pytest.approx(-0.4999999975, rel=1e-9)


def test_synthetic_fixtures_are_reproducible():
    from oracle import synth

    a = synth.calibrated_state(0)
    b = synth.calibrated_state(0)
    assert all(torch.equal(a[k], b[k]) for k in a) and len(a) == 118
    f1, m1 = synth.glottis_clip(3, 64, 64, seed=5)
    f2, _ = synth.glottis_clip(3, 64, 64, seed=5)
    assert np.array_equal(f1, f2) and set(np.unique(m1)) <= {0, 255}


# ------------------------------------------------------------------ callers around the U-Net
def _unpack(bits, shape):
    return np.unpackbits(bits)[: shape[0] * shape[1]].reshape(shape).astype(np.uint8) * 255


def test_letterbox_oracle_matches_reference():
    from oracle import pipeline_oracle as po

    g = np.load(GOLDEN / "crops.npz")
    n = sum(1 for k in g.files if k.startswith("crop"))
    assert n >= 8
    for k in range(n):
        crop = g[f"crop{k}"]
        boxed, pt, pl, ch, cw = po.letterbox_with_info(crop, 256, 0)
        assert [pt, pl, ch, cw] == g[f"geom{k}"].tolist()
        assert np.array_equal(boxed, g[f"boxed{k}"])
        mask_cs = _unpack(g[f"maskcs{k}"], (256, 256))
        back = po.unletterbox(mask_cs, pt, pl, ch, cw, *crop.shape)
        assert np.array_equal(back > 0, _unpack(g[f"back{k}"], crop.shape) > 0)


def test_letterbox_geometry_matches_reference():
    import openglottal_b200 as ogl

    g = np.load(GOLDEN / "crops.npz")
    for k in range(8):
        h, w = g[f"crop{k}"].shape
        assert list(ogl.letterbox_geometry(h, w, 256)) == g[f"geom{k}"].tolist()


def test_metric_oracle_matches_reference():
    from oracle import pipeline_oracle as po
    import openglottal_b200 as ogl

    for case in json.loads((GOLDEN / "metrics.json").read_text()):
        a = _unpack(np.array(case["a"], dtype=np.uint8), (48, 64))
        b = _unpack(np.array(case["b"], dtype=np.uint8), (48, 64))
        assert po.dice(a, b) == case["dice"] and po.iou(a, b) == case["iou"]
        assert ogl.dice(a, b) == pytest.approx(case["dice"], rel=1e-6)
        assert ogl.iou(a, b) == pytest.approx(case["iou"], rel=1e-6)


def test_gated_oracle_matches_reference_features(calibrated_sd):
    import cv2
    from oracle import pipeline_oracle as po, unet_oracle as uo
    from oracle.features_oracle import kinematic_features

    ref = json.loads((GOLDEN / "gated.json").read_text())
    boxes = [None if b is None else tuple(b) for b in ref["boxes"]]
    cap = cv2.VideoCapture(str(GOLDEN / "pipeline_clip.avi"))
    frames = []
    while True:
        ok, frm = cap.read()
        if not ok:
            break
        frames.append(cv2.cvtColor(frm, cv2.COLOR_BGR2GRAY))
    cap.release()
    masks = [uo.segment_frame(calibrated_sd, f) for f in frames]
    wave = po.gated_area_wave(masks, boxes)
    want = ref["features"]["_area"]
    assert np.abs(np.array(wave) - np.array(want)).max() <= 2
    assert all(w == 0.0 for w, b in zip(want, boxes) if b is None)
    feats = kinematic_features(want)
    for k in ("area_mean", "area_std", "open_quotient", "periodicity"):
        assert feats[k] == pytest.approx(ref["features"][k], rel=1e-12, abs=1e-12), k


def test_draw_overlay_matches_reference():
    """openglottal_b200.analysis.draw_overlay vs the reference's _draw_overlay (overlay.npz)."""
    from openglottal_b200.analysis import draw_overlay, features_row, FEATURE_COLS

    g = np.load(GOLDEN / "overlay.npz")
    for style in ("fill", "contour", "none"):
        got = draw_overlay(g["frame"], g["mask"], (30, 20, 100, 80), 1234.0, style)
        assert np.array_equal(got, g[f"out_{style}"]), style
    assert np.array_equal(draw_overlay(g["frame"], None, None, 0.0, "fill"), g["out_nomask"])
    row = features_row("clip", {"area_mean": 1.0, "f0": None, "cv": 2.0})
    assert len(row) == 1 + len(FEATURE_COLS) and row[0] == "clip" and row[5] == ""
    assert features_row("x", None) == ["x"] + [""] * 7


RESIZE_SIZES = [(512, 256), (256, 512), (512, 512), (96, 128), (300, 200), (1024, 1024), (128, 128),
                (257, 255), (480, 640), (64, 64), (256, 256), (320, 256), (256, 1000), (17, 33)]


def test_resize_oracle_u8_is_bit_exact_with_cv2():
    """oracle/resize_oracle.py restates cv2.resize INTER_LINEAR (utils.py:234) for u8 frames:
    bit-exact with cv2 for down-scales, up-scales, the 2x2 area substitution and the identity."""
    import cv2
    from oracle import resize_oracle as ro

    rng = np.random.default_rng(0)
    for hgt, wid in RESIZE_SIZES:
        img = rng.integers(0, 256, (hgt, wid), dtype=np.uint8)
        want = cv2.resize(img, (256, 256), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(ro.resize_u8_linear(img, 256, 256), want), (hgt, wid)


def test_resize_oracle_f32_matches_opencv_arithmetic():
    """f32 probability resize (utils.py:238-240): bit-exact with OpenCV's own INTER_LINEAR code
    (IPP switched off) and within 2e-5 of the IPP routine cv2 dispatches to by default."""
    import cv2
    from oracle import resize_oracle as ro

    rng = np.random.default_rng(1)
    use_ipp = cv2.ipp.useIPP()
    try:
        for hgt, wid in RESIZE_SIZES:
            pr = rng.random((256, 256), dtype=np.float32)
            got = ro.resize_f32_linear(pr, wid, hgt)
            cv2.ipp.setUseIPP(use_ipp)
            dflt = cv2.resize(pr, (wid, hgt), interpolation=cv2.INTER_LINEAR)
            assert np.abs(got - dflt).max() <= 2e-5, (hgt, wid)
            cv2.ipp.setUseIPP(False)
            own = cv2.resize(pr, (wid, hgt), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(got, own), (hgt, wid)
    finally:
        cv2.ipp.setUseIPP(use_ipp)


def test_resize_oracle_reproduces_reference_masks(calibrated_sd):
    """The restated resize pair inside unet_segment_frame's steps against the reference's OWN mask
    (segment_frame.npz: the 96x128 frame goes through both resizes)."""
    from oracle import resize_oracle as ro, unet_oracle as uo

    g = np.load(GOLDEN / "segment_frame.npz")
    frame = g["f96"]
    ref = np.unpackbits(g["m96"])[: 96 * 128].reshape(96, 128).astype(bool)
    small = ro.resize_u8_linear(frame, 256, 256)
    logits = uo.ref_forward(calibrated_sd, uo.frames_to_input(small[None]))[0, 0].numpy()
    got = ro.segment_frame_restated(logits, 96, 128) > 0
    assert (got != ref).sum() <= 2


def test_gaw_oracle_matches_reference_512x256(calibrated_sd):
    """gaw_512x256.json: the reference's extract_gaw_features (scripts/analyze_gaw.py:75-100) on
    BAGLS-shaped frames -- squash to 256 x 256, probability resized back, gated count, f0 in Hz."""
    from oracle import pipeline_oracle as po, synth, unet_oracle as uo
    from oracle.features_oracle import kinematic_features

    ref = json.loads((GOLDEN / "gaw_512x256.json").read_text())
    c = ref["clip"]
    clip, _ = synth.glottis_clip(c["n"], c["height"], c["width"], seed=c["seed"], period=c["period"])
    boxes = [None if b is None else tuple(b) for b in ref["boxes"]]
    masks = [uo.segment_frame(calibrated_sd, f) for f in clip]
    assert masks[0].shape == (c["height"], c["width"])
    wave = po.gated_area_wave(masks, boxes)
    want = ref["features"]
    assert np.abs(np.array(wave) - np.array(want["_area"])).max() <= 2
    feats = kinematic_features(want["_area"])
    assert feats["f0"] * ref["capture_fps"] == pytest.approx(want["f0"], rel=1e-12)
    for k in ("area_mean", "area_std", "open_quotient", "periodicity", "cv"):
        assert feats[k] == pytest.approx(want[k], rel=1e-12, abs=1e-12), k
