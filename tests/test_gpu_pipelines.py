"""GPU parity of the callers around the U-Net against the
oracle (oracle/pipeline_oracle.py, pinned by tests/golden/crops.npz, metrics.json, gated.json):
detection-gated area, the yolo-crop+unet letterbox / un-letterbox geometry, batched Dice / IoU.
All integer / byte work: bit-exact."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


def _boxes(n, hgt, wid, seed, none_every=5):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        if none_every and i % none_every == 2:
            out.append(None)
            continue
        x1, y1 = int(rng.integers(0, wid - 2)), int(rng.integers(0, hgt - 2))
        x2, y2 = int(rng.integers(x1 + 1, wid + 1)), int(rng.integers(y1 + 1, hgt + 1))
        out.append((x1, y1, x2, y2))
    return out


def test_gated_area_bit_exact(lib):
    import openglottal_b200 as ogl
    from oracle import pipeline_oracle as po

    rng = np.random.default_rng(3)
    masks = (rng.random((37, 96, 128)) > 0.7).astype(np.uint8) * 255
    boxes = _boxes(37, 96, 128, seed=4)
    # Python slice semantics: negative, reversed, out-of-range and empty boxes
    boxes[0] = (-20, -10, 128, 96)
    boxes[1] = (5, 5, 5, 50)
    boxes[3] = (100, 90, 400, 300)
    boxes[4] = (60, 40, 30, 20)
    boxes[5] = (0, 0, -1, -1)
    got = ogl.gated_area(torch.from_numpy(masks).cuda(), boxes).cpu().numpy()
    want = np.array(po.gated_area_wave(masks, boxes))
    assert np.array_equal(got, want.astype(np.int64))
    full = ogl.gated_area(torch.from_numpy(masks).cuda(), [(0, 0, 128, 96)] * 37).cpu().numpy()
    assert np.array_equal(full, (masks > 0).reshape(37, -1).sum(1))


def test_letterbox_and_unletterbox_match_golden(lib):
    """The reference's own letterbox_with_info / unletterbox outputs (crops.npz)."""
    import openglottal_b200 as ogl

    g = np.load(GOLDEN / "crops.npz")
    for k in range(8):
        crop = g[f"crop{k}"]
        h, w = crop.shape
        hgt, wid = 352, 352
        frame = np.random.default_rng(k).integers(0, 256, (hgt, wid), dtype=np.uint8)
        y1, x1 = 7 + k, 3 + 2 * k
        frame[y1:y1 + h, x1:x1 + w] = crop
        box = (x1, y1, x1 + w, y1 + h)
        boxed, geom = ogl.letterbox_crops(torch.from_numpy(frame[None]).cuda(), [box], 256)
        assert geom[0, 4:].tolist() == g[f"geom{k}"].tolist()
        assert np.array_equal(boxed[0].cpu().numpy(), g[f"boxed{k}"]), k
        mask_cs = np.unpackbits(g[f"maskcs{k}"])[:256 * 256].reshape(256, 256).astype(np.uint8) * 255
        area, full = ogl.unletterbox_area(torch.from_numpy(mask_cs[None]).cuda(), geom, hgt, wid,
                                          want_full=True)
        back = np.unpackbits(g[f"back{k}"])[:h * w].reshape(h, w).astype(np.uint8) * 255
        want_full = np.zeros((hgt, wid), np.uint8)
        want_full[y1:y1 + h, x1:x1 + w] = back
        assert np.array_equal(full[0].cpu().numpy(), want_full), k
        assert int(area[0]) == int((back > 0).sum())


def test_letterbox_random_boxes_match_oracle(lib):
    import openglottal_b200 as ogl
    from oracle import pipeline_oracle as po

    rng = np.random.default_rng(9)
    n, hgt, wid = 24, 200, 312
    frames = rng.integers(0, 256, (n, hgt, wid), dtype=np.uint8)
    boxes = _boxes(n, hgt, wid, seed=10)
    boxes[0] = (10, 10, 10, 80)          # empty crop -> treated like no box
    boxed, geom = ogl.letterbox_crops(torch.from_numpy(frames).cuda(), boxes, 256)
    boxed = boxed.cpu().numpy()
    masks_cs = (rng.random((n, 256, 256)) > 0.5).astype(np.uint8) * 255
    area, full = ogl.unletterbox_area(torch.from_numpy(masks_cs).cuda(), geom, hgt, wid, want_full=True)
    area, full = area.cpu().numpy(), full.cpu().numpy()
    for i, b in enumerate(boxes):
        it = iter([masks_cs[i]])
        want_area, want_full = po.crop_unet_frame(frames[i], b, lambda boxed_ref: next(it))
        if want_full is None:
            assert not boxed[i].any() and area[i] == 0 and not full[i].any()
            continue
        x1, y1, x2, y2 = b
        ref_boxed = po.letterbox_with_info(frames[i][y1:y2, x1:x2], 256, 0)[0]
        assert np.array_equal(boxed[i], ref_boxed), i
        assert area[i] == want_area and np.array_equal(full[i], want_full), i


def test_dice_iou_batch_matches_reference_metrics(lib):
    import openglottal_b200 as ogl
    from oracle import pipeline_oracle as po

    cases = json.loads((GOLDEN / "metrics.json").read_text())
    a = np.stack([np.unpackbits(np.array(c["a"], np.uint8))[:48 * 64].reshape(48, 64) for c in cases]).astype(np.uint8) * 255
    b = np.stack([np.unpackbits(np.array(c["b"], np.uint8))[:48 * 64].reshape(48, 64) for c in cases]).astype(np.uint8) * 7
    d, j = ogl.dice_iou_batch(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    for k, c in enumerate(cases):
        assert d[k] == c["dice"] and j[k] == c["iou"], k
    rng = np.random.default_rng(1)
    p = (rng.random((40, 256, 256)) > 0.8).astype(np.uint8) * 255
    q = (rng.random((40, 256, 256)) > 0.8).astype(np.uint8)
    d, j = ogl.dice_iou_batch(torch.from_numpy(p).cuda(), torch.from_numpy(q).cuda())
    for k in range(40):
        assert d[k] == po.dice(p[k], q[k]) and j[k] == po.iou(p[k], q[k])


class ScriptedDetector:
    def __init__(self, boxes):
        self.boxes, self.i, self.resets = boxes, 0, 0

    def reset(self):
        self.i = 0
        self.resets += 1

    def detect(self, frame_bgr):
        b = self.boxes[self.i]
        self.i += 1
        return b


def test_gated_pipeline_matches_reference_golden(lib, calibrated_sd):
    """extract_features_unet(clip, detector, ...) against the reference's own output (gated.json);
    64x64 clip -> the reference-resize path. The golden was made with the variance-calibrated
    RANDOM weights, which amplify bf16 rounding, so the network runs in
    the fp32 validation mode here: this test is about the pipeline semantics around it."""
    import openglottal_b200 as ogl

    ref = json.loads((GOLDEN / "gated.json").read_text())
    boxes = [None if b is None else tuple(b) for b in ref["boxes"]]
    m = ogl.UNet().to("cuda")
    m.load_state_dict(calibrated_sd)
    m.eval()
    m.precision = "fp32"
    det = ScriptedDetector(boxes)
    got = ogl.extract_features_unet(str(GOLDEN / "pipeline_clip.avi"), det, m, torch.device("cuda"))
    assert det.resets == 1 and det.i == len(boxes)
    want = ref["features"]
    area = np.array(want["_area"])
    err = np.abs(got["_area"] - area)
    print("gated area max abs err", err.max(), "max area", area.max())
    assert err.max() <= 2.0
    assert all(g == 0 for g, b in zip(got["_area"], boxes) if b is None)
    for k in ("area_mean", "area_std", "open_quotient", "periodicity"):
        assert got[k] == pytest.approx(want[k], rel=5e-3, abs=1e-3), k


def test_gated_pipeline_native_size(lib, native_model, trained_sd, tmp_path):
    """256x256 clip: masks never leave the GPU; compare with the oracle's gated wave."""
    import cv2
    import openglottal_b200 as ogl
    from oracle import pipeline_oracle as po, synth, unet_oracle as uo

    clip, _ = synth.glottis_clip(20, 256, 256, seed=31, period=7.0)
    path = tmp_path / "clip.avi"
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"FFV1"), 25.0, (256, 256))
    for f in clip:
        vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
    vw.release()
    boxes = _boxes(20, 256, 256, seed=32, none_every=6)
    got = ogl.extract_features_unet(str(path), ScriptedDetector(boxes), native_model, None)
    _, ref_masks, _ = uo.batch_masks(trained_sd, clip)
    want = np.array(po.gated_area_wave(ref_masks, boxes))
    rel = np.abs(got["_area"] - want) / np.maximum(want, 1)
    print("gated native: max rel area err", rel.max())
    assert rel.max() <= 0.005 or np.abs(got["_area"] - want).max() <= 2


def test_yolo_crop_unet_pipeline(lib, native_model, trained_sd, tmp_path):
    """scripts/infer.py:222-248 end to end: GPU crop/letterbox/segment/un-letterbox/count vs the
    oracle loop with the fp32 reference forward (area within 0.5 % or 4 boundary pixels -- the
    crops are out of distribution for the full-frame-trained test weights, so a few logits sit
    at the threshold --, Dice >= 0.99 on the pasted masks)."""
    import cv2
    import openglottal_b200 as ogl
    from oracle import pipeline_oracle as po, synth, unet_oracle as uo

    clip, _ = synth.glottis_clip(12, 256, 256, seed=41, period=6.0)
    path = tmp_path / "clip.avi"
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"FFV1"), 25.0, (256, 256))
    for f in clip:
        vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
    vw.release()
    rng = np.random.default_rng(42)
    boxes = []
    for i in range(12):
        if i == 5:
            boxes.append(None)
            continue
        x1, y1 = int(rng.integers(20, 90)), int(rng.integers(20, 90))
        boxes.append((x1, y1, x1 + int(rng.integers(90, 150)), y1 + int(rng.integers(90, 150))))
    got = ogl.extract_features_yolo_crop_unet(str(path), ScriptedDetector(boxes), native_model)
    seg = lambda boxed: uo.segment_frame(trained_sd, boxed)
    want = [po.crop_unet_frame(f, b, seg) for f, b in zip(clip, boxes)]
    want_area = np.array([w[0] for w in want])
    err = np.abs(got["_area"] - want_area)
    print("crop pipeline area", got["_area"][:6], want_area[:6])
    assert (err <= np.maximum(4.0, 0.005 * want_area)).all()
    area, full = ogl.segment_crops(torch.from_numpy(clip).cuda(), boxes, native_model, want_full=True)
    full = full.cpu().numpy()
    for i, (a, fm) in enumerate(want):
        if fm is None:
            assert not full[i].any() and int(area[i]) == 0
        else:
            assert ogl.dice(full[i], fm) >= 0.99, i
    assert np.array_equal(area.cpu().numpy(), (full > 0).reshape(12, -1).sum(1))


def test_gaw_features_and_annotation(lib, native_model, trained_sd, tmp_path):
    """scripts/analyze_gaw.py:75-100 (gated waveform, f0 in Hz) and the unet-only branch of
    scripts/infer.py:212-219 (overlay frames + area), against the oracle."""
    import cv2
    from openglottal_b200.analysis import annotate_unet_only, extract_gaw_features, write_avi
    from oracle import pipeline_oracle as po, synth, unet_oracle as uo
    from oracle.features_oracle import kinematic_features

    clip, _ = synth.glottis_clip(24, 256, 256, seed=81, period=8.0)
    frames_bgr = [cv2.cvtColor(f, cv2.COLOR_GRAY2BGR) for f in clip]
    boxes = _boxes(24, 256, 256, seed=82, none_every=7)
    got = extract_gaw_features(frames_bgr, 4000.0, ScriptedDetector(boxes), native_model)
    _, ref_masks, _ = uo.batch_masks(trained_sd, clip)
    want = kinematic_features(po.gated_area_wave(ref_masks, boxes))
    if want["f0"] is not None:
        assert got["f0"] == pytest.approx(want["f0"] * 4000.0, rel=1e-9)
    assert got["area_mean"] == pytest.approx(want["area_mean"], rel=5e-3)
    annotated, wave = annotate_unet_only(frames_bgr, native_model)
    assert len(annotated) == 24 and annotated[0].shape == (256, 256, 3)
    ref_area = (ref_masks > 0).reshape(24, -1).sum(1)
    assert np.abs(np.array(wave) - ref_area).max() <= np.maximum(2, 0.005 * ref_area).max()
    out = tmp_path / "annotated.avi"
    write_avi(out, annotated, fps=25.0)
    cap = cv2.VideoCapture(str(out))
    ok, frm = cap.read()
    cap.release()
    assert ok and frm.shape == (256, 256, 3)


def test_gaw_features_512x256_match_reference_golden(lib, calibrated_sd):
    """extract_gaw_features on BAGLS-shaped 512(H) x 256(W) frames against the reference's own
    output (gaw_512x256.json, scripts/analyze_gaw.py:75-100): the frames must go through the
    reference's resize -> forward -> resize(prob) -> threshold order (the helpers in
    analysis.py once ran the network at the frames' own size). Calibrated RANDOM weights, hence the
    fp32 validation mode, as for gated.json; annotate_unet_only shares the same masks."""
    import cv2
    import openglottal_b200 as ogl
    from openglottal_b200.analysis import annotate_unet_only, extract_gaw_features
    from oracle import synth

    ref = json.loads((GOLDEN / "gaw_512x256.json").read_text())
    c = ref["clip"]
    clip, _ = synth.glottis_clip(c["n"], c["height"], c["width"], seed=c["seed"], period=c["period"])
    frames_bgr = [cv2.cvtColor(f, cv2.COLOR_GRAY2BGR) for f in clip]
    boxes = [None if b is None else tuple(b) for b in ref["boxes"]]
    m = ogl.UNet().to("cuda")
    m.load_state_dict(calibrated_sd)
    m.eval()
    m.precision = "fp32"
    det = ScriptedDetector(boxes)
    got = extract_gaw_features(frames_bgr, ref["capture_fps"], det, m)
    assert det.resets == 1 and det.i == len(boxes)
    want = ref["features"]
    area = np.array(want["_area"])
    err = np.abs(got["_area"] - area)
    print("gaw 512x256 area max abs err", err.max(), "max area", area.max())
    # fp32 kernels vs torch CPU fp32 on random weights: a handful of pixels of 131 072 sit within
    # rounding of the threshold (running at the frames' own size instead would be off by hundreds)
    assert err.max() <= 2.0      # measured 0.0 on B200 (profiles/pytest_gpu_r02_v4.log)
    assert got["f0"] == pytest.approx(want["f0"], rel=1e-9)          # Hz
    for k in ("area_mean", "area_std", "open_quotient", "periodicity"):
        assert got[k] == pytest.approx(want[k], rel=5e-3, abs=1e-3), k
    # the overlay path segments the same way: its ungated waveform bounds the gated one
    annotated, wave = annotate_unet_only(frames_bgr, m)
    assert annotated[0].shape == (c["height"], c["width"], 3)
    assert all(w + 4.0 >= a for w, a in zip(wave, area))


def test_unet_only_pipeline_on_512x256_video_reference_resize(lib, native_model, trained_sd, tmp_path):
    """BASELINE.json configs[2] through the pipeline: BAGLS-shaped 512(H) x 256(W) frames take the
    reference's resize path (utils.py:234-241: squash to 256x256, upsample the probability), here
    batched; areas against the oracle's per-frame loop."""
    import cv2
    import openglottal_b200 as ogl
    from oracle import synth, unet_oracle as uo

    clip, _ = synth.glottis_clip(6, 512, 256, seed=71, period=5.0)
    path = tmp_path / "bagls.avi"
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"FFV1"), 25.0, (256, 512))
    for f in clip:
        vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
    vw.release()
    got = ogl.extract_features_unet(str(path), None, native_model, torch.device("cuda"))
    want = np.array(uo.area_wave(trained_sd, list(clip)))
    err = np.abs(got["_area"] - want)
    print("512x256 reference-resize areas", got["_area"], want)
    assert (err <= np.maximum(4.0, 0.005 * want)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("fourcc", ["MJPG", "FFV1"])
def test_decode_gray_clip_equals_cv2(tmp_path, fourcc, write_clip):
    """The streaming ingestion (threads decoding frame ranges into pinned chunk buffers, H2D and
    BGR->gray on the GPU) gives exactly cv2.cvtColor of the reference's sequentially decoded
    frames (features.py:226,235), incl. a last partial chunk; a missing file gives None."""
    import cv2

    import openglottal_b200 as ogl
    from openglottal_b200.features import decode_gray_clip

    clip = tmp_path / f"c_{fourcc}.avi"
    write_clip(clip, fourcc, 700, 64, 48)
    want = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in ogl.load_frames_bgr(str(clip))])
    t = {}
    got = decode_gray_clip(str(clip), torch.device("cuda:0"), workers=4, chunk=256, timings=t)
    assert t["mode"] == "parallel" and got.shape == want.shape
    assert np.array_equal(got.cpu().numpy(), want)
    seq = decode_gray_clip(str(clip), torch.device("cuda:0"), workers=1)
    assert np.array_equal(seq.cpu().numpy(), want)
    assert decode_gray_clip(str(tmp_path / "missing.avi"), torch.device("cuda:0")) is None


@pytest.mark.gpu
def test_streamed_extract_features_equals_staged(tmp_path, write_clip, native_model):
    """``extract_features_unet`` on an MJPG file streams: frame ranges decode on several threads
    while earlier chunks are segmented. Its features (and ``_area``) equal those of the staged run
    -- sequential ``load_frames_bgr``, ``cv2.cvtColor``, one ``masks_for_clip`` over the whole clip --
    for a chunk size that does not divide the frame count; 256 x 256 frames take the native
    kernels, 64 x 48 ones the reference-resize path."""
    import cv2

    import openglottal_b200 as ogl

    for hgt, wid, n in ((256, 256, 300), (64, 48, 650)):
        clip = tmp_path / f"s_{hgt}x{wid}.avi"
        write_clip(clip, "MJPG", n, hgt, wid)
        frames = ogl.load_frames_bgr(str(clip))
        gray = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in frames])
        area, _ = ogl.masks_for_clip(torch.from_numpy(gray).cuda(), native_model)
        want = ogl.kinematic_features_device(area)
        got = ogl.extract_features_unet(str(clip), None, native_model, decode_chunk=256)
        assert (want is None) == (got is None)
        if want is not None:
            assert np.array_equal(got["_area"], want["_area"])
            for k in ("area_mean", "area_std", "area_range", "open_quotient", "f0", "periodicity", "cv"):
                assert got[k] == want[k], k
