"""Generates the committed golden fixtures by running the UNMODIFIED reference
(/root/reference, importable only in the build container) on seeded inputs.

    python tests/golden/make_golden.py

Outputs (all small, committed):
  features_kat.json        reference `_kinematic_features` on known waveforms
  unet_forward.npz         reference `UNet.forward` logits, calibrated_state(0), 2 frames 64x64
  segment_frame.npz        reference `unet_segment_frame` masks (256x256 identity path and a
                           96x128 frame through both cv2 resizes), bit-packed
  pipeline_clip.avi        30-frame 64x64 lossless (FFV1) synthetic clip
  pipeline.json            reference `extract_features_unet(clip, None, model, cpu)` output
  crops.npz                reference `letterbox_with_info` / `unletterbox` on seeded gray crops
  metrics.json             reference `dice` / `iou` on seeded mask pairs (incl. empty ones)
  overlay.npz              reference `_draw_overlay` (scripts/infer.py) on a seeded frame and mask
  gated.json               reference `extract_features_unet(clip, detector, model, cpu)` with a
                           scripted detector (boxes listed in the file)
  gaw_512x256.json         reference `extract_gaw_features` (scripts/analyze_gaw.py:75-100) on a
                           seeded 512(H) x 256(W) clip (frames regenerated from the seed), scripted
                           boxes, capture rate 4000 fps: the BAGLS-shaped reference-resize path
                           (`python tests/golden/make_golden.py gaw` writes this file alone)

The state dict is regenerated from its seed at test time (oracle.synth.calibrated_state uses only
a CPU torch.Generator), so no weights are committed.
"""
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
stub = types.ModuleType("ultralytics")   # openglottal/models/detector.py:6 imports it; YOLO unused
stub.YOLO = object
sys.modules.setdefault("ultralytics", stub)
sys.path.insert(0, "/root/reference")

import cv2  # noqa: E402
from openglottal.features import _kinematic_features, extract_features_unet  # noqa: E402
from openglottal.models.unet import UNet  # noqa: E402
from openglottal.utils import (  # noqa: E402
    dice, iou, letterbox_with_info, unet_segment_frame, unletterbox)

from oracle import synth  # noqa: E402


def jsonable(d):
    if d is None:
        return None
    return {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in d.items()}


def features_kat():
    t500 = np.arange(500)
    cases = [
        [0.0] * 10,
        [5.0],
        [5.0, 7.0],
        [5.0, 7.0, 5.0],
        [100.0] * 64,
        np.floor(np.maximum(0, 300 * np.sin(2 * np.pi * t500 / 20))).tolist(),
        np.floor(200 + 150 * np.sin(2 * np.pi * 0.06 * t500 + 0.3)).tolist(),
        np.random.default_rng(1234).integers(0, 65537, 1000).astype(float).tolist(),
        np.floor(np.maximum(0, 800 * np.sin(2 * np.pi * np.arange(2000) / 12.5)) + 40).tolist(),
        (np.random.default_rng(5).random(301) * 1000).tolist(),          # non-integral, odd n
        np.floor(500 + 400 * np.cos(2 * np.pi * np.arange(49) / 7)).tolist(),   # n < 50 lags
    ]
    out = []
    for wave in cases:
        try:
            res = _kinematic_features(wave)
            if res is not None:
                res = {k: v for k, v in jsonable(res).items() if k != "_area"}
            out.append({"input": wave, "raises": False, "output": res})
        except ValueError as e:
            out.append({"input": wave, "raises": True, "error": str(e), "output": None})
    (HERE / "features_kat.json").write_text(json.dumps(out))


class ScriptedDetector:
    """Stands in for TemporalDetector (Ultralytics is absent): replays a list of boxes."""

    def __init__(self, boxes):
        self.boxes, self.i = boxes, 0

    def reset(self):
        self.i = 0

    def detect(self, frame_bgr):
        b = self.boxes[self.i]
        self.i += 1
        return b


def scripted_boxes(n, hgt, wid, seed):
    rng = np.random.default_rng(seed)
    boxes = []
    for i in range(n):
        if i % 7 == 3:
            boxes.append(None)
            continue
        x1, y1 = int(rng.integers(0, wid - 8)), int(rng.integers(0, hgt - 8))
        x2, y2 = int(rng.integers(x1 + 1, wid + 1)), int(rng.integers(y1 + 1, hgt + 1))
        boxes.append((x1, y1, x2, y2))
    return boxes


def crops_and_metrics():
    rng = np.random.default_rng(77)
    out = {}
    sizes = [(256, 256), (40, 57), (300, 123), (17, 301), (255, 256), (1, 9), (128, 64), (333, 334)]
    for k, (h, w) in enumerate(sizes):
        crop = rng.integers(0, 256, (h, w), dtype=np.uint8)
        boxed, pt, pl, ch, cw = letterbox_with_info(crop, 256, value=0)
        mask_cs = (rng.random((256, 256)) > 0.6).astype(np.uint8) * 255
        back = unletterbox(mask_cs, pt, pl, ch, cw, h, w, interp=cv2.INTER_NEAREST)
        out[f"crop{k}"] = crop
        out[f"boxed{k}"] = boxed
        out[f"geom{k}"] = np.array([pt, pl, ch, cw], dtype=np.int32)
        out[f"maskcs{k}"] = np.packbits(mask_cs > 0)
        out[f"back{k}"] = np.packbits(back > 0)
    np.savez_compressed(HERE / "crops.npz", **out)
    cases = []
    for k in range(6):
        a = (rng.random((48, 64)) > (0.3 + 0.1 * k)).astype(np.uint8) * 255
        b = (rng.random((48, 64)) > 0.5).astype(np.uint8) * (k + 1)
        if k == 4:
            a[:] = 0
            b[:] = 0
        if k == 5:
            b[:] = 0
        cases.append({"seed_index": k, "a": np.packbits(a > 0).tolist(), "b": np.packbits(b > 0).tolist(),
                      "dice": dice(a, b), "iou": iou(a, b)})
    (HERE / "metrics.json").write_text(json.dumps(cases))


def overlay_golden():
    """The reference's own _draw_overlay (scripts/infer.py:91-124) on seeded inputs."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_infer", "/root/reference/scripts/infer.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(91)
    frame = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    mask = np.zeros((96, 128), np.uint8)
    cv2.ellipse(mask, (60, 50), (25, 12), 20, 0, 360, 255, -1)
    out = {"frame": frame, "mask": mask}
    for style in ("fill", "contour", "none"):
        out[f"out_{style}"] = mod._draw_overlay(frame, mask, (30, 20, 100, 80), 1234.0, style)
    out["out_nomask"] = mod._draw_overlay(frame, None, None, 0.0, "fill")
    np.savez_compressed(HERE / "overlay.npz", **out)


def gaw_golden():
    """scripts/analyze_gaw.py:75-100 on BAGLS-shaped frames: every frame is squashed to 256 x 256
    and the probability resized back (utils.py:234-240), the area is counted inside the box, f0 is
    converted to Hz."""
    sys.path.insert(0, "/root/reference/scripts")
    from analyze_gaw import extract_gaw_features

    sd = synth.calibrated_state(0)
    model = UNet(1, 1, (32, 64, 128, 256))
    model.load_state_dict(sd)
    model.eval()
    n, hgt, wid, seed, period, fps = 16, 512, 256, 91, 5.0, 4000.0
    clip, _ = synth.glottis_clip(n, hgt, wid, seed=seed, period=period)
    boxes = scripted_boxes(n, hgt, wid, seed=92)
    frames_bgr = [cv2.cvtColor(f, cv2.COLOR_GRAY2BGR) for f in clip]
    feats = extract_gaw_features(frames_bgr, fps, ScriptedDetector(boxes), model, torch.device("cpu"))
    (HERE / "gaw_512x256.json").write_text(json.dumps({
        "clip": {"n": n, "height": hgt, "width": wid, "seed": seed, "period": period},
        "capture_fps": fps, "boxes": boxes, "features": jsonable(feats)}))


def main():
    if sys.argv[1:] == ["gaw"]:
        gaw_golden()
        return
    gaw_golden()
    features_kat()
    crops_and_metrics()
    overlay_golden()
    sd = synth.calibrated_state(0)
    model = UNet(1, 1, (32, 64, 128, 256))
    model.load_state_dict(sd)
    model.eval()

    frames, _ = synth.glottis_clip(2, 64, 64, seed=11, period=5.0)
    with torch.no_grad():
        logits = model(torch.from_numpy(frames.astype("float32") / 255.0).unsqueeze(1)).numpy()
    np.savez_compressed(HERE / "unet_forward.npz", frames=frames, logits=logits)

    f256 = synth.glottis_clip(1, 256, 256, seed=12)[0][0]
    f96 = synth.glottis_clip(1, 96, 128, seed=13)[0][0]
    m256 = unet_segment_frame(f256, model, torch.device("cpu"))
    m96 = unet_segment_frame(f96, model, torch.device("cpu"))
    np.savez_compressed(HERE / "segment_frame.npz", f256=f256, f96=f96,
                        m256=np.packbits(m256 > 0), m96=np.packbits(m96 > 0))

    clip, _ = synth.glottis_clip(30, 64, 64, seed=14, period=6.0)
    path = HERE / "pipeline_clip.avi"
    vw = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"FFV1"), 25.0, (64, 64))
    for f in clip:
        vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
    vw.release()
    cap = cv2.VideoCapture(str(path))
    back = []
    while True:
        ok, frm = cap.read()
        if not ok:
            break
        back.append(cv2.cvtColor(frm, cv2.COLOR_BGR2GRAY))
    cap.release()
    assert len(back) == 30 and np.array_equal(np.stack(back), clip), "codec is not lossless here"
    feats = extract_features_unet(str(path), None, model, torch.device("cpu"))
    (HERE / "pipeline.json").write_text(json.dumps(jsonable(feats)))
    boxes = scripted_boxes(30, 64, 64, seed=15)
    gated = extract_features_unet(str(path), ScriptedDetector(boxes), model, torch.device("cpu"))
    (HERE / "gated.json").write_text(json.dumps({"boxes": boxes, "features": jsonable(gated)}))
    print("golden written:", sorted(p.name for p in HERE.iterdir() if p.is_file()))


if __name__ == "__main__":
    main()
