"""Consumers of the masks / area waveform, mirroring the reference's inference and analysis scripts:

* ``draw_overlay``          <- /root/reference/scripts/infer.py:91-124 (``_draw_overlay``)
* ``annotate_unet_only``    <- /root/reference/scripts/infer.py:212-219 (unet-only branch of
                               ``_run_pipeline``: mask + area of every frame, overlay frames)
* ``write_avi``             <- /root/reference/scripts/infer.py:270-278 (MJPG writer)
* ``extract_gaw_features``  <- /root/reference/scripts/analyze_gaw.py:75-100 (gated area waveform,
                               kinematic features, f0 converted from cycles/frame to Hz)
* ``features_row``          <- the CSV column order of scripts/infer.py:84-85 (FEATURE_COLS)

The segmentation and the gating run on the GPU in batches; drawing and video writing are OpenCV
host code exactly as in the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from .features import gray_clip_from_bgr, kinematic_features_device, masks_for_clip
from .utils import _require_native, gated_area

FEATURE_COLS = ["area_mean", "area_std", "area_range", "open_quotient", "f0", "periodicity", "cv"]
GIRAFE_CAPTURE_FPS = 4000.0


def draw_overlay(frame_bgr: np.ndarray, mask, box, area: float, overlay_style: str = "fill") -> np.ndarray:
    """Copy of ``frame_bgr`` with the mask (``"fill"``: 40 % green fill + outline, ``"contour"``:
    outline only, ``"none"``: ignored), the bbox (if any) and an ``area=N`` label burned in."""
    import cv2

    out = frame_bgr.copy()
    if overlay_style != "none" and mask is not None and mask.any():
        contours, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        if overlay_style == "fill":
            tint = np.zeros_like(out)
            tint[:, :, 1] = mask
            out = cv2.addWeighted(out, 1.0, tint, 0.4, 0)
        cv2.drawContours(out, contours, -1, (0, 255, 0), 1)
    if box is not None:
        x1, y1, x2, y2 = box
        cv2.rectangle(out, (x1, y1), (x2, y2), (0, 220, 255), 1)
    cv2.putText(out, f"area={int(area)}", (4, 14), cv2.FONT_HERSHEY_SIMPLEX, 0.4, (255, 255, 255), 1,
                cv2.LINE_AA)
    return out


def annotate_unet_only(frames_bgr: list, model, overlay_style: str = "fill", batch: int = 512):
    """unet-only branch of the reference's ``_run_pipeline``: returns ``(annotated frames,
    area waveform as a list of floats)``; frames of any size, with the reference's resize semantics
    (``masks_for_clip``)."""
    model = _require_native(model)
    dev = model._device()
    area, masks = masks_for_clip(gray_clip_from_bgr(frames_bgr, dev), model, batch=batch,
                                 want_masks=True)
    area_h = area.cpu().numpy()
    masks_h = masks.cpu().numpy()
    annotated = [draw_overlay(f, m, None, float(a), overlay_style)
                 for f, m, a in zip(frames_bgr, masks_h, area_h)]
    return annotated, [float(a) for a in area_h]


def write_avi(path, frames: list, fps: float = 25.0) -> None:
    """MJPG AVI of BGR frames; nothing is written for an empty list."""
    import cv2

    if not frames:
        return
    hgt, wid = frames[0].shape[:2]
    writer = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"MJPG"), fps, (wid, hgt))
    for f in frames:
        writer.write(f)
    writer.release()


def extract_gaw_features(frames: list, capture_fps: float, detector, unet_model, device=None):
    """YOLO-gated area waveform of BGR frames -> kinematic features with ``f0`` in Hz, or ``None``
    (analyze_gaw.py:75-100). The detector is the caller's; masks and gating are batched."""
    model = _require_native(unet_model)
    dev = model._device()
    detector.reset()
    boxes = [detector.detect(frm) for frm in frames]
    _, masks = masks_for_clip(gray_clip_from_bgr(frames, dev), model, want_masks=True)
    feats = kinematic_features_device(gated_area(masks, boxes))
    if feats is not None and feats.get("f0") is not None:
        feats["f0"] = feats["f0"] * capture_fps
    return feats


def features_row(name: str, feats: dict | None) -> list:
    """One CSV row ``[name, *FEATURE_COLS]``; missing features are written as empty strings."""
    if feats is None:
        return [name] + [""] * len(FEATURE_COLS)
    return [name] + ["" if feats.get(k) is None else feats[k] for k in FEATURE_COLS]
