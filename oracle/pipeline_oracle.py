"""CPU oracle of the callers around the U-Net (TEST INFRASTRUCTURE, see __init__):

  letterbox_with_info  <- /root/reference/openglottal/utils.py:103-131
  unletterbox          <- /root/reference/openglottal/utils.py:170-186
  gated_area_wave      <- /root/reference/openglottal/features.py:238-245
  crop_unet_frame      <- /root/reference/scripts/infer.py:226-244 (yolo-crop+unet, one frame)
  dice / iou           <- /root/reference/openglottal/utils.py:191-206

cv2 and numpy do the arithmetic, exactly as in the reference; pinned by tests/golden/crops.npz
and metrics.json, which tests/golden/make_golden.py wrote by calling the reference's own
functions.
"""
from __future__ import annotations

import numpy as np


def letterbox_with_info(img: np.ndarray, size: int = 256, value: int = 0):
    import cv2

    h, w = img.shape[:2]
    scale = size / max(h, w)
    new_h, new_w = int(round(h * scale)), int(round(w * scale))
    interp = cv2.INTER_LINEAR if img.ndim == 3 else cv2.INTER_NEAREST   # utils.py:122
    resized = cv2.resize(img, (new_w, new_h), interpolation=interp)
    pad_h, pad_w = size - new_h, size - new_w
    top, bottom = pad_h // 2, pad_h - pad_h // 2
    left, right = pad_w // 2, pad_w - pad_w // 2
    val = (value, value, value) if img.ndim == 3 else value
    out = cv2.copyMakeBorder(resized, top, bottom, left, right, cv2.BORDER_CONSTANT, value=val)
    return out, top, left, new_h, new_w


def unletterbox(letterboxed: np.ndarray, pad_top: int, pad_left: int, content_h: int,
                content_w: int, target_h: int, target_w: int) -> np.ndarray:
    import cv2

    crop = letterboxed[pad_top:pad_top + content_h, pad_left:pad_left + content_w]
    if (content_h, content_w) == (target_h, target_w):
        return crop
    return cv2.resize(crop, (target_w, target_h), interpolation=cv2.INTER_NEAREST)


def gated_area_wave(masks, boxes) -> list[float]:
    """features.py:238-245: full-frame count when there is no detector is the caller's business;
    here every frame has a box or None."""
    out = []
    for m, b in zip(masks, boxes):
        if b is None:
            out.append(0.0)
        else:
            x1, y1, x2, y2 = b
            out.append(float(np.sum(m[y1:y2, x1:x2] > 0)))
    return out


def crop_unet_frame(gray: np.ndarray, box, segment, crop_size: int = 256):
    """scripts/infer.py:226-244 for one frame. ``segment(boxed) -> mask`` stands for
    ``unet_segment_frame(boxed, model, device)``. Returns (area, full-size mask or None)."""
    if box is None:
        return 0.0, None
    x1, y1, x2, y2 = box
    crop = gray[y1:y2, x1:x2]
    if crop.size == 0:
        return 0.0, None
    crop_h, crop_w = crop.shape[:2]
    boxed, pad_t, pad_l, content_h, content_w = letterbox_with_info(crop, crop_size, value=0)
    mask_cs = segment(boxed)
    mask_orig = unletterbox(mask_cs, pad_t, pad_l, content_h, content_w, crop_h, crop_w)
    full = np.zeros_like(gray)
    full[y1:y2, x1:x2] = mask_orig
    return float(np.sum(mask_orig > 0)), full


def dice(pred: np.ndarray, gt: np.ndarray) -> float:
    p = (pred > 0).astype(np.float32)
    g = (gt > 0).astype(np.float32)
    inter = (p * g).sum()
    denom = p.sum() + g.sum()
    return float(2 * inter / denom) if denom > 0 else 1.0


def iou(pred: np.ndarray, gt: np.ndarray) -> float:
    p = (pred > 0).astype(np.float32)
    g = (gt > 0).astype(np.float32)
    inter = (p * g).sum()
    union = p.sum() + g.sum() - inter
    return float(inter / union) if union > 0 else 1.0
