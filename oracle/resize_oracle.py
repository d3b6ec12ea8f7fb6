"""NumPy restatement of the two ``cv2.resize(..., interpolation=cv2.INTER_LINEAR)`` calls of
``unet_segment_frame`` (TEST INFRASTRUCTURE, see oracle/__init__.py).

  resize_u8_linear   <- /root/reference/openglottal/utils.py:234 (u8 frame -> 256 x 256)
  resize_f32_linear  <- /root/reference/openglottal/utils.py:238-240 (f32 probability -> (W, H))

The arithmetic is OpenCV's (third party: opencv-python >= 4.8 per the reference's pyproject.toml:26;
4.13.0 here), modules/imgproc/src/resize.cpp, generic INTER_LINEAR path:

* position of destination index d on a source axis: ``f = float32((d + 0.5) * (src / dst) - 0.5)``
  computed in double, ``s = floor(f)``, ``f -= s``;
* horizontal axis: ``s < 0 -> (s, f) = (0, 0)``; ``s >= src - 1 -> (s, f) = (src - 1, 0)``;
* vertical axis: the weights keep the unclamped ``f``; the two row indices are clipped;
* u8: 11-bit fixed-point weights ``cvRound(w * 2048)``, rows as int32, and
  ``(((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2`` vertically; when both axes
  halve exactly cv2 substitutes the 2 x 2 area mean ``(a + b + c + d + 2) >> 2``;
* f32: plain float32 products and sums, horizontal first.

Pinned by tests/test_oracle_golden.py against cv2 itself: the u8 function is bit-exact for every
size pair tried; the f32 function is bit-exact with OpenCV's own code (``cv2.ipp.setUseIPP(False)``)
and within 2e-5 of the Intel IPP routine the default build dispatches f32 resizes to.
"""
from __future__ import annotations

import numpy as np


def _pos(src: int, dst: int):
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * (src / dst) - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    return s, (f - s.astype(np.float32)).astype(np.float32)


def _coeffs_x(src: int, dst: int):
    s, f = _pos(src, dst)
    lo = s < 0
    s[lo], f[lo] = 0, 0
    hi = s >= src - 1
    s[hi], f[hi] = src - 1, 0
    return s, np.minimum(s + 1, src - 1), f


def _coeffs_y(src: int, dst: int):
    s, f = _pos(src, dst)
    return np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1), f


def resize_u8_linear(img: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    h, w = img.shape
    if (w, h) == (dst_w, dst_h):
        return img.copy()
    if w == 2 * dst_w and h == 2 * dst_h:
        a = img.astype(np.int32)
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx0, sx1, fx = _coeffs_x(w, dst_w)
    sy0, sy1, fy = _coeffs_y(h, dst_h)

    def fixed(f):
        one = np.float32(1)
        return (np.rint((one - f) * np.float32(2048)).astype(np.int32),
                np.rint(f * np.float32(2048)).astype(np.int32))

    ax0, ax1 = fixed(fx)
    ay0, ay1 = fixed(fy)
    s = img.astype(np.int32)
    rows = s[:, sx0] * ax0[None, :] + s[:, sx1] * ax1[None, :]
    out = (((ay0[:, None] * (rows[sy0] >> 4)) >> 16) + ((ay1[:, None] * (rows[sy1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def resize_f32_linear(img: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    h, w = img.shape
    if (w, h) == (dst_w, dst_h):
        return img.copy()
    img = img.astype(np.float32, copy=False)
    sx0, sx1, fx = _coeffs_x(w, dst_w)
    sy0, sy1, fy = _coeffs_y(h, dst_h)
    one = np.float32(1)
    rows = img[:, sx0] * (one - fx)[None, :] + img[:, sx1] * fx[None, :]
    return (rows[sy0] * (one - fy)[:, None] + rows[sy1] * fy[:, None]).astype(np.float32)


def segment_frame_restated(logits_256: np.ndarray, hgt: int, wid: int, threshold: float = 0.5) -> np.ndarray:
    """utils.py:237-241 on given 256 x 256 f32 logits: sigmoid, resize to (wid, hgt), threshold."""
    prob = (1.0 / (1.0 + np.exp(-logits_256.astype(np.float32)))).astype(np.float32)
    if (hgt, wid) != logits_256.shape:
        prob = resize_f32_linear(prob, wid, hgt)
    return (prob > np.float32(threshold)).astype(np.uint8) * 255
