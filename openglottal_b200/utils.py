"""Per-frame / batched U-Net helpers mirroring /root/reference/openglottal/utils.py.

``unet_segment_frame`` keeps the reference signature and semantics
(/root/reference/openglottal/utils.py:218-241): squash to 256x256 with cv2 INTER_LINEAR,
forward, sigmoid, bilinear resize of the PROBABILITY map back to (H, W), threshold.
``unet_segment_frames`` is the batched GPU form the pipeline uses.
"""
from __future__ import annotations

import contextlib
import os

import numpy as np
import torch

from .unet import UNet


def _require_native(model) -> UNet:
    if not isinstance(model, UNet):
        raise TypeError(
            "openglottal_b200 helpers need an openglottal_b200.UNet (got "
            f"{type(model).__name__}); use the reference package for other modules")
    return model


@contextlib.contextmanager
def _silence_stderr():
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    os.dup2(devnull, 2)
    try:
        yield
    finally:
        os.dup2(saved, 2)
        os.close(saved)
        os.close(devnull)


def load_frames_bgr(avi_path: str) -> list[np.ndarray]:
    """All frames of a video as BGR uint8 arrays (cv2.VideoCapture; CPU decode stays with
    OpenCV exactly as in /root/reference/openglottal/utils.py:43-54)."""
    import cv2

    frames: list[np.ndarray] = []
    with _silence_stderr():
        cap = cv2.VideoCapture(str(avi_path))
        while True:
            ok, frm = cap.read()
            if not ok:
                break
            frames.append(frm)
        cap.release()
    return frames


def bgr_to_gray(frames_bgr: torch.Tensor) -> torch.Tensor:
    """``(N, H, W, 3)`` uint8 BGR CUDA tensor -> ``(N, H, W)`` uint8 gray, bit-exact with
    ``cv2.cvtColor(..., COLOR_BGR2GRAY)`` (/root/reference/openglottal/features.py:235)."""
    from . import _native

    if frames_bgr.dtype != torch.uint8 or frames_bgr.dim() != 4 or frames_bgr.shape[-1] != 3:
        raise ValueError("expected a (N, H, W, 3) uint8 tensor")
    if frames_bgr.device.type != "cuda":
        raise RuntimeError("bgr_to_gray runs on CUDA only")
    frames_bgr = frames_bgr.contiguous()
    gray = torch.empty(frames_bgr.shape[:3], dtype=torch.uint8, device=frames_bgr.device)
    with torch.cuda.device(frames_bgr.device):
        _native.check(_native.load().ogl_bgr_to_gray(
            frames_bgr.data_ptr(), gray.data_ptr(), gray.numel(),
            torch.cuda.current_stream().cuda_stream))
    return gray


def unet_segment_frames(frames_gray, model, threshold: float = 0.5):
    """Batched segmentation of ``(N, H, W)`` uint8 gray frames with H, W multiples of 16.

    The network runs at the frames' own resolution (what ``UNet.forward`` does in the
    reference; for 256x256 clips this is identical to the reference pipeline because its
    resize is the identity). Returns ``(mask uint8 {0,255} CUDA (N,H,W), area int32 CUDA (N,))``.
    """
    model = _require_native(model)
    dev = model._device()
    if isinstance(frames_gray, np.ndarray):
        frames_gray = torch.from_numpy(np.ascontiguousarray(frames_gray)).to(dev, non_blocking=True)
    _, mask, area = model.run(frames_gray, threshold=threshold)
    return mask, area


def unet_segment_frame(frame_gray: np.ndarray, model, device=None,
                       threshold: float = 0.5) -> np.ndarray:
    """Reference-compatible single-frame call: ``(H, W)`` uint8 -> uint8 mask {0, 255}.

    Same steps as /root/reference/openglottal/utils.py:234-241; the forward pass runs on the
    native kernels, resizes use the same cv2 calls as the reference.
    """
    import cv2

    model = _require_native(model)
    dev = model._device()
    if device is not None and torch.device(device).type != dev.type:
        raise RuntimeError(f"model is on {dev}, requested device {device}: there is no CPU path")
    inp = cv2.resize(frame_gray, (256, 256), interpolation=cv2.INTER_LINEAR)
    hgt, wid = frame_gray.shape
    t = torch.from_numpy(np.ascontiguousarray(inp)).unsqueeze(0).to(dev)
    if (hgt, wid) == (256, 256):
        _, mask, _ = model.run(t, threshold=threshold, want_area=False)
        return mask[0].cpu().numpy()
    logits, _, _ = model.run(t, want_logits=True, want_mask=False, want_area=False)
    prob = torch.sigmoid(logits[0]).cpu().numpy()
    prob = cv2.resize(prob, (wid, hgt), interpolation=cv2.INTER_LINEAR)
    return (prob > threshold).astype(np.uint8) * 255


def dice(pred: np.ndarray, gt: np.ndarray) -> float:
    """Dice of two binary masks, 1.0 when both are empty -- the parity metric
    (same definition as /root/reference/openglottal/utils.py:191-197)."""
    p = pred > 0
    g = gt > 0
    denom = int(p.sum()) + int(g.sum())
    if denom == 0:
        return 1.0
    return 2.0 * int(np.logical_and(p, g).sum()) / denom
