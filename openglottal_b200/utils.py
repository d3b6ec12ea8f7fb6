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


# Codecs whose every frame is a key frame: a seek to frame k lands exactly on frame k, so
# contiguous frame ranges can be decoded by independent VideoCapture instances. (Anything else --
# H.264, MPEG-4 ... -- is decoded sequentially: OpenCV's seek is not frame-exact there.)
_INTRA_ONLY_FOURCC = {"MJPG", "mjpg", "FFV1", "ffv1", "HFYU", "hfyu", "FFVH", "Y800", "GREY",
                      "DIB ", "RAW ", "I420", "IYUV", "YV12", "YUY2", "UYVY", "\0\0\0\0"}


def video_info(avi_path: str) -> dict:
    """``{"frames", "height", "width", "fourcc"}`` of a video as OpenCV reports them (``frames``
    is the container's count and may be 0 / wrong for broken headers)."""
    import cv2

    with _silence_stderr():
        cap = cv2.VideoCapture(str(avi_path))
        code = int(cap.get(cv2.CAP_PROP_FOURCC))
        info = {"frames": int(cap.get(cv2.CAP_PROP_FRAME_COUNT)),
                "height": int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)),
                "width": int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)),
                "fourcc": "".join(chr((code >> (8 * i)) & 0xFF) for i in range(4))}
        cap.release()
    return info


class RangeDecoder:
    """One ``VideoCapture`` that decodes frame ranges straight into caller-owned arrays (no
    per-frame allocation: at ~200 KB per frame every fresh numpy array is an mmap/munmap pair,
    which serialises concurrent decoders on the process's memory map)."""

    def __init__(self, avi_path: str):
        import cv2

        self.cap = cv2.VideoCapture(str(avi_path))
        self.pos = 0

    def read_into(self, start: int, stop: int, dst: np.ndarray) -> int:
        """Frames ``[start, stop)`` into ``dst[0:stop-start]`` (``(.., H, W, 3)`` uint8); returns
        how many were read (short on a failed seek, a read error or a frame of another shape)."""
        import cv2

        if self.pos != start:
            if not self.cap.set(cv2.CAP_PROP_POS_FRAMES, start) or \
                    int(self.cap.get(cv2.CAP_PROP_POS_FRAMES)) != start:
                return 0
            self.pos = start
        for i in range(stop - start):
            view = dst[i]
            ok, frm = self.cap.read(view)
            if not ok or frm.shape != view.shape:
                return i
            if frm is not view:   # OpenCV allocated its own array: copy
                dst[i] = frm
            self.pos += 1
        return stop - start

    def release(self) -> None:
        self.cap.release()


def decode_workers(default: int | None = None) -> int:
    """Decoder threads: ``OGL_DECODE_WORKERS`` or the host cores this process may use (max 32)."""
    env = os.environ.get("OGL_DECODE_WORKERS")
    if env:
        return max(1, int(env))
    if default:
        return default
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    return max(1, min(32, cores))


def load_frames_bgr_parallel(avi_path: str, workers: int | None = None,
                             min_frames: int = 256) -> list[np.ndarray]:
    """``load_frames_bgr`` with the decode spread over ``workers`` threads (OpenCV releases the GIL
    inside ``read``), each decoding one contiguous frame range with its own ``VideoCapture``. Only
    for intra-only codecs (MJPG, FFV1, raw ...), where a seek is frame-exact and every decoder
    produces the very frames the sequential loop of /root/reference/openglottal/utils.py:43-54
    does; anything else, short clips, wrong frame counts or a failed seek fall back to that loop.
    After the GPU path the CPU decode is the end-to-end bottleneck by more than 10x."""
    from concurrent.futures import ThreadPoolExecutor

    workers = decode_workers(workers)
    info = video_info(avi_path)
    n = info["frames"]
    if not parallel_decodable(info, workers, min_frames):
        return load_frames_bgr(avi_path)
    clip = np.empty((n, info["height"], info["width"], 3), np.uint8)
    bounds = [n * w // workers for w in range(workers + 1)]

    def work(w):
        dec = RangeDecoder(avi_path)
        try:
            got = dec.read_into(bounds[w], bounds[w + 1], clip[bounds[w]:bounds[w + 1]])
            if w == workers - 1 and got == bounds[w + 1] - bounds[w]:
                got += dec.cap.read()[0]     # a frame beyond the header's count: not trustworthy
            return got
        finally:
            dec.release()

    with _silence_stderr(), ThreadPoolExecutor(workers) as pool:
        counts = list(pool.map(work, range(workers)))
    if any(c != bounds[w + 1] - bounds[w] for w, c in enumerate(counts)):
        return load_frames_bgr(avi_path)      # header count or seek not trustworthy
    return list(clip)


def seekable_clip(info: dict) -> bool:
    """Whether a clip (``video_info``) can be read by frame ranges: an intra-only codec (a seek is
    frame-exact) and a plausible header."""
    return (info["frames"] > 0 and info["height"] > 0 and info["width"] > 0
            and info["fourcc"] in _INTRA_ONLY_FOURCC)


def parallel_decodable(info: dict, workers: int, min_frames: int = 256) -> bool:
    """Whether a clip is decoded by several ``RangeDecoder``s: seekable and enough frames to be
    worth it."""
    return workers >= 2 and info["frames"] >= max(min_frames, 2 * workers) and seekable_clip(info)


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


NET_SIZE = 256     # utils.py:234 squashes every frame to 256 x 256 before the forward pass


def resize_u8_linear(frames: torch.Tensor, hgt: int = NET_SIZE, wid: int = NET_SIZE) -> torch.Tensor:
    """Batched ``cv2.resize(frame, (wid, hgt), interpolation=cv2.INTER_LINEAR)`` of ``(N, H, W)``
    uint8 CUDA frames (/root/reference/openglottal/utils.py:234), bit-exact with cv2."""
    from . import _native

    if frames.device.type != "cuda" or frames.dtype != torch.uint8 or frames.dim() != 3:
        raise ValueError("expected (N, H, W) uint8 CUDA frames")
    n, sh, sw = frames.shape
    if (sh, sw) == (hgt, wid):
        return frames
    out = torch.empty((n, hgt, wid), dtype=torch.uint8, device=frames.device)
    with torch.cuda.device(frames.device):
        _native.check(_native.load().ogl_resize_u8_linear(
            frames.contiguous().data_ptr(), n, sh, sw, out.data_ptr(), hgt, wid,
            _stream(frames.device)))
    return out


def prob_resize_mask(logits: torch.Tensor, hgt: int, wid: int, threshold: float = 0.5,
                     want_mask: bool = True):
    """/root/reference/openglottal/utils.py:237-241 after the forward pass, batched on the device:
    sigmoid, bilinear resize of the PROBABILITY to ``(hgt, wid)`` (cv2 f32 INTER_LINEAR arithmetic;
    skipped when the size is unchanged), ``> threshold``. Returns ``(mask uint8 {0,255} CUDA
    (N, hgt, wid) | None, area int32 CUDA (N,))``."""
    from . import _native

    if logits.device.type != "cuda" or logits.dtype != torch.float32 or logits.dim() != 3:
        raise ValueError("expected (N, H, W) float32 CUDA logits")
    n, sh, sw = logits.shape
    dev = logits.device
    mask = torch.empty((n, hgt, wid), dtype=torch.uint8, device=dev) if want_mask else None
    area = torch.empty(n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _native.check(_native.load().ogl_prob_resize_mask(
            logits.contiguous().data_ptr(), n, sh, sw, hgt, wid, float(threshold),
            mask.data_ptr() if want_mask else None, area.data_ptr(), _stream(dev)))
    return mask, area


def segment_frames_reference_resize(frames_gray: torch.Tensor, model, threshold: float = 0.5,
                                    want_masks: bool = True, batch: int = 512):
    """The reference's per-frame semantics (/root/reference/openglottal/utils.py:234-241) for a
    batch of equally sized ``(N, H, W)`` uint8 CUDA frames of ANY size, entirely on the device:
    squash to 256 x 256 (cv2 u8 INTER_LINEAR, bit-exact), forward, sigmoid, bilinear resize of the
    probability back to ``(H, W)``, threshold, count. Only masks and areas exist at ``(H, W)``.
    Returns ``(area int32 CUDA (N,), masks uint8 CUDA (N, H, W) | None)``."""
    model = _require_native(model)
    n, hgt, wid = frames_gray.shape
    if (hgt, wid) == (NET_SIZE, NET_SIZE):   # both resizes are the identity: the fused head does it
        _, mask, area = model.run(frames_gray, threshold=threshold, want_mask=want_masks)
        return area, mask
    areas, masks = [], []
    for i0 in range(0, n, batch):
        small = resize_u8_linear(frames_gray[i0:i0 + batch])
        logits, _, _ = model.run(small, want_logits=True, want_mask=False, want_area=False)
        m, a = prob_resize_mask(logits, hgt, wid, threshold, want_mask=want_masks)
        areas.append(a)
        masks.append(m)
    return torch.cat(areas), (torch.cat(masks) if want_masks else None)


def unet_segment_frame(frame_gray: np.ndarray, model, device=None,
                       threshold: float = 0.5) -> np.ndarray:
    """Reference-compatible single-frame call: ``(H, W)`` uint8 -> uint8 mask {0, 255}.

    Same steps as /root/reference/openglottal/utils.py:234-241, all of them on the device (the
    two cv2 resizes are ``ogl_resize_u8_linear`` / ``ogl_prob_resize_mask``); one H2D copy of the
    frame in, one D2H copy of the mask out.
    """
    model = _require_native(model)
    dev = model._device()
    if device is not None and torch.device(device).type != dev.type:
        raise RuntimeError(f"model is on {dev}, requested device {device}: there is no CPU path")
    if frame_gray.ndim != 2 or frame_gray.dtype != np.uint8:
        raise ValueError("expected a (H, W) uint8 gray frame")
    t = torch.from_numpy(np.ascontiguousarray(frame_gray)).unsqueeze(0).to(dev)
    _, masks = segment_frames_reference_resize(t, model, threshold=threshold)
    return masks[0].cpu().numpy()


# ---------------------------------------------------------------------------------------------
# Callers around the U-Net: detection-gated area, the yolo-crop+unet geometry, batched metrics
# ---------------------------------------------------------------------------------------------
def _stream(dev: torch.device):
    return torch.cuda.current_stream(dev).cuda_stream


def gated_area(masks: torch.Tensor, boxes) -> torch.Tensor:
    """Detection-gated areas, /root/reference/openglottal/features.py:240-245, batched.

    ``masks``: ``(N, H, W)`` uint8 CUDA; ``boxes``: sequence of N entries, each ``None`` (the
    detector found nothing: area 0) or ``(x1, y1, x2, y2)`` ints with Python slice semantics
    ``mask[y1:y2, x1:x2]``. Returns int32 CUDA ``(N,)``.
    """
    from . import _native

    if masks.device.type != "cuda" or masks.dtype != torch.uint8 or masks.dim() != 3:
        raise ValueError("expected (N, H, W) uint8 CUDA masks")
    n, hgt, wid = masks.shape
    if len(boxes) != n:
        raise ValueError(f"{len(boxes)} boxes for {n} masks")
    host = np.zeros((n, 4), dtype=np.int32)
    has = np.zeros(n, dtype=np.uint8)
    for i, b in enumerate(boxes):
        if b is not None:
            host[i] = [int(v) for v in b]
            has[i] = 1
    dev = masks.device
    boxes_d = torch.from_numpy(host).to(dev)
    has_d = torch.from_numpy(has).to(dev)
    area = torch.empty(n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _native.check(_native.load().ogl_mask_area_boxes(
            masks.contiguous().data_ptr(), n, hgt, wid, boxes_d.data_ptr(), has_d.data_ptr(),
            area.data_ptr(), _stream(dev)))
    return area


def letterbox_geometry(hgt: int, wid: int, size: int = 256):
    """``(pad_top, pad_left, content_h, content_w)`` of ``letterbox_with_info``
    (/root/reference/openglottal/utils.py:119-126; Python ``round`` = half to even)."""
    scale = size / max(hgt, wid)
    new_h, new_w = int(round(hgt * scale)), int(round(wid * scale))
    return (size - new_h) // 2, (size - new_w) // 2, new_h, new_w


def _crop_geometry(boxes, hgt: int, wid: int, size: int) -> np.ndarray:
    """8 int32 per frame {x1, y1, x2, y2, pad_top, pad_left, content_h, content_w}; all zero for
    frames without a usable crop (box None or empty slice, scripts/infer.py:228-231)."""
    geom = np.zeros((len(boxes), 8), dtype=np.int32)
    for i, b in enumerate(boxes):
        if b is None:
            continue
        x1, y1, x2, y2 = (int(v) for v in b)
        ys, xs = slice(y1, y2).indices(hgt), slice(x1, x2).indices(wid)
        ch, cw = max(0, ys[1] - ys[0]), max(0, xs[1] - xs[0])
        if ch == 0 or cw == 0:
            continue
        pt, pl, nh, nw = letterbox_geometry(ch, cw, size)
        if nh <= 0 or nw <= 0:
            raise ValueError(f"crop {cw}x{ch} collapses under letterboxing (cv2.resize would fail)")
        geom[i] = [xs[0], ys[0], xs[1], ys[1], pt, pl, nh, nw]
    return geom


def letterbox_crops(frames_gray: torch.Tensor, boxes, size: int = 256):
    """Batched ``letterbox_with_info(gray[y1:y2, x1:x2], size)`` for gray frames
    (/root/reference/scripts/infer.py:229-236). Returns ``(boxed uint8 CUDA (N,size,size),
    geometry int32 numpy (N,8))``; frames without a usable box give all-zero images."""
    from . import _native

    if frames_gray.device.type != "cuda" or frames_gray.dtype != torch.uint8 or frames_gray.dim() != 3:
        raise ValueError("expected (N, H, W) uint8 CUDA frames")
    n, hgt, wid = frames_gray.shape
    if len(boxes) != n:
        raise ValueError(f"{len(boxes)} boxes for {n} frames")
    geom = _crop_geometry(boxes, hgt, wid, size)
    dev = frames_gray.device
    geom_d = torch.from_numpy(geom).to(dev)
    out = torch.empty((n, size, size), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _native.check(_native.load().ogl_letterbox_crops(
            frames_gray.contiguous().data_ptr(), n, hgt, wid, geom_d.data_ptr(), size,
            out.data_ptr(), _stream(dev)))
    return out, geom


def unletterbox_area(masks_cs: torch.Tensor, geom: np.ndarray, hgt: int, wid: int,
                     want_full: bool = False):
    """``unletterbox`` (/root/reference/openglottal/utils.py:170-186) of every crop-space mask back
    to its crop size, its area, and optionally the full-size mask with the crop pasted in
    (scripts/infer.py:237-244). Returns ``(area int32 CUDA (N,), full uint8 CUDA (N,H,W) | None)``."""
    from . import _native

    if masks_cs.device.type != "cuda" or masks_cs.dtype != torch.uint8 or masks_cs.dim() != 3:
        raise ValueError("expected (N, size, size) uint8 CUDA masks")
    n, size, size2 = masks_cs.shape
    if size != size2 or geom.shape != (n, 8):
        raise ValueError("geometry does not match the masks")
    dev = masks_cs.device
    geom_d = torch.from_numpy(np.ascontiguousarray(geom, dtype=np.int32)).to(dev)
    area = torch.empty(n, dtype=torch.int32, device=dev)
    full = torch.empty((n, hgt, wid), dtype=torch.uint8, device=dev) if want_full else None
    with torch.cuda.device(dev):
        _native.check(_native.load().ogl_unletterbox_area(
            masks_cs.contiguous().data_ptr(), n, size, geom_d.data_ptr(), hgt, wid,
            full.data_ptr() if want_full else None, area.data_ptr(), _stream(dev)))
    return area, full


def segment_crops(frames_gray, boxes, model, size: int = 256, threshold: float = 0.5,
                  want_full: bool = False, batch: int = 512):
    """The ``yolo-crop+unet`` stage (/root/reference/scripts/infer.py:222-248) for a batch: crop
    each frame at its box, letterbox to ``size`` x ``size``, segment with the (crop-trained)
    U-Net, un-letterbox, count. Boxes come from the caller (the reference's TemporalDetector).
    Returns ``(area int32 CUDA (N,), full-size masks uint8 CUDA (N,H,W) | None)``."""
    model = _require_native(model)
    dev = model._device()
    if isinstance(frames_gray, np.ndarray):
        frames_gray = torch.from_numpy(np.ascontiguousarray(frames_gray))
    frames_gray = frames_gray.to(dev, non_blocking=True)
    n, hgt, wid = frames_gray.shape
    areas, fulls = [], []
    for i0 in range(0, n, batch):
        sl = slice(i0, min(n, i0 + batch))
        boxed, geom = letterbox_crops(frames_gray[sl], boxes[sl], size)
        _, mask_cs, _ = model.run(boxed, threshold=threshold, want_area=False)
        a, f = unletterbox_area(mask_cs, geom, hgt, wid, want_full=want_full)
        areas.append(a)
        fulls.append(f)
    return torch.cat(areas), (torch.cat(fulls) if want_full else None)


def overlap_counts(pred: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """Per-frame ``{|pred & gt|, |pred|, |gt|}`` over ``value > 0`` as int32 CUDA ``(N, 3)``."""
    from . import _native

    if pred.shape != gt.shape or pred.dim() < 2:
        raise ValueError("pred and gt must have the same (N, ...) shape")
    if pred.device.type != "cuda" or gt.device != pred.device:
        raise ValueError("pred and gt must be CUDA tensors on the same device")
    if pred.dtype != torch.uint8 or gt.dtype != torch.uint8:
        raise TypeError("masks must be uint8")
    n = pred.shape[0]
    pixels = pred[0].numel()
    dev = pred.device
    counts = torch.empty((n, 3), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _native.check(_native.load().ogl_mask_overlap_counts(
            pred.contiguous().data_ptr(), gt.contiguous().data_ptr(), n, pixels,
            counts.data_ptr(), _stream(dev)))
    return counts


def dice_iou_batch(pred: torch.Tensor, gt: torch.Tensor):
    """Batched ``dice`` / ``iou`` (/root/reference/openglottal/utils.py:191-206) for ``(N, H, W)``
    uint8 CUDA masks: the counts come from one kernel, the ratios are formed exactly as the
    reference forms them (float32 sums, 1.0 when the denominator is 0). Returns two float64
    numpy arrays of length N."""
    c = overlap_counts(pred, gt).cpu().numpy().astype(np.float32)
    inter, p, g = c[:, 0], c[:, 1], c[:, 2]
    denom = p + g
    union = p + g - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.where(denom > 0, (np.float32(2) * inter) / denom, np.float32(1.0))
        j = np.where(union > 0, inter / union, np.float32(1.0))
    return d.astype(np.float64), j.astype(np.float64)


def iou(pred: np.ndarray, gt: np.ndarray) -> float:
    """Intersection-over-union of two binary masks, 1.0 when both are empty
    (/root/reference/openglottal/utils.py:200-206)."""
    p = pred > 0
    g = gt > 0
    inter = int(np.logical_and(p, g).sum())
    union = int(p.sum()) + int(g.sum()) - inter
    return inter / union if union > 0 else 1.0


def dice(pred: np.ndarray, gt: np.ndarray) -> float:
    """Dice of two binary masks, 1.0 when both are empty -- the parity metric
    (same definition as /root/reference/openglottal/utils.py:191-197)."""
    p = pred > 0
    g = gt > 0
    denom = int(p.sum()) + int(g.sum())
    if denom == 0:
        return 1.0
    return 2.0 * int(np.logical_and(p, g).sum()) / denom
