"""unet-only pipeline and kinematic features, mirroring /root/reference/openglottal/features.py.

* ``_kinematic_features``      <- features.py:38-68  (same dict, same None / ValueError cases)
* ``extract_features_unet``    <- features.py:202-247 (same signature and return value)
* ``extract_features_unet_frames`` is the raw-frame entry point the video wrapper, the bench and
  the multi-GPU launcher share: batches stream from pinned host memory, frames are sharded by
  contiguous ranges over ranks, and only the int32 area vector is gathered (NCCL).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native, sharding
from .unet import UNet
from .utils import (RangeDecoder, _require_native, _silence_stderr, bgr_to_gray, decode_workers,
                    load_frames_bgr, parallel_decodable, seekable_clip, video_info)

_FEATURE_KEYS = ("area_mean", "area_std", "area_range", "open_quotient", "f0", "periodicity", "cv")


def _features_from_device(area_dev: torch.Tensor):
    """Runs the CUDA feature kernels on an int32 or float64 CUDA vector of n >= 2 samples.
    Returns (out8 numpy float64, flags numpy int32) after one D2H copy."""
    lib = _native.load()
    n = area_dev.numel()
    dev = area_dev.device
    nbytes = lib.ogl_features_workspace_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty(8, dtype=torch.float64, device=dev)
    flags = torch.empty(2, dtype=torch.int32, device=dev)
    fn = lib.ogl_features if area_dev.dtype == torch.int32 else lib.ogl_features_f64
    with torch.cuda.device(dev):
        _native.check(fn(area_dev.data_ptr(), n, out.data_ptr(), flags.data_ptr(), ws.data_ptr(),
                         ws.numel(), torch.cuda.current_stream().cuda_stream))
    return out.cpu().numpy(), flags.cpu().numpy()


def _assemble(out8: np.ndarray, flags: np.ndarray, area_f64: np.ndarray) -> dict | None:
    if flags[0]:
        return None  # silent waveform (features.py:45-46)
    return {
        "area_mean": np.float64(out8[0]),
        "area_std": np.float64(out8[1]),
        "area_range": np.float64(out8[2]),
        "open_quotient": float(out8[3]),
        "f0": None if flags[1] else float(out8[4]),
        "periodicity": float(out8[5]),
        "cv": np.float64(out8[6]),
        "_area": area_f64,
    }


def kinematic_features_device(area: torch.Tensor) -> dict | None:
    """Features of an int32 CUDA area vector (the hot-path form). ``_area`` is the float64 copy
    the reference returns."""
    if area.device.type != "cuda" or area.dtype != torch.int32 or area.dim() != 1:
        raise ValueError("expected a 1-D int32 CUDA tensor")
    n = area.numel()
    if n == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    area_host = area.cpu().numpy().astype(np.float64)
    if n == 1:
        if area_host[0] == 0:
            return None
        raise ValueError("attempt to get argmax of an empty sequence")  # as the reference does
    out8, flags = _features_from_device(area.contiguous())
    return _assemble(out8, flags, area_host)


def _kinematic_features(area_wave) -> dict | None:
    """Reference-compatible entry: list/array of per-frame areas -> feature dict or ``None``.

    Edge cases follow /root/reference/openglottal/features.py:44-54 exactly: empty input and
    n == 1 (non-silent) raise ``ValueError`` as numpy does there; an all-zero waveform gives
    ``None``. The arithmetic runs on the GPU in fp64.
    """
    area = np.array(area_wave, dtype=np.float64)
    if area.ndim != 1:
        raise ValueError("area_wave must be one-dimensional")
    if area.size == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    if area.max() == 0:
        return None
    if area.size == 1:
        raise ValueError("attempt to get argmax of an empty sequence")
    if not torch.cuda.is_available():
        raise RuntimeError("openglottal_b200._kinematic_features needs a CUDA device (no CPU path)")
    dev = torch.device("cuda", torch.cuda.current_device())
    exact_int = bool(np.all(area == np.round(area)) and np.abs(area).max() < 2**31)
    if exact_int:
        t = torch.from_numpy(area.astype(np.int32)).to(dev)
    else:
        t = torch.from_numpy(area).to(dev)
    out8, flags = _features_from_device(t)
    return _assemble(out8, flags, area)


def segment_clip(frames_gray, model: UNet, batch: int = 512, threshold: float = 0.5,
                 want_masks: bool = False):
    """Area waveform (and optionally masks) of a clip of ``(N, H, W)`` uint8 gray frames.

    Host input (numpy / CPU tensor) is streamed batch by batch from pinned memory on a copy
    stream, double-buffered against the compute stream; CUDA input is used in place.
    Returns ``(area int32 CUDA (N,), masks uint8 CUDA (N,H,W) | None)``.
    """
    model = _require_native(model)
    dev = model._device()
    if isinstance(frames_gray, np.ndarray):
        frames_gray = torch.from_numpy(np.ascontiguousarray(frames_gray))
    if frames_gray.dtype != torch.uint8 or frames_gray.dim() != 3:
        raise ValueError("expected (N, H, W) uint8 frames")
    n, hgt, wid = frames_gray.shape
    if n == 0:
        raise ValueError("no frames")
    area = torch.empty(n, dtype=torch.int32, device=dev)
    masks = torch.empty((n, hgt, wid), dtype=torch.uint8, device=dev) if want_masks else None
    if frames_gray.device.type == "cuda":
        for i0 in range(0, n, batch):
            _, m, a = model.run(frames_gray[i0:i0 + batch], threshold=threshold,
                                want_mask=want_masks)
            area[i0:i0 + batch] = a
            if want_masks:
                masks[i0:i0 + batch] = m
        return area, masks

    # Pinned input is copied from in place. Pageable input is staged through two pinned batch
    # buffers (pinning a whole 10^6-frame clip would lock 65 GB of host memory): the CPU copy of
    # batch k+1 into its staging buffer overlaps the GPU work of batch k.
    pinned_in = frames_gray.is_pinned()
    stage = None if pinned_in else [
        torch.empty((min(batch, n), hgt, wid), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    compute = torch.cuda.current_stream(dev)
    copy = torch.cuda.Stream(device=dev)
    bufs = [torch.empty((min(batch, n), hgt, wid), dtype=torch.uint8, device=dev) for _ in range(2)]
    # `bufs` come from the caching allocator and may be the blocks an earlier, still queued call on
    # the compute stream reads from: the side stream must not write them before that work is done
    copy.wait_stream(compute)
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]   # H2D of a staging buffer has finished
    starts = list(range(0, n, batch))

    def issue_copy(k: int) -> None:
        i0 = starts[k]
        m = min(batch, n - i0)
        if pinned_in:
            src = frames_gray[i0:i0 + m]
        else:
            if k >= 2:
                copied[k % 2].synchronize()        # the staging buffer is free again
            stage[k % 2][:m].copy_(frames_gray[i0:i0 + m])
            src = stage[k % 2][:m]
        with torch.cuda.stream(copy):
            if k >= 2:
                copy.wait_event(freed[k % 2])
            bufs[k % 2][:m].copy_(src, non_blocking=True)
            ready[k % 2].record(copy)
            copied[k % 2].record(copy)

    issue_copy(0)
    for k, i0 in enumerate(starts):
        if k + 1 < len(starts):
            issue_copy(k + 1)
        m = min(batch, n - i0)
        compute.wait_event(ready[k % 2])
        _, mk, a = model.run(bufs[k % 2][:m], threshold=threshold, want_mask=want_masks)
        freed[k % 2].record(compute)
        area[i0:i0 + m] = a
        if want_masks:
            masks[i0:i0 + m] = mk
    return area, masks


def gray_clip_from_bgr(frames_bgr: list, dev: torch.device, chunk: int = 2048) -> torch.Tensor:
    """List of ``(H, W, 3)`` uint8 BGR frames -> ``(N, H, W)`` uint8 gray CUDA tensor, converted on
    the GPU (bit-exact with ``cv2.COLOR_BGR2GRAY``, features.py:235) ``chunk`` frames at a time so
    that neither the 3-channel clip nor a pinned copy of it has to exist as a whole."""
    n = len(frames_bgr)
    hgt, wid = frames_bgr[0].shape[:2]
    gray = torch.empty((n, hgt, wid), dtype=torch.uint8, device=dev)
    for i0 in range(0, n, chunk):
        part = torch.from_numpy(np.stack(frames_bgr[i0:i0 + chunk])).to(dev)
        gray[i0:i0 + part.shape[0]] = bgr_to_gray(part)
    return gray


class _DecodeFallback(Exception):
    """The parallel decoder met something it does not trust; the caller restarts sequentially."""


def iter_gray_chunks(avi_path: str, dev: torch.device, workers: int | None = None,
                     chunk: int = 1024, out: torch.Tensor | None = None, stats: dict | None = None,
                     frame_range: tuple[int, int] | None = None):
    """Generator over ``(first frame index, (m, H, W) uint8 gray CUDA tensor)`` of a video whose
    codec is intra-only (``parallel_decodable``), in order. ``workers`` threads, each with its own
    ``VideoCapture`` on a contiguous frame range, decode straight into two pinned BGR chunk
    buffers; the chunk crosses PCIe and becomes gray (OpenCV's integer coefficients, bit-exact) on
    a side stream, and the yielded tensor is ordered after that on the CURRENT stream -- so a
    consumer that enqueues GPU work per chunk overlaps it with the decode of the next chunk.
    ``out``: an ``(N, H, W)`` tensor to fill (slices are yielded) instead of one tensor per chunk.
    ``frame_range = (lo, hi)``: only those frames (a rank's shard; indices stay absolute) -- the
    frames outside it are never decoded.
    Raises ``_DecodeFallback`` (before or between chunks) when a read comes back short or the
    container holds more frames than its header says."""
    import time
    from concurrent.futures import ThreadPoolExecutor

    workers = decode_workers(workers)
    info = video_info(avi_path)
    # a rank's shard is read by range even with one decoder thread (the alternative is decoding the
    # whole file on every rank); a whole clip only when several threads make it worth it
    if not (seekable_clip(info) if frame_range is not None else parallel_decodable(info, workers)):
        raise _DecodeFallback("not an intra-only clip with a plausible header")
    n, hgt, wid = info["frames"], info["height"], info["width"]
    lo, hi = (0, n) if frame_range is None else (max(0, frame_range[0]), min(n, frame_range[1]))
    chunk = max(workers, min(chunk, max(hi - lo, 1)))
    bufs = [torch.empty((chunk, hgt, wid, 3), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    views = [b.numpy() for b in bufs]
    copy = torch.cuda.Stream(device=dev)
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    if stats is not None:
        stats.update(frames=n, height=hgt, width=wid, workers=workers, decode_s=0.0)
    with _silence_stderr(), ThreadPoolExecutor(workers) as pool:
        decs = list(pool.map(lambda _: RangeDecoder(avi_path), range(workers)))
        try:
            for k, i0 in enumerate(range(lo, hi, chunk)):
                m, b = min(chunk, hi - i0), k % 2
                if k >= 2:
                    freed[b].synchronize()        # the copy that read this buffer has finished
                cut = [m * w // workers for w in range(workers + 1)]
                t0 = time.perf_counter()
                got = list(pool.map(lambda w: decs[w].read_into(i0 + cut[w], i0 + cut[w + 1],
                                                                views[b][cut[w]:cut[w + 1]]),
                                    range(workers)))
                if stats is not None:
                    stats["decode_s"] += time.perf_counter() - t0
                if any(g != cut[w + 1] - cut[w] for w, g in enumerate(got)):
                    raise _DecodeFallback("short read")
                if i0 + m == n and decs[-1].pos == n and decs[-1].cap.read()[0]:
                    raise _DecodeFallback("frames beyond the header's count")
                cur = torch.cuda.current_stream(dev)
                gray = out[i0:i0 + m] if out is not None else torch.empty(
                    (m, hgt, wid), dtype=torch.uint8, device=dev)
                copy.wait_stream(cur)             # earlier users of this memory are done
                with torch.cuda.stream(copy):
                    gray.copy_(bgr_to_gray(bufs[b][:m].to(dev, non_blocking=True)))
                    freed[b].record(copy)
                cur.wait_stream(copy)
                yield i0, gray
        finally:
            for d in decs:
                d.release()


def decode_gray_clip(avi_path: str, dev: torch.device, workers: int | None = None,
                     chunk: int = 1024, timings: dict | None = None) -> torch.Tensor | None:
    """Video file -> ``(N, H, W)`` uint8 gray CUDA tensor, or ``None`` for a file without frames:
    ``load_frames_bgr`` + ``cvtColor`` of /root/reference/openglottal/features.py:226,235.
    Intra-only codecs (MJPG, FFV1, raw) go through ``iter_gray_chunks`` -- the frames are those
    the reference's sequential loop decodes (same decoder, frame-exact seeks; checked in tests) --
    anything else, or anything that decoder does not trust, through the sequential loop.
    ``timings`` (optional dict) receives ``decode_s`` / ``total_s`` / ``workers`` / ``mode``."""
    import time

    t_start = time.perf_counter()
    stats: dict = {}
    try:
        info = video_info(avi_path)
        gray = torch.empty((max(info["frames"], 0), max(info["height"], 0), max(info["width"], 0)),
                           dtype=torch.uint8, device=dev)
        for _ in iter_gray_chunks(avi_path, dev, workers, chunk, out=gray, stats=stats):
            pass
        if timings is not None:
            torch.cuda.synchronize(dev)
            timings.update(mode="parallel", workers=stats["workers"], decode_s=stats["decode_s"],
                           total_s=time.perf_counter() - t_start)
        return gray
    except _DecodeFallback:
        pass
    frames_bgr = load_frames_bgr(avi_path)
    t_dec = time.perf_counter()
    gray = gray_clip_from_bgr(frames_bgr, dev) if frames_bgr else None
    if timings is not None:
        torch.cuda.synchronize(dev)
        timings.update(mode="sequential", workers=1, decode_s=t_dec - t_start,
                       total_s=time.perf_counter() - t_start)
    return gray


def extract_features_unet_frames(frames_gray, model: UNet, batch: int = 512,
                                 threshold: float = 0.5, group=None) -> dict | None:
    """unet-only pipeline on raw gray frames ``(N, H, W)`` uint8 (H, W multiples of 16).

    With ``torch.distributed`` initialised (one process per GPU) every rank passes the FULL clip
    (or at least its own shard, see ``sharding.shard_range``); each rank segments its contiguous
    frame range and the int32 areas are all-gathered, so every rank returns the same dict.
    """
    n = frames_gray.shape[0]
    rank, world = sharding.rank_world(group)
    lo, hi = sharding.shard_range(n, rank, world)
    if hi > lo:
        local, _ = segment_clip(frames_gray[lo:hi], model, batch=batch, threshold=threshold)
    else:
        local = torch.empty(0, dtype=torch.int32, device=model._device())
    area = sharding.gather_area(local, n, group)
    return kinematic_features_device(area)


def masks_for_clip(frames_gray, model: UNet, threshold: float = 0.5, want_masks: bool = False,
                   batch: int = 512):
    """Area waveform (and masks) of a clip with the REFERENCE's per-frame semantics
    (/root/reference/openglottal/utils.py:218-241): every caller of ``unet_segment_frame`` in the
    reference -- features.py:236, scripts/infer.py:217, scripts/analyze_gaw.py:86 -- squashes the
    frame to 256 x 256 and resizes the probability back. 256 x 256 clips (both resizes are the
    identity) take the streaming native path; other sizes the device resize kernels. No per-frame
    host work in either. Returns ``(area int32 CUDA (N,), masks uint8 CUDA (N, H, W) | None)``."""
    from .utils import NET_SIZE, segment_frames_reference_resize

    model = _require_native(model)
    if isinstance(frames_gray, np.ndarray):
        frames_gray = torch.from_numpy(np.ascontiguousarray(frames_gray))
    if frames_gray.dim() != 3 or frames_gray.dtype != torch.uint8:
        raise ValueError("expected (N, H, W) uint8 frames")
    if tuple(frames_gray.shape[1:]) == (NET_SIZE, NET_SIZE):
        return segment_clip(frames_gray, model, batch=batch, threshold=threshold,
                            want_masks=want_masks)
    dev = model._device()
    areas, masks = [], []
    for i0 in range(0, frames_gray.shape[0], batch):
        part = frames_gray[i0:i0 + batch].to(dev, non_blocking=True)
        a, m = segment_frames_reference_resize(part, model, threshold=threshold,
                                               want_masks=want_masks, batch=batch)
        areas.append(a)
        masks.append(m)
    return torch.cat(areas), (torch.cat(masks) if want_masks else None)


def _local_world(world: int) -> int:
    """Ranks sharing this host (torchrun's LOCAL_WORLD_SIZE; all of them when it is not set)."""
    import os

    try:
        return max(1, min(world, int(os.environ.get("LOCAL_WORLD_SIZE", world))))
    except ValueError:
        return max(1, world)


def extract_features_unet(avi_path: str, detector, model, device=None, *,
                          decode_chunk: int = 1024, group=None) -> dict | None:
    """Drop-in for ``openglottal.extract_features_unet`` (features.py:202-247).

    ``detector is None`` (unet-only) is the accelerated path. With a detector the masks come
    from the native kernels in batches, the detector (the reference's Ultralytics
    ``TemporalDetector``, or anything with ``reset()`` / ``detect(frame_bgr)``) is called once
    per frame in order as the reference does, and the bbox gating of features.py:240-245 runs
    as one CUDA reduction over the batch of masks. ``decode_chunk`` / ``group`` (keyword only,
    not in the reference): frames per pinned staging buffer of the streaming reader; the process
    group the frames are sharded over (default: the world, when torch.distributed is initialised).
    """
    from .utils import gated_area

    model = _require_native(model)
    dev = model._device()
    if detector is None:
        # unet-only: no consumer of the BGR frames on the host, so they are never held as a list.
        # Intra-only clips stream: chunk k is segmented on the GPU while chunk k + 1 decodes.
        # With torch.distributed initialised (one process per GPU) every rank decodes and
        # segments only its contiguous frame range (sharding.shard_range: a seek, the frames of
        # the other ranks are never decoded), with its share of the host's decoder threads; the
        # int32 areas are all-gathered, so every rank returns the same dict.
        rank, world = sharding.rank_world(group)
        fell_back = False
        try:
            n = video_info(avi_path)["frames"]
            lo, hi = sharding.shard_range(n, rank, world)
            local = torch.empty(hi - lo, dtype=torch.int32, device=dev)
            workers = max(1, decode_workers() // _local_world(world))
            for i0, part in iter_gray_chunks(avi_path, dev, workers=workers, chunk=decode_chunk,
                                             frame_range=(lo, hi) if world > 1 else None):
                local[i0 - lo:i0 - lo + part.shape[0]] = masks_for_clip(part, model)[0]
        except _DecodeFallback:
            fell_back = True
        # a short read or a wrong header shows up only on the rank whose range it falls into: the
        # ranks agree before the gather, or they would wait for each other in different collectives
        if not sharding.agree_any(fell_back, dev, group):
            return kinematic_features_device(sharding.gather_area(local, n, group))
        # sequential decode (any codec): every rank reads the file, segments its own range
        gray = decode_gray_clip(avi_path, dev, workers=1)
        if gray is None:
            return None
        n = gray.shape[0]
        lo, hi = sharding.shard_range(n, rank, world)
        if hi > lo:
            local, _ = masks_for_clip(gray[lo:hi], model)
        else:
            local = torch.empty(0, dtype=torch.int32, device=dev)
        return kinematic_features_device(sharding.gather_area(local, n, group))
    frames_bgr = load_frames_bgr(avi_path)
    if not frames_bgr:
        return None
    detector.reset()
    boxes = [detector.detect(frm) for frm in frames_bgr]
    _, masks = masks_for_clip(gray_clip_from_bgr(frames_bgr, dev), model, want_masks=True)
    return kinematic_features_device(gated_area(masks, boxes))


def extract_features_yolo_crop_unet(avi_path: str, detector, model, device=None,
                                    crop_size: int = 256) -> dict | None:
    """Area waveform + features of the ``yolo-crop+unet`` pipeline
    (/root/reference/scripts/infer.py:222-248): per frame, the detector's box is cropped from the
    gray frame, letterboxed to ``crop_size``, segmented by the crop-trained U-Net, un-letterboxed
    and counted; frames without a box count 0. Cropping, letterboxing, un-letterboxing and
    counting run on the GPU for the whole clip; the detector is the caller's."""
    from .utils import segment_crops

    model = _require_native(model)
    dev = model._device()
    frames_bgr = load_frames_bgr(avi_path)
    if not frames_bgr:
        return None
    detector.reset()
    boxes = [detector.detect(frm) for frm in frames_bgr]
    area, _ = segment_crops(gray_clip_from_bgr(frames_bgr, dev), boxes, model, size=crop_size)
    return kinematic_features_device(area)
