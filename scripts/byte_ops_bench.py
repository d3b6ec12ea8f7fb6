"""The byte / integer operators around the U-Net (BGR->gray, detection-gated
area, letterbox / un-letterbox, Dice / IoU counts, the two cv2-compatible resizes) against the HBM
roof, through the C ABI, with inputs larger than the 126 MB L2.

    python scripts/byte_ops_bench.py [--prev other_build.so] [--out gpurun_out/byte_ops.json]

Per operator: algorithmic bytes (what has to be read and written once), CUDA-event time on the
launching stream, GB/s and the fraction of MEASURED_PEAKS.json's copy bandwidth. With ``--prev`` the
same calls go to a second build loaded into the same process: every output must be bit-identical
(this is how a rewritten kernel is checked against the one the parity tests pinned), and its times
are reported beside the new ones. A few ragged cases (unaligned pointers, sizes that are not
multiples of 4 / 16, boxes with negative and empty slices) are compared for equality only.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from openglottal_b200 import _native  # noqa: E402
from openglottal_b200.utils import _crop_geometry  # noqa: E402


def hbm_peak():
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]), "measured hbm_gbs"
    except Exception:
        return 6452.8, "fallback (profiling guide)"


def timed(fn, iters: int = 5) -> float:
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def ck(lib, rc):
    if rc != 0:
        raise RuntimeError((lib.ogl_last_error() or b"?").decode())


def random_boxes(n, hgt, wid, rng, exotic=False, min_side=1):
    """Every ninth box is None. ``min_side`` >= 8 keeps a crop from collapsing under letterboxing
    (cv2.resize to a zero-sized image fails in the reference, _crop_geometry raises)."""
    boxes = []
    for i in range(n):
        if i % 9 == 4:
            boxes.append(None)
            continue
        x1, y1 = int(rng.integers(0, wid - 8)), int(rng.integers(0, hgt - 8))
        x2, y2 = int(rng.integers(x1 + min_side, wid + 1)), int(rng.integers(y1 + min_side, hgt + 1))
        if exotic and i % 5 == 0:
            x1, x2 = x1 - wid, x2 if x2 < wid else wid + 7      # negative start, end past the edge
        if exotic and i % 11 == 3:
            x2 = x1                                               # empty slice
        boxes.append((x1, y1, x2, y2))
    return boxes


class Ops:
    """One build's operators on torch tensors (outputs freshly allocated per call)."""

    def __init__(self, lib):
        self.lib = lib
        self.s = torch.cuda.current_stream().cuda_stream

    def bgr(self, bgr, out=None):
        out = torch.empty(bgr.shape[:-1], dtype=torch.uint8, device=bgr.device) if out is None else out
        ck(self.lib, self.lib.ogl_bgr_to_gray(bgr.data_ptr(), out.data_ptr(), out.numel(), self.s))
        return out

    def bgr_raw(self, src, dst, pixels):      # flat byte buffers (alignment cases)
        ck(self.lib, self.lib.ogl_bgr_to_gray(src.data_ptr(), dst.data_ptr(), pixels, self.s))
        return dst

    def gated(self, masks, boxes_d, has_d):
        n, h, w = masks.shape
        area = torch.empty(n, dtype=torch.int32, device=masks.device)
        ck(self.lib, self.lib.ogl_mask_area_boxes(masks.data_ptr(), n, h, w, boxes_d.data_ptr(),
                                                  has_d.data_ptr(), area.data_ptr(), self.s))
        return area

    def overlap(self, p, g):
        n = p.shape[0]
        out = torch.empty((n, 3), dtype=torch.int32, device=p.device)
        ck(self.lib, self.lib.ogl_mask_overlap_counts(p.data_ptr(), g.data_ptr(), n, p[0].numel(),
                                                      out.data_ptr(), self.s))
        return out

    def letterbox(self, gray, geom_d, size):
        n, h, w = gray.shape
        out = torch.empty((n, size, size), dtype=torch.uint8, device=gray.device)
        ck(self.lib, self.lib.ogl_letterbox_crops(gray.data_ptr(), n, h, w, geom_d.data_ptr(), size,
                                                  out.data_ptr(), self.s))
        return out

    def unletterbox(self, mask_cs, geom_d, h, w):
        n, size, _ = mask_cs.shape
        area = torch.empty(n, dtype=torch.int32, device=mask_cs.device)
        full = torch.empty((n, h, w), dtype=torch.uint8, device=mask_cs.device)
        ck(self.lib, self.lib.ogl_unletterbox_area(mask_cs.data_ptr(), n, size, geom_d.data_ptr(), h, w,
                                                   full.data_ptr(), area.data_ptr(), self.s))
        return area, full

    def resize(self, src, dh, dw):
        n, sh, sw = src.shape
        out = torch.empty((n, dh, dw), dtype=torch.uint8, device=src.device)
        ck(self.lib, self.lib.ogl_resize_u8_linear(src.data_ptr(), n, sh, sw, out.data_ptr(), dh, dw, self.s))
        return out

    def prob(self, logits, dh, dw):
        n, sh, sw = logits.shape
        mask = torch.empty((n, dh, dw), dtype=torch.uint8, device=logits.device)
        area = torch.empty(n, dtype=torch.int32, device=logits.device)
        ck(self.lib, self.lib.ogl_prob_resize_mask(logits.data_ptr(), n, sh, sw, dh, dw, 0.5,
                                                   mask.data_ptr(), area.data_ptr(), self.s))
        return mask, area


def same(a, b) -> bool:
    if isinstance(a, tuple):
        return all(same(x, y) for x, y in zip(a, b))
    return bool(torch.equal(a, b))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--prev", default=None)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "byte_ops.json"))
    ap.add_argument("--frames", type=int, default=2048)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    new = Ops(_native.load())
    prev = Ops(_native.load_path(args.prev)) if args.prev else None
    peak, peak_src = hbm_peak()
    rng = np.random.default_rng(7)
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    n, rows = args.frames, []

    def u8(shape, sparse=False):
        t = torch.randint(0, 256, shape, dtype=torch.uint8, device=dev, generator=g)
        if sparse:      # masks: {0, 255} with a few other non-zero values (the operators test "> 0")
            t = torch.where(t > 140, t | 1, torch.zeros_like(t))
        return t

    def geom_for(boxes, h, w, size):
        geom = _crop_geometry(boxes, h, w, size)
        return geom, torch.from_numpy(geom).to(dev)

    def boxes_dev(boxes):
        host = np.zeros((len(boxes), 4), np.int32)
        has = np.zeros(len(boxes), np.uint8)
        for i, b in enumerate(boxes):
            if b is not None:
                host[i], has[i] = b, 1
        return torch.from_numpy(host).to(dev), torch.from_numpy(has).to(dev)

    def case(name, shape, nbytes, call, timing=True):
        row = {"op": name, "shape": shape, "algorithmic_bytes": int(nbytes)}
        try:
            got = call(new)
            if timing:
                ms = timed(lambda: call(new))
                row.update(ms=ms, gbs=nbytes / ms / 1e6, frac_of_hbm_peak=nbytes / ms / 1e6 / peak)
            if prev is not None:
                row["equal_to_prev_build"] = same(got, call(prev))
                if timing:
                    ms_p = timed(lambda: call(prev))
                    row.update(prev_ms=ms_p, prev_gbs=nbytes / ms_p / 1e6)
        except Exception as e:      # keep measuring the other operators
            row["error"] = f"{type(e).__name__}: {e}"
        rows.append(row)
        print(json.dumps(row), flush=True)

    # ---- timed cases: aligned, BASELINE-shaped, larger than L2
    bgr = u8((n // 2, 256, 256, 3))
    case("bgr_to_gray", f"{n // 2} x 256x256x3", bgr.numel() + bgr.numel() // 3, lambda o: o.bgr(bgr))
    del bgr

    masks = u8((n, 256, 256), sparse=True)
    full_boxes = boxes_dev([(0, 0, 256, 256)] * n)
    case("mask_area_boxes (whole frame)", f"{n} x 256x256", masks.numel(),
         lambda o: o.gated(masks, *full_boxes))
    rb = random_boxes(n, 256, 256, rng, exotic=True)
    rb_d = boxes_dev(rb)
    inside = sum(max(0, len(range(*slice(b[0], b[2]).indices(256)))) *
                 max(0, len(range(*slice(b[1], b[3]).indices(256)))) for b in rb if b is not None)
    case("mask_area_boxes (random boxes, Python slices)", f"{n} x 256x256", inside,
         lambda o: o.gated(masks, *rb_d))
    gt = u8((n, 256, 256), sparse=True)
    case("mask_overlap_counts", f"2 x {n} x 256x256", 2 * masks.numel(), lambda o: o.overlap(masks, gt))
    del gt

    gray = u8((n, 256, 256))
    lb_boxes = random_boxes(n, 256, 256, rng, min_side=8)
    geom, geom_d = geom_for(lb_boxes, 256, 256, 256)
    crop_px = int(((geom[:, 2] - geom[:, 0]) * (geom[:, 3] - geom[:, 1])).sum())
    cont_px = int((geom[:, 6] * geom[:, 7]).sum())
    case("letterbox_crops", f"{n} x 256x256 -> 256x256", n * 256 * 256 + min(crop_px, cont_px),
         lambda o: o.letterbox(gray, geom_d, 256))
    case("unletterbox_area (+ full mask)", f"{n} x 256x256 -> 256x256", n * 256 * 256 + min(crop_px, cont_px),
         lambda o: o.unletterbox(masks, geom_d, 256, 256))
    del gray, masks

    tall = u8((n, 512, 256))
    case("resize_u8_linear 512x256 -> 256x256", f"{n} frames", tall.numel() + n * 65536,
         lambda o: o.resize(tall, 256, 256))
    del tall
    big = u8((n // 2, 512, 512))
    case("resize_u8_linear 512x512 -> 256x256 (2x2 area)", f"{n // 2} frames", big.numel() + (n // 2) * 65536,
         lambda o: o.resize(big, 256, 256))
    del big
    logits = torch.randn((n // 2, 256, 256), device=dev, generator=g) * 4
    case("prob_resize_mask 256x256 -> 512x256", f"{n // 2} frames", logits.numel() * 4 + (n // 2) * 131072,
         lambda o: o.prob(logits, 512, 256))
    case("prob_resize_mask 256x256 -> 256x256 (identity)", f"{n // 2} frames", logits.numel() * 4 + (n // 2) * 65536,
         lambda o: o.prob(logits, 256, 256))

    # ---- ragged cases: equality with the previous build only
    if prev is not None:
        small = torch.randn((5, 96, 128), device=dev, generator=g) * 4
        case("prob_resize_mask 96x128 -> 250x300 / 64x64 -> 100x36", "ragged", 0,
             lambda o: (o.prob(small, 250, 300), o.prob(small[:, :64, :64].contiguous(), 100, 36)), timing=False)
        odd = u8((7, 250, 300))
        case("resize_u8_linear 250x300 -> 256x256 / 96x100 / 125x150", "ragged", 0,
             lambda o: (o.resize(odd, 256, 256), o.resize(odd, 96, 100), o.resize(odd, 125, 150)), timing=False)
        m_odd = u8((9, 250, 300), sparse=True)
        ob = random_boxes(9, 250, 300, rng, exotic=True)
        case("mask_area_boxes 250x300 (scalar path) / 64x48 (vector path)", "ragged", 0,
             lambda o: (o.gated(m_odd, *boxes_dev(ob)),
                        o.gated(m_odd[:, :64, :48].contiguous(), *boxes_dev(random_boxes(9, 64, 48, np.random.default_rng(3), True)))),
             timing=False)
        case("mask_overlap_counts 250x300 / 9 x 75000 px", "ragged", 0,
             lambda o: o.overlap(m_odd, m_odd.flip(0).contiguous()), timing=False)
        g_odd = u8((6, 250, 300))
        for size in (256, 250, 64):
            gm, gm_d = geom_for(random_boxes(6, 250, 300, rng, min_side=8), 250, 300, size)
            cs = u8((6, size, size), sparse=True)
            case(f"letterbox_crops / unletterbox_area 250x300, size {size}", "ragged", 0,
                 lambda o: (o.letterbox(g_odd, gm_d, size), o.unletterbox(cs, gm_d, 250, 300)), timing=False)
        for (sh, sw), (dh, dw) in (((256, 256), (512, 256)), ((96, 128), (384, 512)), ((480, 640), (120, 160)),
                                   ((300, 200), (256, 256)), ((64, 64), (1024, 64)), ((1024, 64), (64, 64))):
            lg = torch.randn((3, sh, sw), device=dev, generator=g) * 4
            lg[0, 0, :7] = torch.tensor([float("inf"), float("-inf"), float("nan"), 0.0, -0.0, 88.0, -104.0],
                                        device=dev)[: min(7, sw)]
            case(f"prob_resize_mask {sh}x{sw} -> {dh}x{dw} (incl. inf / nan logits)", "ragged", 0,
                 lambda o: o.prob(lg, dh, dw), timing=False)
            src = u8((3, sh, sw))
            case(f"resize_u8_linear {sh}x{sw} -> {dh}x{dw}", "ragged", 0, lambda o: o.resize(src, dh, dw),
                 timing=False)
        flat = u8((3 * 100_003 + 64,))
        for off_s, off_d, px in ((0, 0, 100_003), (1, 0, 99_999), (0, 3, 4096), (16, 16, 15), (0, 0, 16)):
            def call(o, off_s=off_s, off_d=off_d, px=px):
                dst = torch.zeros(px + 64, dtype=torch.uint8, device=dev)
                o.bgr_raw(flat[off_s:], dst[off_d:], px)
                return dst
            case(f"bgr_to_gray {px} px, src +{off_s} B, dst +{off_d} B", "ragged", 0, call, timing=False)

    out = {"hbm_peak_gbs": peak, "peak_source": peak_src, "frames": n, "gpu": torch.cuda.get_device_name(0),
           "prev_build": args.prev, "rows": rows}
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(out, indent=1))
    bad = [r["op"] for r in rows if r.get("error") or r.get("equal_to_prev_build") is False]
    print("byte_ops: " + ("ALL EQUAL / OK" if not bad else f"PROBLEMS: {bad}"))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
