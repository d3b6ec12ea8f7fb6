"""Per-kernel parity of the tcgen05 implicit-GEMM layers against torch.nn.functional on the
same bf16-rounded operands (fp32 accumulate on both sides), called through the C ABI
(ogl_debug_tc_layer). Tolerance: the kernel stores bf16, so |err| <= 2^-8 * |ref| + 1e-3."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


@pytest.fixture(scope="module")
def handle(lib):
    from openglottal_b200 import _native

    h = C.c_void_p()
    _native.check(lib.ogl_unet_create(C.byref(h), 0))
    yield h
    lib.ogl_unet_destroy(h)


def _run_layer(lib, handle, kind, x0, x1, w, b, cout):
    from openglottal_b200 import _native

    n, c0, hgt, wid = x0.shape
    c1 = 0 if x1 is None else x1.shape[1]
    oh, ow = (2 * hgt, 2 * wid) if kind == 3 else (hgt, wid)
    out = torch.full((n, cout, oh, ow), float("nan"), device="cuda")
    pool = torch.full((n, cout, hgt // 2, wid // 2), float("nan"), device="cuda") if kind == 1 else None
    wh = w.contiguous().cpu()
    bh = b.contiguous().cpu()
    rc = lib.ogl_debug_tc_layer(
        handle, kind, x0.data_ptr(), c0, None if x1 is None else x1.data_ptr(), c1,
        C.cast(wh.data_ptr(), C.POINTER(C.c_float)), C.cast(bh.data_ptr(), C.POINTER(C.c_float)),
        cout, n, hgt, wid, out.data_ptr(), None if pool is None else pool.data_ptr(), None)
    _native.check(rc)
    torch.cuda.synchronize()
    return out, pool


def _report(name, got, ref):
    err = (got - ref).abs()
    tol = ref.abs() * 2.0 ** -7 + 2e-3
    bad = err > tol
    nan = torch.isnan(got).sum().item()
    msg = (f"{name}: max|err|={err.max().item():.4g} ref_rms={ref.pow(2).mean().sqrt().item():.4g} "
           f"bad={bad.sum().item()}/{bad.numel()} nan={nan}")
    if bad.any():
        idx = bad.nonzero()
        msg += f" first_bad={idx[0].tolist()} per-channel bad={bad.sum((0, 2, 3)).tolist()[:16]}"
        msg += f" per-row bad={bad.sum((0, 1, 3)).tolist()[:20]} per-col bad={bad.sum((0, 1, 2)).tolist()[:20]}"
    print(msg)
    assert nan == 0 and not bad.any(), msg


CONV_CASES = [
    # kind, c0, c1, cout, n, H, W
    (0, 32, 0, 32, 2, 16, 16),
    (0, 32, 0, 32, 2, 32, 48),
    (0, 32, 0, 64, 3, 32, 32),
    (1, 64, 0, 64, 2, 32, 32),
    (0, 32, 32, 32, 2, 32, 32),
    (0, 128, 128, 128, 1, 32, 32),
    (0, 256, 0, 256, 3, 16, 16),
    (1, 256, 0, 256, 2, 32, 32),
    (0, 256, 0, 512, 5, 16, 16),
    (0, 512, 0, 512, 2, 16, 16),
]


@pytest.mark.parametrize("kind,c0,c1,cout,n,hgt,wid", CONV_CASES)
def test_conv3x3_tc(lib, handle, kind, c0, c1, cout, n, hgt, wid):
    g = torch.Generator().manual_seed(c0 * 7 + cout + hgt)
    cin = c0 + c1
    x = _bf(torch.randn(n, cin, hgt, wid, generator=g))
    w = _bf(torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1)).float()
    xc = x.cuda()
    x0 = xc[:, :c0].contiguous()
    x1 = xc[:, c0:].contiguous() if c1 else None
    out, pool = _run_layer(lib, handle, kind, x0, x1, w, b, cout)
    _report(f"conv {c0}+{c1}->{cout} {n}x{hgt}x{wid}", out.cpu(), ref)
    if kind == 1:
        _report("pooled", pool.cpu(), F.max_pool2d(_bf(ref), 2, 2))


@pytest.mark.parametrize("cin,cout,n,hgt,wid", [(64, 32, 2, 16, 16), (512, 256, 3, 16, 16),
                                                (128, 64, 1, 32, 48)])
def test_convt2x2_tc(lib, handle, cin, cout, n, hgt, wid):
    g = torch.Generator().manual_seed(cin + cout)
    x = _bf(torch.randn(n, cin, hgt, wid, generator=g))
    w = _bf(torch.randn(cin, cout, 2, 2, generator=g) * (1.0 / cin) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.conv_transpose2d(x.double(), w.double(), b.double(), stride=2).float()
    out, _ = _run_layer(lib, handle, 3, x.cuda(), None, w, b, cout)
    _report(f"convT {cin}->{cout} {n}x{hgt}x{wid}", out.cpu(), ref)


@pytest.mark.parametrize("cin,cout,n,hgt,wid", [(128, 64, 3, 32, 48), (512, 256, 5, 16, 16),
                                                (256, 128, 2, 64, 32), (128, 64, 1, 16, 16)])
def test_convt_cta_pairs_bit_identical(lib, handle, cin, cout, n, hgt, wid):
    """Transposed conv on CTA pairs (one sub-tile per CTA, double-buffered accumulators; only
    with OGL_CONVT_PAIR=1, else this compares the one-CTA form with itself): same MMAs on the
    same operands, so the same bits, odd tile counts included."""
    from openglottal_b200 import _native

    g = torch.Generator().manual_seed(cin + cout + hgt)
    x = _bf(torch.randn(n, cin, hgt, wid, generator=g))
    w = _bf(torch.randn(cin, cout, 2, 2, generator=g) * (1.0 / cin) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    try:
        _native.check(lib.ogl_unet_set_cta_pairs(handle, 1))
        out1, _ = _run_layer(lib, handle, 3, x.cuda(), None, w, b, cout)
        _native.check(lib.ogl_unet_set_cta_pairs(handle, 3))
        out2, _ = _run_layer(lib, handle, 3, x.cuda(), None, w, b, cout)
    finally:
        _native.check(lib.ogl_unet_set_cta_pairs(handle, 2))
    ref = F.conv_transpose2d(x.double(), w.double(), b.double(), stride=2).float()
    _report(f"pair convT {cin}->{cout} {n}x{hgt}x{wid}", out2.cpu(), ref)
    assert not torch.isnan(out2).any()
    assert torch.equal(out1, out2)


# ------------------------------------------------------------------ space-to-depth layers
def _run_s2d(lib, handle, kind, x, below, w3, b3, wt, bt):
    from openglottal_b200 import _native

    fp = C.POINTER(C.c_float)
    n, cin_s, hgt, wid = x.shape
    out = torch.full((n, 32, hgt, wid), float("nan"), device="cuda")
    pool = torch.full((n, 32, hgt // 2, wid // 2), float("nan"), device="cuda") if kind == 1 else None
    host = [t.contiguous().cpu().float() if t is not None else None for t in (w3, b3, wt, bt)]
    ptr = lambda t: None if t is None else C.cast(t.data_ptr(), fp)
    xd = x.cuda().contiguous()
    bd = None if below is None else below.cuda().contiguous()
    rc = lib.ogl_debug_s2d_layer(handle, kind, xd.data_ptr(), cin_s,
                                 None if bd is None else bd.data_ptr(), ptr(host[0]), ptr(host[1]),
                                 ptr(host[2]), ptr(host[3]), n, hgt, wid, out.data_ptr(),
                                 None if pool is None else pool.data_ptr(), None)
    _native.check(rc)
    torch.cuda.synchronize()
    return out, pool


@pytest.mark.parametrize("kind,cin_s,n,hgt,wid", [
    (0, 32, 2, 16, 16), (0, 32, 3, 32, 48), (1, 32, 2, 48, 32), (0, 16, 1, 16, 32),
    (1, 32, 5, 64, 64), (0, 32, 40, 64, 64), (1, 32, 2, 256, 256)])
def test_s2d_conv3x3(lib, handle, kind, cin_s, n, hgt, wid):
    g = torch.Generator().manual_seed(cin_s * 3 + hgt + n)
    x = _bf(torch.randn(n, cin_s, hgt, wid, generator=g))
    w = _bf(torch.randn(32, cin_s, 3, 3, generator=g) * (2.0 / (cin_s * 9)) ** 0.5)
    b = torch.randn(32, generator=g) * 0.1
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1)).float()
    out, pool = _run_s2d(lib, handle, kind, x, None, w, b, None, None)
    _report(f"s2d conv {cin_s}->32 {n}x{hgt}x{wid}", out.cpu(), ref)
    if kind == 1:
        _report("s2d pooled", pool.cpu(), F.max_pool2d(_bf(ref), 2, 2))


@pytest.mark.parametrize("n,hgt,wid", [(2, 32, 32), (3, 48, 80), (2, 256, 256)])
def test_s2d_conv3x3_with_composed_convt(lib, handle, n, hgt, wid):
    """ups.6 + cat + ups.7.net.0 (unet.py:82-87) in one launch. The composed weights are
    rounded to bf16 once (instead of `up` being rounded), hence the wider tolerance."""
    g = torch.Generator().manual_seed(hgt + wid)
    skip = _bf(torch.randn(n, 32, hgt, wid, generator=g))
    below = _bf(torch.randn(n, 64, hgt // 2, wid // 2, generator=g))
    w3 = _bf(torch.randn(32, 64, 3, 3, generator=g) * (2.0 / (64 * 9)) ** 0.5)
    b3 = torch.randn(32, generator=g) * 0.1
    wt = _bf(torch.randn(64, 32, 2, 2, generator=g) * (1.0 / 64) ** 0.5)
    bt = torch.randn(32, generator=g) * 0.5
    up = F.conv_transpose2d(below.double(), wt.double(), bt.double(), stride=2)
    ref = F.relu(F.conv2d(torch.cat([skip.double(), up], 1), w3.double(), b3.double(), padding=1)).float()
    out, _ = _run_s2d(lib, handle, 0, skip, below, w3, b3, wt, bt)
    got = out.cpu()
    err = (got - ref).abs()
    msg = f"s2d composed {n}x{hgt}x{wid}: max|err|={err.max().item():.4g} mean={err.mean().item():.3g}"
    print(msg)
    assert not torch.isnan(got).any()
    assert err.max() <= 2e-2 * max(1.0, ref.abs().max().item() / 2) and err.mean() <= 3e-3, msg
    # border pixels carry the bias table: check them separately
    edge = torch.ones_like(err, dtype=torch.bool)
    edge[:, :, 1:-1, 1:-1] = False
    assert err[edge].max() <= 2e-2 * max(1.0, ref.abs().max().item() / 2), msg


# ------------------------------------------------------------------ CTA pairs (cta_group::2)
@pytest.mark.parametrize("kind,c0,c1,cout,n,hgt,wid", [c for c in CONV_CASES if c[3] >= 64] + [
    (0, 64, 0, 64, 9, 64, 64), (1, 128, 0, 128, 7, 48, 32), (0, 64, 64, 64, 5, 32, 32),
    (0, 512, 0, 256, 3, 32, 32)])
def test_conv3x3_cta_pairs_bit_identical(lib, handle, kind, c0, c1, cout, n, hgt, wid):
    """The CTA-pair form issues the same MMAs on the same operands (256 rows at a time), so its
    output must equal the one-CTA form bit for bit; and both match the fp64 reference."""
    from openglottal_b200 import _native

    g = torch.Generator().manual_seed(c0 * 5 + cout + wid)
    cin = c0 + c1
    x = _bf(torch.randn(n, cin, hgt, wid, generator=g))
    w = _bf(torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    xc = x.cuda()
    x0 = xc[:, :c0].contiguous()
    x1 = xc[:, c0:].contiguous() if c1 else None
    try:
        _native.check(lib.ogl_unet_set_cta_pairs(handle, 1))
        out1, pool1 = _run_layer(lib, handle, kind, x0, x1, w, b, cout)
        _native.check(lib.ogl_unet_set_cta_pairs(handle, 3))
        out2, pool2 = _run_layer(lib, handle, kind, x0, x1, w, b, cout)
    finally:
        _native.check(lib.ogl_unet_set_cta_pairs(handle, 2))
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1)).float()
    _report(f"pair conv {c0}+{c1}->{cout} {n}x{hgt}x{wid}", out2.cpu(), ref)
    assert torch.equal(out1, out2)
    if kind == 1:
        assert torch.equal(pool1, pool2)


@pytest.mark.parametrize("kind,composed,n,hgt,wid", [(0, False, 3, 32, 48), (1, False, 5, 64, 64),
                                                     (0, True, 3, 48, 80), (1, False, 2, 256, 256),
                                                     (0, True, 2, 256, 256), (0, False, 1, 16, 16)])
def test_s2d_cta_pairs_bit_identical(lib, handle, kind, composed, n, hgt, wid):
    """Space-to-depth layers on CTA pairs (each CTA keeps half of every weight block): same bits
    as the one-CTA form, odd tile counts included (the last pair lacks a tile)."""
    from openglottal_b200 import _native

    g = torch.Generator().manual_seed(hgt * 3 + wid + n)
    x = _bf(torch.randn(n, 32, hgt, wid, generator=g))
    cin3 = 64 if composed else 32
    w3 = _bf(torch.randn(32, cin3, 3, 3, generator=g) * (2.0 / (cin3 * 9)) ** 0.5)
    b3 = torch.randn(32, generator=g) * 0.1
    below = wt = bt = None
    if composed:
        below = _bf(torch.randn(n, 64, hgt // 2, wid // 2, generator=g))
        wt = _bf(torch.randn(64, 32, 2, 2, generator=g) * (1.0 / 64) ** 0.5)
        bt = torch.randn(32, generator=g) * 0.5
    try:
        _native.check(lib.ogl_unet_set_cta_pairs(handle, 1))
        out1, pool1 = _run_s2d(lib, handle, kind, x, below, w3, b3, wt, bt)
        _native.check(lib.ogl_unet_set_cta_pairs(handle, 3))
        out2, pool2 = _run_s2d(lib, handle, kind, x, below, w3, b3, wt, bt)
    finally:
        _native.check(lib.ogl_unet_set_cta_pairs(handle, 2))
    assert not torch.isnan(out2).any()
    assert torch.equal(out1, out2)
    if kind == 1:
        assert torch.equal(pool1, pool2)


UPCAT_CASES = [
    # f, n, H, W (the level's own resolution)
    (64, 2, 32, 32),
    (64, 3, 48, 80),      # partial tiles in x and y
    (64, 5, 128, 128),    # several tiles per frame
    (128, 2, 32, 32),
    (128, 3, 64, 16),
    (256, 3, 32, 32),     # two output-channel passes
    (256, 2, 16, 48),
    (64, 1, 2, 2),        # a single half-resolution position
]


def _run_upcat(lib, handle, skip, below, w3, b3, wt, bt, f):
    from openglottal_b200 import _native

    n, _, hgt, wid = skip.shape
    out = torch.full((n, f, hgt, wid), float("nan"), device="cuda")
    host = [t.contiguous().cpu() for t in (w3, b3, wt, bt)]
    fp = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_float))
    _native.check(lib.ogl_debug_upcat_layer(handle, skip.data_ptr(), below.data_ptr(), fp(host[0]),
                                            fp(host[1]), fp(host[2]), fp(host[3]), f, n, hgt, wid,
                                            out.data_ptr(), None))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("f,n,hgt,wid", UPCAT_CASES)
def test_upcat_composed_convT_cat_conv(lib, handle, f, n, hgt, wid):
    """ConvTranspose2d + cat + conv3x3 + bias + ReLU as ONE launch (unet.py:82-87) against torch in
    fp64 on the same bf16-rounded activations; one CTA per tile and CTA pairs give the same bits."""
    from openglottal_b200 import _native

    g = torch.Generator().manual_seed(f * 3 + hgt + wid)
    skip = _bf(torch.randn(n, f, hgt, wid, generator=g))
    below = _bf(torch.randn(n, 2 * f, hgt // 2, wid // 2, generator=g))
    w3 = torch.randn(f, 2 * f, 3, 3, generator=g) * (2.0 / (2 * f * 9)) ** 0.5
    b3 = torch.randn(f, generator=g) * 0.1
    wt = torch.randn(2 * f, f, 2, 2, generator=g) * (1.0 / (2 * f)) ** 0.5
    bt = torch.randn(f, generator=g) * 0.5
    up = F.conv_transpose2d(below.double(), wt.double(), bt.double(), stride=2)
    ref = F.relu(F.conv2d(torch.cat([skip.double(), up], 1), w3.double(), b3.double(), padding=1)).float()
    outs = []
    try:
        for mode in (1, 3):
            _native.check(lib.ogl_unet_set_cta_pairs(handle, mode))
            outs.append(_run_upcat(lib, handle, skip.cuda(), below.cuda(), w3, b3, wt, bt, f))
    finally:
        _native.check(lib.ogl_unet_set_cta_pairs(handle, 2))
    # weights are rounded to bf16 AFTER the composition: |err| <= 2^-7 |ref| + a few weight ulps
    err = (outs[0].cpu() - ref).abs()
    tol = ref.abs() * 2.0 ** -7 + 2e-2
    bad = err > tol
    print(f"upcat f={f} {n}x{hgt}x{wid}: max|err|={err.max().item():.4g} mean={err.mean().item():.4g} "
          f"nan={torch.isnan(outs[0]).sum().item()} bad={bad.sum().item()}")
    assert not torch.isnan(outs[0]).any() and not bad.any()
    assert err.mean().item() <= 3e-3
    assert torch.equal(outs[0], outs[1])
