"""CPU check of the operands build_upcat_host packs for the composed decoder layer of levels 1-3
(openglottal_b200/csrc/upcat_tc.cu): a NumPy emulation of the kernel's MMA program -- the same
(source phase, halo cell) arithmetic per tap, the same (phase, offset) pairs of the composed
ConvTranspose2d, the same bias classes -- on the packed bf16 blobs must reproduce
relu(conv3x3(cat([skip, conv_transpose2d(below)])) + b) of /root/reference/openglottal/models/unet.py:82-87
(torch, fp64). Runs without a GPU; the kernel itself is tested in tests/test_gpu_layers.py."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _bf16_to_f32(u16: np.ndarray) -> np.ndarray:
    return (u16.astype(np.uint32) << 16).view(np.float32)


def _program(lib, w3, b3, wt, bt, f):
    from openglottal_b200 import _native

    N = min(f, 128)
    ws = np.zeros(f * f * 9, np.uint16)
    wsp = np.zeros_like(ws)
    wb = np.zeros(32 * f * f, np.uint16)
    wbp = np.zeros_like(wb)
    bt9 = np.zeros(9 * f, np.float32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    fp = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_float))
    _native.check(lib.ogl_debug_upcat_program(fp(w3), fp(b3), fp(wt), fp(bt), f, ptr(ws), ptr(wsp), ptr(wb),
                                              ptr(wbp), ptr(bt9)))
    npass = f // N
    ws = _bf16_to_f32(ws).reshape(npass, f // 32, 9, 4, N, 8)
    wb = _bf16_to_f32(wb).reshape(npass, 2 * f // 32, 16, 4, N, 8)
    wsp = _bf16_to_f32(wsp).reshape(npass, f // 32, 2, 9, 4, N // 2, 8)
    wbp = _bf16_to_f32(wbp).reshape(npass, 2 * f // 32, 2, 16, 4, N // 2, 8)
    # the CTA-pair forms hold the same numbers, rank r = columns [r N/2, (r+1) N/2)
    assert np.array_equal(np.concatenate([wsp[:, :, 0], wsp[:, :, 1]], axis=4), ws)
    assert np.array_equal(np.concatenate([wbp[:, :, 0], wbp[:, :, 1]], axis=4), wb)
    # -> [tap][ci][co]
    Ws = ws.transpose(2, 1, 3, 5, 0, 4).reshape(9, f, f)          # tap, (kb, c4, e), (pass, n)
    Wb = wb.transpose(2, 1, 3, 5, 0, 4).reshape(16, 2 * f, f)
    return Ws, Wb, bt9.reshape(3, 3, f)


def _emulate(Ws, Wb, btab, skip, below, f):
    """skip [f][H][W], below [2f][H/2][W/2] -> out [f][H][W], following issue_half / the epilogue."""
    _, H, W = skip.shape
    H2, W2 = H // 2, W // 2
    src_phase = lambda p, d: (p + d - 1) & 1
    src_cell = lambda p, d: (p + d + 1) >> 1
    s2d = np.zeros((2, 2, f, H2 + 2, W2 + 2), np.float64)          # [qy][qx][c][halo rows][halo cols]
    for qy in range(2):
        for qx in range(2):
            s2d[qy, qx, :, 1:-1, 1:-1] = skip[:, qy::2, qx::2]
    bel = np.zeros((2 * f, H2 + 2, W2 + 2), np.float64)
    bel[:, 1:-1, 1:-1] = below
    out = np.zeros((f, H, W), np.float64)
    for py in range(2):
        for px in range(2):
            acc = np.zeros((f, H2, W2), np.float64)
            for dy in range(3):
                for dx in range(3):
                    a = s2d[src_phase(py, dy), src_phase(px, dx)][
                        :, src_cell(py, dy):src_cell(py, dy) + H2, src_cell(px, dx):src_cell(px, dx) + W2]
                    acc += np.einsum("chw,co->ohw", a, Ws[dy * 3 + dx].astype(np.float64))
            for o4 in range(4):
                cy, cx = py + (o4 >> 1), px + (o4 & 1)
                a = bel[:, cy:cy + H2, cx:cx + W2]
                acc += np.einsum("chw,co->ohw", a, Wb[(px * 2 + py) * 4 + o4].astype(np.float64))
            ys = 2 * np.arange(H2) + py
            xs = 2 * np.arange(W2) + px
            ry = np.where(ys == 0, 0, np.where(ys == H - 1, 2, 1))
            rx = np.where(xs == 0, 0, np.where(xs == W - 1, 2, 1))
            bias = btab[ry[:, None], rx[None, :]].transpose(2, 0, 1)   # [f][H2][W2]
            out[:, py::2, px::2] = np.maximum(acc + bias, 0.0)
    return out


@pytest.mark.parametrize("f,hgt,wid", [(64, 8, 12), (128, 6, 4), (256, 4, 4)])
def test_upcat_program_reproduces_convT_cat_conv(lib, f, hgt, wid):
    g = torch.Generator().manual_seed(f + hgt)
    skip = _bf(torch.randn(1, f, hgt, wid, generator=g))
    below = _bf(torch.randn(1, 2 * f, hgt // 2, wid // 2, generator=g))
    w3 = torch.randn(f, 2 * f, 3, 3, generator=g) * (2.0 / (2 * f * 9)) ** 0.5
    b3 = torch.randn(f, generator=g) * 0.1
    wt = torch.randn(2 * f, f, 2, 2, generator=g) * (1.0 / (2 * f)) ** 0.5
    bt = torch.randn(f, generator=g) * 0.5           # large: a wrong bias class shows
    up = F.conv_transpose2d(below.double(), wt.double(), bt.double(), stride=2)
    ref = F.relu(F.conv2d(torch.cat([skip.double(), up], 1), w3.double(), b3.double(), padding=1))[0].numpy()
    Ws, Wb, btab = _program(lib, w3.contiguous(), b3, wt.contiguous(), bt, f)
    got = _emulate(Ws, Wb, btab, skip[0].double().numpy(), below[0].double().numpy(), f)
    err = np.abs(got - ref)
    print(f"f={f}: max|err| {err.max():.4g}, ref rms {np.sqrt((ref ** 2).mean()):.4g}")
    # the only difference left is the rounding of the packed weights to bf16
    assert err.max() <= 2e-2 and err.mean() <= 2e-3
