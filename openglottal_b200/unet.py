"""Drop-in for ``openglottal.UNet`` whose forward runs on hand-written sm_100a kernels.

Mirrors /root/reference/openglottal/models/unet.py:36-88: same constructor signature, the same
parameter/buffer tree (118 state-dict entries: ``downs.i.net.{0,1,3,4}.*``, ``ups.{0..7}.*``,
``bottleneck.net.*``, ``head.*``) so ``load_state_dict(torch.load(path, weights_only=True))``
(/root/reference/openglottal/cli.py:61-65) works unchanged. The forward pass is NOT PyTorch:
BatchNorm is folded into the convolutions at pack time and the whole network runs through
``libopenglottal_b200.so`` (tcgen05 implicit-GEMM convs, fused pool/head epilogues).
There is no CPU path and no training path.
"""
from __future__ import annotations

import ctypes as C
import os
import math

import torch
import torch.nn as nn

from . import _native

_FEATURES = (32, 64, 128, 256)


def _conv_bn_relu_x2(cin: int, cout: int) -> nn.Module:
    """Container with the reference's parameter names: ``net.0`` conv, ``net.1`` BN, ``net.3``
    conv, ``net.4`` BN (indices 2 and 5 are the parameter-free ReLUs)."""
    block = nn.Module()
    block.net = nn.Sequential(
        nn.Conv2d(cin, cout, 3, padding=1, bias=False),
        nn.BatchNorm2d(cout),
        nn.ReLU(inplace=True),
        nn.Conv2d(cout, cout, 3, padding=1, bias=False),
        nn.BatchNorm2d(cout),
        nn.ReLU(inplace=True),
    )
    return block


class UNet(nn.Module):
    """B200-native U-Net for binary glottal segmentation (1-channel in, 1-channel logits out).

    Parameters match ``openglottal.UNet``; only ``UNet(1, 1, (32, 64, 128, 256))`` -- the
    architecture every shipped checkpoint and the CLI use -- is implemented natively.
    """

    def __init__(self, in_ch: int = 1, out_ch: int = 1,
                 features: tuple[int, ...] = _FEATURES) -> None:
        super().__init__()
        if in_ch != 1 or out_ch != 1 or tuple(features) != _FEATURES:
            raise NotImplementedError(
                "openglottal_b200.UNet implements UNet(1, 1, (32, 64, 128, 256)) only "
                f"(got in_ch={in_ch}, out_ch={out_ch}, features={tuple(features)})")
        self.downs = nn.ModuleList()
        self.ups = nn.ModuleList()
        self.pool = nn.MaxPool2d(2, 2)  # parameter-free; kept so the module tree matches
        ch = in_ch
        for f in features:
            self.downs.append(_conv_bn_relu_x2(ch, f))
            ch = f
        self.bottleneck = _conv_bn_relu_x2(ch, 2 * ch)
        for f in reversed(features):
            self.ups.append(nn.ConvTranspose2d(2 * f, f, kernel_size=2, stride=2))
            self.ups.append(_conv_bn_relu_x2(2 * f, f))
        self.head = nn.Conv2d(features[0], out_ch, 1)

        self.precision = "bf16"      # "bf16" (tensor cores, the default), "fp16" (the same kernels
                                     # with f16 operands) or "fp32" (validation mode)
        self.max_batch = 512         # frames per native call; larger inputs are chunked
        self.schedule = "s2d"        # full-resolution level: "s2d" (space-to-depth GEMMs, the
                                     # default) or "direct" (per-tap form of the other levels)
        self.cta_pairs = int(os.environ.get("OGL_CG", "2"))   # 1: one CTA per tile; 2: CTA pairs
                                     # (tcgen05 cta_group::2) for the Cout >= 64 conv layers when
                                     # a launch has a tile per SM; 3: pairs whenever possible
        self.fuse_stem = int(os.environ.get("OGL_FUSE_STEM", "3"))   # stem inside downs.0.net.3:
                                     # 0 separate kernel, 1 in-kernel on the CUDA cores (fp32),
                                     # 2 in-kernel on the tensor cores (bf16 hi + lo weights, fp32
                                     # accumulation; 8 stem warps), 3 (default) the same GEMM with
                                     # 16 stem warps and an f16 im2col operand
        self.compose_up = os.environ.get("OGL_COMPOSE", "1") != "0"   # decoder levels 1-3: every
                                     # ConvTranspose2d composed into the conv after it (17 launches,
                                     # no `up` tensor); False: the round-1 schedule (20 launches)
        self.use_graphs = True       # small batches: replay the forward's launches as one CUDA graph
        self.graph_max_batch = 128   # (a forward is ~20 launches; at batch 32 they take ~1 ms of
                                     # GPU time, the same order as enqueueing them one by one)
        self.graph_cache_entries = 16   # captured call signatures kept (each owns static buffers)
        self._handle = None
        self._handle_device = None
        self._packed_sig = None
        self._sig_tensors = None
        self._f16_ready = False
        self._graphs = {}            # key -> [calls seen, graph, static input, static outputs]
        self._workspace = None
        self._keepalive = None

    # ------------------------------------------------------------------ native plumbing
    def _device(self) -> torch.device:
        return self.head.weight.device

    def _ensure_handle(self) -> C.c_void_p:
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError(
                "openglottal_b200.UNet runs on CUDA (B200) only; move the module with "
                ".to('cuda'). There is no CPU fallback -- use the reference for CPU.")
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        if self._handle is not None and self._handle_device == index:
            return self._handle
        self._release()
        lib = _native.load()
        h = C.c_void_p()
        _native.check(lib.ogl_unet_create(C.byref(h), index))
        self._handle, self._handle_device, self._packed_sig = h, index, None
        return h

    def _release(self) -> None:
        if self._handle is not None:
            _native.load().ogl_unet_destroy(self._handle)
        self._handle = None
        self._packed_sig = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self._release()
        except Exception:
            pass

    def _signature(self):
        # the tensor list is cached (building a state dict costs more than a small forward); it is
        # dropped whenever the module tree can have changed (_apply: .to() / .cuda() / .half();
        # load_state_dict copies in place and shows up in _version)
        ts = self._sig_tensors
        if ts is None:
            ts = self._sig_tensors = list(self.state_dict(keep_vars=True).values())
        return tuple((t.data_ptr(), t._version) for t in ts)

    def _apply(self, fn, *args, **kwargs):
        self._sig_tensors = None
        self._packed_sig = None
        self._graphs = {}
        return super()._apply(fn, *args, **kwargs)

    def _pack_if_needed(self) -> None:
        h = self._ensure_handle()
        sig = self._signature()
        if sig == self._packed_sig:
            return
        keep = []

        def host(t: torch.Tensor):
            a = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
            keep.append(a)
            return C.cast(a.data_ptr(), C.POINTER(C.c_float))

        def conv_bn(dst, seq, ci: int, bi: int) -> None:
            conv, bn = seq[ci], seq[bi]
            dst.weight = host(conv.weight)
            dst.bn_weight = host(bn.weight)
            dst.bn_bias = host(bn.bias)
            dst.running_mean = host(bn.running_mean)
            dst.running_var = host(bn.running_var)

        st = _native.UNetState()
        for i in range(4):
            conv_bn(st.downs[i][0], self.downs[i].net, 0, 1)
            conv_bn(st.downs[i][1], self.downs[i].net, 3, 4)
            st.up_t[i].weight = host(self.ups[2 * i].weight)
            st.up_t[i].bias = host(self.ups[2 * i].bias)
            conv_bn(st.up_c[i][0], self.ups[2 * i + 1].net, 0, 1)
            conv_bn(st.up_c[i][1], self.ups[2 * i + 1].net, 3, 4)
        conv_bn(st.bottleneck[0], self.bottleneck.net, 0, 1)
        conv_bn(st.bottleneck[1], self.bottleneck.net, 3, 4)
        st.head_weight = host(self.head.weight)
        st.head_bias = host(self.head.bias)
        st.bn_eps = float(self.downs[0].net[1].eps)
        _native.check(_native.load().ogl_unet_load_state(h, C.byref(st)))
        self._packed_sig = sig
        self._f16_ready = False
        self._graphs = {}

    def _precision_code(self) -> int:
        if self.precision == "bf16":
            return _native.PRECISION_BF16
        if self.precision == "fp32":
            return _native.PRECISION_F32
        if self.precision == "fp16":
            return _native.PRECISION_F16
        raise ValueError(f"precision must be 'bf16', 'fp16' or 'fp32', got {self.precision!r}")

    def _get_workspace(self, nbytes: int, device: torch.device) -> torch.Tensor:
        ws = self._workspace
        if ws is None or ws.numel() < nbytes or ws.device != device:
            self._workspace = None
            self._graphs = {}        # captured launches hold the old workspace's addresses
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._workspace = ws
        return ws

    # ------------------------------------------------------------------ public API
    def run(self, frames: torch.Tensor, threshold: float = 0.5, want_logits: bool = False,
            want_mask: bool = True, want_area: bool = True):
        """Batched hot path: gray frames -> (logits | None, mask | None, area | None).

        ``frames``: CUDA tensor ``(N, H, W)`` uint8 in 0..255 (scaled by 1/255 inside the
        kernel, /root/reference/openglottal/utils.py:235) or float32 already scaled.
        Returns f32 logits ``(N, H, W)``, u8 masks in {0, 255}
        (utils.py:241) and int32 per-frame areas (features.py:238).
        The module owns ONE activation workspace: drive it from one stream at a time (calls on
        the same stream queue up correctly; concurrent streams need one module each).
        """
        if self.training:
            raise RuntimeError("openglottal_b200.UNet is inference-only: call .eval() first "
                               "(training stays with the reference implementation)")
        if frames.dim() != 3:
            raise ValueError(f"expected (N, H, W) frames, got shape {tuple(frames.shape)}")
        if frames.dtype in (torch.bfloat16, torch.float16):
            frames = frames.float()      # exact widening; the stem computes in fp32 anyway
        if frames.dtype == torch.uint8:
            in_dtype = _native.DTYPE_U8
        elif frames.dtype == torch.float32:
            in_dtype = _native.DTYPE_F32
        else:
            raise TypeError(f"frames must be uint8, float32, bfloat16 or float16, got {frames.dtype}")
        dev = self._device()
        if frames.device != dev:
            raise RuntimeError(f"frames are on {frames.device} but the model is on {dev}")
        n, hgt, wid = frames.shape
        if n == 0:
            raise ValueError("no frames")
        if hgt % 16 or wid % 16:
            raise ValueError(f"H and W must be multiples of 16 (got {hgt}x{wid}); resize first")
        if not (0.0 < threshold < 1.0) or math.isnan(threshold):
            raise ValueError("threshold must be in (0, 1)")
        if self.schedule not in ("s2d", "direct"):
            raise ValueError(f"schedule must be 's2d' or 'direct', got {self.schedule!r}")
        self._pack_if_needed()
        lib = _native.load()
        _native.check(lib.ogl_unet_set_schedule(self._handle, 1 if self.schedule == "s2d" else 0))
        _native.check(lib.ogl_unet_set_cta_pairs(self._handle, int(self.cta_pairs)))
        _native.check(lib.ogl_unet_set_fused_stem(self._handle, int(self.fuse_stem)))
        _native.check(lib.ogl_unet_set_compose(self._handle, int(bool(self.compose_up))))
        frames = frames.contiguous()
        prec = self._precision_code()
        if prec == _native.PRECISION_F16 and not self._f16_ready:
            _native.check(lib.ogl_unet_prepare(self._handle, prec))
            self._f16_ready = True
        chunk = min(n, self.max_batch if prec != _native.PRECISION_F32 else min(self.max_batch, 16))
        nbytes = lib.ogl_unet_workspace_bytes(self._handle, chunk, hgt, wid, prec)
        ws = self._get_workspace(nbytes, dev)

        def enqueue(src, logits, mask, area, i0, m):
            _native.check(lib.ogl_unet_forward(
                self._handle, src[i0:i0 + m].data_ptr(), in_dtype, m, hgt, wid,
                ws.data_ptr(), ws.numel(),
                logits[i0:i0 + m].data_ptr() if want_logits else None,
                mask[i0:i0 + m].data_ptr() if want_mask else None,
                area[i0:i0 + m].data_ptr() if want_area else None,
                float(threshold), prec, torch.cuda.current_stream().cuda_stream))

        def outputs():
            return (torch.empty((n, hgt, wid), dtype=torch.float32, device=dev) if want_logits else None,
                    torch.empty((n, hgt, wid), dtype=torch.uint8, device=dev) if want_mask else None,
                    torch.empty((n,), dtype=torch.int32, device=dev) if want_area else None)

        with torch.cuda.device(dev):
            if (self.use_graphs and n <= min(self.graph_max_batch, chunk)
                    and prec != _native.PRECISION_F32 and not torch.cuda.is_current_stream_capturing()):
                # Small batch: the ~20 launches of a forward (each with its tensor maps built on the
                # host) are captured once per call signature and replayed; inputs and outputs of the
                # captured launches are static buffers, copied from / into fresh tensors.
                key = (n, hgt, wid, in_dtype, float(threshold), want_logits, want_mask, want_area, prec,
                       self.schedule, int(self.cta_pairs), int(self.fuse_stem), bool(self.compose_up),
                       ws.data_ptr())
                ent = self._graphs.get(key)
                if ent is None:
                    if len(self._graphs) >= self.graph_cache_entries:
                        # every entry owns static input / output buffers: a caller that walks many
                        # batch sizes must not accumulate them (oldest signature goes first)
                        self._graphs.pop(next(iter(self._graphs)))
                    ent = self._graphs[key] = [0, None, None, None]
                ent[0] += 1
                if ent[0] == 2:          # second call with this signature: capture
                    ent[2] = torch.empty_like(frames)
                    ent[3] = outputs()
                    ent[2].copy_(frames)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        enqueue(ent[2], *ent[3], 0, n)
                    ent[1] = graph
                if ent[1] is not None:
                    ent[2].copy_(frames)
                    ent[1].replay()
                    return tuple(None if t is None else t.clone() for t in ent[3])
            logits, mask, area = outputs()
            for i0 in range(0, n, chunk):
                enqueue(frames, logits, mask, area, i0, min(chunk, n - i0))
        return logits, mask, area

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """``(N, 1, H, W)`` float32 (or uint8) -> raw logits ``(N, 1, H, W)`` float32
        (/root/reference/openglottal/models/unet.py:74-88)."""
        if x.dim() != 4 or x.shape[1] != 1:
            raise ValueError(f"expected (N, 1, H, W) input, got shape {tuple(x.shape)}")
        logits, _, _ = self.run(x[:, 0], want_logits=True, want_mask=False, want_area=False)
        return logits.unsqueeze(1)
