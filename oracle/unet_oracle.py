"""CPU oracle of UNet.forward and the per-frame wrapper (TEST INFRASTRUCTURE, see __init__).

The reference's arithmetic for this path lives in PyTorch (Conv2d / BatchNorm2d / ReLU /
MaxPool2d / ConvTranspose2d / cat / sigmoid), OpenCV (resize) and NumPy; the oracle restates
the reference's *composition* of those ops as plain functions of a state dict:

  ref_forward        <- /root/reference/openglottal/models/unet.py:18-33 (DoubleConv), :74-88
  fold_state         <- eval-mode BatchNorm folded into the conv
  bitmodel_forward   <- the same network with the kernels' bf16 rounding points, fp32 accumulate
  segment_frame      <- /root/reference/openglottal/utils.py:218-241
  area_wave          <- /root/reference/openglottal/features.py:234-238 (detector is None)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

FEATURES = (32, 64, 128, 256)
BN_EPS = 1e-5


def _double_conv(sd: dict, prefix: str, x: torch.Tensor) -> torch.Tensor:
    # unet.py:24-29: Conv3x3(no bias) -> BN(eval) -> ReLU, twice
    for conv_i, bn_i in ((0, 1), (3, 4)):
        x = F.conv2d(x, sd[f"{prefix}.net.{conv_i}.weight"], None, padding=1)
        x = F.batch_norm(
            x,
            sd[f"{prefix}.net.{bn_i}.running_mean"],
            sd[f"{prefix}.net.{bn_i}.running_var"],
            sd[f"{prefix}.net.{bn_i}.weight"],
            sd[f"{prefix}.net.{bn_i}.bias"],
            training=False,
            eps=BN_EPS,
        )
        x = F.relu(x)
    return x


@torch.no_grad()
def ref_forward(sd: dict, x: torch.Tensor) -> torch.Tensor:
    """fp32 logits (N,1,H,W) for x (N,1,H,W) fp32 -- unet.py:74-88 (H, W multiples of 16)."""
    skips = []
    for i in range(4):                                   # unet.py:76-79
        x = _double_conv(sd, f"downs.{i}", x)
        skips.append(x)
        x = F.max_pool2d(x, 2, 2)
    x = _double_conv(sd, "bottleneck", x)                # unet.py:80
    for k in range(4):                                   # unet.py:81-87
        x = F.conv_transpose2d(x, sd[f"ups.{2 * k}.weight"], sd[f"ups.{2 * k}.bias"], stride=2)
        x = torch.cat([skips[3 - k], x], dim=1)          # skip channels first (unet.py:86)
        x = _double_conv(sd, f"ups.{2 * k + 1}", x)
    return F.conv2d(x, sd["head.weight"], sd["head.bias"])  # unet.py:88


def fold_state(sd: dict) -> dict:
    """BN folded in fp64: W' = W*g/sqrt(v+eps), b' = beta - mu*g/sqrt(v+eps); fp32 results."""
    out = {}
    prefixes = [f"downs.{i}" for i in range(4)] + ["bottleneck"] + [f"ups.{2 * k + 1}" for k in range(4)]
    for p in prefixes:
        for conv_i, bn_i in ((0, 1), (3, 4)):
            w = sd[f"{p}.net.{conv_i}.weight"].double()
            s = sd[f"{p}.net.{bn_i}.weight"].double() / torch.sqrt(
                sd[f"{p}.net.{bn_i}.running_var"].double() + BN_EPS)
            out[f"{p}.{conv_i}.w"] = (w * s[:, None, None, None]).float()
            out[f"{p}.{conv_i}.b"] = (sd[f"{p}.net.{bn_i}.bias"].double()
                                      - sd[f"{p}.net.{bn_i}.running_mean"].double() * s).float()
    for k in range(4):
        out[f"ups.{2 * k}.w"] = sd[f"ups.{2 * k}.weight"].float()
        out[f"ups.{2 * k}.b"] = sd[f"ups.{2 * k}.bias"].float()
    out["head.w"] = sd["head.weight"].float()
    out["head.b"] = sd["head.bias"].float()
    return out


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def _f16(t: torch.Tensor) -> torch.Tensor:
    return t.clamp(-65504.0, 65504.0).to(torch.float16).to(torch.float32)   # cvt.rn.satfinite


@torch.no_grad()
def folded_forward(sd: dict, x: torch.Tensor, bf16: bool, composed_level0: bool = True,
                   operand: str = "bf16", composed_up: bool = True) -> torch.Tensor:
    """Network on BN-folded weights. ``bf16=False``: the fp32 validation mode's arithmetic.
    ``bf16=True``: bit-model of the tensor-core path -- fp32 stem from the fp32 input, bf16
    weights, every stored activation rounded to bf16 (after bias+ReLU; convT after bias),
    fp32 accumulation, and the last conv's ReLU output fed to the 1x1 head unrounded.
    ``composed_level0`` (the default "s2d" schedule): the last ConvTranspose2d is composed into
    the conv that follows it, so its output is never rounded (the composed weights are; that
    difference is inside the test tolerance and not modelled).
    ``composed_up`` (the default decoder): the same at levels 1-3 (``compose_up`` of the model).
    ``operand="fp16"`` (with ``bf16=True``): the same rounding points with f16 operands -- the
    kernels' f16 precision mode."""
    fs = fold_state(sd)
    r = (_f16 if operand == "fp16" else _bf16) if bf16 else (lambda t: t)

    def conv(x, key, relu=True, round_out=True, round_w=True):
        w = r(fs[f"{key}.w"]) if round_w else fs[f"{key}.w"]
        y = F.conv2d(x, w, fs[f"{key}.b"], padding=1)
        if relu:
            y = F.relu(y)
        return r(y) if round_out else y

    skips = []
    for i in range(4):
        x = conv(x, f"downs.{i}.0", round_w=(i != 0))     # stem stays fp32
        x = conv(x, f"downs.{i}.3")
        skips.append(x)
        x = F.max_pool2d(x, 2, 2)
    x = conv(x, "bottleneck.0")
    x = conv(x, "bottleneck.3")
    for k in range(4):
        up = F.conv_transpose2d(x, r(fs[f"ups.{2 * k}.w"]), fs[f"ups.{2 * k}.b"], stride=2)
        x = up if ((k == 3 and composed_level0) or (k < 3 and composed_up)) else r(up)
        x = torch.cat([skips[3 - k], x], dim=1)
        x = conv(x, f"ups.{2 * k + 1}.0")
        x = conv(x, f"ups.{2 * k + 1}.3", round_out=(k != 3))
    return F.conv2d(x, fs["head.w"], fs["head.b"])


def frames_to_input(frames_u8: np.ndarray) -> torch.Tensor:
    """(N,H,W) u8 -> (N,1,H,W) fp32 exactly as utils.py:235 (float32 / 255.0)."""
    return torch.from_numpy(frames_u8.astype("float32") / 255.0).unsqueeze(1)


@torch.no_grad()
def segment_frame(sd: dict, frame_gray: np.ndarray, threshold: float = 0.5,
                  forward=ref_forward) -> np.ndarray:
    """utils.py:218-241: resize to 256x256, forward, sigmoid, resize prob back, threshold."""
    import cv2

    inp = cv2.resize(frame_gray, (256, 256), interpolation=cv2.INTER_LINEAR)
    t = torch.from_numpy(inp.astype("float32") / 255.0).unsqueeze(0).unsqueeze(0)
    prob = torch.sigmoid(forward(sd, t)).squeeze().numpy()
    hgt, wid = frame_gray.shape
    if (hgt, wid) != (256, 256):
        prob = cv2.resize(prob, (wid, hgt), interpolation=cv2.INTER_LINEAR)
    return (prob > threshold).astype(np.uint8) * 255


def area_wave(sd: dict, frames_gray, forward=ref_forward) -> list[float]:
    """features.py:234-238 with detector None: per-frame count of mask pixels."""
    return [float(np.sum(segment_frame(sd, f, forward=forward) > 0)) for f in frames_gray]


@torch.no_grad()
def batch_masks(sd: dict, frames_u8: np.ndarray, forward=ref_forward, chunk: int = 16):
    """Batched form for frames whose size the network takes natively (no resize):
    returns (logits f32 (N,H,W), masks u8 {0,255}, areas int64)."""
    outs = []
    for i in range(0, len(frames_u8), chunk):
        outs.append(forward(sd, frames_to_input(frames_u8[i:i + chunk]))[:, 0])
    logits = torch.cat(outs).numpy()
    prob = torch.sigmoid(torch.from_numpy(logits)).numpy()
    masks = (prob > 0.5).astype(np.uint8) * 255
    return logits, masks, (masks > 0).reshape(len(masks), -1).sum(1)
