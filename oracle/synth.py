"""Synthetically trained state dicts (TEST INFRASTRUCTURE, see oracle/__init__.py).

``trained_state`` follows the reference recipe (/root/reference/scripts/train_unet.py:155-181:
AdamW lr 1e-3, 0.5*BCE + 0.5*Dice) for ~150 steps on a synthetic clip, giving trained-like
weights for the bf16 tolerance tests. The seeded clip / random-state generators
live in ``synthdata`` and are re-exported here.
"""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

from synthdata import _empty_state, calibrated_state, glottis_clip  # noqa: F401

from .unet_oracle import BN_EPS


def _train_forward(params: dict, buffers: dict, x: torch.Tensor) -> torch.Tensor:
    """Training-mode functional forward (batch statistics, running stats updated in place)."""
    def dc(prefix, x):
        for ci, bi in ((0, 1), (3, 4)):
            x = F.conv2d(x, params[f"{prefix}.net.{ci}.weight"], None, padding=1)
            x = F.batch_norm(x, buffers[f"{prefix}.net.{bi}.running_mean"],
                             buffers[f"{prefix}.net.{bi}.running_var"],
                             params[f"{prefix}.net.{bi}.weight"], params[f"{prefix}.net.{bi}.bias"],
                             training=True, momentum=0.1, eps=BN_EPS)
            x = F.relu(x)
        return x

    skips = []
    for i in range(4):
        x = dc(f"downs.{i}", x)
        skips.append(x)
        x = F.max_pool2d(x, 2, 2)
    x = dc("bottleneck", x)
    for k in range(4):
        x = F.conv_transpose2d(x, params[f"ups.{2 * k}.weight"], params[f"ups.{2 * k}.bias"], stride=2)
        x = torch.cat([skips[3 - k], x], dim=1)
        x = dc(f"ups.{2 * k + 1}", x)
    return F.conv2d(x, params["head.weight"], params["head.bias"])


def trained_state(seed: int = 0, steps: int = 150, size: int = 128, batch: int = 8,
                  cache_dir: str | os.PathLike | None = None, device: str | None = None) -> dict:
    """Synthetically *trained* state dict. Cached as a plain
    ``torch.save(state_dict)`` file -- the format of /root/reference/scripts/train_unet.py:207."""
    path = None
    if cache_dir is not None:
        path = Path(cache_dir) / f"trained_seed{seed}_s{steps}_r{size}.pt"
        if path.exists():
            return torch.load(path, map_location="cpu", weights_only=True)
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    torch.manual_seed(seed)
    sd = _empty_state()
    g = torch.Generator().manual_seed(1000 + seed)
    for k, v in sd.items():     # PyTorch-default-like init: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
        if v.dim() == 4:
            fan_in = (v.shape[1] if ".net." in k or k.startswith("head") else v.shape[1]) * v.shape[2] * v.shape[3]
            bound = (1.0 / fan_in) ** 0.5
            sd[k] = (torch.rand(v.shape, generator=g) * 2 - 1) * bound
        elif k in ("head.bias",) or (k.startswith("ups") and k.endswith(".bias") and ".net." not in k):
            sd[k] = (torch.rand(v.shape, generator=g) * 2 - 1) * 0.05
    params = {k: v.to(device).requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and "running_" not in k}
    buffers = {k: v.to(device) for k, v in sd.items() if "running_" in k}
    opt = torch.optim.AdamW(list(params.values()), lr=1e-3)
    frames, masks = glottis_clip(max(64, 2 * batch), 256, 256, seed=500 + seed, period=16.0, jitter=6.0)
    rng = np.random.default_rng(seed)
    for _ in range(steps):
        idx = rng.integers(0, len(frames), batch)
        y0 = rng.integers(0, 256 - size + 1, batch) if size < 256 else np.zeros(batch, int)
        x0 = rng.integers(0, 256 - size + 1, batch) if size < 256 else np.zeros(batch, int)
        # bias the crops towards the glottis so positives are present
        y0 = np.clip((y0 + (128 - size // 2)) // 2, 0, 256 - size)
        x0 = np.clip((x0 + (128 - size // 2)) // 2, 0, 256 - size)
        xb = np.stack([frames[i, y:y + size, x:x + size] for i, y, x in zip(idx, y0, x0)])
        yb = np.stack([masks[i, y:y + size, x:x + size] for i, y, x in zip(idx, y0, x0)])
        xt = torch.from_numpy(xb.astype("float32") / 255.0).unsqueeze(1).to(device)
        yt = torch.from_numpy((yb > 0).astype("float32")).unsqueeze(1).to(device)
        logits = _train_forward(params, buffers, xt)
        p = torch.sigmoid(logits)
        dice_l = 1 - (2 * (p * yt).sum() + 1e-6) / (p.sum() + yt.sum() + 1e-6)
        loss = 0.5 * F.binary_cross_entropy_with_logits(logits, yt) + 0.5 * dice_l
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    out = {}
    for k in sd:
        if k in params:
            out[k] = params[k].detach().cpu().contiguous()
        elif k in buffers:
            out[k] = buffers[k].detach().cpu().contiguous()
        else:
            out[k] = torch.tensor(steps, dtype=torch.long)
    if path is not None:
        path.parent.mkdir(parents=True, exist_ok=True)
        tmp = path.with_suffix(".tmp")
        torch.save(out, tmp)
        os.replace(tmp, path)
    return out
