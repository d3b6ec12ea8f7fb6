"""Seeded synthetic inputs shared by tests/, bench.py and oracle/ (no reference arithmetic here).

The reference's weights and datasets are not in the snapshot, and a freshly initialised UNet is
degenerate, so tests and the bench use:

* ``glottis_clip``      -- seeded synthetic HSV clip with a periodically opening dark ellipse and
                           its ground-truth mask (known f0 = 1/period);
* ``calibrated_state``  -- seeded, variance-calibrated random state dict: cheap, bit-reproducible,
                           non-degenerate; used to pin the oracle against the reference and for
                           kernel-vs-bit-model tests (chaotic under bf16, so NOT for Dice bars);
* ``oracle.synth.trained_state`` (test infrastructure) trains on these clips to obtain
  trained-like weights for the bf16 tolerance tests.
"""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np
import torch

FEATURES = (32, 64, 128, 256)


def glottis_clip(n: int, hgt: int = 256, wid: int = 256, seed: int = 0, period: float = 20.0,
                 jitter: float = 0.0):
    """(frames u8 (n,H,W), masks u8 {0,255} (n,H,W)). Bright textured background, dark ellipse
    whose half-width follows max(0, sin(2*pi*t/period)); pixel noise; all seeded."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:hgt, 0:wid].astype(np.float32)
    bg = 150.0 + 20.0 * np.sin(xx / 17.0) + 15.0 * np.cos(yy / 23.0)
    frames = np.empty((n, hgt, wid), np.uint8)
    masks = np.empty((n, hgt, wid), np.uint8)
    half_h = 0.17 * hgt
    for t in range(n):
        a = 0.012 * wid + 0.07 * wid * max(0.0, np.sin(2 * np.pi * t / period))
        cx = wid / 2 + (rng.normal(0, jitter) if jitter else 0.0)
        cy = hgt / 2 + (rng.normal(0, jitter) if jitter else 0.0)
        inside = ((xx - cx) / a) ** 2 + ((yy - cy) / half_h) ** 2 <= 1.0
        img = bg + rng.normal(0, 6.0, (hgt, wid)).astype(np.float32)
        dark = 20.0 + rng.normal(0, 4.0, (hgt, wid)).astype(np.float32)
        img = np.where(inside, dark, img)
        frames[t] = np.clip(img, 0, 255).astype(np.uint8)
        masks[t] = inside.astype(np.uint8) * 255
    return frames, masks


def _empty_state() -> dict:
    sd = {}

    def block(prefix, cin, cout):
        for ci, bi, c_in in ((0, 1, cin), (3, 4, cout)):
            sd[f"{prefix}.net.{ci}.weight"] = torch.zeros(cout, c_in, 3, 3)
            sd[f"{prefix}.net.{bi}.weight"] = torch.ones(cout)
            sd[f"{prefix}.net.{bi}.bias"] = torch.zeros(cout)
            sd[f"{prefix}.net.{bi}.running_mean"] = torch.zeros(cout)
            sd[f"{prefix}.net.{bi}.running_var"] = torch.ones(cout)
            sd[f"{prefix}.net.{bi}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    ch = 1
    for i, f in enumerate(FEATURES):
        block(f"downs.{i}", ch, f)
        ch = f
    for k, f in enumerate(reversed(FEATURES)):
        sd[f"ups.{2 * k}.weight"] = torch.zeros(2 * f, f, 2, 2)
        sd[f"ups.{2 * k}.bias"] = torch.zeros(f)
        block(f"ups.{2 * k + 1}", 2 * f, f)
    block("bottleneck", 256, 512)
    sd["head.weight"] = torch.zeros(1, 32, 1, 1)
    sd["head.bias"] = torch.zeros(1)
    return sd


def calibrated_state(seed: int = 0) -> dict:
    """Seeded random state dict with O(1) activations at every depth: He-normal convs,
    BN gamma ~ U(0.5, 1.5), beta ~ N(0, 0.2), running_mean ~ N(0, 0.1), running_var ~ U(0.5, 1.5).
    Uses only a CPU torch.Generator, so it is identical on every machine with this torch."""
    g = torch.Generator().manual_seed(seed)
    sd = _empty_state()
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(100, dtype=torch.long)
        elif k.startswith("head"):
            sd[k] = torch.randn(v.shape, generator=g) * (0.5 if k.endswith("weight") else 0.1)
        elif v.dim() == 4 and k.startswith("ups") and ".net." not in k:   # convT (cin,cout,2,2)
            sd[k] = torch.randn(v.shape, generator=g) * (1.0 / v.shape[0]) ** 0.5
        elif v.dim() == 4:                                                # conv3x3
            fan_in = v.shape[1] * 9
            sd[k] = torch.randn(v.shape, generator=g) * (2.0 / fan_in) ** 0.5
        elif k.endswith("running_var"):
            sd[k] = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith("running_mean"):
            sd[k] = torch.randn(v.shape, generator=g) * 0.1
        elif k.endswith(".weight"):                                       # BN gamma
            sd[k] = torch.rand(v.shape, generator=g) + 0.5
        else:                                                             # BN beta / convT bias
            sd[k] = torch.randn(v.shape, generator=g) * 0.2
    return sd


