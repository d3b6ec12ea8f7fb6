"""Frame-range sharding of a clip over the GPUs of one box (one process per GPU).

Frames are independent through the U-Net, threshold and popcount
(/root/reference/openglottal/features.py:234-238 has no cross-frame state in unet-only), so the
data path needs no collective; the only exchange is an all-gather of the per-shard int32 area
vector (4 B/frame) before ``_kinematic_features`` (features.py:247).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def rank_world(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_size(n: int, world: int) -> int:
    """Frames per rank: ceil(n / world) (the last ranks may get fewer or none)."""
    return (n + world - 1) // world


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous frame range [lo, hi) of ``rank``: rank r owns frames
    [r*ceil(n/R), min(n, (r+1)*ceil(n/R)))."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    per = shard_size(n, world)
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


def agree_any(flag: bool, device, group=None) -> bool:
    """True on every rank when ``flag`` is true on ANY rank (one 4-byte all-reduce; no collective
    with a single rank). Ranks that must take the same branch before a collective -- e.g. "my part
    of the file could not be read by range, decode sequentially" -- agree through this."""
    _, world = rank_world(group)
    if world == 1:
        return bool(flag)
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return bool(t.item())


def gather_area(local: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """All-gather the per-rank int32 area shards into the full ``(n,)`` waveform, in frame
    order, on every rank. Shards are padded to ceil(n/R) so one ``all_gather_into_tensor``
    (NCCL over NVLink on the GPU box, gloo in CPU tests) suffices."""
    rank, world = rank_world(group)
    if world == 1:
        if local.numel() != n:
            raise ValueError("single-rank gather: shard does not cover the clip")
        return local
    per = shard_size(n, world)
    lo, hi = shard_range(n, rank, world)
    if local.numel() != hi - lo:
        raise ValueError(f"rank {rank}: shard has {local.numel()} frames, expected {hi - lo}")
    padded = torch.zeros(per, dtype=torch.int32, device=local.device)
    padded[: hi - lo] = local
    full = torch.empty(per * world, dtype=torch.int32, device=local.device)
    dist.all_gather_into_tensor(full, padded, group=group)
    if per * world == n:
        return full
    # drop the padding of each shard (only trailing shards are short)
    pieces = []
    for r in range(world):
        rlo, rhi = shard_range(n, r, world)
        pieces.append(full[r * per: r * per + (rhi - rlo)])
    return torch.cat(pieces)
