"""Sharding of independent work across the GPUs of one box.

The path has no exchange step: camera views and LOS rays are independent, the grid / atlas /
segment list are replicated (a few MB), and the only collective is the gather of COMPACT
results (per-view checksum + convergence, 1 byte per ray).  These helpers are pure index
arithmetic plus a torch.distributed gather, so they run under gloo on CPU as well as nccl.
"""
from __future__ import annotations

import numpy as np


def view_shard(n_views: int, rank: int, world: int) -> np.ndarray:
    """Round-robin view indices of `rank` (views i % world == rank), SURVEY §8d C4."""
    return np.arange(rank, n_views, world, dtype=np.int64)


def ray_shard(n_rays: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) ray range of `rank`; sizes differ by at most one."""
    base, rem = divmod(n_rays, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_views(local: "torch.Tensor", n_views: int, rank: int, world: int):
    """all_gather per-view results sharded round-robin back into view order.
    local: [ceil-ish(n_views/world), ...] tensor holding this rank's views in shard order."""
    import torch
    import torch.distributed as dist

    per = (n_views + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = torch.zeros((n_views,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = torch.arange(r, n_views, world, device=local.device)
        out[idx] = parts[r][: idx.numel()]
    return out


def gather_rays(local: "torch.Tensor", n_rays: int, rank: int, world: int):
    """all_gather contiguous ray shards (sizes differ by <= 1) back into one [n_rays] tensor."""
    import torch
    import torch.distributed as dist

    per = (n_rays + world - 1) // world
    pad = torch.zeros(per, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = torch.empty(n_rays, dtype=local.dtype, device=local.device)
    for r in range(world):
        lo, hi = ray_shard(n_rays, r, world)
        out[lo:hi] = parts[r][: hi - lo]
    return out
