"""Batch sharding of the alignment path across GPUs.

Utterances are independent (reference core.pyx:41-42 loops over the batch with no cross-talk), so the
path shards by utterance with NO data-path collective: under DDP every rank simply aligns its own
per-replica batch (reference train.py:99-106).  The helpers here are for the standalone benchmark and
for tests: contiguous shard bounds, and an all-gather of the compact per-frame index (int32 [B, T_y],
1/T_x of the dense path) that is used only to VERIFY results, never inside a timed region.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of `n` utterances owned by `rank`; the first n % world ranks get one more."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_index(index: torch.Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather per-rank index shards ([n_rank, T_y] int32, shards from `shard_bounds`) into [n_total, T_y].
    Shards may differ by one row: they are padded to the largest shard for the collective."""
    world = dist.get_world_size(group)
    T_y = index.shape[1]
    per = (n_total + world - 1) // world
    pad = torch.full((per, T_y), -2, dtype=torch.int32, device=index.device)
    pad[: index.shape[0]] = index
    out = torch.empty((world * per, T_y), dtype=torch.int32, device=index.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    rows = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        rows.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(rows, 0)


def check_index(index, t_ys, t_xs) -> bool:
    """Size-independent invariants of a monotonic alignment in index form: starts at 0, ends at t_x-1,
    non-decreasing with steps in {0,1}, -1 on padded frames."""
    idx = np.asarray(index.cpu() if isinstance(index, torch.Tensor) else index)
    for b in range(idx.shape[0]):
        ty, tx = int(t_ys[b]), int(t_xs[b])
        row = idx[b, :ty]
        if ty < 1 or row[0] != 0 or row[-1] != tx - 1:
            return False
        d = np.diff(row)
        if not ((d == 0) | (d == 1)).all() or not (idx[b, ty:] == -1).all():
            return False
    return True


def maximum_path_sharded(neg_cent: torch.Tensor, mask: torch.Tensor, group: Optional[dist.ProcessGroup] = None):
    """Align this rank's contiguous shard of a replicated batch; returns (path_shard, (lo, hi))."""
    from .monotonic_align import maximum_path
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(neg_cent.shape[0], rank, world)
    return maximum_path(neg_cent[lo:hi].contiguous(), mask[lo:hi]), (lo, hi)
