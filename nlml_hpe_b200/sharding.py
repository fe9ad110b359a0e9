"""Sample-index sharding of a feature batch across the GPUs of one box.

Both halves of the hot path treat samples independently (no cross-sample term in
TD_Tester.py:110-125 or NLML_HPE_Model_Builder.py:115-126), so the batch is cut into contiguous
slices, one per rank (one process per GPU), constants are replicated, and nothing is exchanged on
the compute path.  The only communication is the optional end-of-run gather of the [N,k] results.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n, world_size, rank):
    """Contiguous [lo, hi) slice of n samples for `rank`; sizes differ by at most one."""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local, n_total, group=None):
    """All-gather ragged row shards back into one [n_total, k] tensor on every rank (end of run only)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], 0)


def run_sharded(fn, X_full_rows, n_total, group=None):
    """Apply fn to this rank's slice (X_full_rows(lo, hi) -> tensor) and gather the results."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_bounds(n_total, world, rank)
    return gather_rows(fn(X_full_rows(lo, hi)), n_total, group)
