"""Batch sharding of the warp path across ranks (one process per GPU).

Every output element of every op depends only on its own sample b
(block_extractor_kernel.cu:52,62-63; resample2d_kernel.cu:42,47-49), so the
path shards over batch with NO data-path collective: rank r owns a contiguous
slice of samples.  The only collective a training job needs is the all-reduce
of the *model's* parameter gradients (the ops have no parameters), which is
DDP's business, not this library's.
"""
from __future__ import annotations


def shard_bounds(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of `n_items` owned by `rank`: contiguous, sizes differ by at most 1,
    earlier ranks take the remainder."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, world_size: int, rank: int):
    """Slice dim 0 of each tensor to this rank's share."""
    lo, hi = shard_bounds(tensors[0].shape[0], world_size, rank)
    return [t[lo:hi] for t in tensors]


def reduce_max_time(ms: float, device=None) -> float:
    """max over ranks of a per-rank device time (ms); identity when not distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
