"""Batch sharding across GPUs: one process per GPU, the batch dimension split contiguously, and ONE
exchange step -- an all-reduce (sum) of the small float64 / int64 statistics buffer between the
statistics pass and the closed forms / gradient pass.

The sums are additive across shards; every loss is a non-linear function of the GLOBAL sums, so the
result equals the single-device full-batch value (not a mean of per-shard losses).  Gradients are
w.r.t. each rank's own activations: no gradient exchange.
"""
from __future__ import annotations

import torch

WORLD = "world"  # pass as ``group`` to use the default process group


def _resolve(group):
    import torch.distributed as dist
    return dist.group.WORLD if (group is WORLD or group is True) else group


def world_size(group=None) -> int:
    import torch.distributed as dist
    if group is None or not dist.is_available() or not dist.is_initialized():
        return 1
    return dist.get_world_size(_resolve(group))


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous [lo, hi) slice of a batch of n images owned by ``rank`` (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad shard spec world={world} rank={rank}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t: torch.Tensor, world: int, rank: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], world, rank)
    return t[lo:hi]


def allreduce_sums_(sums: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of a statistics buffer over ``group`` (no-op when group is None or has one rank).
    Stream-ordered on the current CUDA stream for NCCL; works on CPU tensors over gloo as well."""
    if group is None:
        return sums
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("a process group was requested but torch.distributed is not initialised")
    pg = _resolve(group)
    if dist.get_world_size(pg) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=pg)
    return sums
