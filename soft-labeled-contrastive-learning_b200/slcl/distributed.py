"""Data-parallel plumbing of the path: one process per GPU, the
batch dimension is sharded, and the ONLY data-path collective is one all-reduce
(sum) of the per-class feature sums / weight sums ``[sets*K, C+1]`` float64
between the class-sum kernel and its finaliser.  Per-pixel losses need no
exchange beyond their scalar numerator/denominator."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of ``n_items`` for ``rank`` (sizes differ by at most 1)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """Sum ``sums`` over the ranks of ``group`` (NCCL over NVLink on GPUs, gloo on
    CPU in the tests).  Returns the reduced tensor; a no-op when torch.distributed
    is not initialised or the world has one rank."""
    if not dist.is_available() or not dist.is_initialized():
        return sums
    if dist.get_world_size(group if group is not True else None) == 1:
        return sums
    out = sums.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=None if group is True else group)
    return out


def all_reduce_mean_loss(numerator: torch.Tensor, denominator: torch.Tensor, group=None) -> torch.Tensor:
    """Global ratio of sums for a sharded per-pixel loss: each rank passes its
    local numerator (sum of weighted row losses) and denominator (sum of weights)."""
    pair = torch.stack([numerator.reshape(()), denominator.reshape(())]).to(torch.float64)
    pair = all_reduce_sums(pair, group)
    return (pair[0] / pair[1]).to(numerator.dtype)
