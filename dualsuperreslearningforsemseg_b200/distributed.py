"""Batch sharding helpers for the hot path.  One process per GPU; torch.distributed (NCCL on the
B200 box, gloo in CPU tests) is only plumbing.

* FA loss: samples are independent -- no gradient exchange.  Inside DDP training nothing is needed at all
  (equal per-rank batches; DDP averages parameter gradients; the reference logs rank 0's local loss,
  train_or_resume.py:448-472).  ``all_reduce_mean_loss`` gives the global mean for reporting: one scalar.
* Metrics: ``mIoU.sync()`` / ``Accuracy.sync()`` exchange the per-update int64 count rows (exact at any world size).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized()


def shard_slice(n: int, rank: int, world: int) -> slice:
    """Contiguous shard of n samples for `rank` (first n % world ranks get one extra)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return slice(lo, lo + base + (1 if rank < rem else 0))


def all_reduce_mean_loss(loss: torch.Tensor, group=None) -> torch.Tensor:
    """Global mean of per-rank mean losses (equal per-rank batch sizes), accumulated in float64."""
    if not is_dist():
        return loss.detach().clone()
    t = loss.detach().to(torch.float64).clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return (t / dist.get_world_size(group)).to(loss.dtype)
