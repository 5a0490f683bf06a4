"""View-sharded data parallelism: one process per GPU, each back-projects views r, r+W, ...;
one all-reduce sums the accumulators.  The path has no other exchange step, so
there is no data-path collective inside the view loop.

num and den are plain sums over views (backproject.py:149-150), so the reduction is exact up
to fp32 summation order.  The reference's 1e-12 den initialiser must be counted once, not once
per rank: ranks > 0 contribute den - 1e-12.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .backproject import DEN_EPS


def shard_views(n_views: int, rank: int, world: int) -> range:
    """Interleaved assignment keeps per-rank load even along a camera path."""
    return range(rank, n_views, world)


def allreduce_accumulators(num: torch.Tensor, den: torch.Tensor, group=None) -> None:
    """In place: every rank ends with the global (num, den).  Works with nccl (GPU tensors) and
    gloo (CPU tensors, used by the world_size-2 tests)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    if dist.get_rank(group) != 0:
        den.sub_(DEN_EPS)
    dist.all_reduce(num, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(den, op=dist.ReduceOp.SUM, group=group)


def reduce_scatter_accumulators(num: torch.Tensor, den: torch.Tensor, group=None):
    """Half the NVLink traffic of the all-reduce when every rank only needs ITS rows of the result
    (each rank then finalises / saves its own shard of features_*.pt): rank r receives the global sums of
    rows [lo, hi) = [r*ceil(N/W), ...).  Returns (num_shard, den_shard, lo, hi); the inputs are left
    unchanged except for the 1e-12 bookkeeping on den."""
    n, d = num.shape
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return num, den, 0, n
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if rank != 0:
        den.sub_(DEN_EPS)
    per = -(-n // world)
    lo, hi = min(rank * per, n), min((rank + 1) * per, n)
    if dist.get_backend(group) == "gloo" or n % world:
        # gloo has no reduce-scatter; ragged N would need padding: fall back to all-reduce + slice
        dist.all_reduce(num, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(den, op=dist.ReduceOp.SUM, group=group)
        return num[lo:hi], den[lo:hi], lo, hi
    num_out = torch.empty(per, d, dtype=num.dtype, device=num.device)
    den_out = torch.empty(per, dtype=den.dtype, device=den.device)
    dist.reduce_scatter_tensor(num_out, num, op=dist.ReduceOp.SUM, group=group)
    dist.reduce_scatter_tensor(den_out, den, op=dist.ReduceOp.SUM, group=group)
    return num_out, den_out, lo, hi
