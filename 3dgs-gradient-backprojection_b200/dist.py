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


def shard_rows(n: int, rank: int, world: int):
    """Row range [lo, hi) of the feature field that rank `rank` owns after the reduce-scatter: n // world rows each,
    the remainder (< world rows) goes to the last rank."""
    per = n // world
    lo = rank * per
    hi = n if rank == world - 1 else lo + per
    return lo, hi


def reduce_scatter_accumulators(num: torch.Tensor, den: torch.Tensor, group=None):
    """Half the NVLink traffic of the all-reduce when every rank only needs ITS rows of the result (each rank then
    finalises / saves its own shard of features_*.pt): rank r receives the global sums of rows shard_rows(n, r, W).
    Returns (num_shard, den_shard, lo, hi); the inputs are left unchanged except for the 1e-12 bookkeeping on den.
    N need not divide by W: the first (N // W) * W rows go through one reduce-scatter, the < W remaining rows through a
    tiny all-reduce and belong to the last rank."""
    n, d = num.shape
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return num, den, 0, n
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if rank != 0:
        den.sub_(DEN_EPS)
    lo, hi = shard_rows(n, rank, world)
    if dist.get_backend(group) == "gloo":  # gloo (CPU tests) has no reduce-scatter: all-reduce + slice
        dist.all_reduce(num, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(den, op=dist.ReduceOp.SUM, group=group)
        return num[lo:hi], den[lo:hi], lo, hi
    per, main = n // world, (n // world) * world
    num_out = torch.empty(hi - lo, d, dtype=num.dtype, device=num.device)
    den_out = torch.empty(hi - lo, dtype=den.dtype, device=den.device)
    if per:
        dist.reduce_scatter_tensor(num_out[:per], num[:main], op=dist.ReduceOp.SUM, group=group)
        dist.reduce_scatter_tensor(den_out[:per], den[:main], op=dist.ReduceOp.SUM, group=group)
    if main < n:  # ragged tail
        tail_n, tail_d = num[main:].clone(), den[main:].clone()
        dist.all_reduce(tail_n, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(tail_d, op=dist.ReduceOp.SUM, group=group)
        if rank == world - 1:
            num_out[per:] = tail_n
            den_out[per:] = tail_d
    return num_out, den_out, lo, hi


def finalize_sharded(bp, group=None):
    """The closing step of a view-sharded job: reduce-scatter (num, den), then every rank finalises ITS rows
    (backproject.py:166-169).  Returns (features_shard [hi-lo, D], keep_shard [hi-lo] bool, lo, hi)."""
    from .engine import finalize as _finalize

    num, den = bp.raw()
    ns, ds, lo, hi = reduce_scatter_accumulators(num, den, group)
    feats = _finalize(ns.contiguous(), ds.contiguous())
    return feats, ds > DEN_EPS, lo, hi


def save_sharded(bp, path: str, group=None, gather: bool = False):
    """Sharded `features_*.pt`: rank r writes `<path>.shard<r>of<W>.pt` = {"features": [kept rows of its range, D],
    "kept": global Gaussian indices of those rows}; concatenating the shards in rank order gives exactly the tensor
    the single-GPU job saves (rows follow prune_by_gradients' mask, utils.py:257-268).  gather=True additionally
    collects everything on rank 0 and writes the reference's single-tensor file `path` there."""
    feats, keep, lo, hi = finalize_sharded(bp, group)
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    kept = torch.nonzero(keep).flatten() + lo
    shard = {"features": feats[keep].contiguous().cpu(), "kept": kept.cpu(), "rows": (lo, hi), "world": world}
    torch.save(shard, f"{path}.shard{rank}of{world}.pt")
    if gather:
        parts = [None] * world if rank == 0 else None
        if world > 1:
            dist.gather_object(shard, parts, dst=0, group=group)
        else:
            parts = [shard]
        if rank == 0:
            torch.save(torch.cat([p["features"] for p in parts], 0), path)
            torch.save(torch.cat([p["kept"] for p in parts], 0), path + ".kept.pt")
    return shard
