"""View-sharded data parallelism: one process per GPU, each back-projects views r, r+W, ...;
one exchange sums the accumulators at the end.  The path has no other exchange step, so
there is no data-path collective inside the view loop.

Three closing steps, same result up to fp32 summation order:
  allreduce_accumulators       NCCL all-reduce of (num, den): every rank ends with the whole field
  reduce_scatter_accumulators  NCCL reduce-scatter: every rank ends with ITS rows (half the traffic)
  PeerExchange.reduce_finalize ONE hand-written kernel per rank over NVLink peer memory (CUDA IPC mappings of every
                               rank's accumulators): the owner of a row pulls only the rows its peers actually touched,
                               adds them in rank order and finalises -- sparse, deterministic, fused with the normalise
`finalize_sharded(bp)` picks the peer exchange when it can be set up and says which path ran.

num and den are plain sums over views (backproject.py:149-150), so the reduction is exact up
to fp32 summation order.  The reference's 1e-12 den initialiser must be counted once, not once
per rank: ranks > 0 contribute den - 1e-12.
"""
from __future__ import annotations

import os

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


class PeerExchange:
    """NVLink peer mappings of every rank's (num [N,D], den [N]) + the fused sparse closing kernel
    (gwbp_peer_reduce_finalize).  Collective: every rank of `group` constructs it with its own accumulators (same
    shapes) -- handles are exchanged with all_gather_object -- and every rank calls reduce_finalize() the same number
    of times.  Raises RuntimeError if peer memory cannot be set up (more than 8 ranks, D not covered, memory not
    exportable, e.g. expandable segments); `finalize_sharded` then uses NCCL."""

    def __init__(self, num: torch.Tensor, den: torch.Tensor, group=None):
        import ctypes as C

        from . import _lib as L

        assert dist.is_initialized(), "PeerExchange needs an initialised process group"
        assert num.is_cuda and den.is_cuda and num.dtype == torch.float32 and den.dtype == torch.float32
        assert num.is_contiguous() and den.is_contiguous() and num.dim() == 2 and den.shape == (num.shape[0],)
        self.group, self.num, self.den = group, num, den
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n, self.d = num.shape
        self._bases = []
        lib = L.lib()
        ok = bool(lib.gwbp_peer_reduce_supported(self.world, self.d))
        mine = None
        if ok:
            try:
                mine = []
                with torch.cuda.device(num.device):
                    for t in (num, den):
                        h = (C.c_ubyte * L.IPC_HANDLE_BYTES)()
                        off = C.c_int64(0)
                        L.check(lib.gwbp_ipc_export(t.data_ptr(), h, C.byref(off)), "gwbp_ipc_export")
                        mine.append((bytes(h), int(off.value)))
            except RuntimeError as ex:
                mine, self._why = None, str(ex)
        else:
            self._why = f"world={self.world}, D={self.d} not covered by the peer kernel"
        everyone = [None] * self.world
        dist.all_gather_object(everyone, (mine, tuple(num.shape), os.getpid()), group=group)
        if any(e[0] is None for e in everyone):
            raise RuntimeError("peer exchange unavailable: " + getattr(self, "_why", "a peer could not export its memory"))
        if any(e[1] != tuple(num.shape) for e in everyone):
            raise RuntimeError("peer exchange: accumulator shapes differ between ranks")
        self._num_ptrs = (C.c_void_p * self.world)()
        self._den_ptrs = (C.c_void_p * self.world)()
        failed = None
        with torch.cuda.device(num.device):
            for r, (handles, _, pid) in enumerate(everyone):
                if r == self.rank:
                    self._num_ptrs[r], self._den_ptrs[r] = num.data_ptr(), den.data_ptr()
                    continue
                ptrs, opened = [], {}  # num and den may live in ONE allocation of the peer: map it once
                for hbytes, off in handles:
                    if hbytes not in opened:
                        base = C.c_void_p()
                        rc = lib.gwbp_ipc_open(hbytes, C.byref(base))
                        if rc != 0:
                            failed = failed or f"gwbp_ipc_open(rank {r}): {L.last_error()}"
                            ptrs.append(0)
                            continue
                        self._bases.append(base.value)
                        opened[hbytes] = base.value
                    ptrs.append(opened[hbytes] + off)
                self._num_ptrs[r], self._den_ptrs[r] = ptrs
        flags = [None] * self.world
        dist.all_gather_object(flags, failed, group=group)
        if any(f is not None for f in flags):
            self.close()
            raise RuntimeError("peer exchange unavailable: " + next(f for f in flags if f is not None))

    def _barrier(self) -> None:
        torch.cuda.current_stream(self.num.device).synchronize()
        dist.barrier(group=self.group)

    def reduce_finalize(self, want_num: bool = False):
        """Rows shard_rows(N, rank, W) of the field: (features [rows, D], den [rows], lo, hi[, num [rows, D]]).
        Barriers on both sides: every rank's views are complete before any row is pulled, and nobody touches its
        accumulators again before every peer has finished reading them."""
        from . import _lib as L

        lo, hi = shard_rows(self.n, self.rank, self.world)
        rows = hi - lo
        dev = self.num.device
        feats = torch.empty(rows, self.d, dtype=torch.float32, device=dev)
        den = torch.empty(rows, dtype=torch.float32, device=dev)
        num = torch.empty(rows, self.d, dtype=torch.float32, device=dev) if want_num else None
        self._barrier()
        with torch.cuda.device(dev):
            L.check(L.lib().gwbp_peer_reduce_finalize(
                self._num_ptrs, self._den_ptrs, self.world, lo, rows, self.d, DEN_EPS, feats.data_ptr(),
                num.data_ptr() if num is not None else None, den.data_ptr(),
                int(torch.cuda.current_stream(dev).cuda_stream)), "gwbp_peer_reduce_finalize")
        self._barrier()
        return (feats, den, lo, hi, num) if want_num else (feats, den, lo, hi)

    def close(self) -> None:
        from . import _lib as L

        for b in self._bases:
            L.lib().gwbp_ipc_close(b)
        self._bases = []


_peer_cache = {}


def peer_exchange_for(bp, group=None):
    """The (cached) PeerExchange of a BackProjector, or None when peer memory is unavailable; collective."""
    key = (id(bp), bp.num.data_ptr(), bp.den.data_ptr())
    if key not in _peer_cache:
        try:
            _peer_cache[key] = PeerExchange(bp.num, bp.den, group)
        except RuntimeError as ex:
            _peer_cache[key] = str(ex)
    px = _peer_cache[key]
    return px if isinstance(px, PeerExchange) else None


def finalize_sharded(bp, group=None, exchange: str = "auto"):
    """The closing step of a view-sharded job: every rank ends with ITS finalised rows (backproject.py:166-169).
    exchange="peer": the fused sparse kernel over NVLink peer memory; "nccl": reduce-scatter + finalise; "auto": peer
    when it can be set up.  Returns (features_shard [hi-lo, D], keep_shard [hi-lo] bool, lo, hi)."""
    from .engine import finalize as _finalize

    assert exchange in ("auto", "peer", "nccl"), exchange
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    if multi and exchange != "nccl":
        bp.flush()
        px = peer_exchange_for(bp, group)
        if px is not None:
            feats, den, lo, hi = px.reduce_finalize()
            return feats, den > DEN_EPS, lo, hi
        if exchange == "peer":
            raise RuntimeError("peer exchange requested but unavailable: " + str(_peer_cache.get(
                (id(bp), bp.num.data_ptr(), bp.den.data_ptr()))))
    num, den = bp.raw()
    ns, ds, lo, hi = reduce_scatter_accumulators(num, den, group)
    feats = _finalize(ns.contiguous(), ds.contiguous())
    return feats, ds > DEN_EPS, lo, hi


def save_sharded(bp, path: str, group=None, gather: bool = False, exchange: str = "auto"):
    """Sharded `features_*.pt`: rank r writes `<path>.shard<r>of<W>.pt` = {"features": [kept rows of its range, D],
    "kept": global Gaussian indices of those rows}; concatenating the shards in rank order gives exactly the tensor
    the single-GPU job saves (rows follow prune_by_gradients' mask, utils.py:257-268).  gather=True additionally
    collects everything on rank 0 and writes the reference's single-tensor file `path` there."""
    feats, keep, lo, hi = finalize_sharded(bp, group, exchange)
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
