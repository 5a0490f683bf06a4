"""The fused multi-GPU closing step (gwbp_peer_reduce_finalize over CUDA IPC peer mappings), exercised with TWO
PROCESSES ON ONE GPU: same code path as one process per GPU over NVLink (IPC export / open, pointer tables, the sparse
pull kernel, barriers), gloo for the host-side handle exchange because NCCL refuses two ranks on one device.
Reference: one process back-projecting all views and finalising (backproject.py:166-169)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, out):
    sys.path.insert(0, ROOT)
    import gwbp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    S = gwbp.scene
    W, H, n_views = 96, 64, 6
    sc = S.make_scene(n, 11)
    vm, K = S.make_cameras(n_views, W, H, 11)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    feats = [dev(np.ascontiguousarray(S.make_feature_map_np(v, d, H, W, 11))) for v in range(n_views)]
    mk = lambda: gwbp.BackProjector(dev(sc.means), dev(sc.quats), dev(sc.scales), dev(sc.opacities), d)  # noqa: E731
    bp = mk()
    for v in gwbp.dist.shard_views(n_views, rank, world):
        bp.add_view(vm[v], K, W, H, feats[v])
    ref = mk()  # the single-process job
    for v in range(n_views):
        ref.add_view(vm[v], K, W, H, feats[v])
    f_ref, keep_ref = ref.finalize(), ref.prune_mask()
    failed = []

    def check(name, cond):
        if not bool(cond):
            failed.append(name)

    def rows_close(a, b, scale, tol):
        """max_j |a - b| <= tol * scale per row: the two jobs add the same fp32 terms in different orders (per-rank
        partial sums, atomics), so the error bound is relative to the size of the TERMS (the row's den, features being
        unit vectors), not to the possibly cancelled sum."""
        return float(((a - b).abs().amax(dim=1) / scale.clamp_min(1e-30)).max()) <= tol

    f, keep, lo, hi = gwbp.dist.finalize_sharded(bp, exchange="peer")
    check("row range", (lo, hi) == gwbp.dist.shard_rows(n, rank, world) and f.shape == (hi - lo, d))
    check("prune mask", torch.equal(keep, keep_ref[lo:hi]))
    seen = ref.den[lo:hi] > 1e-6
    rel = (f[seen] - f_ref[lo:hi][seen]).norm(dim=1) / f_ref[lo:hi][seen].norm(dim=1).clamp_min(1e-12)
    check(f"feature rows rel-err {float(rel.max()) if rel.numel() else 0:.2e}", rel.numel() == 0 or float(rel.max()) <= 1e-4)
    check("untouched rows are zero", float(f[~keep].abs().max() if (~keep).any() else 0.0) == 0.0)
    # raw sums too, twice (the exchange is repeatable and deterministic: rank-ordered sums, accumulators left untouched)
    px = gwbp.dist.peer_exchange_for(bp)
    for rep in range(2):
        f2, den2, lo2, hi2, num2 = px.reduce_finalize(want_num=True)
        check(f"repeat {rep}: bit-identical features", torch.equal(f2, f))
        check(f"repeat {rep}: num", rows_close(num2, ref.num[lo:hi], ref.den[lo:hi], 2e-6))
        check(f"repeat {rep}: den", torch.allclose(den2, ref.den[lo:hi], atol=1e-12, rtol=2e-6))
    out[rank] = (failed, float((bp.den > gwbp.DEN_EPS).float().mean()))
    px.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,d", [(3001, 32), (2500, 512), (64, 768)])
def test_peer_exchange_two_processes_one_gpu(n, d):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as m:
        out = m.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n, d, out)) for r in range(2)]
        [p.start() for p in procs]
        [p.join(300) for p in procs]
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        res = dict(out)
        assert res[0][0] == [] and res[1][0] == [], res
        print(f"[peer exchange] rows touched per rank: {res[0][1]:.3f}, {res[1][1]:.3f}")
