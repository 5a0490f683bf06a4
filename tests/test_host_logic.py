"""CPU tests of the host logic: scene determinism, view sharding, and the N>1 reduction path
(world_size-2 gloo, as the N-GPU path uses nccl for the same call)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_scene_generator_is_deterministic(gwbp):
    S = gwbp.scene
    a, b = S.make_scene(1000, 3), S.make_scene(1000, 3)
    assert all(np.array_equal(getattr(a, k), getattr(b, k)) for k in ("means", "quats", "scales", "opacities"))
    assert not np.array_equal(a.means, S.make_scene(1000, 4).means)
    vm, K = S.make_cameras(5, 128, 96)
    for v in vm:  # rigid world->camera matrices looking at the origin
        assert np.allclose(v[:3, :3] @ v[:3, :3].T, np.eye(3), atol=1e-5)
        assert v[2, 3] > 3.0  # origin is in front of the camera
    f = S.make_feature_map_np(0, 8, 20, 30)
    assert f.shape == (20, 30, 8) and f.strides[2] > f.strides[1]  # permuted view of a planar buffer
    assert np.abs(np.linalg.norm(S.make_text_queries(3, 16), axis=1) - 1).max() < 1e-6


def test_shard_rows_partition(gwbp):
    for n, world in [(5_800_000, 8), (7, 2), (3, 4), (0, 2), (10, 3)]:
        ranges = [gwbp.dist.shard_rows(n, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))


def test_shard_views_partition(gwbp):
    for n_views, world in [(185, 8), (7, 2), (3, 4), (0, 2)]:
        seen = sorted(v for r in range(world) for v in gwbp.dist.shard_views(n_views, r, world))
        assert seen == list(range(n_views))
        sizes = [len(gwbp.dist.shard_views(n_views, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import gwbp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, d, n_views = 64, 4, 5
    g = torch.Generator().manual_seed(0)
    per_view_num = torch.rand(n_views, n, d, generator=g)
    per_view_den = torch.rand(n_views, n, generator=g)
    num = torch.zeros(n, d)
    den = torch.full((n,), gwbp.DEN_EPS)
    for v in gwbp.dist.shard_views(n_views, rank, world):
        num += per_view_num[v]
        den += per_view_den[v]
    gwbp.dist.allreduce_accumulators(num, den)
    ref_num, ref_den = per_view_num.sum(0), per_view_den.sum(0) + gwbp.DEN_EPS
    ok = torch.allclose(num, ref_num, atol=1e-6) and torch.allclose(den, ref_den, atol=1e-6)
    # reduce-scatter variant: every rank ends with the global sums of its own row range
    num2 = torch.zeros(n, d)
    den2 = torch.full((n,), gwbp.DEN_EPS)
    for v in gwbp.dist.shard_views(n_views, rank, world):
        num2 += per_view_num[v]
        den2 += per_view_den[v]
    ns, ds, lo, hi = gwbp.dist.reduce_scatter_accumulators(num2, den2)
    ok = ok and (lo, hi) == gwbp.dist.shard_rows(n, rank, world) and torch.allclose(ns, ref_num[lo:hi], atol=1e-6) and \
        torch.allclose(ds, ref_den[lo:hi], atol=1e-6)
    # ragged N (not a multiple of the world size): the last rank owns the remainder
    n3 = 7
    num3 = torch.full((n3, d), float(rank + 1))
    den3 = torch.full((n3,), gwbp.DEN_EPS + float(rank + 1))
    ns3, ds3, lo3, hi3 = gwbp.dist.reduce_scatter_accumulators(num3, den3)
    tot = float(sum(range(1, world + 1)))
    ok = ok and (lo3, hi3) == gwbp.dist.shard_rows(n3, rank, world) and (hi3 == n3 if rank == world - 1 else True)
    ok = ok and torch.allclose(ns3, torch.full((hi3 - lo3, d), tot)) and torch.allclose(ds3, torch.full((hi3 - lo3,), tot), atol=1e-6)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_allreduce_accumulators_world2_gloo():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as m:
        out = m.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
        [p.start() for p in procs]
        [p.join(120) for p in procs]
        assert all(p.exitcode == 0 for p in procs)
        assert dict(out) == {0: True, 1: True}


def test_sh_colours_and_feature_maps_refuse_cpu_and_wrong_inputs(gwbp):
    """No CPU fallback anywhere: the SH colour stage (backproject.py:88-100) is a CUDA kernel behind gwbp_sh_colors and
    refuses host tensors; BackProjector validates a feature map BEFORE raw pointers reach the C ABI."""
    means, coeffs = torch.randn(5, 3), torch.randn(5, 16, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        gwbp.sh_colors(3, means, coeffs, torch.eye(4))
    # the map checks themselves are pure host logic: exercise them on an object that skips the CUDA constructor
    bp = gwbp.BackProjector.__new__(gwbp.BackProjector)
    bp.d, bp.device = 8, torch.device("cuda", 0)
    cam = gwbp.make_camera(torch.eye(4), torch.eye(3), 32, 16)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        bp._check_map(torch.zeros(16, 32, 8), cam, False)
    with pytest.raises(AssertionError):
        bp.add_view_host(torch.eye(4), torch.eye(3), 32, 16, torch.zeros(8, 16, 31))       # wrong width
    with pytest.raises(AssertionError):
        bp.add_view_host(torch.eye(4), torch.eye(3), 32, 16, torch.zeros(8, 16, 32).half())  # wrong dtype
    with pytest.raises(AssertionError):
        bp.add_view_host(torch.eye(4), torch.eye(3), 32, 16, torch.zeros(7, 16, 32))       # wrong D


def test_splats_data_contract(gwbp):
    """gsplat checkpoint <-> `splats` dict, activations, pruning, COLMAP pose -> viewmat."""
    import torch
    S = gwbp.splats
    n = 7
    g = torch.Generator().manual_seed(0)
    ckpt = {"splats": {"means": torch.randn(n, 3, generator=g), "quats": torch.randn(n, 4, generator=g),
                       "scales": torch.randn(n, 3, generator=g), "opacities": torch.randn(n, generator=g),
                       "sh0": torch.randn(n, 1, 3, generator=g), "shN": torch.randn(n, 15, 3, generator=g)}}
    ckpt["splats"]["means"].requires_grad_(True)
    sp = S.splats_from_gsplat_checkpoint(ckpt)
    assert set(sp) >= {"means", "rotation", "scaling", "opacity", "features_dc", "features_rest"}
    assert not sp["means"].requires_grad and sp["active_sh_degree"] == 3      # utils.py:11-17,66
    back = S.gsplat_checkpoint_from_splats(sp)
    assert all(torch.equal(back["splats"][k], ckpt["splats"][k].detach()) for k in ckpt["splats"])
    means, quats, scales, opac = S.activated(sp)
    assert torch.equal(scales, torch.exp(sp["scaling"])) and torch.equal(opac, torch.sigmoid(sp["opacity"]))
    assert quats is sp["rotation"] and means is sp["means"]
    keep = torch.tensor([True, False, True, True, False, False, True])
    sp["camera_matrix"] = S.camera_matrix(1000.0, 900.0, 640.0, 360.0, data_factor=4)
    pr = S.prune_splats(sp, keep)
    assert pr["means"].shape[0] == 4 and pr["features_rest"].shape == (4, 15, 3)
    assert torch.equal(pr["opacity"], sp["opacity"][keep]) and pr["camera_matrix"] is sp["camera_matrix"]
    assert torch.allclose(sp["camera_matrix"], torch.tensor([[250.0, 0, 160], [0, 225.0, 90], [0, 0, 1]]))
    R = torch.tensor([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]])
    vm = S.viewmat_from_rotation_translation(R.numpy(), [1.0, 2.0, 3.0])
    assert vm.shape == (4, 4) and torch.equal(vm[:3, :3], R) and vm[:3, 3].tolist() == [1.0, 2.0, 3.0]
    assert vm[3].tolist() == [0.0, 0.0, 0.0, 1.0]
