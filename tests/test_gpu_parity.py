"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Everything goes through the C ABI of
lib/libgwbp.so (via the ctypes binding) and is compared with the CPU oracle on the same seeded
inputs.  Bars (BASELINE.json north_star / BASELINE.md §3):
  * integer stages (radii, isect_ids, flatten_ids, isect_offsets, gaussian_ids): BIT-EXACT
  * per-Gaussian normalised features: row rel-err <= 1e-4, cosine >= 0.9999 (rows with den > 1e-6)
  * identical den>0 (prune) mask; identical segmentation masks (ties aside)
Threshold discontinuities: `alpha < 1/255 -> skip` and `T(1-alpha) <= 1e-4 -> stop` are hard cut-offs, and the GPU's
ex2.approx differs from libm's expf in the last bits, so a (pixel, Gaussian) pair sitting ON a threshold can fall on
the other side.  There is NO blanket outlier allowance: every row above the bar must be PROVEN to be such a flip --
the oracle reports, per row, the smallest relative distance to a threshold among the tests that can change it
(oracle.c::orc_view_margins) and that distance must be < FLIP_MARGIN; the count is printed beside the base rate."""
import numpy as np
import pytest
import torch

from helpers import oracle_job, oracle_margins, row_cosine, row_rel_err, small_case

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4      # north_star: rel-err <= 1e-4 (fp32 accumulate)
COS_TOL = 0.9999    # north_star: cosine >= 0.9999
FLIP_MARGIN = 1e-5  # a row may exceed the bars only if one of its threshold tests is closer than this (relative)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _feat_dev(f):
    """Upload a [H,W,D] numpy view keeping the reference's layout: a permuted view of a planar
    [D,H,W] buffer (backproject.py:113)."""
    planar = np.ascontiguousarray(np.transpose(f, (2, 0, 1)))
    return torch.from_numpy(planar).cuda().permute(1, 2, 0)


def _gpu_job(gwbp, sc, vm, K, W, H, feats, d, kernel="simt", contiguous=False):
    bp = gwbp.BackProjector(_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities), d, kernel=kernel,
                            collect_stats=True)
    for v in range(vm.shape[0]):
        f = _feat_dev(feats[v])
        bp.add_view(vm[v], K, W, H, f.contiguous() if contiguous else f)
    return bp


def _check_features(bp, num_o, den_o, noracle, margin=None, tag=""):
    """Bars of BASELINE.json on every row with den > 1e-6: rel-err <= 1e-4, cosine >= 0.9999, den rel-err <= 1e-4,
    identical prune mask.  `margin` (helpers.oracle_margins): rows above a bar must be proven threshold flips."""
    num = bp.num.double().cpu().numpy()
    den = (bp.den.double().cpu().numpy() - 1e-12)
    return _check_arrays(num, den, bp.finalize().double().cpu().numpy(), num_o, den_o, noracle, margin, tag)


def _check_arrays(num, den, f_gpu, num_o, den_o, noracle, margin, tag=""):
    mask_diff = (den > 5e-13) != (den_o > 0)
    sel = den_o > 1e-6
    f_ref = noracle.finalize(num_o[sel], den_o[sel] + 1e-12)
    return _check_rows(f_gpu[sel], den[sel], f_ref, den_o[sel], None if margin is None else margin[sel], mask_diff,
                       None if margin is None else margin[mask_diff], tag)


def _check_rows(f_gpu, den, f_ref, den_o, m, mask_diff, m_mask_diff, tag=""):
    """Rows already restricted to den_o > 1e-6; `m` = their threshold margins; mask_diff / m_mask_diff = the rows whose
    den > 0 decision differs and their margins."""
    rel, _ = row_rel_err(f_gpu, f_ref)
    cos, _ = row_cosine(f_gpu, f_ref)
    den_rel = np.abs(den - den_o) / den_o
    bad = (rel > REL_TOL) | (cos < COS_TOL) | (den_rel > REL_TOL)
    if m is None:
        assert not bad.any(), f"{int(bad.sum())} rows above the bars (max rel {rel.max():.3e}) and no margins supplied"
        assert not mask_diff.any(), "prune mask differs"
    else:
        unexplained = bad & ~(m < FLIP_MARGIN)
        assert not unexplained.any(), (f"{int(unexplained.sum())} rows above the bars are NOT threshold flips: "
                                       f"max rel {rel[unexplained].max():.3e}, margins {m[unexplained][:5]}")
        # a den > 0 <-> den == 0 disagreement is the same thing: the row's only pair sits on a threshold
        assert not (m_mask_diff >= FLIP_MARGIN).any(), "prune mask differs away from thresholds"
        base = float((m < FLIP_MARGIN).mean())
        print(f"[parity{tag}] rows {rel.size}: {int(bad.sum())} above the bars, all with a threshold test closer than "
              f"{FLIP_MARGIN:g} (such rows are {100 * base:.2f} % of all rows); prune-mask flips {int(mask_diff.sum())}; "
              f"other rows: rel max {rel[~bad].max():.2e}, cos min {cos[~bad].min():.7f}, den rel max {den_rel[~bad].max():.2e}")
    return rel, cos


@pytest.fixture(scope="module")
def case(gwbp):
    return small_case(gwbp.scene)


def test_native_library_is_loaded(gwbp):
    import ctypes
    assert torch.cuda.is_available()
    assert gwbp._lib.lib().gwbp_abi_version() == gwbp._lib.ABI_VERSION
    with open("/proc/self/maps") as f:
        assert "libgwbp.so" in f.read()


@pytest.mark.parametrize("cull", [False, True])
@pytest.mark.parametrize("counting_bin", [False, True])
def test_integer_stages_bit_exact(gwbp, coracle, case, cull, counting_bin):
    """cull=False: gsplat-1.4.0 isect_tiles semantics; cull=True: the exact tile-culling extension.
    counting_bin=False: emit + radix sort on the tile id (default); True: the hand-written sort-free counting path."""
    sc, vm, K, _ = case
    scene = gwbp.PackedScene(_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities))
    for v in range(vm.shape[0]):
        view = gwbp.View(scene, gwbp.make_camera(vm[v], K, 96, 64), tile_cull=cull, counting_bin=counting_bin)
        assert view.info.tile_key_bytes == (0 if counting_bin else 2)
        e = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[v], K, 96, 64, cull=cull).export()
        m = view.meta()
        gids = m["gaussian_ids"].cpu().numpy()
        assert np.array_equal(gids, e["gaussian_ids"])
        assert np.array_equal(m["radii"].cpu().numpy(), e["radii"][gids])
        assert np.array_equal(m["isect_ids"].cpu().numpy(), e["isect_ids"])
        assert np.array_equal(m["flatten_ids"].cpu().numpy(), e["flatten_ids"])
        assert np.array_equal(m["isect_offsets"].cpu().numpy()[0], e["isect_offsets"])
        # float outputs of the -fmad=false projection are bit-exact too
        assert np.array_equal(m["means2d"].cpu().numpy().view(np.int32), e["means2d"][gids].view(np.int32))
        assert np.array_equal(m["conics"].cpu().numpy().view(np.int32), e["conics"][gids].view(np.int32))
        assert np.array_equal(m["depths"].cpu().numpy().view(np.int32), e["depths"][gids].view(np.int32))


def test_integer_stages_bit_exact_config_S_and_odd_sizes(gwbp, coracle):
    S = gwbp.scene
    # (3000, 4112, 4100): 257 x 257 = 66 049 tiles -- beyond 16-bit tile keys (32-bit key path) and beyond the counting
    # path (which then silently uses the radix sort); (40_000, 1920, 1080): config M's 8 160 tiles (5 counting warps per
    # CTA); the small scenes at large images hold rectangles of more than 64 tiles (warp-cooperative emission)
    for (n, W, H, seed) in [(50_000, 256, 256, 0), (20_000, 333, 211, 3), (5_000, 17, 15, 4), (2_000, 1297, 840, 5),
                            (40_000, 1920, 1080, 8), (700, 2200, 1400, 9), (3_000, 4112, 4100, 6)]:
        sc = S.make_scene(n, seed)
        vm, K = S.make_cameras(2, W, H, seed)
        scene = gwbp.PackedScene(_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities))
        for cull in (False, True):
            e = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[1], K, W, H, cull=cull).export()
            for counting_bin in (False, True):
                view = gwbp.View(scene, gwbp.make_camera(vm[1], K, W, H), tile_cull=cull, counting_bin=counting_bin)
                tiles = view.info.tile_w * view.info.tile_h
                want = 0 if (counting_bin and tiles <= 12288) else (2 if tiles <= 65536 else 4)
                assert view.info.tile_key_bytes == want, (n, W, H, view.info.tile_key_bytes)
                m = view.meta()
                assert np.array_equal(m["isect_ids"].cpu().numpy(), e["isect_ids"]), (n, W, H, cull, counting_bin)
                assert np.array_equal(m["flatten_ids"].cpu().numpy(), e["flatten_ids"]), (n, W, H, cull, counting_bin)
                assert np.array_equal(m["isect_offsets"].cpu().numpy()[0], e["isect_offsets"]), (n, W, H, cull, counting_bin)


@pytest.mark.parametrize("kernel", ["simt", "auto"])
def test_backprojection_small(gwbp, coracle, noracle, case, kernel):
    sc, vm, K, feats = case
    num_o, den_o, st = oracle_job(coracle, sc, vm, K, 96, 64, feats, 8)
    margin = oracle_margins(coracle, sc, vm, K, 96, 64, den_o)
    bp = _gpu_job(gwbp, sc, vm, K, 96, 64, feats, 8, kernel)
    _check_features(bp, num_o, den_o, noracle, margin, f" small/{kernel}")
    s = bp.stats()
    assert abs(s["rows_nonzero"] - sum(x["rows_nonzero"] for x in st)) <= 4  # threshold flips only


@pytest.mark.parametrize("kernel", ["simt", "auto"])
def test_backprojection_config_S(gwbp, coracle, noracle, kernel):
    """BASELINE config 1: 50k Gaussians, 8 views 256x256, 64-d."""
    S = gwbp.scene
    c = S.CONFIGS["S"]
    sc = S.make_scene(c["n"], 0)
    vm, K = S.make_cameras(c["views"], c["width"], c["height"], 0)
    feats = [S.make_feature_map_np(v, c["d"], c["height"], c["width"], 0) for v in range(c["views"])]
    num_o, den_o, _ = oracle_job(coracle, sc, vm, K, c["width"], c["height"], feats, c["d"])
    margin = oracle_margins(coracle, sc, vm, K, c["width"], c["height"], den_o)
    bp = _gpu_job(gwbp, sc, vm, K, c["width"], c["height"], feats, c["d"], kernel)
    _check_features(bp, num_o, den_o, noracle, margin, f" config S/{kernel}")


def test_tensor_core_and_cuda_core_kernels_flip_identically(gwbp):
    """Both kernels evaluate a (pixel, Gaussian) pair with the SAME roundings (common.cuh pair_sigma / pair_alpha /
    composite_step), so their weights are bit-identical: same den up to the summation order, same non-zero rows --
    the tcgen05 path adds no threshold flips of its own (config S)."""
    S = gwbp.scene
    c = S.CONFIGS["S"]
    sc = S.make_scene(c["n"], 0)
    vm, K = S.make_cameras(c["views"], c["width"], c["height"], 0)
    feats = [S.make_feature_map_np(v, c["d"], c["height"], c["width"], 0) for v in range(c["views"])]
    a = _gpu_job(gwbp, sc, vm, K, c["width"], c["height"], feats, c["d"], "simt")
    b = _gpu_job(gwbp, sc, vm, K, c["width"], c["height"], feats, c["d"], "tc")
    assert a.stats()["rows_nonzero"] == b.stats()["rows_nonzero"]  # (entries_walked counts whole batches: 32 vs 128)
    assert torch.equal(a.den > 1e-12, b.den > 1e-12)
    assert torch.allclose(a.den, b.den, rtol=2e-6, atol=1e-12)  # fp32 sums of identical terms in a different order
    seen = a.den > 1e-6
    err = (a.num[seen] - b.num[seen]).norm(dim=1) / a.num[seen].norm(dim=1).clamp_min(1e-9)
    assert float(err.max()) < 2e-5, float(err.max())  # split-bf16 contraction vs fp32 FMA, no outliers at all


@pytest.mark.parametrize("d", [1, 3, 16, 100, 512, 768, 1024])
def test_backprojection_feature_dims(gwbp, coracle, noracle, d):
    S = gwbp.scene
    sc, vm, K, _ = small_case(S, n=1500, views=1, width=64, height=48, d=8, seed=7)
    feats = [S.make_feature_map_np(0, d, 48, 64, 7, enc_res=10)]
    num_o, den_o, _ = oracle_job(coracle, sc, vm, K, 64, 48, feats, d)
    margin = oracle_margins(coracle, sc, vm, K, 64, 48, den_o)
    for kernel in ("simt", "auto"):
        bp = _gpu_job(gwbp, sc, vm, K, 64, 48, feats, d, kernel)
        _check_features(bp, num_o, den_o, noracle, margin, f" D={d}/{kernel}")


@pytest.mark.parametrize("mode,d,eh,ew,W,H,adjoint", [
    ("nearest", 1024, 64, 64, 422, 274, True),    # DINOv2: 64 x 64 tokens, [N,1024] accumulators (backproject.py:206-249)
    ("bilinear", 512, 60, 60, 330, 230, True),    # LSeg-shaped, windows of up to 6 x 5 texels (3 column chunks)
    ("bilinear", 100, 24, 31, 211, 137, True),    # D not a multiple of 16, non-square map
    ("nearest", 48, 9, 9, 211, 137, True),
    ("bilinear", 16, 7, 7, 211, 137, True),
    ("bilinear", 64, 64, 64, 211, 137, False),    # mild up-sampling: windows too large -> re-layout fallback
])
def test_encoder_resolution_maps_against_the_oracle(gwbp, coracle, noracle, mode, d, eh, ew, W, H, adjoint):
    """The DINOv2 variant (backproject.py:206-211,236-249: [N,1024] accumulators, 64x64 tokens upsampled with
    mode="nearest") and the LSeg one (:108-112, bilinear), fed with the ENCODER-resolution map: the oracle
    upsamples with its own restatement of F.interpolate (oracle/gsplat_oracle.py::upsample) and back-projects the
    full-resolution map; the GPU gets the low-res map (adjoint kernel: down-sampled weights x low-res map; and the
    fused-upsample re-layout) and, for reference, the materialised full-resolution map."""
    S = gwbp.scene
    sc = S.make_scene(6000, 21)
    vm, K = S.make_cameras(2, W, H, 21)
    rng = np.random.default_rng(d + eh)
    lows = []
    for _ in range(2):
        low = rng.standard_normal((eh, ew, d)).astype(np.float32)
        lows.append(low / np.linalg.norm(low, axis=2, keepdims=True))
    feats = [noracle.upsample(low, H, W, mode) for low in lows]
    num_o, den_o, _ = oracle_job(coracle, sc, vm, K, W, H, feats, d)
    margin = oracle_margins(coracle, sc, vm, K, W, H, den_o)
    assert bool(gwbp._lib.lib().gwbp_lowres_adjoint_supported(W, H, eh, ew, d, int(mode == "nearest"))) == adjoint
    args = (_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities), d)
    jobs = {"adjoint": gwbp.BackProjector(*args, kernel="tc", collect_stats=True),
            "upsample": gwbp.BackProjector(*args, kernel="tc", collect_stats=True),
            "materialised": gwbp.BackProjector(*args, kernel="tc", collect_stats=True)}
    jobs["upsample"].lowres_impl = "upsample"
    for v in range(2):
        planar_low = _dev(np.ascontiguousarray(np.transpose(lows[v], (2, 0, 1))))  # encoder output [D,h,w]
        for name in ("adjoint", "upsample"):
            jobs[name].add_view_lowres(vm[v], K, W, H, planar_low.permute(1, 2, 0), mode=mode)
        jobs["materialised"].add_view(vm[v], K, W, H, _feat_dev(feats[v]))
    for name, bp in jobs.items():
        _check_features(bp, num_o, den_o, noracle, margin, f" lowres {mode} D={d} {name}")
        assert bp.stats() == jobs["materialised"].stats(), name  # same weights, same rows
    # the single C-ABI call without the separate pack (gwbp_backproject_view_lowres packs the map itself, or takes the
    # fused-upsample fallback when the geometry is not covered): same accumulators as the two-call path above
    from gwbp.engine import View, make_camera, fpack_bytes
    bp1 = gwbp.BackProjector(*args, kernel="tc")
    fp = torch.empty(fpack_bytes(W, H, d), dtype=torch.uint8, device="cuda")
    for v in range(2):
        planar_low = _dev(np.ascontiguousarray(np.transpose(lows[v], (2, 0, 1))))
        view = View(bp1.scene, make_camera(vm[v], K, W, H), tile_cull=True, supertile=True)
        view.backproject_lowres(planar_low.permute(1, 2, 0), mode == "nearest", bp1.num, bp1.den, fp)
    ref = jobs["adjoint"] if adjoint else jobs["upsample"]
    assert torch.equal(bp1.den > 1e-12, ref.den > 1e-12)
    assert torch.allclose(bp1.den, ref.den, rtol=2e-6, atol=1e-12)
    seen = ref.den > 1e-6
    err = (bp1.num[seen] - ref.num[seen]).abs().amax(dim=1) / ref.den[seen]
    assert float(err.max()) < 2e-6, float(err.max())


def test_tile_culling_does_not_change_the_accumulators(gwbp, case):
    """Culled (Gaussian, tile) pairs have zero weight on every pixel: same rows, same sums."""
    sc, vm, K, feats = case
    out = []
    for cull in (False, True):
        bp = gwbp.BackProjector(_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities), 8, kernel="simt",
                                collect_stats=True, tile_cull=cull)
        for v in range(vm.shape[0]):
            bp.add_view(vm[v], K, 96, 64, _feat_dev(feats[v]))
        out.append((bp.num.clone(), bp.den.clone(), bp.stats()))
    (n0, d0, s0), (n1, d1, s1) = out
    assert torch.allclose(n0, n1, rtol=1e-5, atol=1e-6) and torch.allclose(d0, d1, rtol=1e-5, atol=1e-7)
    assert s0["rows_nonzero"] == s1["rows_nonzero"] and s1["entries_walked"] < s0["entries_walked"]
    assert torch.equal(d0 > 1e-12, d1 > 1e-12)


@pytest.mark.parametrize("cull", [False, True])
@pytest.mark.parametrize("scale_mul,W,H", [(1.0, 256, 256), (1.0, 422, 274), (6.0, 330, 230), (25.0, 211, 137),
                                           (3.0, 2304, 1296)])  # 18 x 21 = 378 supertiles: 16-bit ids, two radix passes
def test_supertile_lists_give_the_per_tile_results(gwbp, scale_mul, W, H, cull):
    """BackProjector bins views for the tcgen05 kernels into 8 x 4-tile supertiles (one entry per Gaussian and
    supertile + a 32-bit tile mask, filtered per tile by the kernels' lister warp).  Same Gaussians in the same order as
    the per-tile lists: identical walk / row counters, identical intersection counts, accumulators equal up to the
    order of the fp32 atomic adds.  scale_mul blows the Gaussians up so that rectangles of more than 64 tiles (the
    warp-cooperative path) and masks spanning several supertiles occur; odd image sizes give partial supertiles."""
    S = gwbp.scene
    sc = S.make_scene(20000, 3)
    scales = (sc.scales * scale_mul).astype(np.float32)
    vm, K = S.make_cameras(3, W, H, 3)
    d = 32
    if W * H > 500_000:  # large image: generate on the device (the numpy up-sampler would need GBs of host temporaries)
        g = torch.Generator(device="cuda").manual_seed(3)
        feats = [torch.rand(d, H, W, generator=g, device="cuda").permute(1, 2, 0) for _ in range(3)]
    else:
        feats = [_feat_dev(S.make_feature_map_np(v, d, H, W, 3)) for v in range(3)]
    out = []
    for sup in (False, True):
        bp = gwbp.BackProjector(_dev(sc.means), _dev(sc.quats), _dev(scales), _dev(sc.opacities), d, kernel="tc",
                                collect_stats=True, tile_cull=cull, supertile=sup)
        counts = []
        for v in range(3):
            view = bp.add_view(vm[v], K, W, H, feats[v])
            assert view.info.list_kind == int(sup)
            counts.append((view.n_vis, view.n_isects))
            if sup:
                assert 0 < view.n_entries <= view.n_isects
        out.append((bp.num.clone(), bp.den.clone(), bp.stats(), counts))
    (n0, d0, s0, c0), (n1, d1, s1, c1) = out
    assert c0 == c1 and s0 == s1, (c0, c1, s0, s1)
    assert torch.equal(d0 > 1e-12, d1 > 1e-12)
    assert torch.allclose(d0, d1, rtol=2e-6, atol=1e-12)
    seen = d0 > 1e-6
    err = (n0[seen] - n1[seen]).norm(dim=1) / n0[seen].norm(dim=1).clamp_min(1e-9)
    assert float(err.max()) < 2e-5, float(err.max())


def test_tile_culling_config_S_identical_results(gwbp):
    """Tile culling is ON by default in BackProjector and is not part of gsplat: at config S (50k Gaussians, 8 views
    of 256x256) the culled and the gsplat-exact intersection lists must give the same non-zero rows, the same prune
    mask and the same accumulators (only the order of the fp32 atomic adds may differ)."""
    S = gwbp.scene
    c = S.CONFIGS["S"]
    sc = S.make_scene(c["n"], 0)
    vm, K = S.make_cameras(c["views"], c["width"], c["height"], 0)
    feats = [_feat_dev(S.make_feature_map_np(v, c["d"], c["height"], c["width"], 0)) for v in range(c["views"])]
    for kernel in ("simt", "tc"):
        out = []
        for cull in (False, True):
            bp = gwbp.BackProjector(_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities), c["d"],
                                    kernel=kernel, collect_stats=True, tile_cull=cull)
            isects = 0
            for v in range(c["views"]):
                isects += bp.add_view(vm[v], K, c["width"], c["height"], feats[v]).n_isects
            out.append((bp.num.clone(), bp.den.clone(), bp.stats(), isects))
        (n0, d0, s0, i0), (n1, d1, s1, i1) = out
        assert s0["rows_nonzero"] == s1["rows_nonzero"] and i1 < 0.8 * i0 and s1["entries_walked"] < s0["entries_walked"]
        assert torch.equal(d0 > 1e-12, d1 > 1e-12)
        assert torch.allclose(d0, d1, rtol=2e-6, atol=1e-12)
        seen = d0 > 1e-6
        err = (n0[seen] - n1[seen]).norm(dim=1) / n0[seen].norm(dim=1).clamp_min(1e-9)
        assert float(err.max()) < 2e-5, (kernel, float(err.max()))


@pytest.mark.parametrize("mode", ["bilinear", "nearest"])
def test_fused_lowres_upsample_matches_materialised(gwbp, mode):
    """add_view_lowres(enc_out) == add_view(F.interpolate(enc_out)) (backproject.py:110-113, :245-249)."""
    S = gwbp.scene
    W, H, d = 211, 137, 32
    sc = S.make_scene(6000, 9)
    vm, K = S.make_cameras(2, W, H, 9)
    g = torch.Generator(device="cuda").manual_seed(3)
    args = (_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities), d)
    a, b = gwbp.BackProjector(*args, kernel="tc"), gwbp.BackProjector(*args, kernel="tc")
    for v in range(2):
        low = torch.nn.functional.normalize(torch.randn(1, d, 24, 31, generator=g, device="cuda"), dim=1)
        full = torch.nn.functional.interpolate(low, size=(H, W), mode=mode)[0].permute(1, 2, 0)
        a.add_view(vm[v], K, W, H, full)
        b.add_view_lowres(vm[v], K, W, H, low[0].permute(1, 2, 0), mode=mode)
    assert torch.allclose(a.den, b.den, rtol=1e-5, atol=1e-9)  # same weights; only the atomic order differs
    fa, fb = a.finalize(), b.finalize()
    seen = a.den > 1e-6
    assert float((fa[seen] - fb[seen]).norm(dim=1).max()) < 2e-5


def test_feature_strides_do_not_matter(gwbp, case):
    sc, vm, K, feats = case
    a = _gpu_job(gwbp, sc, vm, K, 96, 64, feats, 8, "simt", contiguous=False)
    b = _gpu_job(gwbp, sc, vm, K, 96, 64, feats, 8, "simt", contiguous=True)
    assert torch.allclose(a.num, b.num, rtol=1e-5, atol=1e-6) and torch.allclose(a.den, b.den, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("d", [48, 272])
def test_tcgen05_feature_layouts_agree(gwbp, d):
    """The tensor-core path re-lays-out F itself: the reference's planar [D,H,W] view (lane = pixel kernel), a
    contiguous [H,W,D] tensor and an arbitrary strided view (generic kernel) must give the same accumulators, and
    agree with the CUDA-core kernel that reads F in place.  Odd image size: partial tiles in x and y; d = 272: a
    second, 16-column chunk."""
    S = gwbp.scene
    W, H = 211, 137
    sc = S.make_scene(5000, 13)
    vm, K = S.make_cameras(2, W, H, 13)
    g = torch.Generator(device="cuda").manual_seed(5)
    args = (_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities), d)
    planar = [torch.randn(d, H, W, generator=g, device="cuda") for _ in range(2)]
    wide = [torch.zeros(H, 2 * W, d + 3, device="cuda") for _ in range(2)]
    for v in range(2):
        wide[v][:, ::2, 1:d + 1] = planar[v].permute(1, 2, 0)
    layouts = {
        "planar": [p.permute(1, 2, 0) for p in planar],
        "contiguous": [p.permute(1, 2, 0).contiguous() for p in planar],
        "strided": [w[:, ::2, 1:d + 1] for w in wide],
    }
    ref = gwbp.BackProjector(*args, kernel="simt")
    for v in range(2):
        ref.add_view(vm[v], K, W, H, layouts["planar"][v])
    seen = ref.den > 1e-6
    first = None
    for name, maps in layouts.items():
        assert (maps[0].stride(1) == 1) == (name == "planar")
        bp = gwbp.BackProjector(*args, kernel="tc")
        for v in range(2):
            bp.add_view(vm[v], K, W, H, maps[v])
        if first is None:  # tensor cores vs CUDA cores: split-bf16 contraction, threshold flips aside
            first = bp
            err = (bp.num[seen] - ref.num[seen]).norm(dim=1) / ref.num[seen].norm(dim=1).clamp_min(1e-6)
            assert float(err.max()) < 2e-5, float(err.max())  # identical weights (common.cuh): no flips, no outliers
        else:  # same packed operand whatever the input layout: only the order of the atomic adds differs
            assert torch.allclose(bp.den, first.den, rtol=1e-5, atol=1e-9), name
            assert torch.allclose(bp.num, first.num, rtol=1e-4, atol=1e-5), name



def test_rasterization_autograd_reproduces_reference_loop(gwbp, coracle, noracle, case):
    """The reference's own code shape (backproject.py:62-72,115-151) on our `rasterization`."""
    sc, vm, K, feats = case
    rasterization = gwbp.rasterization
    means, quats, scales, opac = _dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities)
    Kt = _dev(K)
    n = sc.n
    gaussian_features = torch.zeros(n, 8, device="cuda")
    gaussian_denoms = torch.ones(n, device="cuda") * 1e-12
    colors_feats = torch.zeros(n, 8, device="cuda", requires_grad=True)
    colors_feats_0 = torch.zeros(n, 3, device="cuda", requires_grad=True)
    for v in range(vm.shape[0]):
        viewmat = _dev(vm[v])
        f = _feat_dev(feats[v])
        out, _, meta = rasterization(means, quats, scales, opac, colors_feats, viewmat[None], Kt[None], width=96, height=64)
        assert float(out.abs().max()) == 0.0  # zero colours render zero
        target = (out[0] * f).sum()
        target.backward()
        copy = colors_feats.grad.clone()
        colors_feats.grad.zero_()
        out0, _, _ = rasterization(means, quats, scales, opac, colors_feats_0, viewmat[None], Kt[None], width=96, height=64)
        out0[0].sum().backward()
        gaussian_features += copy
        gaussian_denoms += colors_feats_0.grad[:, 0]
        g = colors_feats_0.grad
        assert torch.equal(g[:, 0], g[:, 1]) or torch.allclose(g[:, 0], g[:, 2], rtol=1e-6)  # channel independence
        colors_feats_0.grad.zero_()
    f_ref_loop = gaussian_features / gaussian_denoms[..., None]
    f_ref_loop = f_ref_loop / f_ref_loop.norm(dim=-1, keepdim=True)
    f_ref_loop[torch.isnan(f_ref_loop)] = 0
    num_o, den_o, _ = oracle_job(coracle, sc, vm, K, 96, 64, feats, 8)
    f_o = noracle.finalize(num_o, den_o + 1e-12)
    sel = den_o > 1e-6
    rel, _ = row_rel_err(f_ref_loop.double().cpu().numpy()[sel], f_o[sel])
    assert np.percentile(rel, 99.9) <= REL_TOL
    assert "means2d" in meta and "gaussian_ids" in meta  # affordance demo :392-395


def test_rasterization_strided_colors_and_float_sizes(gwbp, case):
    """utils.py:238-249: colors[:,0,:] slice of [N,16,3] with grad; width/height as 0-dim CUDA floats."""
    sc, vm, K, _ = case
    means, quats, scales, opac = _dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities)
    colors = torch.rand(sc.n, 16, 3, device="cuda", requires_grad=True)
    Kt = _dev(K)
    out, alphas, _ = gwbp.rasterization(means, quats, scales, opac, colors[:, 0, :], viewmats=_dev(vm[0])[None],
                                        Ks=Kt[None], width=Kt[0, 2] * 2, height=Kt[1, 2] * 2)
    assert out.shape == (1, 64, 96, 3) and alphas.shape == (1, 64, 96, 1)
    pseudo_loss = ((out.detach() + 1 - out) ** 2).mean()
    pseudo_loss.backward()
    assert colors.grad.shape == (sc.n, 16, 3)
    assert float(colors.grad[:, 1:].abs().max()) == 0.0 and float(colors.grad[:, 0].abs().max()) > 0.0


def test_forward_render_and_masks(gwbp, coracle, noracle, case):
    sc, vm, K, feats = case
    num_o, den_o, _ = oracle_job(coracle, sc, vm, K, 96, 64, feats, 8)
    f_o = noracle.finalize(num_o, den_o + 1e-12)
    text = gwbp.scene.make_text_queries(3, 8, 0)
    m_o, score_o = noracle.mask3d(f_o, text, 1)
    f_dev = _dev(f_o.astype(np.float32))
    m_gpu, _ = gwbp.get_mask3d(f_dev, _dev(text), 1)
    margin = np.abs(score_o[:, 0] - score_o[:, 1:].max(1))
    differ = (m_gpu.cpu().numpy() != m_o)
    assert not differ[margin > 1e-5].any(), "3-D mask differs away from ties"
    # threshold variant (segment.py:56-57)
    m_thr, _ = gwbp.get_mask3d(f_dev, _dev(text), 1, threshold=0.1)
    m_thr_o, _ = noracle.mask3d(f_o, text, 1, threshold=0.1)
    assert ((m_thr.cpu().numpy() != m_thr_o) & (np.abs(score_o[:, 0] - 0.1) > 1e-5) & (margin > 1e-5)).sum() == 0
    # forward feature render (segment.py:209-220)
    scene = gwbp.PackedScene(_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities))
    render, alpha = gwbp.render_features(scene, f_dev, vm[1], K, 96, 64)
    cv = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[1], K, 96, 64)
    r_o, a_o = cv.render(f_o.astype(np.float32))
    err = np.abs(render.double().cpu().numpy() - r_o)
    assert np.percentile(err, 99.9) < 1e-5 and np.abs(alpha.double().cpu().numpy() - a_o).max() < 1e-3
    # 2-D mask, both evaluation orders
    m2_o, s2_o = noracle.mask2d(r_o, text, 1)
    margin2 = np.abs(s2_o[..., 0] - s2_o[..., 1:].max(-1))
    covered = a_o > 1e-3
    for exact in (True, False):
        m2 = gwbp.render_mask_2d(scene, f_dev, _dev(text), 1, vm[1], K, 96, 64, exact_render=exact).cpu().numpy()
        bad = (m2 != m2_o) & (margin2 > 1e-4) & covered
        assert bad.sum() == 0, (exact, int(bad.sum()))


def test_edge_cases_gpu(gwbp):
    S = gwbp.scene
    vm, K = S.make_cameras(1, 40, 24, 0)
    # empty scene
    z = torch.zeros(0, 3, device="cuda")
    bp = gwbp.BackProjector(z, torch.zeros(0, 4, device="cuda"), z, torch.zeros(0, device="cuda"), 4)
    v = bp.add_view(vm[0], K, 40, 24, torch.ones(24, 40, 4, device="cuda"))
    assert v.n_vis == 0 and bp.finalize().shape == (0, 4)
    # nothing visible (all behind the camera) + degenerate inputs
    means = torch.tensor([[100.0, 100.0, 100.0], [0, 0, 0.1], [0, 0, 0.2], [0, 0, 0]], device="cuda")
    quats = torch.tensor([[1.0, 0, 0, 0], [0, 0, 0, 0], [2, 0, 0, 0], [1, 0, 0, 0]], device="cuda")
    scales = torch.tensor([[0.1, 0.1, 0.1], [0.1, 0.1, 0.1], [0, 0, 0], [0.1, 0.1, 0.1]], device="cuda")
    opac = torch.tensor([0.9, 0.9, 0.0, 0.9], device="cuda")
    bp = gwbp.BackProjector(means, quats, scales, opac, 2)
    bp.add_view(vm[0], K, 40, 24, torch.ones(24, 40, 2, device="cuda"))
    assert torch.isfinite(bp.num).all() and bp.den[3] > 1e-6 and float(bp.den[2]) <= 1.1e-12
    f = bp.finalize()
    assert torch.isfinite(f).all() and float(f[0].abs().sum()) == 0.0  # NaN -> 0
    # capacity growth: start with a tiny intersection capacity
    sc = S.make_scene(4000, 2)
    vm, K = S.make_cameras(1, 128, 128, 2)
    bp = gwbp.BackProjector(_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities), 4, cap_isects=16)
    v = bp.add_view(vm[0], K, 128, 128, torch.ones(128, 128, 4, device="cuda"))
    assert v.n_isects > 16 and bp.cap >= v.n_isects
    # errors: CPU tensors are refused loudly
    with pytest.raises(RuntimeError):
        gwbp.rasterization(means.cpu(), quats.cpu(), scales.cpu(), opac.cpu(), torch.zeros(4, 3),
                           torch.eye(4)[None], torch.eye(3)[None], 8, 8)


def test_known_answers_gpu(gwbp):
    """The hand-derived known-answer cases of tests/test_oracle.py::test_known_answers_from_published_semantics,
    through the C ABI: radius 7 and a 2 x 2 tile rectangle, alpha = o at the centre pixel, front-to-back
    compositing 1 - (1 - 0.5)(1 - 0.8), the 0.999 clamp + T <= 1e-4 stop rule, and the alpha < 1/255 skip."""
    W = H = 32
    K = np.array([[32, 0, 16.5], [0, 32, 16.5], [0, 0, 1]], np.float32)
    vm = np.eye(4, dtype=np.float32)

    def run(opacities, depths):
        n = len(opacities)
        z = np.asarray(depths, np.float32)
        means = np.stack([np.zeros(n, np.float32), np.zeros(n, np.float32), z], 1)
        quats = np.tile(np.array([[1, 0, 0, 0]], np.float32), (n, 1))
        scales = (0.25 * z / 4.0)[:, None].repeat(3, 1).astype(np.float32)
        bp = gwbp.BackProjector(_dev(means), _dev(quats), _dev(scales), _dev(np.asarray(opacities, np.float32)), 1,
                                kernel="simt", tile_cull=False)
        view = bp.add_view(vm, K, W, H, torch.ones(H, W, 1, device="cuda"))
        _, alpha = view.render(torch.ones(n, 1, device="cuda"))
        return view, alpha.cpu().numpy(), (bp.den - 1e-12).cpu().numpy()

    view, a, den = run([0.5], [4.0])
    m = view.meta()
    assert m["radii"].cpu().tolist() == [7] and view.n_isects == 4
    assert np.allclose(m["means2d"].cpu().numpy()[0], [16.5, 16.5], atol=1e-6)
    assert np.allclose(m["conics"].cpu().numpy()[0], [1 / 4.3, 0.0, 1 / 4.3], atol=1e-6)
    assert sorted((m["isect_ids"] >> 32).cpu().tolist()) == [0, 1, 2, 3]
    assert abs(a[16, 16] - 0.5) < 1e-6 and abs(a[16, 17] - 0.5 * np.exp(-0.5 / 4.3)) < 1e-6
    assert abs(den[0] - 0.5 * 2 * np.pi * 4.3 * (1 - 1 / 127.5)) < 0.02 * den[0]
    # off-axis beyond the frustum clamp (lim_x+ = 0.634375): radius 29, conic from the clamped Jacobian
    bp = gwbp.BackProjector(_dev(np.array([[4.0, 0.0, 4.0]], np.float32)), _dev(np.array([[1, 0, 0, 0]], np.float32)),
                            _dev(np.ones((1, 3), np.float32)), _dev(np.array([0.5], np.float32)), 1, kernel="simt",
                            tile_cull=False)
    m = bp.add_view(vm, K, W, H, torch.ones(H, W, 1, device="cuda")).meta()
    assert m["radii"].cpu().tolist() == [29]
    assert np.allclose(m["conics"].cpu().numpy()[0], [1 / (64.0 * (1.0 + 0.634375 ** 2) + 0.3), 0.0, 1 / 64.3],
                       rtol=2e-6, atol=1e-9)
    _, a, _ = run([0.8, 0.5], [5.0, 4.0])
    assert abs(a[16, 16] - 0.9) < 1e-6
    _, a, _ = run([1.0, 1.0, 1.0], [4.0, 5.0, 6.0])
    assert abs(a[16, 16] - 0.999) < 1e-6
    _, a, den = run([0.00393, 0.00391], [4.0, 5.0])
    assert abs(den[0] - 0.00393) < 1e-7 and den[1] == 0.0
    assert abs(a[16, 16] - 0.00393) < 1e-7 and a[16, 17] == 0.0


def test_full_size_properties_config_G_shape(gwbp):
    """Size-independent properties at the benchmark's image size (1297x840, D=512) on a lighter
    scene: sum(den_v) == sum(alpha_v), constant features -> num == den*c, adjointness."""
    S = gwbp.scene
    W, H, d = 1297, 840, 512
    sc = S.make_scene(300_000, 11)
    vm, K = S.make_cameras(1, W, H, 11)
    means, quats, scales, opac = _dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities)
    c = torch.linspace(-1, 1, d, device="cuda")
    F = c.expand(H, W, d)
    for kernel in ("simt", "auto"):
        bp = gwbp.BackProjector(means, quats, scales, opac, d, kernel=kernel)
        view = bp.add_view(vm[0], K, W, H, F.contiguous())
        den = bp.den - 1e-12
        _, alpha = view.render(torch.ones(sc.n, 1, device="cuda"))
        assert abs(float(den.double().sum()) - float(alpha.double().sum())) <= 2e-4 * float(alpha.double().sum())
        seen = den > 1e-5
        ratio = bp.num[seen] / den[seen, None]
        assert float((ratio - c[None]).abs().max()) < 2e-3
        X = torch.randn(sc.n, 4, device="cuda")
        G = torch.randn(H, W, 4, device="cuda")
        r, _ = view.render(X)
        num4 = torch.zeros(sc.n, 4, device="cuda")
        den4 = torch.zeros(sc.n, device="cuda")
        view.backproject(G, num4, den4, gwbp.KERNEL_SIMT)
        lhs, rhs = float((r.double() * G.double()).sum()), float((X.double() * num4.double()).sum())
        assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs), float((r.double() * G.double()).abs().sum()) * 1e-2)


def test_full_size_config_G_properties(gwbp):
    """BASELINE config[1] at FULL size (5.8 M Gaussians, 1297x840, D=512, tcgen05 path), checked through
    size-independent properties: sum(den_v) == sum(alpha_v); constant features => num == den * c;
    linearity in F; tile culling keeps every contributing row."""
    S = gwbp.scene
    cfg = S.CONFIGS["G"]
    W, H, d = cfg["width"], cfg["height"], cfg["d"]
    sc = S.make_scene(cfg["n"], 0)
    vm, K = S.make_cameras(cfg["views"], W, H, 0)
    args = (_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities), d)
    c = torch.linspace(-1, 1, d, device="cuda")
    bp = gwbp.BackProjector(*args, kernel="tc", collect_stats=True)
    view = bp.add_view(vm[7], K, W, H, c.expand(H, W, d))
    den = bp.den - 1e-12
    _, alpha = view.render(torch.ones(sc.n, 1, device="cuda"))
    tot_a, tot_d = float(alpha.double().sum()), float(den.double().sum())
    assert abs(tot_a - tot_d) <= 2e-4 * tot_a, (tot_a, tot_d)
    seen = den > 1e-5
    assert int(seen.sum()) > 50_000
    assert float((bp.num[seen] / den[seen, None] - c[None]).abs().max()) < 2e-3
    st = bp.stats()
    # exact (gsplat) tile list gives the same live rows
    bp2 = gwbp.BackProjector(*args, kernel="tc", collect_stats=True, tile_cull=False)
    v2 = bp2.add_view(vm[7], K, W, H, c.expand(H, W, d))
    assert bp2.stats()["rows_nonzero"] == st["rows_nonzero"] and v2.n_isects > view.n_isects
    assert torch.allclose(bp2.den, bp.den, rtol=1e-4, atol=1e-9)
    # linearity: F -> 2F doubles num, leaves den
    bp3 = gwbp.BackProjector(*args, kernel="tc")
    bp3.add_view(vm[7], K, W, H, (2 * c).expand(H, W, d))
    assert torch.allclose(bp3.num[seen], 2 * bp.num[seen], rtol=2e-4, atol=1e-6)
    del bp2, bp3
    # forward render on tensor cores at full size.  (a) constant colours render to alpha * c;
    # (b) adjointness with the back-projection: <render(X), G> == <X, backproject(G)>  (the two tcgen05 kernels
    # are transposes of each other: segment.py:209-220 vs backproject.py:127-131)
    r_c, a_c = view.render(c.expand(sc.n, d).contiguous(), None, gwbp.KERNEL_TC)
    da = (a_c - alpha).abs()  # ex2-based vs __expf-based alpha: equal up to threshold flips (alpha ~ 1/255, T ~ 1e-4)
    assert float((da > 1e-5).float().mean()) < 1e-4 and float(da.max()) < 2e-2, (float(da.max()), int((da > 1e-5).sum()))
    assert float((r_c - a_c[..., None] * c).abs().max()) < 2e-4
    del r_c
    g = torch.Generator(device="cuda").manual_seed(3)
    X = torch.randn(sc.n, d, device="cuda", generator=g)
    G = torch.randn(H, W, d, device="cuda", generator=g)
    r, _ = view.render(X, None, gwbp.KERNEL_TC)
    lhs = float((r.double() * G.double()).sum())
    scale = float(torch.linalg.vector_norm(r.double()) * torch.linalg.vector_norm(G.double()))
    del r
    bp.reset()
    bp.add_view(vm[7], K, W, H, G)
    rhs = float((X.double() * bp.num.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * scale, (lhs, rhs, scale)


@pytest.mark.parametrize("config,view", [("G", 7), ("M", 500)])
def test_full_size_view_against_the_oracle(gwbp, coracle, noracle, config, view):
    """ONE view of BASELINE config[1] (5.8 M Gaussians, 1297x840, D = 512) and of config[4] (6 M, 1920x1080, D = 768)
    at FULL size, tcgen05 path with the default tile culling, against the C oracle on the same inputs: bit-exact
    intersection counts (gsplat-exact list), identical prune mask, and the rel-err / cosine / den bars on every row
    with den > 1e-6 (rows above a bar must be proven threshold flips)."""
    S = gwbp.scene
    cfg = S.CONFIGS[config]
    W, H, d = cfg["width"], cfg["height"], cfg["d"]
    sc = S.make_scene(cfg["n"], 0)
    vm, K = S.make_cameras(cfg["views"], W, H, 0)
    F = S.make_feature_map_torch(view, d, H, W, "cuda", 0, enc_res=240)  # permuted view of a planar [D,H,W] buffer
    bp = gwbp.BackProjector(_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities), d, kernel="tc",
                            collect_stats=True)
    v_gpu = bp.add_view(vm[view], K, W, H, F)
    F_host = F.permute(2, 0, 1).contiguous().cpu().numpy().transpose(1, 2, 0)
    del F
    cv = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[view], K, W, H)
    cvc = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[view], K, W, H, cull=True)
    assert (v_gpu.n_vis, v_gpu.n_isects) == (cvc.n_vis, cvc.n_isects)  # the culled list, bit-exact counts
    cvc.close()
    num_o = np.zeros((sc.n, d), np.float64)  # calloc: only touched rows are committed
    den_o = np.zeros(sc.n, np.float64)
    st = cv.backproject(F_host, num_o, den_o)
    margin = np.full(sc.n, np.inf, np.float32)
    cv.margins(den_o, margin)
    cv.close()
    assert abs(bp.stats()["rows_nonzero"] - st["rows_nonzero"]) <= max(4, 1e-4 * st["rows_nonzero"])
    den = bp.den.double().cpu().numpy() - 1e-12
    mask_diff = (den > 5e-13) != (den_o > 0)
    idx = np.nonzero(den_o > 1e-6)[0]
    f_ref = noracle.finalize(num_o[idx], den_o[idx] + 1e-12)
    del num_o
    f_gpu = bp.finalize()[torch.from_numpy(idx).cuda()].double().cpu().numpy()
    _check_rows(f_gpu, den[idx], f_ref, den_o[idx], margin[idx], mask_diff, margin[mask_diff], f" config {config} full size")


def test_full_size_config_M_properties(gwbp):
    """BASELINE config[4] at FULL size on one GPU (6 M Gaussians, 1920x1080, D = 768: three column chunks on the
    tcgen05 path), through size-independent properties: sum(den_v) == sum(alpha_v); constant features =>
    num == den * c; and view-shard additivity -- two ranks' (num, den) add up to the single-process result,
    which is all the closing all-reduce of the multi-GPU job relies on (dist.py)."""
    S = gwbp.scene
    cfg = S.CONFIGS["M"]
    W, H, d = cfg["width"], cfg["height"], cfg["d"]
    sc = S.make_scene(cfg["n"], 0)
    vm, K = S.make_cameras(cfg["views"], W, H, 0)
    args = (_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities), d)
    c = torch.linspace(-1, 1, d, device="cuda")
    planar = c[:, None, None].expand(d, H, W).contiguous()  # the reference's layout: [D,H,W] viewed as [H,W,D]
    F = planar.permute(1, 2, 0)
    both = gwbp.BackProjector(*args, kernel="tc")
    both.add_view(vm[0], K, W, H, F)
    view = both.add_view(vm[500], K, W, H, F)
    r0, r1 = gwbp.BackProjector(*args, kernel="tc"), gwbp.BackProjector(*args, kernel="tc")
    r0.add_view(vm[0], K, W, H, F)
    v1 = r1.add_view(vm[500], K, W, H, F)
    den1 = r1.den - 1e-12
    _, alpha = v1.render(torch.ones(sc.n, 1, device="cuda"))
    tot_a, tot_d = float(alpha.double().sum()), float(den1.double().sum())
    assert abs(tot_a - tot_d) <= 2e-4 * tot_a, (tot_a, tot_d)
    seen = den1 > 1e-5
    assert int(seen.sum()) > 50_000
    assert float((r1.num[seen] / den1[seen, None] - c[None]).abs().max()) < 2e-3
    den_sum = r0.den + r1.den - 1e-12  # the initial 1e-12 is counted once (dist.allreduce_accumulators)
    assert torch.allclose(den_sum, both.den, rtol=1e-5, atol=1e-9)
    r0.num += r1.num
    assert torch.allclose(r0.num, both.num, rtol=1e-4, atol=1e-6)
    assert view.n_isects == v1.n_isects



def test_rasterization_backgrounds_depth_modes_and_sh(gwbp, coracle, noracle, case, tmp_path):
    """The remaining kwargs the reference passes to `rasterization`: backgrounds (affordance demo :918),
    render_mode="RGB+D" (click_and_segment.py:251) / "RGB+ED", sh_degree=3 (backproject.py:99)."""
    sc, vm, K, _ = case
    means, quats, scales, opac = _dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities)
    rng = np.random.default_rng(5)
    cols = rng.uniform(0, 1, (sc.n, 3)).astype(np.float32)
    cv = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[0], K, 96, 64)
    r_o, a_o = cv.render(cols)
    bg = torch.tensor([[0.2, 0.4, 0.6]], device="cuda")
    out, alphas, _ = gwbp.rasterization(means, quats, scales, opac, _dev(cols), _dev(vm[0])[None], _dev(K)[None], 96, 64,
                                        backgrounds=bg)
    want = r_o + (1.0 - a_o)[..., None] * np.array([0.2, 0.4, 0.6])
    assert np.abs(out[0].double().cpu().numpy() - want).max() < 2e-4
    assert np.abs(alphas[0, ..., 0].double().cpu().numpy() - a_o).max() < 2e-4
    # depth channel = composited camera-space z; ED divides by alpha
    z = (sc.means @ vm[0][2, :3] + vm[0][2, 3]).astype(np.float32)
    rz_o, _ = cv.render(z[:, None])
    out_d, _, _ = gwbp.rasterization(means, quats, scales, opac, _dev(cols), _dev(vm[0])[None], _dev(K)[None], 96, 64,
                                     render_mode="RGB+D")
    assert out_d.shape == (1, 64, 96, 4)
    assert np.abs(out_d[0, ..., 3].double().cpu().numpy() - rz_o[..., 0]).max() < 1e-3
    out_ed, _, _ = gwbp.rasterization(means, quats, scales, opac, _dev(cols), _dev(vm[0])[None], _dev(K)[None], 96, 64,
                                      render_mode="RGB+ED")
    want_ed = rz_o[..., 0] / np.maximum(a_o, 1e-10)
    cover = a_o > 0.05
    assert np.abs(out_ed[0, ..., 3].double().cpu().numpy() - want_ed)[cover].max() < 1e-2
    # SH colours, degree 0..4 (the reference uses 3: backproject.py:99), against the oracle's independent basis
    # (associated Legendre functions, oracle/gsplat_oracle.py::sh_basis) composited by the C oracle
    sh_np = rng.standard_normal((sc.n, 25, 3)).astype(np.float32)  # DC ~ N(0,1): the clamp at 0 bites at every degree
    for degree in (0, 1, 2, 3, 4):
        kk = (degree + 1) ** 2
        coeffs = _dev(sh_np)[:, :max(kk, 16)] if degree < 4 else _dev(sh_np)
        out_sh, _, _ = gwbp.rasterization(means, quats, scales, opac, coeffs, _dev(vm[0])[None], _dev(K)[None], 96, 64,
                                          sh_degree=degree)
        cols_o = noracle.sh_colors(degree, sc.means, sh_np, vm[0])
        assert (cols_o == 0).any() and (cols_o > 0).any()  # the clamp at 0 is exercised
        cols_gpu = gwbp.sh_colors(degree, means, coeffs, _dev(vm[0])).double().cpu().numpy()
        assert np.abs(cols_gpu - cols_o).max() < 5e-6 * max(1.0, np.abs(cols_o).max()), degree
        r_sh, _ = cv.render(cols_o.astype(np.float32))
        assert np.abs(out_sh[0].double().cpu().numpy() - r_sh).max() < 2e-4, degree
    # feature-field file in the reference's format
    bp = gwbp.BackProjector(means, quats, scales, opac, 8)
    for v in range(vm.shape[0]):
        bp.add_view(vm[v], K, 96, 64, torch.rand(64, 96, 8, device="cuda"))
    saved = bp.save(str(tmp_path / "features_lseg.pt"))
    loaded = torch.load(str(tmp_path / "features_lseg.pt"))
    kept = torch.load(str(tmp_path / "features_lseg.pt.kept.pt"))
    assert loaded.shape == (int(bp.prune_mask().sum()), 8) and loaded.dtype == torch.float32
    assert torch.equal(loaded.cpu(), saved.cpu()) and kept.numel() == loaded.shape[0]
    assert torch.allclose(loaded.norm(dim=1), torch.ones(loaded.shape[0], device=loaded.device), atol=1e-4)
    # rows for an externally pruned checkpoint (the mask prune_by_gradients returned, utils.py:257-268)
    ext = torch.zeros(sc.n, dtype=torch.bool)
    ext[::3] = True
    saved_ext = bp.save(str(tmp_path / "features_ext.pt"), keep=ext)
    assert saved_ext.shape == (int(ext.sum()), 8)
    assert torch.equal(torch.load(str(tmp_path / "features_ext.pt.kept.pt")), torch.nonzero(ext).flatten())
    assert torch.equal(saved_ext.cpu(), bp.finalize()[ext.cuda()].cpu())


def test_probe_pixel_render_and_click_prompt(gwbp, coracle, noracle, case):
    """gwbp_render_pixels == the full RGB+D render read at the clicked pixels (click_and_segment.py:241-275)."""
    sc, vm, K, _ = case
    rng = np.random.default_rng(5)
    feats = rng.standard_normal((sc.n, 24)).astype(np.float32)
    feats /= np.linalg.norm(feats, axis=1, keepdims=True)
    z = (sc.means @ vm[0][2, :3] + vm[0][2, 3]).astype(np.float32)
    cv = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[0], K, 96, 64)
    r_o, a_o = cv.render(np.concatenate([feats, z[:, None]], 1))
    xy = np.array([[0, 0], [95, 63], [48, 32], [17, 40], [80, 5], [33, 33], [-1, 3], [96, 10]], np.int32)
    scene = gwbp.PackedScene(_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities))
    for cull in (False, True):
        view = gwbp.View(scene, gwbp.make_camera(vm[0], K, 96, 64), tile_cull=cull)
        out, alpha = view.render_pixels(_dev(feats), _dev(xy), extra=_dev(z))
        out, alpha = out.double().cpu().numpy(), alpha.double().cpu().numpy()
        assert out.shape == (8, 25)
        for i, (x, y) in enumerate(xy):
            if 0 <= x < 96 and 0 <= y < 64:
                assert np.abs(out[i] - r_o[y, x]).max() < 1e-5 * max(1.0, np.abs(r_o[y, x]).max()), (i, cull)
                assert abs(alpha[i] - a_o[y, x]) < 1e-5
            else:
                assert np.abs(out[i]).max() == 0.0 and alpha[i] == 0.0
        # against our own full-frame render: same arithmetic
        full, _ = view.render(torch.cat([_dev(feats), _dev(z)[:, None]], 1))
        full = full.double().cpu().numpy()
        for i, (x, y) in enumerate(xy[:6]):
            assert np.abs(out[i] - full[y, x]).max() < 1e-6 * max(1.0, np.abs(full[y, x]).max())
    # no `extra`: D columns
    out2, _ = view.render_pixels(_dev(feats), _dev(xy[:3]))
    assert out2.shape == (3, 24) and np.allclose(out2.cpu().numpy(), out[:3, :24], atol=1e-6)
    # click prompt: normalised feature + un-projected world point, then the 3-D mask compare (:317-321)
    covered = [i for i in range(6) if a_o[xy[i, 1], xy[i, 0]] > 0.5]
    assert len(covered) >= 2
    prompt, world, _ = gwbp.click_prompt(scene, _dev(feats), vm[0], K, 96, 64, xy[covered])
    for j, i in enumerate(covered):
        p_o, w_o = noracle.click_prompt(r_o, vm[0], K, xy[i])
        assert np.abs(prompt[j].double().cpu().numpy() - p_o).max() < 1e-5
        assert np.abs(world[j].double().cpu().numpy() - w_o).max() < 1e-3 * max(1.0, np.abs(w_o).max())
    m = gwbp.click_mask3d(_dev(feats), prompt[:1], prompt[1:]).cpu().numpy()
    s = feats.astype(np.float64) @ prompt.double().cpu().numpy().T
    m_o = s[:, 0] > s[:, 1:].max(1)
    assert ((m != m_o) & (np.abs(s[:, 0] - s[:, 1:].max(1)) > 1e-5)).sum() == 0


@pytest.mark.parametrize("kernel,d", [("simt", 8), ("tc", 32)])
def test_per_view_ratio_accumulation(gwbp, coracle, noracle, kernel, d):
    """accumulate="per_view_ratio": features += grad/(grad0 + 1e-12) per view
    (affordance_transfer/demo_affordance_transfer.py:768-800)."""
    sc, vm, K, feats = small_case(gwbp.scene, n=3000, views=3, d=d, seed=4)
    per_view = []
    for v in range(3):
        num = np.zeros((sc.n, d), np.float64)
        den = np.zeros(sc.n, np.float64)
        cv = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[v], K, 96, 64)
        cv.backproject(np.asarray(feats[v]), num, den)
        per_view.append((num, den))
    acc_o, f_o = noracle.ratio_backproject(per_view, 96, 64, d)
    bp = gwbp.BackProjector(_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities), d, kernel=kernel,
                            accumulate="per_view_ratio")
    for v in range(3):
        bp.add_view(vm[v], K, 96, 64, _feat_dev(feats[v]))
    assert float(bp.num_v.abs().max()) == 0.0 and float(bp.den_v.abs().max()) == 0.0  # scratch left clean
    den_total = sum(dv for _, dv in per_view)
    assert np.array_equal(bp.prune_mask().cpu().numpy(), den_total > 0)
    f = bp.finalize().double().cpu().numpy()
    # rows whose per-view den is comparable to eps*(H*W*3) are dominated by the epsilon: keep well-seen rows
    sel = np.all([(dv == 0) | (dv > 1e-4) for _, dv in per_view], axis=0) & (den_total > 1e-4)
    rel, _ = row_rel_err(f[sel], f_o[sel])
    assert sel.sum() > 100 and np.percentile(rel, 99.9) <= REL_TOL, (sel.sum(), np.percentile(rel, 99.9))
    assert np.abs(f[den_total == 0]).max() == 0.0  # never-seen rows: 0 (the reference leaves NaN)


@pytest.mark.parametrize("d,W,H", [(64, 96, 64), (128, 96, 64), (512, 100, 70), (100, 41, 37), (768, 64, 48)])
def test_tcgen05_forward_render(gwbp, coracle, d, W, H):
    """The tcgen05 forward render against the CPU oracle and the fp32 CUDA-core kernel (segment.py:209-220)."""
    sc = gwbp.scene.make_scene(3000, 3)
    vm, K = gwbp.scene.make_cameras(2, W, H, 3)
    rng = np.random.default_rng(d)
    feats = rng.standard_normal((sc.n, d)).astype(np.float32)
    feats /= np.linalg.norm(feats, axis=1, keepdims=True)
    bg = rng.uniform(0, 1, d).astype(np.float32)
    scene = gwbp.PackedScene(_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities))
    for v, cull in ((0, False), (1, True)):
        view = gwbp.View(scene, gwbp.make_camera(vm[v], K, W, H), tile_cull=cull)
        r_tc, a_tc = view.render(_dev(feats), None, gwbp.KERNEL_TC)
        r_si, a_si = view.render(_dev(feats), None, gwbp.KERNEL_SIMT)
        r_o, a_o = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[v], K, W, H).render(feats)
        scale = np.abs(r_o).max()
        err = np.abs(r_tc.double().cpu().numpy() - r_o)
        assert np.percentile(err, 99.9) < 1e-5 * max(1.0, scale) and err.max() < 1e-2, (err.max(), np.percentile(err, 99.9))
        assert np.abs(a_tc.double().cpu().numpy() - a_o).max() < 1e-3
        assert float((r_tc - r_si).abs().max()) < 2e-5 * max(1.0, scale)   # same weights, split-bf16 vs fp32 contraction
        assert float((a_tc - a_si).abs().max()) < 1e-5  # ex2-based vs __expf-based alpha
        # background term: render += T * bg
        r_bg, _ = view.render(_dev(feats), _dev(bg), gwbp.KERNEL_TC)
        want = r_tc + (1.0 - a_tc)[..., None] * _dev(bg)
        assert float((r_bg - want).abs().max()) < 1e-5
    # AUTO picks the tensor-core kernel for wide features and still renders RGB (D = 3) on CUDA cores
    rgb, _ = view.render(_dev(feats[:, :3].copy()))
    assert rgb.shape == (H, W, 3)


def test_pipelined_host_upload_matches_device_path(gwbp, case):
    """add_view_host (copy stream + two staging buffers + deferred accumulation) == add_view on device tensors."""
    sc, vm, K, feats = case
    ref = _gpu_job(gwbp, sc, vm, K, 96, 64, feats, 8, kernel="simt")
    bp = gwbp.BackProjector(_dev(sc.means), _dev(sc.quats), _dev(sc.scales), _dev(sc.opacities), 8, kernel="simt",
                            collect_stats=True)
    for rep in range(2):  # 4 uploads through 2 staging buffers
        for v in range(vm.shape[0]):
            planar = torch.from_numpy(np.ascontiguousarray(np.transpose(feats[v], (2, 0, 1)))).pin_memory()
            bp.add_view_host(vm[v], K, 96, 64, planar)
    num, den = bp.raw()  # flushes the deferred view
    assert bp.n_views == 2 * vm.shape[0]
    assert torch.allclose(num, 2 * ref.num, rtol=1e-5, atol=1e-6)
    assert torch.allclose(den - 1e-12, 2 * (ref.den - 1e-12), rtol=1e-5, atol=1e-9)
    assert bp.stats()["entries_walked"] == 2 * ref.stats()["entries_walked"]
