"""CPU tests: the oracle against itself (numpy restatement vs threaded C restatement), against the
algebraic invariants of SURVEY.md §4, and against the committed golden fixtures.

There are no reference-owned golden vectors for this path (gsplat is un-vendored, the repo has no
tests: SURVEY.md §8c) -- parity is "unpinned"; these tests pin the two restatements to each other and
to the invariants the reference itself asserts (utils.py:353-355, affordance demo :384-386)."""
import os

import numpy as np
import pytest

from helpers import oracle_job, small_case

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "small_case.npz")


@pytest.fixture(scope="module")
def case(gwbp):
    return small_case(gwbp.scene)


def test_numpy_and_c_oracle_agree_bit_exact_on_integer_stages(case, noracle, coracle):
    sc, vm, K, _ = case
    for v in range(vm.shape[0]):
        proj, isect = noracle.view_geometry(sc.means, sc.quats, sc.scales, vm[v], K, 96, 64)
        cv = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[v], K, 96, 64)
        e = cv.export()
        assert np.array_equal(e["radii"], proj["radii"])
        assert np.array_equal(e["depths"].view(np.int32), proj["depths"].view(np.int32))
        vis = proj["radii"] > 0
        assert np.array_equal(e["means2d"][vis].view(np.int32), proj["means2d"][vis].view(np.int32))
        assert np.array_equal(e["conics"][vis].view(np.int32), proj["conics"][vis].view(np.int32))
        for k in ("gaussian_ids", "isect_ids", "flatten_ids", "isect_offsets"):
            assert np.array_equal(e[k], isect[k]), k
        assert np.array_equal(coracle.covar(sc.quats, sc.scales).view(np.int32),
                              noracle.quat_scale_to_covar(sc.quats, sc.scales).view(np.int32))


def test_tile_culling_extension(case, noracle, coracle):
    """The exact tile-culling extension: both restatements agree bit for bit, the culled list is a
    sub-list of gsplat's, and no (pixel, Gaussian) pair with non-zero weight is lost."""
    sc, vm, K, feats = case
    for v in range(vm.shape[0]):
        _, full = noracle.view_geometry(sc.means, sc.quats, sc.scales, vm[v], K, 96, 64)
        _, cut = noracle.view_geometry(sc.means, sc.quats, sc.scales, vm[v], K, 96, 64, sc.opacities, True)
        cv = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[v], K, 96, 64, cull=True)
        e = cv.export()
        for k in ("isect_ids", "flatten_ids", "isect_offsets"):
            assert np.array_equal(e[k], cut[k]), k
        assert cut["n_isects"] < full["n_isects"]
        pairs_full = set(zip(full["isect_ids"].tolist(), full["flatten_ids"].tolist()))
        assert set(zip(cut["isect_ids"].tolist(), cut["flatten_ids"].tolist())) <= pairs_full
        a0 = np.zeros((sc.n, 8)); d0 = np.zeros(sc.n)
        a1 = np.zeros((sc.n, 8)); d1 = np.zeros(sc.n)
        s0 = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[v], K, 96, 64).backproject(np.asarray(feats[v]), a0, d0)
        s1 = cv.backproject(np.asarray(feats[v]), a1, d1)
        assert s0["rows_nonzero"] == s1["rows_nonzero"] and s0["pairs"] == s1["pairs"]
        assert np.allclose(a0, a1, rtol=1e-12, atol=1e-14) and np.allclose(d0, d1, rtol=1e-12, atol=1e-14)
    x = np.array([1.0001, 1.5, 2.0, 3.7, 100.0, 254.9], np.float32)
    assert np.abs(noracle.ln_approx(x) - np.log(x.astype(np.float64))).max() < 2e-5


def test_numpy_and_c_oracle_agree_on_accumulators(case, noracle, coracle):
    sc, vm, K, feats = case
    num_c, den_c, _ = oracle_job(coracle, sc, vm, K, 96, 64, feats, 8)
    num_n = np.zeros_like(num_c)
    den_n = np.zeros_like(den_c)
    for v in range(vm.shape[0]):
        a, b = noracle.backproject_view(sc.means, sc.quats, sc.scales, sc.opacities, vm[v], K, 96, 64, feats[v])
        num_n += a
        den_n += b
    # expf (glibc) vs np.exp differ in the last ulp of fp32 -> ~1e-7 relative
    assert np.abs(num_c - num_n).max() <= 2e-6 * max(1.0, np.abs(num_n).max())
    assert np.abs(den_c - den_n).max() <= 2e-6 * max(1.0, den_n.max())
    assert np.array_equal(den_c > 0, den_n > 0)


def test_fp32_mirror_close_to_fp64_truth(case, noracle):
    sc, vm, K, feats = case
    a32, d32 = noracle.backproject_view(sc.means, sc.quats, sc.scales, sc.opacities, vm[0], K, 96, 64, feats[0])
    a64, d64 = noracle.backproject_view(sc.means, sc.quats, sc.scales, sc.opacities, vm[0], K, 96, 64, feats[0],
                                        dtype=np.float64)
    # threshold flips (alpha ~ 1/255, T ~ 1e-4) are possible but rare; bulk error is fp32 rounding
    rel = np.abs(d32 - d64) / np.maximum(d64, 1e-6)
    assert np.percentile(rel, 99) < 1e-5


def test_invariant_sum_den_equals_sum_alpha(case, noracle, coracle):
    """sum_g w(g,p) = 1 - T_final(p)  =>  sum_g den_v[g] = sum_p alpha_v(p)."""
    sc, vm, K, feats = case
    _, den, _ = oracle_job(coracle, sc, vm[:1], K, 96, 64, feats[:1], 8)
    _, alpha = noracle.render_view(sc.means, sc.quats, sc.scales, sc.opacities, np.ones((sc.n, 1), np.float32),
                                   vm[0], K, 96, 64)
    assert abs(den.sum() - alpha.sum()) <= 1e-5 * alpha.sum()
    cv = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[0], K, 96, 64)
    _, alpha_c = cv.render(np.ones((sc.n, 1), np.float32))
    assert abs(alpha_c.sum() - alpha.sum()) <= 1e-5 * alpha.sum()


def test_invariant_constant_features(case, noracle, coracle):
    """F == c  =>  num[g,:] = den[g] * c  =>  finalised row = c/||c||."""
    sc, vm, K, _ = case
    c = np.array([0.3, -1.0, 2.0, 0.5], np.float32)
    F = np.broadcast_to(c, (64, 96, 4)).copy()
    num, den, _ = oracle_job(coracle, sc, vm[:1], K, 96, 64, [F], 4)
    assert np.allclose(num, den[:, None] * c[None, :].astype(np.float64), rtol=1e-12, atol=1e-12)
    f = noracle.finalize(num, den + 1e-12)
    seen = den > 0
    assert np.allclose(f[seen], (c / np.linalg.norm(c))[None, :], atol=1e-6)
    assert np.all(f[~seen] == 0)  # NaN -> 0 branch (backproject.py:169)


def test_invariant_adjoint_of_forward_render(case, noracle, coracle):
    """<render(X), F> == <X, backproject(F)> -- the back-projection IS the
    gradient of the render w.r.t. colours (backproject.py:115-131)."""
    sc, vm, K, feats = case
    rng = np.random.default_rng(0)
    X = rng.standard_normal((sc.n, 8)).astype(np.float32)
    cv = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[1], K, 96, 64)
    render, _ = cv.render(X)
    num = np.zeros((sc.n, 8))
    den = np.zeros(sc.n)
    cv.backproject(np.asarray(feats[1]), num, den)
    lhs = float((render * np.asarray(feats[1], np.float64)).sum())
    rhs = float((X.astype(np.float64) * num).sum())
    assert abs(lhs - rhs) <= 1e-9 * max(1.0, abs(lhs))


def test_invariant_linearity_and_channel_independence(case, coracle):
    """SURVEY §4 inv.4 (the reference's own assert: affordance demo :384-386)."""
    sc, vm, K, feats = case
    F = np.ascontiguousarray(feats[0])
    G = np.ascontiguousarray(feats[1])
    cv = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[0], K, 96, 64)

    def bp(feat):
        num = np.zeros((sc.n, feat.shape[2]))
        den = np.zeros(sc.n)
        cv.backproject(feat, num, den)
        return num, den

    nf, _ = bp(F)
    ng, _ = bp(G)
    nfg, _ = bp((2.0 * F + G).astype(np.float32))
    assert np.allclose(nfg, 2.0 * nf + ng, rtol=1e-5, atol=1e-6)
    ones, den = bp(np.ones((64, 96, 3), np.float32))
    assert np.allclose(ones[:, 0], ones[:, 2]) and np.allclose(ones[:, 0], den)


def test_view_sharding_sums_to_single_job(case, coracle):
    """SURVEY §4 inv.5."""
    sc, vm, K, feats = case
    num, den, _ = oracle_job(coracle, sc, vm, K, 96, 64, feats, 8)
    n0, d0, _ = oracle_job(coracle, sc, vm[0::2], K, 96, 64, feats[0::2], 8)
    n1, d1, _ = oracle_job(coracle, sc, vm[1::2], K, 96, 64, feats[1::2], 8)
    assert np.allclose(num, n0 + n1, rtol=1e-12, atol=1e-12) and np.allclose(den, d0 + d1, rtol=1e-12, atol=1e-12)


def test_pruned_gaussians_do_not_change_the_render(case, noracle, coracle):
    """The reference's own self-check (utils.py:292-360): dropping den==0 Gaussians changes no pixel."""
    sc, vm, K, feats = case
    _, den, _ = oracle_job(coracle, sc, vm, K, 96, 64, feats, 8)
    keep = noracle.prune_mask(den)
    assert 0 < keep.sum() < sc.n
    rng = np.random.default_rng(1)
    cols = rng.uniform(0, 1, (sc.n, 3)).astype(np.float32)
    for v in range(vm.shape[0]):
        full, _ = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[v], K, 96, 64).render(cols)
        pruned, _ = coracle.View(sc.means[keep], sc.quats[keep], sc.scales[keep], sc.opacities[keep], vm[v], K, 96,
                                 64).render(cols[keep])
        assert np.abs(full - pruned).max() < 1.0 / (255 * 2)  # utils.py:353-355


def test_per_view_ratio_and_click_prompt_restatements(case, noracle, coracle):
    """8f row 4.  (a) with ONE view the per-view ratio is a positive per-row rescale of num/den, so its
    normalised rows equal finalize(num, den) wherever den >> eps*(H*W*3)
    (demo_affordance_transfer.py:768-800 vs backproject.py:166-169).  (b) the un-projected click point
    re-projects onto the clicked pixel at the rendered depth (click_and_segment.py:254-269)."""
    sc, vm, K, feats = case
    num = np.zeros((sc.n, 8), np.float64)
    den = np.zeros(sc.n, np.float64)
    cv = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[0], K, 96, 64)
    cv.backproject(np.asarray(feats[0]), num, den)
    _, f_ratio = noracle.ratio_backproject([(num, den)], 96, 64, 8)
    f_sum = noracle.finalize(num, den + 1e-12)
    sel = den > 1e-3
    assert sel.sum() > 100 and np.abs(f_ratio[sel] - f_sum[sel]).max() < 1e-9
    assert np.isnan(f_ratio[den == 0]).all()  # the reference leaves never-seen rows NaN in this mode
    z = (sc.means @ vm[0][2, :3] + vm[0][2, 3]).astype(np.float32)
    rgbd, alpha = cv.render(np.concatenate([sc.means.astype(np.float32), z[:, None]], 1))
    ys, xs = np.nonzero(alpha > 0.9)
    x, y = int(xs[len(xs) // 2]), int(ys[len(ys) // 2])
    prompt, world = noracle.click_prompt(rgbd, vm[0], K, (x, y))
    assert abs(np.linalg.norm(prompt) - 1.0) < 1e-12
    cam = vm[0].astype(np.float64) @ np.append(world, 1.0)
    assert abs(cam[2] - rgbd[y, x, -1]) < 1e-9
    assert abs(K[0, 0] * cam[0] / cam[2] + K[0, 2] - x) < 1e-6 and abs(K[1, 1] * cam[1] / cam[2] + K[1, 2] - y) < 1e-6


def _kat_scene(opacities, depths=None):
    """Isotropic Gaussians on the optical axis of an identity camera whose principal point is the centre of pixel
    (16,16): every quantity below has a closed form.  f = 32, z = 4, s = 0.25 => cov2d = (f s / z)^2 I = 4 I."""
    n = len(opacities)
    z = np.asarray(depths if depths is not None else [4.0] * n, np.float32)
    means = np.stack([np.zeros(n, np.float32), np.zeros(n, np.float32), z], 1)
    quats = np.tile(np.array([[1, 0, 0, 0]], np.float32), (n, 1))
    scales = (0.25 * z / 4.0)[:, None].repeat(3, 1).astype(np.float32)  # same 2-D footprint at every depth
    K = np.array([[32, 0, 16.5], [0, 32, 16.5], [0, 0, 1]], np.float32)
    return means, quats, scales, np.asarray(opacities, np.float32), np.eye(4, dtype=np.float32), K


def test_known_answers_from_published_semantics(noracle, coracle):
    """Hand-derived known-answer tests: the reference holds no golden vectors for this path (gsplat is un-vendored,
    SURVEY 8c), so both restatements are pinned to numbers worked out on paper from gsplat-1.4.0's published
    arithmetic: EWA blur 0.3, radius = ceil(3 sqrt(lambda_max)) with the 0.01 floor, tile rectangle,
    alpha = min(0.999, o exp(-sigma)), the alpha < 1/255 skip and the T(1-alpha) <= 1e-4 stop rule."""
    W = H = 32
    # -- projection + binning: cov2d = 4 I + 0.3 I; conic = 1/4.3; lambda = 4.3 + sqrt(max(0.01, 0)) = 4.4;
    #    radius = ceil(3 sqrt(4.4)) = ceil(6.2929) = 7; tiles: floor(1.03125 - 0.4375) = 0 .. ceil(1.46875) = 2 -> 2 x 2
    means, quats, scales, opac, vm, K = _kat_scene([0.5])
    proj, isect = noracle.view_geometry(means, quats, scales, vm, K, W, H)
    assert proj["radii"].tolist() == [7]
    assert np.allclose(proj["means2d"][0], [16.5, 16.5], atol=1e-6)
    assert np.allclose(proj["conics"][0], [1 / 4.3, 0.0, 1 / 4.3], atol=1e-6)
    assert np.allclose(proj["depths"], [4.0])
    assert int(isect["n_isects"]) == 4
    tiles = sorted(int(k) >> 32 for k in isect["isect_ids"])
    assert tiles == [0, 1, 2, 3]
    assert all((int(k) & 0xFFFFFFFF) == int(np.float32(4.0).view(np.int32)) for k in isect["isect_ids"])
    cv = coracle.View(means, quats, scales, opac, vm, K, W, H)
    e = cv.export()
    assert e["radii"].tolist() == [7] and np.array_equal(np.sort(e["isect_ids"]), np.sort(isect["isect_ids"]))

    # -- off-axis Gaussian beyond the frustum clamp: x/z = 1 > lim_x+ = (W - cx)/fx + 0.3 * (0.5 W/fx) = 0.634375, so the
    #    Jacobian uses tx = 0.634375 z: cov2d_xx = s^2 (fx/z)^2 (1 + 0.634375^2) + 0.3 = 90.0556, cov2d_yy = 64.3,
    #    radius = ceil(3 sqrt(90.0556)) = 29 (34 without the clamp); mean2d itself is not clamped: 32 * 1 + 16.5
    m1 = np.array([[4.0, 0.0, 4.0]], np.float32)
    p1, _ = noracle.view_geometry(m1, quats, np.ones((1, 3), np.float32), vm, K, W, H)
    xx = 64.0 * (1.0 + 0.634375 ** 2) + 0.3
    assert p1["radii"].tolist() == [29] and np.allclose(p1["means2d"][0], [48.5, 16.5], atol=1e-5)
    assert np.allclose(p1["conics"][0], [1 / xx, 0.0, 1 / 64.3], rtol=2e-6, atol=1e-9)
    c1 = coracle.View(m1, quats, np.ones((1, 3), np.float32), opac, vm, K, W, H).export()
    assert c1["radii"].tolist() == [29]

    def alpha_c(opacities, depths=None):
        m, q, s, o, v, k = _kat_scene(opacities, depths)
        ones = np.ones((len(opacities), 1), np.float32)
        a_np = noracle.render_view(m, q, s, o, ones, v, k, W, H)[1]
        view = coracle.View(m, q, s, o, v, k, W, H)
        a_c = view.render(ones)[1]
        num, den = np.zeros((len(opacities), 1)), np.zeros(len(opacities))
        view.backproject(np.ones((H, W, 1), np.float32), num, den)
        assert np.abs(a_np - a_c).max() < 2e-6
        return a_np, den

    # -- one Gaussian, sigma = 0 at pixel (16,16): alpha = o; one pixel to the right sigma = 0.5/4.3
    a, den = alpha_c([0.5])
    assert abs(a[16, 16] - 0.5) < 1e-6
    assert abs(a[16, 17] - 0.5 * np.exp(-0.5 / 4.3)) < 1e-6
    # den = sum over the footprint ~ o * 2 pi * 4.3 minus the tail below alpha = 1/255 (sigma > ln(127.5))
    assert abs(den[0] - 0.5 * 2 * np.pi * 4.3 * (1 - 1 / 127.5)) < 0.02 * den[0]
    # -- front-to-back compositing of two Gaussians: 1 - (1 - 0.5)(1 - 0.8) = 0.9, nearer one first
    a, den = alpha_c([0.8, 0.5], depths=[5.0, 4.0])
    assert abs(a[16, 16] - 0.9) < 1e-6
    # -- alpha is clamped to 0.999; after one such Gaussian T = 1e-3, the second would leave 1e-6 <= 1e-4:
    #    the pixel stops and the second Gaussian is NOT composited there
    a, _ = alpha_c([1.0, 1.0, 1.0], depths=[4.0, 5.0, 6.0])
    assert abs(a[16, 16] - 0.999) < 1e-6
    # -- alpha < 1/255 is skipped: o just above 1/255 contributes at the centre pixel only (next pixel:
    #    o exp(-0.116) < 1/255), o just below contributes nowhere and is pruned (utils.py:257)
    a, den = alpha_c([0.00393, 0.00391], depths=[4.0, 5.0])
    assert abs(den[0] - 0.00393) < 1e-8 and den[1] == 0.0
    assert abs(a[16, 16] - 0.00393) < 1e-7 and a[16, 17] == 0.0  # alpha = 1 - T, one fp32 rounding


def test_edge_cases(gwbp, noracle, coracle):
    S = gwbp.scene
    vm, K = S.make_cameras(1, 40, 24, 0)
    # empty scene
    e = np.zeros((0, 3), np.float32)
    cv = coracle.View(e, np.zeros((0, 4), np.float32), e, np.zeros(0, np.float32), vm[0], K, 40, 24)
    assert cv.n_vis == 0 and cv.n_isects == 0
    # behind the camera / zero scale / zero opacity / zero quaternion are legal inputs
    means = np.array([[0, 0, 0], [100, 100, 100], [0, 0, 0.1], [0, 0, 0.2]], np.float32)
    quats = np.array([[1, 0, 0, 0], [1, 0, 0, 0], [0, 0, 0, 0], [2, 0, 0, 0]], np.float32)
    scales = np.array([[0.1, 0.1, 0.1], [0.1, 0.1, 0.1], [0.1, 0.1, 0.1], [0, 0, 0]], np.float32)
    opac = np.array([0.9, 0.9, 0.9, 0.0], np.float32)
    proj, isect = noracle.view_geometry(means, quats, scales, vm[0], K, 40, 24)
    assert proj["radii"][0] > 0 and proj["radii"][2] == 0  # NaN quaternion is culled, not propagated
    cv = coracle.View(means, quats, scales, opac, vm[0], K, 40, 24)
    assert np.array_equal(cv.export()["radii"], proj["radii"])
    num = np.zeros((4, 2))
    den = np.zeros(4)
    cv.backproject(np.ones((24, 40, 2), np.float32), num, den)
    assert den[0] > 0 and den[3] == 0 and np.isfinite(num).all()


def test_golden_fixture(gwbp, coracle, noracle):
    """Committed fixture (tests/golden/make_golden.py): pins both restatements against drift."""
    g = np.load(GOLDEN)
    sc, vm, K, feats = small_case(gwbp.scene, **{k: int(g["cfg_" + k]) for k in ("n", "views", "width", "height", "d", "seed")})
    W, H, d = int(g["cfg_width"]), int(g["cfg_height"]), int(g["cfg_d"])
    cv = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[0], K, W, H)
    e = cv.export()
    assert np.array_equal(e["isect_ids"], g["isect_ids_v0"])
    assert np.array_equal(e["flatten_ids"], g["flatten_ids_v0"])
    assert np.array_equal(e["isect_offsets"], g["isect_offsets_v0"])
    assert np.array_equal(e["radii"], g["radii_v0"])
    num, den, _ = oracle_job(coracle, sc, vm, K, W, H, feats, d)
    assert np.allclose(num, g["num"], rtol=0, atol=2e-6 * np.abs(g["num"]).max())
    assert np.allclose(den, g["den"], rtol=0, atol=2e-6 * g["den"].max())
    f = noracle.finalize(num, den + 1e-12)
    assert np.abs(f - g["features"]).max() < 1e-5
    t = gwbp.scene.make_text_queries(3, d, 0)
    m, _ = noracle.mask3d(f, t, 1)
    assert (m != g["mask3d"]).sum() <= 1


def test_sh_basis_from_first_principles_matches_the_3dgs_tables(noracle):
    """oracle/gsplat_oracle.py::sh_basis (associated Legendre functions) against the real-SH polynomial tables of the
    3DGS code base that csrc/sh.cu hard-codes: degree 0..3 constants and signs, parity of the bands, and the colour
    rule max(SH + 0.5, 0) with dirs = mean - camera centre (backproject.py:88-100)."""
    rng = np.random.default_rng(0)
    d = rng.standard_normal((500, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    x, y, z = d.T
    B = noracle.sh_basis(4, d)
    C1, C2 = 0.4886025119029199, 1.0925484305920792
    want = {0: 0.28209479177387814 + 0 * x, 1: -C1 * y, 2: C1 * z, 3: -C1 * x, 4: C2 * x * y, 5: -C2 * y * z,
            6: 0.31539156525252005 * (2 * z * z - x * x - y * y), 7: -C2 * x * z, 8: 0.5462742152960396 * (x * x - y * y),
            9: -0.5900435899266435 * y * (3 * x * x - y * y), 10: 2.890611442640554 * x * y * z,
            15: -0.5900435899266435 * x * (x * x - 3 * y * y), 20: 0.10578554691520431 * (z * z * (35 * z * z - 30) + 3)}
    for k, v in want.items():
        assert np.abs(B[:, k] - v).max() < 1e-12, k
    Bm = noracle.sh_basis(4, -d)  # band l has parity (-1)^l
    for l in range(5):
        assert np.abs(Bm[:, l * l:(l + 1) ** 2] - (-1) ** l * B[:, l * l:(l + 1) ** 2]).max() < 1e-12
    # orthonormality on the sphere (Monte-Carlo, 4 pi / n weights)
    dd = rng.standard_normal((200_000, 3))
    G = noracle.sh_basis(3, dd)
    gram = 4 * np.pi * (G.T @ G) / dd.shape[0]
    assert np.abs(gram - np.eye(16)).max() < 0.03
    means = rng.standard_normal((50, 3)).astype(np.float32)
    coeffs = (rng.standard_normal((50, 16, 3)) * 0.3).astype(np.float32)
    vm = np.eye(4, dtype=np.float32)
    vm[:3, 3] = [0.1, -0.2, 4.0]
    c0 = noracle.sh_colors(0, means, coeffs, vm)
    assert np.allclose(c0, np.maximum(0.28209479177387814 * coeffs[:, 0] + 0.5, 0), atol=1e-7)
    c3 = noracle.sh_colors(3, means, coeffs, vm)
    assert c3.shape == (50, 3) and (c3 >= 0).all() and (c3 == 0).any()


def test_upsample_restates_torch_interpolate(noracle):
    """oracle/gsplat_oracle.py::upsample == torch.nn.functional.interpolate (CPU kernel), the call the reference makes
    on the encoder output (backproject.py:110-112 bilinear, :245-249 nearest)."""
    import torch

    rng = np.random.default_rng(1)
    for (h, w, H, W) in [(24, 31, 137, 211), (64, 64, 137, 211), (12, 12, 64, 96), (30, 40, 17, 15)]:
        low = rng.standard_normal((h, w, 5)).astype(np.float32)
        for mode in ("bilinear", "nearest"):
            ref = torch.nn.functional.interpolate(torch.from_numpy(low).permute(2, 0, 1)[None], size=(H, W), mode=mode)
            ref = ref[0].permute(1, 2, 0).numpy()
            up = noracle.upsample(low, H, W, mode)
            assert up.shape == (H, W, 5) and up.dtype == np.float32
            assert np.abs(up - ref).max() < (1e-5 if mode == "bilinear" else 0.0) + 1e-12, (h, w, H, W, mode)


def test_threshold_margins(coracle):
    """oracle.c::orc_view_margins: a Gaussian whose centre-pixel alpha sits exactly on 1/255 gets margin ~0, one far
    from every threshold gets a large margin, and a row behind a near-threshold alpha test inherits the small margin
    on the pixel that carries its weight."""
    W = H = 32
    K = np.array([[32, 0, 16.5], [0, 32, 16.5], [0, 0, 1]], np.float32)
    vm = np.eye(4, dtype=np.float32)

    def run(opacities, depths, scale=0.25):
        n = len(opacities)
        z = np.asarray(depths, np.float32)
        means = np.stack([np.zeros(n, np.float32), np.zeros(n, np.float32), z], 1)
        quats = np.tile(np.array([[1, 0, 0, 0]], np.float32), (n, 1))
        scales = (scale * z / 4.0)[:, None].repeat(3, 1).astype(np.float32)
        v = coracle.View(means, quats, scales, np.asarray(opacities, np.float32), vm, K, W, H)
        num, den = np.zeros((n, 1)), np.zeros(n)
        v.backproject(np.ones((H, W, 1), np.float32), num, den)
        m = np.full(n, np.inf, np.float32)
        v.margins(den, m)
        return m, den

    m, den = run([0.5], [4.0])
    assert den[0] > 1 and 1e-4 < m[0] < 1.0       # ordinary Gaussian: some ring pixel is within 1e-4..1 of alpha = 1/255
    on = np.float32(1.0 / 255.0)
    m, den = run([on], [4.0])
    assert m[0] < 1e-6                            # centre pixel: alpha == opacity == 1/255 exactly
    # a Gaussian exactly on the alpha threshold in front of an ordinary one: the back row inherits that margin through
    # the centre pixel, which carries ~3.7 % of its weight (a flip there would move it by 0.4 % x 3.7 % > 1e-4) ...
    m, den = run([on, 0.5], [4.0, 5.0])
    assert m[0] < 1e-6 and m[1] < 1e-6
    # ... but not when only pixels carrying at least half of the row's weight are allowed to pass a margin on
    z = np.array([4.0, 5.0], np.float32)
    means = np.stack([np.zeros(2, np.float32), np.zeros(2, np.float32), z], 1)
    quats = np.tile(np.array([[1, 0, 0, 0]], np.float32), (2, 1))
    scales = (0.25 * z / 4.0)[:, None].repeat(3, 1).astype(np.float32)
    v = coracle.View(means, quats, scales, np.array([on, 0.5], np.float32), vm, K, W, H)
    mm = np.full(2, np.inf, np.float32)
    v.margins(den, mm, frac=0.5)
    assert mm[0] < 1e-6 and 1e-4 < mm[1] < 1.0
