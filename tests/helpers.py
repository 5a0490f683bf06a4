"""Shared test helpers: seeded scenes + oracle drivers (tests are the only place besides
bench.py's cpu_baseline and smoke() that may touch oracle/)."""
import numpy as np


def small_case(scene_mod, n=3000, views=2, width=96, height=64, d=8, seed=1):
    sc = scene_mod.make_scene(n, seed)
    vm, K = scene_mod.make_cameras(views, width, height, seed)
    feats = [scene_mod.make_feature_map_np(v, d, height, width, seed, enc_res=12) for v in range(views)]
    return sc, vm, K, feats


def oracle_job(coracle, sc, vm, K, width, height, feats, d):
    """(num fp64 [N,d], den fp64 [N] without the 1e-12 initialiser, per-view stats)"""
    num = np.zeros((sc.n, d), np.float64)
    den = np.zeros(sc.n, np.float64)
    stats = []
    for v in range(vm.shape[0]):
        view = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[v], K, width, height)
        stats.append(view.backproject(np.asarray(feats[v]), num, den))
        view.close()
    return num, den, stats


def row_rel_err(a, b):
    """per-row ||a-b|| / ||b|| over rows where b != 0"""
    nb = np.linalg.norm(b, axis=1)
    ok = nb > 0
    return np.linalg.norm(a[ok] - b[ok], axis=1) / nb[ok], ok


def row_cosine(a, b):
    na, nb = np.linalg.norm(a, axis=1), np.linalg.norm(b, axis=1)
    ok = (na > 0) & (nb > 0)
    return (a[ok] * b[ok]).sum(1) / (na[ok] * nb[ok]), ok


def oracle_margins(coracle, sc, vm, K, width, height, den_total, views=None, frac=1e-3):
    """Per-Gaussian smallest relative distance to a compositing threshold (alpha = 1/255, T(1-alpha) = 1e-4)
    over the given views (oracle.c::orc_view_margins); +inf for rows no test ever touched."""
    margin = np.full(sc.n, np.inf, np.float32)
    for v in (range(vm.shape[0]) if views is None else views):
        view = coracle.View(sc.means, sc.quats, sc.scales, sc.opacities, vm[v], K, width, height)
        view.margins(den_total, margin, frac)
        view.close()
    return margin
