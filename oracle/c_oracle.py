"""ctypes binding of oracle/oracle.c (CPU ORACLE -- test infrastructure, see oracle.c header).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_view_create.restype = C.c_void_p
        L.orc_view_create.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_int]
        L.orc_view_destroy.argtypes = [C.c_void_p]
        L.orc_view_counts.argtypes = [C.c_void_p] * 5
        L.orc_view_export.argtypes = [C.c_void_p] * 9
        L.orc_view_backproject.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_view_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_covar.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_view_margins.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    """OpenMP team size of the oracle (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib().orc_set_num_threads(int(n))


class View:
    """One camera's projected + binned + sorted scene."""

    def __init__(self, means, quats, scales, opacities, viewmat, K, width, height,
                 near_plane=0.01, far_plane=1e10, radius_clip=0.0, eps2d=0.3, cull=False):
        self._keep = [_f32(means), _f32(quats), _f32(scales), _f32(opacities), _f32(viewmat), _f32(K)]
        self.n = self._keep[0].shape[0]
        self.width, self.height = int(width), int(height)
        self._h = lib().orc_view_create(self.n, *[_p(a) for a in self._keep], self.width, self.height,
                                        near_plane, far_plane, radius_clip, eps2d, int(bool(cull)))
        nv, ni, tw, th = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
        lib().orc_view_counts(self._h, C.byref(nv), C.byref(ni), C.byref(tw), C.byref(th))
        self.n_vis, self.n_isects, self.tile_width, self.tile_height = nv.value, ni.value, tw.value, th.value

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_view_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    def export(self):
        n, nv, ni = self.n, self.n_vis, self.n_isects
        out = dict(radii=np.empty(n, np.int32), means2d=np.empty((n, 2), np.float32), depths=np.empty(n, np.float32),
                   conics=np.empty((n, 3), np.float32), gaussian_ids=np.empty(nv, np.int32),
                   isect_ids=np.empty(ni, np.int64), flatten_ids=np.empty(ni, np.int32),
                   isect_offsets=np.empty((self.tile_height, self.tile_width), np.int32))
        lib().orc_view_export(self._h, *[_p(out[k]) for k in ("radii", "means2d", "depths", "conics", "gaussian_ids",
                                                              "isect_ids", "flatten_ids", "isect_offsets")])
        return out

    def backproject(self, feats, num, den):
        """num [N,D] fp64 += sum_p w F ; den [N] fp64 += sum_p w.  feats: [H,W,D] fp32, any strides."""
        assert feats.dtype == np.float32 and feats.shape[:2] == (self.height, self.width)
        assert num.dtype == np.float64 and num.flags.c_contiguous and den.dtype == np.float64
        sH, sW, sD = (s // 4 for s in feats.strides)
        stats = np.zeros(4, np.int64)
        lib().orc_view_backproject(self._h, _p(feats), sH, sW, sD, feats.shape[2], _p(num), _p(den), _p(stats))
        return dict(rows_nonzero=int(stats[0]), pairs=int(stats[1]), entries_walked=int(stats[2]),
                    n_vis=self.n_vis, n_isects=self.n_isects)

    def margins(self, den_total, margin, frac=1e-3):
        """min-update margin [N] fp32 (caller initialises to +inf) with every row's smallest relative distance to a
        compositing threshold (alpha = 1/255, T(1-alpha) = 1e-4) in this view; see oracle.c::orc_view_margins."""
        assert margin.dtype == np.float32 and margin.shape == (self.n,) and margin.flags.c_contiguous
        den_total = np.ascontiguousarray(den_total, dtype=np.float64)
        assert den_total.shape == (self.n,)
        lib().orc_view_margins(self._h, _p(den_total), float(frac), _p(margin))

    def render(self, colors):
        colors = _f32(colors)
        d = colors.shape[1]
        out = np.zeros((self.height, self.width, d), np.float64)
        alpha = np.zeros((self.height, self.width), np.float64)
        lib().orc_view_render(self._h, _p(colors), d, _p(out), _p(alpha))
        return out, alpha


def covar(quats, scales):
    q, s = _f32(quats), _f32(scales)
    out = np.empty((q.shape[0], 6), np.float32)
    lib().orc_covar(q.shape[0], _p(q), _p(s), _p(out))
    return out
