"""ctypes binding of lib/libgwbp.so (the C ABI declared in include/gwbp.h).

There is NO fallback: if the CUDA library is missing or fails to load, importing the compute
API raises.  (The oracle under oracle/ is test infrastructure and is never imported here.)
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgwbp.so")
if os.environ.get("GWBP_LIB_VARIANT") == "exp":  # tools/ only: the -DGWBP_EXPERIMENTS build (timing knobs)
    LIB_PATH = os.path.join(_HERE, "lib", "libgwbp_exp.so")

KERNEL_AUTO, KERNEL_SIMT, KERNEL_TC = 0, 1, 2
ABI_VERSION = 12
KERNEL_FPACK_READY = 0x100
PREPARE_GSPLAT_EXACT, PREPARE_TILE_CULL, PREPARE_COUNTING_BIN, PREPARE_SUPERTILE = 0, 1, 2, 4


class Scene(C.Structure):
    _fields_ = [("n", C.c_int64), ("geo", C.c_void_p)]


class Camera(C.Structure):
    _fields_ = [("viewmat", C.c_float * 16), ("K", C.c_float * 9), ("width", C.c_int32), ("height", C.c_int32),
                ("near_plane", C.c_float), ("far_plane", C.c_float), ("radius_clip", C.c_float), ("eps2d", C.c_float)]


class WsLayout(C.Structure):
    _fields_ = [(k, C.c_size_t) for k in ("total", "cnt", "scan", "rec", "mask", "grec", "erec", "radii",
                                          "tiles_per_gauss", "dkeys0", "dkeys1", "dvals0", "dvals1", "cnt2", "base2",
                                          "tkeys0", "tkeys1", "tvals0", "tvals1", "offsets", "stats", "bin_counts", "bin_seg",
                                          "bin_tot", "spg", "svals", "front", "sort_tmp",
                                          "sort_tmp_bytes")]


class ViewInfo(C.Structure):
    _fields_ = [("n_vis", C.c_int64), ("n_isects", C.c_int64), ("cap_isects", C.c_int64), ("tile_w", C.c_int32),
                ("tile_h", C.c_int32), ("sorted_buf", C.c_int32), ("tile_key_bytes", C.c_int32),
                ("list_kind", C.c_int32), ("super_w", C.c_int32), ("super_h", C.c_int32), ("reserved", C.c_int32),
                ("n_entries", C.c_int64)]


# name -> (restype, argtypes); kept in one table so tests can check it against include/gwbp.h
SIGNATURES = {
    "gwbp_abi_version": (C.c_int, []),
    "gwbp_launch_count": (C.c_ulonglong, []),
    "gwbp_last_error": (C.c_char_p, []),
    "gwbp_workspace_layout": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.POINTER(WsLayout)]),
    "gwbp_pack_scene": (C.c_int, [C.c_int64] + [C.c_void_p] * 6),
    "gwbp_view_prepare": (C.c_int, [C.POINTER(Scene), C.POINTER(Camera), C.c_void_p, C.c_size_t, C.c_int64,
                                    C.c_int32, C.c_void_p, C.POINTER(ViewInfo)]),
    "gwbp_profile_enable": (C.c_int, [C.c_int]),
    "gwbp_profile_read": (C.c_int, [C.POINTER(C.c_float), C.c_int]),
    "gwbp_debug_set_trace": (C.c_int, [C.c_void_p, C.c_size_t]),
    "gwbp_fpack_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "gwbp_pack_features": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                     C.c_void_p, C.c_void_p]),
    "gwbp_pack_features_lowres": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                            C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "gwbp_backproject_view": (C.c_int, [C.POINTER(Scene), C.POINTER(Camera), C.c_void_p, C.POINTER(ViewInfo),
                                        C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gwbp_lowres_adjoint_supported": (C.c_int, [C.c_int32] * 6),
    "gwbp_pack_lowres_adjoint": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                           C.c_void_p, C.c_void_p]),
    "gwbp_backproject_view_lowres": (C.c_int, [C.POINTER(Scene), C.POINTER(Camera), C.c_void_p, C.POINTER(ViewInfo),
                                               C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64,
                                               C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p]),
    "gwbp_render_view": (C.c_int, [C.POINTER(Scene), C.POINTER(Camera), C.c_void_p, C.POINTER(ViewInfo), C.c_void_p,
                                   C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "gwbp_render_pixels": (C.c_int, [C.POINTER(Scene), C.POINTER(Camera), C.c_void_p, C.POINTER(ViewInfo), C.c_void_p,
                                     C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "gwbp_ratio_accumulate": (C.c_int, [C.POINTER(Scene), C.POINTER(Camera), C.c_void_p, C.POINTER(ViewInfo),
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_float,
                                        C.c_float, C.c_void_p]),
    "gwbp_sh_colors": (C.c_int, [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "gwbp_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "gwbp_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "gwbp_ipc_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "gwbp_ipc_close": (C.c_int, [C.c_void_p]),
    "gwbp_peer_reduce_supported": (C.c_int, [C.c_int32, C.c_int32]),
    "gwbp_peer_reduce_finalize": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_int64,
                                            C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gwbp_mask3d": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_float,
                              C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gwbp_mask2d": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                              C.c_void_p]),
}

IPC_HANDLE_BYTES, MAX_PEERS = 64, 8
LOWRES_PACKED = 2

PROFILE_STAGES = ("project", "count_scan_and_readback", "compact", "depth_sort", "tile_binning", "feature_relayout",
                  "backproject")

_lib = None


def lib():
    """Load libgwbp.so (once).  Raises RuntimeError -- never falls back to anything."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc); there is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if L.gwbp_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libgwbp.so ABI {L.gwbp_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
        _lib = L
    return _lib


def last_error() -> str:
    return lib().gwbp_last_error().decode("utf-8", "replace")


class CapacityError(RuntimeError):
    """The view has more tile intersections than the workspace was sized for."""


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = f"{what}: {last_error()} (rc={rc})"
    if rc == -2:
        raise CapacityError(msg)
    raise RuntimeError(msg)
