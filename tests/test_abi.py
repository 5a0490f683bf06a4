"""CPU tests of the drop-in boundary: lib/libgwbp.so loads without a GPU, exports exactly the
symbols include/gwbp.h declares, validates arguments, and the Python mirror keeps gsplat's
signature.  No compute call is made here."""
import ctypes
import inspect
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gwbp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gwbp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(gwbp):
    lib = gwbp._lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 11
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in gwbp.h but not exported by libgwbp.so"
        assert name in gwbp._lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(gwbp._lib.SIGNATURES) == declared


def test_layout_and_argument_validation_without_gpu(gwbp):
    L = gwbp._lib
    lib = L.lib()
    lay = L.WsLayout()
    assert lib.gwbp_workspace_layout(1000, 64, 48, 5000, ctypes.byref(lay)) == 0
    offs = [lay.cnt, lay.scan, lay.rec, lay.mask, lay.grec, lay.erec, lay.radii, lay.tiles_per_gauss, lay.dkeys0,
            lay.dkeys1, lay.dvals0, lay.dvals1, lay.cnt2, lay.base2, lay.tkeys0, lay.tkeys1, lay.tvals0, lay.tvals1,
            lay.offsets, lay.stats, lay.bin_counts, lay.bin_seg, lay.bin_tot, lay.sort_tmp]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs) and lay.total >= lay.sort_tmp + lay.sort_tmp_bytes
    assert lib.gwbp_workspace_layout(-5, 64, 48, 10, ctypes.byref(lay)) < 0
    assert "n out of range" in L.last_error()
    assert lib.gwbp_workspace_layout(10, 0, 48, 10, ctypes.byref(lay)) < 0
    assert lib.gwbp_finalize(None, None, None, 5, 4, None) < 0 and "NULL" in L.last_error()
    assert lib.gwbp_fpack_bytes(0, 10, 64) == 0
    with pytest.raises(RuntimeError):
        L.check(-1, "x")
    with pytest.raises(L.CapacityError):
        L.check(-2, "x")


def test_struct_sizes_match_header(gwbp):
    L = gwbp._lib
    assert ctypes.sizeof(L.Scene) == 16
    assert ctypes.sizeof(L.Camera) == 16 * 4 + 9 * 4 + 2 * 4 + 4 * 4
    assert ctypes.sizeof(L.ViewInfo) == 64
    assert ctypes.sizeof(L.WsLayout) == 29 * ctypes.sizeof(ctypes.c_size_t)


def test_rasterization_signature_matches_gsplat(gwbp):
    """Positional order used at backproject.py:89-100 and keywords used at utils.py:238-249."""
    p = list(inspect.signature(gwbp.rasterization).parameters)
    assert p[:9] == ["means", "quats", "scales", "opacities", "colors", "viewmats", "Ks", "width", "height"]
    for kw in ("near_plane", "far_plane", "radius_clip", "eps2d", "sh_degree", "packed", "tile_size", "backgrounds",
               "render_mode", "sparse_grad", "absgrad", "rasterize_mode", "channel_chunk", "distributed",
               "camera_model", "covars"):
        assert kw in p, kw


def test_no_cpu_fallback(gwbp):
    import torch
    with pytest.raises(RuntimeError, match="CUDA"):
        gwbp.PackedScene(torch.zeros(1, 3), torch.zeros(1, 4), torch.zeros(1, 3), torch.zeros(1), device="cpu")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            gwbp.rasterization(torch.zeros(1, 3), torch.zeros(1, 4), torch.zeros(1, 3), torch.zeros(1),
                               torch.zeros(1, 3), torch.eye(4)[None], torch.eye(3)[None], 8, 8)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "3dgs-gradient-backprojection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f
