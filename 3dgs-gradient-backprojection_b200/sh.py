"""Spherical-harmonics colour evaluation (degree <= 4), the `sh_degree=3` branch of `rasterization` that
produces the RGB image fed to the 2-D encoder (backproject.py:88-100, segment.py:197-208).

gsplat-1.4.0 semantics: dirs = mean - camera_position (normalised inside the kernel),
colour = max(SH(dirs) + 0.5, 0).  The arithmetic is csrc/sh.cu behind `gwbp_sh_colors`; this file only
computes the camera centre on the host and hands over pointers and strides.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from .engine import _f32c, _host_floats, _require_cuda, _stream_ptr


def sh_colors(degree: int, means: torch.Tensor, coeffs: torch.Tensor, viewmat) -> torch.Tensor:
    """coeffs [N,K,3] (any strides, K >= (degree+1)^2), viewmat [4,4] world->camera -> colours [N,3] fp32."""
    _require_cuda(means, "means")
    _require_cuda(coeffs, "colors (SH coefficients)")
    if not 0 <= int(degree) <= 4:
        raise NotImplementedError("sh_degree must be 0..4 (the reference uses 3)")
    n = means.shape[0]
    assert coeffs.dim() == 3 and coeffs.shape[0] == n and coeffs.shape[2] == 3, coeffs.shape
    assert (degree + 1) ** 2 <= coeffs.shape[1], (degree, coeffs.shape)
    means = _f32c(means, means.device)
    coeffs = coeffs.detach()
    if coeffs.dtype != torch.float32:
        coeffs = coeffs.float()
    vm = _host_floats(viewmat, 16).reshape(4, 4).astype(np.float64)
    cam_pos = np.ascontiguousarray(-(vm[:3, :3].T @ vm[:3, 3]), dtype=np.float32)  # camtoworld[:3, 3]
    out = torch.empty(n, 3, dtype=torch.float32, device=means.device)
    sN, sK, sC = coeffs.stride()
    with torch.cuda.device(means.device):
        L.check(L.lib().gwbp_sh_colors(n, int(degree), means.data_ptr(), coeffs.data_ptr(), sN, sK, sC,
                                       cam_pos.ctypes.data, out.data_ptr(), _stream_ptr(means.device)), "gwbp_sh_colors")
    return out


eval_sh_colors = sh_colors  # earlier name
