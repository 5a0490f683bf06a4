"""Drop-in for `gsplat.rasterization` on the path the reference uses it for.

Same name, argument order and return triple as gsplat-1.4.0, so that

    from gsplat import rasterization          # backproject.py:7, utils.py:5, segment.py:9

can be switched to this module (or to shims/gsplat on PYTHONPATH) without touching the scripts:

    render, alphas, meta = rasterization(means, quats, scales, opacities, colors,
                                         viewmats[C,4,4], Ks[C,3,3], width, height, ...)

Differentiable w.r.t. `colors` only -- which is all the reference needs: every geometry tensor it
passes is detached (utils.py:11-17,90) and the back-projection *is* d/d(colors)
(backproject.py:115-131).  The backward pass calls the fused back-projection kernel with the
upstream gradient as the feature map.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from .engine import PackedScene, View, fpack_bytes, make_camera
from .sh import sh_colors

_FPACK_CACHE = {}  # device -> scratch for the tcgen05 path's feature re-layout (reused across calls)

# The reference calls `rasterization` three times per view with the SAME geometry and camera (RGB render for the
# encoder, 512-channel pass, 3-channel pass: backproject.py:89,115,133), and gsplat projects + sorts three times.
# Here the packed scene and the prepared view of the last call are kept and reused when nothing changed: tensors are
# identified by (storage pointer, shape, in-place version counter), the camera by its bytes.
_SCENE_CACHE = {"key": None, "scene": None}
_VIEW_CACHE = {"key": None, "view": None}


def cache_clear() -> None:
    """Drop the cached scene / view (e.g. after modifying a geometry tensor through `.data`, which does not bump
    the version counter)."""
    _SCENE_CACHE.update(key=None, scene=None)
    _VIEW_CACHE.update(key=None, view=None)


def _tensor_key(t: torch.Tensor):
    return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype, t._version, t.device)


def _cached_scene(means, quats, scales, opacities) -> PackedScene:
    key = tuple(_tensor_key(t) for t in (means, quats, scales, opacities))
    if _SCENE_CACHE["key"] != key:
        cache_clear()
        _SCENE_CACHE.update(key=key, scene=PackedScene(means, quats, scales, opacities))
    return _SCENE_CACHE["scene"]


def _cached_view(scene: PackedScene, cam) -> View:
    key = (id(scene), bytes(cam))
    if _VIEW_CACHE["key"] != key:
        # a fresh workspace per camera: an autograd graph may still hold the previous view for its backward pass
        _VIEW_CACHE.update(key=None, view=None)
        _VIEW_CACHE.update(key=key, view=View(scene, cam))
    return _VIEW_CACHE["view"]


def _fpack_buffer(device, nbytes: int) -> torch.Tensor:
    buf = _FPACK_CACHE.get(device)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _FPACK_CACHE[device] = buf
    return buf


class _CompositeColors(torch.autograd.Function):
    """render = sum_g w(g,p) colors[g];  d(loss)/d(colors)[g] = sum_p w(g,p) v_render[p]."""

    @staticmethod
    def forward(ctx, colors: torch.Tensor, view: View, background):
        render, alpha = view.render(colors.detach(), background)
        ctx.view = view
        ctx.cshape = colors.shape
        ctx.mark_non_differentiable(alpha)
        return render, alpha

    @staticmethod
    def backward(ctx, v_render, _v_alpha):
        view: View = ctx.view
        n, d = ctx.cshape
        num = torch.zeros(n, d, dtype=torch.float32, device=v_render.device)
        den = torch.zeros(n, dtype=torch.float32, device=v_render.device)
        if view.n_isects:
            # v_render arrives with whatever strides autograd produced (often an expanded /
            # permuted view, backproject.py:113,127); the kernel takes strides as they are,
            # except for broadcast (stride-0) dims, which must be materialised.
            g = v_render.to(torch.float32)
            if 0 in g.stride():
                g = g.contiguous()
            need = fpack_bytes(view.cam.width, view.cam.height, d)  # > 0 iff the tensor-core kernel takes this D
            view.backproject(g, num, den, L.KERNEL_AUTO, _fpack_buffer(g.device, need) if need else None)
        return num, None, None


def rasterization(
    means: torch.Tensor,
    quats: torch.Tensor,
    scales: torch.Tensor,
    opacities: torch.Tensor,
    colors: torch.Tensor,
    viewmats: torch.Tensor,
    Ks: torch.Tensor,
    width,
    height,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = True,
    tile_size: int = 16,
    backgrounds: Optional[torch.Tensor] = None,
    render_mode: str = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    rasterize_mode: str = "classic",
    channel_chunk: int = 32,
    distributed: bool = False,
    camera_model: str = "pinhole",
    covars: Optional[torch.Tensor] = None,
):
    """See module docstring.  Returns (render [C,H,W,X], alphas [C,H,W,1], meta)."""
    if not torch.cuda.is_available() or not means.is_cuda:
        raise RuntimeError("CUDA is required for rasterization (there is no CPU path)")  # backproject.py:314-315
    n = means.shape[0]
    assert means.shape == (n, 3), means.shape
    assert quats.shape == (n, 4), quats.shape
    assert scales.shape == (n, 3), scales.shape
    assert opacities.shape == (n,), opacities.shape
    assert viewmats.dim() == 3 and viewmats.shape[1:] == (4, 4), viewmats.shape
    c = viewmats.shape[0]
    assert Ks.shape == (c, 3, 3), Ks.shape
    assert render_mode in ("RGB", "D", "ED", "RGB+D", "RGB+ED"), render_mode
    if tile_size != 16:
        raise NotImplementedError("tile_size must be 16 (the reference never changes it)")
    if rasterize_mode != "classic" or distributed or camera_model != "pinhole" or covars is not None or absgrad:
        raise NotImplementedError("rasterize_mode/antialiased, distributed, non-pinhole cameras, covars and absgrad "
                                  "are used only by the out-of-scope f3dgs trainer")
    width, height = int(width), int(height)

    scene = _cached_scene(means, quats, scales, opacities)
    vm_host = viewmats.detach().to("cpu", torch.float32)
    k_host = Ks.detach().to("cpu", torch.float32)

    renders, alphas, metas = [], [], []
    for ci in range(c):
        cam = make_camera(vm_host[ci], k_host[ci], width, height, near_plane, far_plane, radius_clip, eps2d)
        view = _cached_view(scene, cam)
        if sh_degree is not None:
            # colors: [N,K,3] or [C,N,K,3] SH coefficients -> view-dependent RGB (backproject.py:88-100)
            coeffs = colors[ci] if colors.dim() == 4 else colors
            assert coeffs.dim() == 3 and coeffs.shape[0] == n and coeffs.shape[2] == 3, coeffs.shape
            assert (sh_degree + 1) ** 2 <= coeffs.shape[1], (sh_degree, coeffs.shape)
            cols = sh_colors(sh_degree, means, coeffs, vm_host[ci])
        else:
            cols = colors[ci] if colors.dim() == 3 else colors
            assert cols.dim() == 2 and cols.shape[0] == n, f"colors must be [N,D] or [C,N,D], got {tuple(colors.shape)}"
        cols = cols.to(torch.float32)
        bg = None
        if backgrounds is not None:
            bg = backgrounds[ci] if backgrounds.dim() == 2 else backgrounds
        if render_mode != "RGB":
            # depth channel (click_and_segment.py:241-254 uses RGB+D): camera-space z per Gaussian
            z = (means.detach() @ vm_host[ci, 2, :3].to(means.device) + vm_host[ci, 2, 3].to(means.device))[:, None]
            cols = z if render_mode in ("D", "ED") else torch.cat([cols, z], dim=1)
            if bg is not None:
                bg = torch.cat([bg.reshape(-1), bg.new_zeros(1)]) if render_mode in ("RGB+D", "RGB+ED") else bg.new_zeros(1)
        render, alpha = _CompositeColors.apply(cols, view, bg)
        if render_mode in ("ED", "RGB+ED"):
            render = torch.cat([render[..., :-1], render[..., -1:] / alpha[..., None].clamp(min=1e-10)], dim=-1)
        renders.append(render)
        alphas.append(alpha[..., None])
        metas.append(view.meta())
    meta = metas[0] if c == 1 else _merge_meta(metas, n)
    meta["n_cameras"] = c
    return torch.stack(renders, 0), torch.stack(alphas, 0), meta


def _merge_meta(metas, n):
    out = dict(metas[-1])
    for key in ("gaussian_ids", "radii", "means2d", "depths", "conics", "opacities", "tiles_per_gauss"):
        out[key] = torch.cat([m[key] for m in metas], 0)
    out["camera_ids"] = torch.cat([torch.full_like(m["camera_ids"], i) for i, m in enumerate(metas)], 0)
    out["isect_offsets"] = torch.cat([m["isect_offsets"] for m in metas], 0)
    # per-camera sorted lists are kept separately (the camera bits of gsplat's keys are not rebuilt)
    out["isect_ids"] = [m["isect_ids"] for m in metas]
    out["flatten_ids"] = [m["flatten_ids"] for m in metas]
    return out
