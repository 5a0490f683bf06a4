"""Host-side plumbing over the C ABI: device buffers (torch), streams, workspaces.

PyTorch is used only for memory, streams and dtype views; every number is produced by
lib/libgwbp.so.  Nothing here calls into oracle/.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib as L


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (CUDA is required; there is no CPU path) -- got {t.device}")


def _f32c(t, device) -> torch.Tensor:
    t = torch.as_tensor(t, device=device) if not isinstance(t, torch.Tensor) else t
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _host_floats(x, count: int) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        x = x.detach().to("cpu", torch.float32).numpy()
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float32)).reshape(-1)
    assert a.size == count, f"expected {count} floats, got {a.size}"
    return a


class PackedScene:
    """Gaussians repacked once into the SoA layout the projection kernel streams (40 B each):
    world covariance replaces (quat, scale).  Mirrors the per-call preamble of the reference
    loop (backproject.py:55-57) but hoisted out of it."""

    def __init__(self, means, quats, scales, opacities, device=None):
        if device is None:
            device = means.device if isinstance(means, torch.Tensor) and means.is_cuda else torch.device("cuda")
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("CUDA is required for feature back-projection")  # backproject.py:314-315
        self.device = device
        means, quats = _f32c(means, device), _f32c(quats, device)
        scales, opacities = _f32c(scales, device), _f32c(opacities, device)
        n = means.shape[0]
        assert means.shape == (n, 3), means.shape
        assert quats.shape == (n, 4), quats.shape
        assert scales.shape == (n, 3), scales.shape
        assert opacities.shape == (n,), opacities.shape
        self.n = n
        self.means = means  # kept for per-Gaussian camera depth (render_mode "+D", click prompts)
        self.geo = torch.empty(max(40 * n, 16), dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            L.check(L.lib().gwbp_pack_scene(n, means.data_ptr(), quats.data_ptr(), scales.data_ptr(),
                                            opacities.data_ptr(), self.geo.data_ptr(), _stream_ptr(device)),
                    "gwbp_pack_scene")
        self.c = L.Scene(n, self.geo.data_ptr())


def make_camera(viewmat, K, width, height, near_plane=0.01, far_plane=1e10, radius_clip=0.0, eps2d=0.3) -> L.Camera:
    cam = L.Camera()
    cam.viewmat[:] = _host_floats(viewmat, 16).tolist()
    cam.K[:] = _host_floats(K, 9).tolist()
    cam.width, cam.height = int(width), int(height)  # 0-dim CUDA float tensors are legal (utils.py:247-248)
    cam.near_plane, cam.far_plane = float(near_plane), float(far_plane)
    cam.radius_clip, cam.eps2d = float(radius_clip), float(eps2d)
    return cam


def workspace_layout(n: int, width: int, height: int, cap: int) -> L.WsLayout:
    lay = L.WsLayout()
    L.check(L.lib().gwbp_workspace_layout(n, width, height, cap, C.byref(lay)), "gwbp_workspace_layout")
    return lay


class View:
    """One camera's projected, binned and depth-sorted scene, living in a caller-owned workspace
    (= everything one `rasterization()` call computes before compositing)."""

    def __init__(self, scene: PackedScene, cam: L.Camera, cap_isects: Optional[int] = None,
                 workspace: Optional[torch.Tensor] = None, tile_cull: bool = False, counting_bin: bool = False,
                 supertile: bool = False):
        """tile_cull=False: the intersection list is gsplat-1.4.0's (bit-exact `meta`).
        tile_cull=True : pairs that cannot reach alpha >= 1/255 on the tile are dropped before the sort
        (identical accumulators, less work) -- the BackProjector default.
        counting_bin=True selects the hand-written sort-free tile binning (images of <= 12 288 tiles) instead of
        emit + radix sort; both give the same flatten_ids / isect_offsets (the radix path is faster: DESIGN.md).
        supertile=True bins into 8 x 4-tile supertiles (one entry per Gaussian and supertile with a 32-bit tile mask;
        one radix pass): only the tcgen05 back-projection kernels read such lists (backproject_packed /
        backproject_lowres); meta(), render*() and the CUDA-core kernel need the per-tile lists."""
        self.scene, self.cam = scene, cam
        self._flags = ((L.PREPARE_TILE_CULL if tile_cull else L.PREPARE_GSPLAT_EXACT)
                       | (L.PREPARE_COUNTING_BIN if counting_bin else 0) | (L.PREPARE_SUPERTILE if supertile else 0))
        self._prepare(cap_isects, workspace)

    def _prepare(self, cap_isects, workspace) -> None:
        scene, cam = self.scene, self.cam
        n = scene.n
        cap = int(cap_isects) if cap_isects else max(1 << 16, 8 * n)
        while True:
            lay = workspace_layout(n, cam.width, cam.height, cap)
            if workspace is None or workspace.numel() < lay.total:
                workspace = torch.empty(lay.total, dtype=torch.uint8, device=scene.device)
            info = L.ViewInfo()
            with torch.cuda.device(scene.device):
                rc = L.lib().gwbp_view_prepare(C.byref(scene.c), C.byref(cam), workspace.data_ptr(), workspace.numel(),
                                               cap, self._flags, _stream_ptr(scene.device), C.byref(info))
            if rc == -2:  # capacity: the library told us the exact need; grow once and redo
                cap = int(info.n_isects * 1.25) + 1024
                workspace = None
                continue
            L.check(rc, "gwbp_view_prepare")
            break
        self.ws, self.layout, self.info, self.cap = workspace, lay, info, cap

    def _ensure_tile_lists(self) -> None:
        """Supertile lists are read by the tcgen05 back-projection kernels only.  Anything else (meta(), render*(), the
        CUDA-core kernel) needs gsplat's per-tile lists: re-bin this view in place (same workspace, same camera)."""
        if self.info.list_kind != 0:
            self._flags &= ~L.PREPARE_SUPERTILE
            self._prepare(self.cap, self.ws)

    # typed windows into the workspace (zero-copy; valid while self.ws is alive)
    def _win(self, off: int, count: int, dtype) -> torch.Tensor:
        nbytes = count * torch.empty((), dtype=dtype).element_size()
        return self.ws[off:off + nbytes].view(dtype)

    @property
    def n_vis(self) -> int:
        return int(self.info.n_vis)

    @property
    def n_isects(self) -> int:
        return int(self.info.n_isects)

    @property
    def n_entries(self) -> int:
        """Sorted list entries: n_isects for per-tile lists, (Gaussian, supertile) pairs for supertile lists."""
        return int(self.info.n_entries)

    def grec(self) -> torch.Tensor:
        return self._win(self.layout.grec, 8 * self.n_vis, torch.float32).view(self.n_vis, 8)

    def meta(self) -> dict:
        """The gsplat `meta` dict (packed=True semantics; the reference reads `means2d` and
        `gaussian_ids`: affordance_transfer/demo_affordance_transfer.py:392-395)."""
        self._ensure_tile_lists()
        g = self.grec()
        nv, ni = self.n_vis, self.n_isects
        sb = self.info.sorted_buf
        th, tw = self.info.tile_h, self.info.tile_w
        if self.info.tile_key_bytes == 0:
            # sort-free binning: sorted tile ids are never materialised; position p belongs to the tile whose
            # [offsets[t], offsets[t+1]) range holds it
            off = self._win(self.layout.offsets, th * tw + 1, torch.int32).to(torch.int64)
            tiles = torch.repeat_interleave(torch.arange(th * tw, device=self.ws.device), off[1:] - off[:-1])
        elif self.info.tile_key_bytes == 2:  # uint16 tile ids
            tiles = self._win(self.layout.tkeys1 if sb else self.layout.tkeys0, ni, torch.int16).to(torch.int64) & 0xFFFF
        else:
            tiles = self._win(self.layout.tkeys1 if sb else self.layout.tkeys0, ni, torch.int32)
        vals = self._win(self.layout.tvals1 if sb else self.layout.tvals0, ni, torch.int32)
        # gsplat's int64 keys (tile << 32 | depth bits), rebuilt from the two-stage sort's outputs
        depth_bits = g[:, 7].contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        keys = (tiles.to(torch.int64) << 32) | depth_bits[vals.to(torch.int64)]
        return {
            "camera_ids": torch.zeros(nv, dtype=torch.int64, device=self.ws.device),
            "gaussian_ids": g[:, 3].contiguous().view(torch.int32).to(torch.int64),
            "radii": self._win(self.layout.radii, nv, torch.int32),
            "means2d": g[:, 0:2],
            "depths": g[:, 7],
            "conics": g[:, 4:7],
            "opacities": g[:, 2],
            "tiles_per_gauss": self._win(self.layout.tiles_per_gauss, nv, torch.int32),
            "isect_ids": keys,
            "flatten_ids": vals,
            "isect_offsets": self._win(self.layout.offsets, th * tw, torch.int32).view(1, th, tw),
            "tile_width": tw, "tile_height": th, "tile_size": 16,
            "width": self.cam.width, "height": self.cam.height, "n_cameras": 1,
        }

    def backproject(self, feats: torch.Tensor, num: torch.Tensor, den: torch.Tensor, kernel: int = L.KERNEL_AUTO,
                    fpack: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None) -> None:
        """num[N,D] += sum_p w F[p,:]; den[N] += sum_p w.  feats: [H,W,D] fp32 CUDA, ANY strides."""
        _require_cuda(feats, "feats")
        assert feats.dtype == torch.float32, feats.dtype
        assert feats.dim() == 3 and feats.shape[0] == self.cam.height and feats.shape[1] == self.cam.width, \
            f"feats must be [H={self.cam.height}, W={self.cam.width}, D], got {tuple(feats.shape)}"
        d = feats.shape[2]
        assert num.shape == (self.scene.n, d) and num.dtype == torch.float32 and num.is_contiguous()
        assert den.shape == (self.scene.n,) and den.dtype == torch.float32 and den.is_contiguous()
        sH, sW, sD = feats.stride()
        if (kernel & 0xFF) != L.KERNEL_TC:
            self._ensure_tile_lists()
        with torch.cuda.device(self.scene.device):
            L.check(L.lib().gwbp_backproject_view(
                C.byref(self.scene.c), C.byref(self.cam), self.ws.data_ptr(), C.byref(self.info), feats.data_ptr(),
                sH, sW, sD, d, num.data_ptr(), den.data_ptr(), kernel,
                fpack.data_ptr() if fpack is not None else None,
                stats.data_ptr() if stats is not None else None, _stream_ptr(self.scene.device)),
                "gwbp_backproject_view")

    def backproject_packed(self, d: int, num: torch.Tensor, den: torch.Tensor, kernel: int, fpack: torch.Tensor,
                           stats: Optional[torch.Tensor] = None) -> None:
        """Same as backproject() when `fpack` was already filled by gwbp_pack_features[_lowres]."""
        assert kernel & L.KERNEL_FPACK_READY and fpack is not None
        assert num.shape == (self.scene.n, d) and num.dtype == torch.float32 and num.is_contiguous()
        assert den.shape == (self.scene.n,) and den.dtype == torch.float32 and den.is_contiguous()
        with torch.cuda.device(self.scene.device):
            L.check(L.lib().gwbp_backproject_view(
                C.byref(self.scene.c), C.byref(self.cam), self.ws.data_ptr(), C.byref(self.info), fpack.data_ptr(),
                0, 0, 0, d, num.data_ptr(), den.data_ptr(), kernel, fpack.data_ptr(),
                stats.data_ptr() if stats is not None else None, _stream_ptr(self.scene.device)),
                "gwbp_backproject_view")

    def backproject_lowres(self, feats_low: torch.Tensor, nearest: bool, num: torch.Tensor, den: torch.Tensor,
                           fpack: torch.Tensor, stats: Optional[torch.Tensor] = None, packed: bool = False) -> None:
        """backproject(interpolate(feats_low)) for an encoder-resolution map [h,w,D] (any strides), without the
        full-resolution map: gwbp_backproject_view_lowres.  packed=True: gwbp_pack_lowres_adjoint already filled `fpack`."""
        _require_cuda(feats_low, "feats_low")
        assert feats_low.dtype == torch.float32 and feats_low.dim() == 3
        d = feats_low.shape[2]
        assert num.shape == (self.scene.n, d) and num.dtype == torch.float32 and num.is_contiguous()
        assert den.shape == (self.scene.n,) and den.dtype == torch.float32 and den.is_contiguous()
        sH, sW, sD = feats_low.stride()
        with torch.cuda.device(self.scene.device):
            L.check(L.lib().gwbp_backproject_view_lowres(
                C.byref(self.scene.c), C.byref(self.cam), self.ws.data_ptr(), C.byref(self.info), feats_low.data_ptr(),
                feats_low.shape[0], feats_low.shape[1], sH, sW, sD, (1 if nearest else 0) | (L.LOWRES_PACKED if packed else 0),
                d, num.data_ptr(),
                den.data_ptr(), fpack.data_ptr(), stats.data_ptr() if stats is not None else None,
                _stream_ptr(self.scene.device)), "gwbp_backproject_view_lowres")

    def render(self, colors: torch.Tensor, background: Optional[torch.Tensor] = None, kernel: int = L.KERNEL_AUTO):
        """(render [H,W,D], alpha [H,W]) for colors [N,D] (row stride free, unit inner stride).
        kernel: KERNEL_AUTO (tcgen05 for D >= 64), KERNEL_SIMT (fp32 CUDA cores) or KERNEL_TC."""
        _require_cuda(colors, "colors")
        assert colors.dtype == torch.float32 and colors.dim() == 2 and colors.shape[0] == self.scene.n, \
            f"colors must be [N={self.scene.n}, D] float32, got {tuple(colors.shape)} {colors.dtype}"
        if colors.stride(1) != 1 or colors.stride(0) < colors.shape[1]:
            colors = colors.contiguous()
        d = colors.shape[1]
        H, W = self.cam.height, self.cam.width
        self._ensure_tile_lists()
        alloc = torch.empty if self.n_isects else torch.zeros  # the kernels write every pixel
        out = alloc(H, W, d, dtype=torch.float32, device=colors.device)
        alpha = alloc(H, W, dtype=torch.float32, device=colors.device)
        if background is not None:
            background = _f32c(background, colors.device).reshape(-1)
            assert background.numel() == d
            if self.n_isects == 0:
                out += background
        if self.n_isects:
            with torch.cuda.device(self.scene.device):
                L.check(L.lib().gwbp_render_view(
                    C.byref(self.scene.c), C.byref(self.cam), self.ws.data_ptr(), C.byref(self.info),
                    colors.data_ptr(), colors.stride(0), d,
                    background.data_ptr() if background is not None else None, out.data_ptr(), alpha.data_ptr(),
                    int(kernel), _stream_ptr(self.scene.device)), "gwbp_render_view")
        return out, alpha


    def render_pixels(self, colors: torch.Tensor, xy: torch.Tensor, extra: Optional[torch.Tensor] = None):
        """The composite of `render` at probe pixels only: (out [k, D(+1)], alpha [k]) for xy [k,2] = (x, y).
        `extra` [N] appends one more composited channel (camera depth -> render_mode="RGB+D").  Replaces the
        513-channel full-frame render of the click prompt, of which one pixel is read
        (click_and_segment.py:241-262)."""
        _require_cuda(colors, "colors")
        assert colors.dtype == torch.float32 and colors.dim() == 2 and colors.shape[0] == self.scene.n, \
            f"colors must be [N={self.scene.n}, D] float32, got {tuple(colors.shape)} {colors.dtype}"
        if colors.stride(1) != 1 or colors.stride(0) < colors.shape[1]:
            colors = colors.contiguous()
        d = colors.shape[1]
        xy = torch.as_tensor(xy, device=colors.device).to(torch.int32).reshape(-1, 2).contiguous()
        k = xy.shape[0]
        if extra is not None:
            extra = _f32c(extra, colors.device).reshape(-1)
            assert extra.numel() == self.scene.n, "extra must be [N]"
        out = torch.zeros(k, d + (extra is not None), dtype=torch.float32, device=colors.device)
        alpha = torch.zeros(k, dtype=torch.float32, device=colors.device)
        self._ensure_tile_lists()
        if k and self.n_isects:
            with torch.cuda.device(self.scene.device):
                L.check(L.lib().gwbp_render_pixels(
                    C.byref(self.scene.c), C.byref(self.cam), self.ws.data_ptr(), C.byref(self.info),
                    colors.data_ptr(), colors.stride(0), d, extra.data_ptr() if extra is not None else None,
                    xy.data_ptr(), k, out.data_ptr(), alpha.data_ptr(), _stream_ptr(self.scene.device)),
                    "gwbp_render_pixels")
        return out, alpha

    def ratio_accumulate(self, num_v: torch.Tensor, den_v: torch.Tensor, acc: torch.Tensor, num_scale: float,
                         den_scale: float, eps: float = 1e-12, den_acc: Optional[torch.Tensor] = None) -> None:
        """acc += (num_scale*num_v) / (den_scale*den_v + eps) on the rows this view saw; resets those rows of
        (num_v, den_v) to zero (affordance_transfer/demo_affordance_transfer.py:768-796); den_acc [N] += den_v if given."""
        n = self.scene.n
        checks = [("num_v", num_v, (n, acc.shape[1])), ("den_v", den_v, (n,)), ("acc", acc, acc.shape)]
        if den_acc is not None:
            checks.append(("den_acc", den_acc, (n,)))
        for name, t_, shape in checks:
            _require_cuda(t_, name)
            assert t_.dtype == torch.float32 and t_.is_contiguous() and tuple(t_.shape) == tuple(shape), \
                f"{name}: expected contiguous float32 {tuple(shape)}, got {tuple(t_.shape)} {t_.dtype}"
        assert acc.shape[0] == n
        if self.n_vis:
            with torch.cuda.device(self.scene.device):
                L.check(L.lib().gwbp_ratio_accumulate(
                    C.byref(self.scene.c), C.byref(self.cam), self.ws.data_ptr(), C.byref(self.info),
                    num_v.data_ptr(), den_v.data_ptr(), acc.data_ptr(),
                    den_acc.data_ptr() if den_acc is not None else None, acc.shape[1], float(num_scale),
                    float(den_scale), float(eps), _stream_ptr(self.scene.device)), "gwbp_ratio_accumulate")


def fpack_bytes(width: int, height: int, d: int) -> int:
    return int(L.lib().gwbp_fpack_bytes(int(width), int(height), int(d)))


def finalize(num: torch.Tensor, den: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """backproject.py:166-169 in one pass."""
    _require_cuda(num, "num")
    assert num.dtype == torch.float32 and num.is_contiguous() and den.is_contiguous() and den.dtype == torch.float32
    n, d = num.shape
    assert den.shape == (n,)
    if out is None:
        out = torch.empty_like(num)
    with torch.cuda.device(num.device):
        L.check(L.lib().gwbp_finalize(num.data_ptr(), den.data_ptr(), out.data_ptr(), n, d, _stream_ptr(num.device)),
                "gwbp_finalize")
    return out


def cosine_mask(x: torch.Tensor, text: torch.Tensor, n_pos: int, threshold: Optional[float] = None,
                return_score: bool = False):
    """mask = max_pos(cos) > max_neg(cos) [and cos_0 > threshold] over the rows of x [..., D]
    (segment.py:52-58 / 221-224)."""
    _require_cuda(x, "features")
    d = x.shape[-1]
    lead = x.shape[:-1]
    x2 = x.detach().to(torch.float32).reshape(-1, d).contiguous()
    text = _f32c(text, x.device)
    assert text.dim() == 2 and text.shape[1] == d, f"text must be [P,{d}], got {tuple(text.shape)}"
    rows, p = x2.shape[0], text.shape[0]
    mask = torch.zeros(rows, dtype=torch.uint8, device=x.device)
    score = torch.empty(rows, p, dtype=torch.float32, device=x.device) if return_score else None
    with torch.cuda.device(x.device):
        L.check(L.lib().gwbp_mask3d(x2.data_ptr(), rows, d, text.data_ptr(), p, int(n_pos),
                                    float(threshold) if threshold is not None else 0.0,
                                    1 if threshold is not None else 0, mask.data_ptr(),
                                    score.data_ptr() if score is not None else None, _stream_ptr(x.device)),
                "gwbp_mask3d")
    mask = mask.view(torch.bool).reshape(lead)
    return (mask, score.reshape(*lead, p)) if return_score else mask
