"""The hot path as an object: per-view fused back-projection into resident (num, den)
accumulators -- the loop body of `create_feature_field_lseg` / `_dino`
(backproject.py:74-165, 214-289) and of `backproject_compressed.py:90-179`, minus the 2-D encoder.

Reference, per view:  3 x rasterization() + 2 x backward() + clone/zero_/+= over dense [N,D]
Here, per view:       1 x (project + bin + sort)  +  1 fused composite/contract/accumulate kernel

    bp = BackProjector(means, quats, scales, opacities, feature_dim=512)
    for viewmat, feats in views:            # feats [H,W,D] fp32 CUDA, any strides
        bp.add_view(viewmat, K, width, height, feats)
    features = bp.finalize()                # == create_feature_field_lseg(...) return value

`raw()` exposes (num, den) for the multi-GPU all-reduce (dist.py); `prune_mask()` is
`prune_by_gradients` (utils.py:222-271) for free: a Gaussian survives iff den received weight.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from .engine import PackedScene, View, finalize as _finalize, fpack_bytes, make_camera

DEN_EPS = 1e-12  # backproject.py:63


class BackProjector:
    def __init__(self, means, quats, scales, opacities, feature_dim: int, device=None, kernel: str = "auto",
                 cap_isects: Optional[int] = None, collect_stats: bool = False, tile_cull: bool = True,
                 accumulate: str = "sum", supertile: bool = True):
        """accumulate="sum": num += num_v, den += den_v, features = num/den (backproject.py:149-150,166).
        accumulate="per_view_ratio": features += mean-scaled num_v / (mean-scaled den_v + 1e-12) per view
        (affordance_transfer/demo_affordance_transfer.py:768-800); `num` then holds that sum."""
        assert accumulate in ("sum", "per_view_ratio"), accumulate
        self.accumulate = accumulate
        self.scene = PackedScene(means, quats, scales, opacities, device)
        self.device = self.scene.device
        self.d = int(feature_dim)
        n = self.scene.n
        self.num = torch.zeros(n, self.d, dtype=torch.float32, device=self.device)       # backproject.py:62
        self.den = torch.full((n,), DEN_EPS, dtype=torch.float32, device=self.device)    # backproject.py:63
        if accumulate == "per_view_ratio":  # per-view scratch pair, kept all-zero between views
            self.num_v = torch.zeros(n, self.d, dtype=torch.float32, device=self.device)
            self.den_v = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.kernel = {"auto": L.KERNEL_AUTO, "simt": L.KERNEL_SIMT, "tc": L.KERNEL_TC}[kernel]
        self.cap = cap_isects
        self.tile_cull = bool(tile_cull)
        # views handed to the tcgen05 kernels are binned into 8 x 4-tile supertiles (~2.4x fewer list entries, one
        # radix pass); views that fall back to the CUDA-core kernel keep per-tile lists
        self.supertile = bool(supertile)
        self._ws: Optional[torch.Tensor] = None
        self._fpack: Optional[torch.Tensor] = None
        self._stats = torch.zeros(4, dtype=torch.int64, device=self.device) if collect_stats else None
        self.n_views = 0
        self.last_view: Optional[View] = None
        self.kernel_events = None  # set to [] to record (start, end) CUDA events around the fused kernel
        # The feature re-layout depends only on F, so it runs on a second stream next to projection/binning of the
        # same view.  Measured on B200 (config G, profiles/r02_overlap_pack.txt): 2.80 -> 2.65 ms per view.  The gain is
        # bounded because the projection kernel fills every SM's register file, so the two only share SMs at the
        # geometry pipeline's small sorts and at kernel tails; a persistent re-layout grid meant to co-reside with it was
        # 1.5x slower on its own and still serialised (an SM's L1/shared split only changes when the SM is empty).
        self.overlap_pack = True
        # encoder-resolution maps: "adjoint" = down-sampled weights x low-res map (gwbp_backproject_view_lowres, no
        # full-resolution intermediate); "upsample" = fused upsample into the packed operand + the full-resolution kernel
        self.lowres_impl = "adjoint"
        self._side = None
        self._main = None
        self._copy_stream = None  # add_view_host(): upload stream, two staging buffers, the deferred view
        self._stage = None
        self._pending = None
        self._host_seq = 0

    @classmethod
    def from_splats(cls, splats: dict, feature_dim: int, **kw) -> "BackProjector":
        """From the reference's `splats` dict (utils.load_checkpoint): log-scales and opacity logits are
        activated as in backproject.py:55-57."""
        from .splats import activated
        return cls(*activated(splats), feature_dim=feature_dim, **kw)

    # -- one view -------------------------------------------------------------------------
    def add_view(self, viewmat, K, width, height, feats: torch.Tensor, **cam_kw) -> View:
        """feats: the full-resolution [H,W,D] fp32 map (any strides), as the reference builds it."""
        assert feats.shape[-1] == self.d, f"feature dim {feats.shape[-1]} != {self.d}"
        return self._add(viewmat, K, width, height, feats, None, cam_kw)

    def add_view_lowres(self, viewmat, K, width, height, feats_low: torch.Tensor, mode: str = "bilinear",
                        **cam_kw) -> View:
        """feats_low: the ENCODER-resolution map [h,w,D] (any strides, e.g. `net_out[0].permute(1,2,0)`).
        Equivalent to add_view(interpolate(feats_low, (H,W), mode)) (backproject.py:110-113 / :245-249) but
        the upsample is fused into the feature re-layout pass and the [H,W,D] tensor is never built."""
        assert mode in ("bilinear", "nearest"), mode
        self._check_map(feats_low, make_camera(viewmat, K, width, height, **cam_kw), True)
        if not fpack_bytes(int(width), int(height), self.d) or self.kernel == L.KERNEL_SIMT:
            up = torch.nn.functional.interpolate(feats_low.permute(2, 0, 1)[None], size=(int(height), int(width)),
                                                 mode=mode)[0].permute(1, 2, 0)
            return self._add(viewmat, K, width, height, up, None, cam_kw)
        return self._add(viewmat, K, width, height, feats_low, mode, cam_kw)

    # -- host-resident feature maps: pipelined upload ------------------------------------
    def add_view_host(self, viewmat, K, width, height, feats_host: torch.Tensor, lowres_mode: Optional[str] = None,
                      **cam_kw) -> None:
        """feats_host: a (pinned) HOST tensor in the reference's planar layout [D,H,W] (backproject.py:110-113), or,
        with lowres_mode="bilinear"/"nearest", the encoder-resolution map [D,h,w] (:109).  The upload runs on a
        copy stream into one of two device staging buffers while the PREVIOUS view is being back-projected, so the
        PCIe transfer and the kernels overlap; the view is accumulated at the next add_view_host()/flush() (every
        accessor of the accumulators flushes)."""
        assert not feats_host.is_cuda and feats_host.dim() == 3 and feats_host.shape[0] == self.d, \
            f"host map must be a CPU tensor [D={self.d}, h, w], got {tuple(feats_host.shape)} on {feats_host.device}"
        assert feats_host.dtype == torch.float32, f"host map must be float32, got {feats_host.dtype}"
        assert lowres_mode in (None, "bilinear", "nearest"), lowres_mode
        if lowres_mode is None:
            assert tuple(feats_host.shape[1:]) == (int(height), int(width)), \
                f"host map must be [D, H={int(height)}, W={int(width)}], got {tuple(feats_host.shape)}"
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        if self._stage is None:
            self._stage = [None, None]
        slot = self._host_seq & 1
        self._host_seq += 1
        buf = self._stage[slot]
        if buf is None or buf.shape != feats_host.shape:
            buf = self._stage[slot] = torch.empty(feats_host.shape, dtype=torch.float32, device=self.device)
        cur = torch.cuda.current_stream(self.device)
        self._copy_stream.wait_stream(cur)  # the kernels that last read this staging buffer are enqueued on `cur`
        with torch.cuda.stream(self._copy_stream):
            buf.copy_(feats_host, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        prev, self._pending = self._pending, (viewmat, K, width, height, buf, lowres_mode, cam_kw, ready)
        if prev is not None:
            self._run_pending(prev)

    def _run_pending(self, item) -> None:
        viewmat, K, width, height, buf, lowres_mode, cam_kw, ready = item
        torch.cuda.current_stream(self.device).wait_event(ready)
        feats = buf.permute(1, 2, 0)
        if lowres_mode is None:
            self.add_view(viewmat, K, width, height, feats, **cam_kw)
        else:
            self.add_view_lowres(viewmat, K, width, height, feats, lowres_mode, **cam_kw)

    def flush(self) -> None:
        """Accumulate the view add_view_host() is still holding back."""
        prev, self._pending = self._pending, None
        if prev is not None:
            self._run_pending(prev)

    def _check_map(self, feats: torch.Tensor, cam, lowres: bool) -> None:
        """The packed (tcgen05) path hands raw pointers and strides to the C ABI, so everything View.backproject()
        would have asserted is checked HERE, before any kernel sees the tensor: a half-precision (autocast) encoder
        output, a CPU tensor or a wrong-size map must raise, not be reinterpreted as fp32 [H,W,D]."""
        if not isinstance(feats, torch.Tensor) or not feats.is_cuda:
            raise RuntimeError("feature map must be a CUDA tensor (CUDA is required; there is no CPU path) -- got "
                               f"{getattr(feats, 'device', type(feats))}; use add_view_host() for host-resident maps")
        if feats.device != self.device:
            raise RuntimeError(f"feature map lives on {feats.device}, the accumulators on {self.device}")
        assert feats.dtype == torch.float32, f"feature map must be float32, got {feats.dtype} (cast it: .float())"
        assert feats.dim() == 3 and feats.shape[2] == self.d, \
            f"feature map must be [h,w,D={self.d}], got {tuple(feats.shape)}"
        if lowres:
            assert feats.shape[0] >= 1 and feats.shape[1] >= 1, tuple(feats.shape)
        else:
            assert feats.shape[0] == cam.height and feats.shape[1] == cam.width, \
                f"feature map must be [H={cam.height}, W={cam.width}, D], got {tuple(feats.shape)}"

    def _add(self, viewmat, K, width, height, feats, lowres_mode, cam_kw) -> View:
        cam = make_camera(viewmat, K, width, height, **cam_kw)
        self._check_map(feats, cam, lowres_mode is not None)
        cur = torch.cuda.current_stream(self.device)
        fp, kernel = None, self.kernel
        if self.kernel != L.KERNEL_SIMT:
            need = fpack_bytes(cam.width, cam.height, self.d)
            if need:
                if self._fpack is None or self._fpack.numel() < need:
                    self._fpack = torch.empty(need, dtype=torch.uint8, device=self.device)
                fp = self._fpack
            elif self.kernel == L.KERNEL_TC:
                raise RuntimeError(f"tcgen05 kernel does not support D={self.d}")
        adjoint = fp is not None and lowres_mode is not None and self.lowres_impl == "adjoint"
        if adjoint and not L.lib().gwbp_lowres_adjoint_supported(cam.width, cam.height, feats.shape[0], feats.shape[1],
                                                                  self.d, 1 if lowres_mode == "nearest" else 0):
            adjoint = False  # window too large for the adjoint kernel: fused upsample + the full-resolution kernel
        overlap = fp is not None and self.overlap_pack
        if overlap and self._side is None:
            # The feature re-layout (a 69k-CTA, HBM-bound grid; for the adjoint path the 39 us bf16 copy of the low-res
            # map) depends only on F, the geometry pipeline (a dozen small kernels) only on the camera: run them
            # concurrently.  The geometry stream gets the higher priority, otherwise its small grids queue behind the
            # big one and nothing overlaps.
            self._side = torch.cuda.Stream(self.device, priority=0)
            self._main = torch.cuda.Stream(self.device, priority=-1)
        main = self._main if overlap else cur
        if overlap:
            main.wait_stream(cur)
        if fp is not None:  # re-layout first (own entry point, so the fused kernel can be timed on its own)
            side = self._side if overlap else main
            if overlap:
                side.wait_stream(main)  # previous view's kernel has finished reading fpack; F is ready
            sH, sW, sD = feats.stride()
            with torch.cuda.device(self.device):
                if adjoint:
                    L.check(L.lib().gwbp_pack_lowres_adjoint(feats.data_ptr(), feats.shape[0], feats.shape[1], sH, sW, sD,
                                                             self.d, fp.data_ptr(), int(side.cuda_stream)),
                            "gwbp_pack_lowres_adjoint")
                elif lowres_mode is None:
                    L.check(L.lib().gwbp_pack_features(cam.width, cam.height, feats.data_ptr(), sH, sW, sD, self.d,
                                                       fp.data_ptr(), int(side.cuda_stream)), "gwbp_pack_features")
                else:
                    L.check(L.lib().gwbp_pack_features_lowres(
                        cam.width, cam.height, feats.data_ptr(), feats.shape[0], feats.shape[1], sH, sW, sD,
                        1 if lowres_mode == "nearest" else 0, self.d, fp.data_ptr(), int(side.cuda_stream)),
                        "gwbp_pack_features_lowres")
            if overlap:
                feats.record_stream(side)
            if not adjoint:
                kernel = (L.KERNEL_TC if kernel == L.KERNEL_AUTO else kernel) | L.KERNEL_FPACK_READY
        with torch.cuda.stream(main):
            view = View(self.scene, cam, self.cap, self._ws, self.tile_cull,
                        supertile=self.supertile and fp is not None and self.kernel != L.KERNEL_SIMT)
            self._ws, self.cap = view.ws, view.cap  # keep (possibly grown) workspace for the next view
            if overlap:
                main.wait_stream(self._side)
            if self.kernel_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(main)
            ratio = self.accumulate == "per_view_ratio"
            num, den = (self.num_v, self.den_v) if ratio else (self.num, self.den)
            if adjoint:
                view.backproject_lowres(feats, lowres_mode == "nearest", num, den, fp, self._stats, packed=True)
            elif kernel & L.KERNEL_FPACK_READY:
                view.backproject_packed(self.d, num, den, kernel, fp, self._stats)
            else:
                view.backproject(feats, num, den, kernel, fp, self._stats)
            if ratio:  # `.mean()` losses: 1/(H*W*D) on num, 1/(H*W*3) on den (demo_affordance_transfer.py:768,790)
                hw = float(cam.width) * float(cam.height)
                view.ratio_accumulate(num, den, self.num, 1.0 / (hw * self.d), 1.0 / (hw * 3.0), DEN_EPS, self.den)
            if self.kernel_events is not None:
                e1.record(main)
                self.kernel_events.append((e0, e1))
        if overlap:
            cur.wait_stream(main)  # results are ordered on the caller's stream, as for any other op
        self.n_views += 1
        self.last_view = view
        return view

    # -- results --------------------------------------------------------------------------
    def raw(self):
        """(num [N,D], den [N]) -- den includes the reference's 1e-12 initial value."""
        self.flush()
        return self.num, self.den

    def stats(self) -> dict:
        """Counters summed over all views so far (device -> host read)."""
        if self._stats is None:
            return {}
        self.flush()
        s = self._stats.tolist()
        return {"rows_nonzero": s[0], "entries_walked": s[1]}

    def reset(self) -> None:
        self._pending = None
        self.num.zero_()
        self.den.fill_(DEN_EPS)
        if self._stats is not None:
            self._stats.zero_()
        self.n_views = 0

    def prune_mask(self) -> torch.Tensor:
        """== `gaussian_grads > 0` of prune_by_gradients (utils.py:257)."""
        self.flush()
        return self.den > DEN_EPS

    def finalize(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """backproject.py:166-169.  In per_view_ratio mode: L2-normalised rows of the ratio sum
        (demo_affordance_transfer.py:800), with never-seen rows 0 instead of the reference's NaN."""
        self.flush()
        if self.accumulate == "per_view_ratio":
            return _finalize(self.num, torch.ones_like(self.den), out)
        return _finalize(self.num, self.den, out)

    def save(self, path: str, prune: bool = True, with_index: bool = True,
             keep: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Write the feature field the way the reference does (backproject.py:330): ONE float32 tensor
        [N_pruned, D] whose rows follow `prune_by_gradients`' mask (utils.py:257-268), so segment.py can
        `torch.load` it unchanged.  with_index additionally writes `<path>.kept.pt` (the kept Gaussian
        indices) so consumers need not re-derive the mask.

        The mask is `den > 0` over the views ADDED TO THIS OBJECT: it equals prune_by_gradients' mask only if every
        COLMAP image that function renders was added (the reference prunes first and back-projects the pruned set; a
        zero-weight Gaussian that terminates a pixel here is absent there, which moves later weights by a threshold
        flip at most -- INTEGRATION.md).  Pass `keep` [N] bool (e.g. the mask prune_by_gradients returned for these
        splats) to make the rows align with an externally pruned checkpoint regardless of the views added."""
        feats = self.finalize()
        if keep is not None:
            keep = torch.as_tensor(keep, device=self.device).to(torch.bool).reshape(-1)
            assert keep.numel() == self.scene.n, f"keep must have {self.scene.n} entries, got {keep.numel()}"
        else:
            keep = self.prune_mask() if prune else torch.ones_like(self.den, dtype=torch.bool)
        out = feats[keep].contiguous()
        torch.save(out, path)
        if with_index:
            torch.save(torch.nonzero(keep).flatten().to(torch.int64).cpu(), path + ".kept.pt")
        return out


def create_feature_field(means, quats, scales, opacities, K, width, height, views, feature_dim: int, **kw):
    """Functional mirror of `create_feature_field_lseg(splats)` with the encoder factored out:
    `views` yields (viewmat[4,4], feats[H,W,D]).  Returns the normalised [N,D] tensor that the
    reference saves as features_*.pt (backproject.py:330)."""
    bp = BackProjector(means, quats, scales, opacities, feature_dim, **kw)
    for viewmat, feats in views:
        bp.add_view(viewmat, K, width, height, feats)
    return bp.finalize()
