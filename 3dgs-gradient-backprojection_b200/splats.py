"""The reference's data contract for this path -- tensors only, no file parsing of
COLMAP / PLY (those loaders are out of scope):

  * `splats` dict as `utils.load_checkpoint` returns it (utils.py:47-107): means [N,3], rotation [N,4] (wxyz,
    un-normalised), scaling [N,3] (LOG scales), opacity [N] (LOGITS), features_dc [N,1,3], features_rest [N,15,3],
    camera_matrix [3,3].
  * gsplat-format checkpoint `{"splats": {means, quats, scales, opacities, sh0, shN}}` (utils.py:58-70;
    written back by segment.py:240-256).
  * `features_*.pt`: one float32 [N_pruned, D] tensor (backproject.py:330) -- BackProjector.save().

Pure tensor plumbing: importable (and tested) without the CUDA library.
"""
from __future__ import annotations

from typing import Dict

import torch

_GSPLAT_TO_SPLATS = {"means": "means", "sh0": "features_dc", "shN": "features_rest", "scales": "scaling",
                     "quats": "rotation", "opacities": "opacity"}
_PER_GAUSSIAN = ("means", "features_dc", "features_rest", "scaling", "rotation", "opacity")


def splats_from_gsplat_checkpoint(model: dict) -> Dict[str, torch.Tensor]:
    """`model = torch.load(ckpt)` in gsplat format -> the reference's `splats` dict (utils.py:58-70), detached."""
    params = model["splats"]
    splats = {dst: params[src].detach() for src, dst in _GSPLAT_TO_SPLATS.items()}
    splats["active_sh_degree"] = 3
    return splats


def gsplat_checkpoint_from_splats(splats: dict) -> dict:
    """Inverse mapping: what `save_to_ckpt` writes (segment.py:240-256)."""
    return {"splats": {src: splats[dst] for src, dst in _GSPLAT_TO_SPLATS.items()}}


def activated(splats: dict):
    """(means, quats, scales, opacities) exactly as every script feeds `rasterization`
    (backproject.py:55-57, utils.py:228-231): exp on the log-scales, sigmoid on the opacity logits."""
    return (splats["means"], splats["rotation"], torch.exp(splats["scaling"]), torch.sigmoid(splats["opacity"]))


def prune_splats(splats: dict, keep: torch.Tensor) -> dict:
    """Row-select every per-Gaussian tensor with `keep` (bool [N] or indices), like the tail of
    prune_by_gradients (utils.py:257-268); other entries (camera_matrix, colmap handles) pass through."""
    out = dict(splats)
    for k in _PER_GAUSSIAN:
        if k in splats:
            out[k] = splats[k][keep.to(splats[k].device)]
    return out


def viewmat_from_rotation_translation(R, t) -> torch.Tensor:
    """get_viewmat_from_colmap_image (utils.py:215-219): world->camera [4,4] from COLMAP's R (3x3) and t (3)."""
    viewmat = torch.eye(4, dtype=torch.float32)
    viewmat[:3, :3] = torch.as_tensor(R, dtype=torch.float32)
    viewmat[:3, 3] = torch.as_tensor(t, dtype=torch.float32)
    return viewmat


def camera_matrix(fx: float, fy: float, cx: float, cy: float, data_factor: float = 1.0) -> torch.Tensor:
    """The `camera_matrix` entry of `splats` (utils.py:92-103): pinhole K with the first two rows divided by
    the image down-scale factor."""
    K = torch.tensor([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], dtype=torch.float32)
    K[:2, :3] /= data_factor
    return K
