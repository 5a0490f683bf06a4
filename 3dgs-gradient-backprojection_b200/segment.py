"""Query side of the path (segment.py:26-61, 145-240 of the reference), encoder factored out:
the caller supplies CLIP text embeddings `text_feat [P,D]` (rows 0..n_pos-1 positive prompts).

    mask3d, mask3d_inv = get_mask3d(features, text_feat, n_pos)            # segment.py:52-61
    mask2d = render_mask_2d(scene, features, text_feat, n_pos, viewmat, K, W, H)   # segment.py:209-224
"""
from __future__ import annotations

from typing import Optional

import torch

from .engine import PackedScene, View, cosine_mask, make_camera


def get_mask3d(features: torch.Tensor, text_feat: torch.Tensor, n_pos: int, threshold: Optional[float] = None):
    """mask = max(pos scores) > max(neg scores) [& score[:,0] > threshold]; returns (mask, ~mask)."""
    m = cosine_mask(features, text_feat, n_pos, threshold)
    return m, ~m


def render_features(scene: PackedScene, features: torch.Tensor, viewmat, K, width, height, **cam_kw):
    """Forward render of per-Gaussian features -> ([H,W,D], alpha [H,W])  (segment.py:209-220)."""
    view = View(scene, make_camera(viewmat, K, width, height, **cam_kw))
    return view.render(features)


def gaussian_scores(features: torch.Tensor, text_feat: torch.Tensor) -> torch.Tensor:
    """[N,P] per-Gaussian scores f_g . normalise(t_j): view-independent, compute ONCE per query and pass it
    to render_mask_2d(scores=...) for every view."""
    t = torch.nn.functional.normalize(text_feat.to(features.device, torch.float32), dim=1)
    return (features.to(torch.float32) @ t.T).contiguous()


def render_mask_2d(scene: PackedScene, features: torch.Tensor, text_feat: torch.Tensor, n_pos: int, viewmat, K,
                   width, height, exact_render: bool = True, scores: Optional[torch.Tensor] = None,
                   **cam_kw) -> torch.Tensor:
    """Per-pixel mask of one view (segment.py:209-224).

    exact_render=True  : render all D channels, then normalise / score / compare per pixel, exactly
                         the reference's order of operations.
    exact_render=False : use linearity -- render the P per-Gaussian scores f_g . t_j instead of the
                         D features (P << D); the per-pixel normalisation is a positive scale common
                         to all P scores, so the compare is unchanged in exact arithmetic
                         (SURVEY.md §9.7).  Pass `scores=gaussian_scores(features, text_feat)` to reuse
                         them across views."""
    # the mask does not need gsplat-exact `meta`, so the culled (shorter) intersection list is used
    view = View(scene, make_camera(viewmat, K, width, height, **cam_kw), tile_cull=True)
    if exact_render:
        render, _ = view.render(features)
        return cosine_mask(render, text_feat, n_pos)
    if scores is None:
        scores = gaussian_scores(features, text_feat)
    rs, _ = view.render(scores)
    return rs[..., :n_pos].max(dim=2)[0] > rs[..., n_pos:].max(dim=2)[0]
