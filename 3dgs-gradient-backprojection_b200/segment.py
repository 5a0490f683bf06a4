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


def click_prompt(scene: PackedScene, features: torch.Tensor, viewmat, K, width, height, xy, **cam_kw):
    """What one click of click_and_segment.py extracts (:241-275): the L2-normalised rendered feature at
    pixel(s) xy [k,2] = (x, y) (the prompt vector) and the world-space point under the click, un-projected
    with the rendered depth channel of `render_mode="RGB+D"`.  The reference renders all 513 channels of the
    whole frame for this; here only the clicked pixels are composited (gwbp_render_pixels).

    Returns (prompt [k,D], xyz_world [k,3], alpha [k])."""
    vm = torch.as_tensor(viewmat, dtype=torch.float32).cpu()
    Kc = torch.as_tensor(K, dtype=torch.float32).cpu()
    view = View(scene, make_camera(vm, Kc, width, height, **cam_kw), tile_cull=True)
    dev = features.device
    z_g = scene.means @ vm[2, :3].to(dev) + vm[2, 3].to(dev)  # camera-space depth per Gaussian (the +D channel)
    xy = torch.as_tensor(xy, device=dev).reshape(-1, 2)
    out, alpha = view.render_pixels(features, xy, extra=z_g)
    prompt = torch.nn.functional.normalize(out[:, :-1], dim=-1)          # click_and_segment.py:275
    Z = out[:, -1]
    fx, fy, cx, cy = Kc[0, 0].item(), Kc[1, 1].item(), Kc[0, 2].item(), Kc[1, 2].item()
    cam_pt = torch.stack([(xy[:, 0].float() - cx) / fx * Z, (xy[:, 1].float() - cy) / fy * Z, Z,
                          torch.ones_like(Z)], dim=1)                    # :262-267
    world = (torch.inverse(vm).to(dev) @ cam_pt.T).T[:, :3]              # :269
    return prompt, world, alpha


def click_mask3d(features: torch.Tensor, positive_prompts: torch.Tensor, negative_prompts: torch.Tensor):
    """mask_3d of click_and_segment.py:317-321: max_j f.pos_j > max_j f.neg_j.  Prompts are unit vectors
    (click_prompt) and feature rows are unit or zero, so the cosine compare of get_mask3d is the same test."""
    text = torch.cat([positive_prompts, negative_prompts], 0)
    return cosine_mask(features, text, positive_prompts.shape[0])
