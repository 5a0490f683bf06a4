"""Synthetic scenes for the back-projection path (SURVEY.md §8d generator spec).

The reference ships no data (its checkpoints / COLMAP dirs are git-ignored), so every
test and benchmark in this repo runs on seeded synthetic scenes that have the *shape*
of what `utils.py:load_checkpoint` (utils.py:20-109) hands to
`create_feature_field_lseg` (backproject.py:25-172):

  splats = {means[N,3], rotation[N,4] (wxyz, un-normalised), scaling[N,3] (log),
            opacity[N] (logit)}   +   per-view viewmat[4,4] (world->cam, OpenCV) and K[3,3].

This module is data generation only -- it holds none of the path's arithmetic -- so it is
shared by the product, the tests and the oracle.  numpy only (it must run on the GPU box
and in the CPU container identically: same seed -> same bytes).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass
class Scene:
    """Activated splat parameters, as the reference passes them to `rasterization`
    (backproject.py:55-57: opacities = sigmoid(.), scales = exp(.), quats raw)."""

    means: np.ndarray      # [N,3] float32
    quats: np.ndarray      # [N,4] float32, wxyz, NOT normalised
    scales: np.ndarray     # [N,3] float32, linear
    opacities: np.ndarray  # [N]   float32 in (0,1)

    @property
    def n(self) -> int:
        return int(self.means.shape[0])


def make_scene(n: int, seed: int = 0) -> Scene:
    """70 % object/ground core + 30 % background shell."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_core = int(round(0.7 * n))
    n_shell = n - n_core

    core = rng.standard_normal((n_core, 3)).astype(np.float32) * np.array([1.0, 1.0, 0.35], np.float32)
    r = np.linalg.norm(core, axis=1, keepdims=True)
    core = core * np.minimum(1.0, 3.0 / np.maximum(r, 1e-6)).astype(np.float32)

    d = rng.standard_normal((n_shell, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rad = rng.uniform(6.0, 20.0, size=(n_shell, 1))
    shell = (d * rad).astype(np.float32)
    means = np.concatenate([core, shell], 0).astype(np.float32)

    v_scene = 4.0 / 3.0 * math.pi * 2.0 * 2.0 * 0.7
    s0 = 0.6 * (v_scene / max(n, 1)) ** (1.0 / 3.0)
    log_s = rng.standard_normal((n, 3)) * 0.6 + math.log(s0)
    log_s[n_core:] += math.log(6.0)
    scales = np.exp(log_s).astype(np.float32)

    quats = rng.standard_normal((n, 4)).astype(np.float32)
    logit = rng.standard_normal(n) * 2.0 + 0.5
    opac = (1.0 / (1.0 + np.exp(-logit))).astype(np.float32)

    perm = rng.permutation(n)  # checkpoints are not ordered core-then-shell
    return Scene(means[perm].copy(), quats[perm].copy(), scales[perm].copy(), opac[perm].copy())


def make_cameras(n_views: int, width: int, height: int, seed: int = 0):
    """V poses on a jittered circle (radius 4.5, height 1.5) looking at the origin.

    Returns (viewmats[V,4,4], K[3,3]) float32; same conventions as
    `get_viewmat_from_colmap_image` (utils.py:215-219): world->camera, +z forward.
    """
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    viewmats = np.zeros((n_views, 4, 4), np.float32)
    for v in range(n_views):
        az = 2.0 * math.pi * v / max(n_views, 1) + rng.uniform(-0.2, 0.2)
        rad = 4.5 + rng.uniform(-0.3, 0.3)
        eye = np.array([rad * math.cos(az), rad * math.sin(az), 1.5])
        fwd = -eye / np.linalg.norm(eye)
        up = np.array([0.0, 0.0, 1.0])
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        rot = np.stack([right, down, fwd], 0)  # rows = camera axes in world coords
        viewmats[v, :3, :3] = rot
        viewmats[v, :3, 3] = -rot @ eye
        viewmats[v, 3, 3] = 1.0
    f = 0.5 * width / math.tan(math.radians(25.0))
    K = np.array([[f, 0, width / 2.0], [0, f, height / 2.0], [0, 0, 1]], np.float32)
    return viewmats, K


def make_feature_map_np(view: int, d: int, height: int, width: int, seed: int = 0, enc_res: int = 24):
    """Small-scene feature map, numpy only: randn(d, enc, enc) -> L2-normalise over d ->
    bilinear upsample (align_corners=False, like F.interpolate in backproject.py:110-112)
    -> returned as the [H,W,d] *permuted view* of a channel-planar buffer
    (backproject.py:113 `feats.permute(1, 2, 0)`)."""
    rng = np.random.Generator(np.random.PCG64(seed * 1000003 + view * 101 + d))
    low = rng.standard_normal((d, enc_res, enc_res)).astype(np.float32)
    low /= np.maximum(np.linalg.norm(low, axis=0, keepdims=True), 1e-12)

    def src_index(n_out, n_in):
        x = (np.arange(n_out, dtype=np.float64) + 0.5) * (n_in / n_out) - 0.5
        x = np.clip(x, 0.0, None)
        i0 = np.minimum(np.floor(x).astype(np.int64), n_in - 1)
        i1 = np.minimum(i0 + 1, n_in - 1)
        w1 = (x - i0).astype(np.float32)
        return i0, i1, w1

    y0, y1, wy = src_index(height, enc_res)
    x0, x1, wx = src_index(width, enc_res)
    top = low[:, y0][:, :, x0] * (1 - wx) + low[:, y0][:, :, x1] * wx
    bot = low[:, y1][:, :, x0] * (1 - wx) + low[:, y1][:, :, x1] * wx
    planar = np.ascontiguousarray((top * (1 - wy)[None, :, None] + bot * wy[None, :, None]).astype(np.float32))  # [d,H,W]
    return np.transpose(planar, (1, 2, 0))  # strided view, NOT contiguous


def make_feature_map_torch(view: int, d: int, height: int, width: int, device, seed: int = 0, enc_res: int = 240,
                           mode: str = "bilinear"):
    """Benchmark-scale feature map built on `device` with torch (2.2 GB at 1297x840x512 —
    too big to ship from the host every time).  LSeg-shaped: encoder-resolution randn,
    L2-normalised over channels, bilinear-upsampled, exposed as the permuted view."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed * 1000003 + view * 101 + d)
    low = torch.randn(1, d, enc_res, enc_res, generator=g, device=device, dtype=torch.float32)
    low = torch.nn.functional.normalize(low, dim=1)
    up = torch.nn.functional.interpolate(low, size=(height, width), mode=mode)[0]
    return up.permute(1, 2, 0)


def make_text_queries(p: int, d: int, seed: int = 0) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed + 424243))
    t = rng.standard_normal((p, d)).astype(np.float32)
    return t / np.linalg.norm(t, axis=1, keepdims=True)


# BASELINE.json configs (name -> shape).  "S" is the CPU-runnable parity config,
# "G" the north-star benchmark shape.
CONFIGS = {
    "S": dict(n=50_000, views=8, width=256, height=256, d=64),
    "G": dict(n=5_800_000, views=185, width=1297, height=840, d=512),
    "C": dict(n=5_800_000, views=185, width=1297, height=840, d=64),
    "C16": dict(n=5_800_000, views=185, width=1297, height=840, d=16),
    "M": dict(n=6_000_000, views=1000, width=1920, height=1080, d=768),
    # the reference's second model family (backproject.py:214-289): DINOv2 1024-d tokens, 64 x 64, nearest up-sampling
    "D": dict(n=5_800_000, views=185, width=1297, height=840, d=1024, enc=64, mode="nearest"),
}
