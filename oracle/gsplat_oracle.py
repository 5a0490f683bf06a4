"""CPU ORACLE (test infrastructure, not product code) -- numpy restatement of the
gradient-weighted feature back-projection path of JojiJoseph/3dgs-gradient-backprojection.

PARITY UNPINNED: the path's arithmetic lives in the un-vendored third-party wheel
`gsplat==1.4.0` (reference `requirements.txt:1`), which is neither under /root/reference nor
installable here, and the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md §4, §8c).  This file therefore restates gsplat-1.4.0's *published* algorithm
(`gsplat/rendering.py::rasterization`, `gsplat/cuda/csrc/{fully_fused_projection_packed_fwd,
isect_tiles,rasterize_to_pixels_fwd,rasterize_to_pixels_bwd}.cu`, as summarised in SURVEY.md §9)
and anchors on the reference's own call sites and driver math:

  * rasterization(...) call shape ......... backproject.py:89-100,115-125,133-143; utils.py:238-249
  * num += colors.grad, den += grad[:,0] .. backproject.py:127-131,145-151
  * den starts at 1e-12 .................... backproject.py:63
  * finalise (num/den, L2-norm, NaN->0) .... backproject.py:166-169
  * prune mask == (sum_v |grad| > 0) ....... utils.py:236-257
  * 3-D mask ............................... segment.py:52-58
  * 2-D mask ............................... segment.py:209-224

In the absence of reference-owned vectors the restatement is checked against known answers worked
out by hand from that published arithmetic (tests/test_oracle.py::
test_known_answers_from_published_semantics: EWA blur, radius, tile rectangle, alpha clamp, the
1/255 skip, the 1e-4 stop rule, the frustum clamp of the Jacobian), against its C twin
(bit-exact integer stages) and against the algebraic invariants the reference relies on.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path never does.

Precision contract
------------------
Projection + tile binning are restated operation-by-operation in IEEE fp32 *without* fused
multiply-add, in a fixed evaluation order (see `project`).  The CUDA projection kernel is
compiled with -fmad=false and follows the same order, so radii, tile ranges, depth bits,
isect_ids, flatten_ids and isect_offsets are compared BIT-EXACT.  Compositing is fp32 by
default ("mirror") with an fp64 switch ("truth"); the per-Gaussian contraction sum_p w*F is
always accumulated in fp64 so the oracle is the more accurate side of every comparison.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
TILE = 16
ALPHA_MIN = F32(1.0 / 255.0)
ALPHA_MAX = F32(0.999)
T_MIN = F32(1e-4)


# --------------------------------------------------------------------------------------
# 1. projection  (gsplat fully_fused_projection_packed_fwd)
# --------------------------------------------------------------------------------------
def quat_scale_to_covar(quats: np.ndarray, scales: np.ndarray) -> np.ndarray:
    """[N,4] wxyz (un-normalised) + [N,3] -> 6 unique entries of R S S^T R^T, fp32, fixed order.
    Returns [N,6] = (c00,c01,c02,c11,c12,c22)."""
    q = quats.astype(F32, copy=False)
    s = scales.astype(F32, copy=False)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    with np.errstate(all="ignore"):
        n2 = ((w * w + x * x) + y * y) + z * z
        inv = F32(1.0) / np.sqrt(n2)
        w, x, y, z = w * inv, x * inv, y * inv, z * inv
        x2, y2, z2 = x * x, y * y, z * z
        xy, xz, yz = x * y, x * z, y * z
        wx, wy, wz = w * x, w * y, w * z
        one, two = F32(1.0), F32(2.0)
        R = [
            [one - two * (y2 + z2), two * (xy - wz), two * (xz + wy)],
            [two * (xy + wz), one - two * (x2 + z2), two * (yz - wx)],
            [two * (xz - wy), two * (yz + wx), one - two * (x2 + y2)],
        ]
        M = [[R[i][j] * s[:, j] for j in range(3)] for i in range(3)]

        def dot(i, j):
            return (M[i][0] * M[j][0] + M[i][1] * M[j][1]) + M[i][2] * M[j][2]

        cov = np.stack([dot(0, 0), dot(0, 1), dot(0, 2), dot(1, 1), dot(1, 2), dot(2, 2)], 1)
    return cov.astype(F32)


def project(means, covars6, viewmat, K, width, height, near_plane=0.01, far_plane=1e10,
            radius_clip=0.0, eps2d=0.3):
    """EWA projection of N Gaussians for ONE camera.  Returns a dict of *unpacked* [N] arrays:
    radii int32 (0 = culled), means2d [N,2], depths [N], conics [N,3], all fp32.
    Every line is one IEEE-fp32 rounding; the CUDA kernel mirrors it 1:1 (csrc/project.cu)."""
    m = means.astype(F32, copy=False)
    c = covars6.astype(F32, copy=False)
    V = viewmat.astype(F32)
    fx, fy, cx, cy = F32(K[0, 0]), F32(K[1, 1]), F32(K[0, 2]), F32(K[1, 2])
    Wf, Hf = F32(width), F32(height)
    mx, my, mz = m[:, 0], m[:, 1], m[:, 2]
    with np.errstate(all="ignore"):
        # world -> camera
        p = [((V[i, 0] * mx + V[i, 1] * my) + V[i, 2] * mz) + V[i, 3] for i in range(3)]
        S = [[c[:, 0], c[:, 1], c[:, 2]], [c[:, 1], c[:, 3], c[:, 4]], [c[:, 2], c[:, 4], c[:, 5]]]
        Tm = [[(V[i, 0] * S[0][j] + V[i, 1] * S[1][j]) + V[i, 2] * S[2][j] for j in range(3)] for i in range(3)]

        def cc(i, j):
            return (Tm[i][0] * V[j, 0] + Tm[i][1] * V[j, 1]) + Tm[i][2] * V[j, 2]

        C00, C01, C02, C11, C12, C22 = cc(0, 0), cc(0, 1), cc(0, 2), cc(1, 1), cc(1, 2), cc(2, 2)
        x, y, z = p
        # perspective (pinhole) with the 1.4.0 asymmetric frustum clamp
        tanx = (F32(0.5) * Wf) / fx
        tany = (F32(0.5) * Hf) / fy
        lim_xp = (Wf - cx) / fx + F32(0.3) * tanx
        lim_xn = cx / fx + F32(0.3) * tanx
        lim_yp = (Hf - cy) / fy + F32(0.3) * tany
        lim_yn = cy / fy + F32(0.3) * tany
        rz = F32(1.0) / z
        rz2 = rz * rz
        tx = z * np.minimum(lim_xp, np.maximum(-lim_xn, x * rz))
        ty = z * np.minimum(lim_yp, np.maximum(-lim_yn, y * rz))
        J00 = fx * rz
        J11 = fy * rz
        J02 = -((fx * tx) * rz2)
        J12 = -((fy * ty) * rz2)
        a0 = J00 * C00 + J02 * C02
        a1 = J00 * C01 + J02 * C12
        a2 = J00 * C02 + J02 * C22
        b1 = J11 * C11 + J12 * C12
        b2 = J11 * C12 + J12 * C22
        s00 = a0 * J00 + a2 * J02
        s01 = a1 * J11 + a2 * J12
        s11 = b1 * J11 + b2 * J12
        m2x = (fx * x) * rz + cx
        m2y = (fy * y) * rz + cy
        # blur + inverse
        e = F32(eps2d)
        A = s00 + e
        Cc = s11 + e
        det = A * Cc - s01 * s01
        inv_det = F32(1.0) / det
        con_x = Cc * inv_det
        con_y = -(s01 * inv_det)
        con_z = A * inv_det
        b = F32(0.5) * (A + Cc)
        v1 = b + np.sqrt(np.maximum(F32(0.01), b * b - det))
        rad_f = np.ceil(F32(3.0) * np.sqrt(v1))
        ok = (z >= F32(near_plane)) & (z <= F32(far_plane)) & (det > 0) & np.isfinite(rad_f)
        ok &= rad_f > F32(radius_clip)
        ok &= (m2x + rad_f > 0) & (m2x - rad_f < Wf) & (m2y + rad_f > 0) & (m2y - rad_f < Hf)
        # non-finite geometry never survives (gsplat would emit NaNs; we cull)
        ok &= np.isfinite(m2x) & np.isfinite(m2y) & np.isfinite(con_x) & np.isfinite(con_y) & np.isfinite(con_z)
        rad_f = np.where(ok, np.minimum(rad_f, F32(1 << 24)), F32(0))
    radii = rad_f.astype(np.int32)
    return dict(
        radii=radii,
        means2d=np.stack([m2x, m2y], 1).astype(F32),
        depths=z.astype(F32),
        conics=np.stack([con_x, con_y, con_z], 1).astype(F32),
    )


# --------------------------------------------------------------------------------------
# 2. tile binning + sort + offsets  (gsplat isect_tiles / radix sort / isect_offset_encode; §9.3)
# --------------------------------------------------------------------------------------
def tile_bounds(means2d, radii, width, height):
    tw = (width + TILE - 1) // TILE
    th = (height + TILE - 1) // TILE
    tr = radii.astype(F32) / F32(TILE)
    txc = means2d[:, 0] / F32(TILE)
    tyc = means2d[:, 1] / F32(TILE)
    x0 = np.clip(np.floor(txc - tr), 0, tw).astype(np.int32)
    x1 = np.clip(np.ceil(txc + tr), 0, tw).astype(np.int32)
    y0 = np.clip(np.floor(tyc - tr), 0, th).astype(np.int32)
    y1 = np.clip(np.ceil(tyc + tr), 0, th).astype(np.int32)
    return x0, x1, y0, y1, tw, th


def ln_approx(x):
    """Fixed-order fp32 series for ln(x), x > 0 (no libm, so CPU and GPU agree bit for bit):
    x = m 2^e, s = (m-1)/(m+1), ln x = e ln2 + 2 s (1 + s^2/3 + s^4/5 + s^6/7)."""
    x = np.asarray(x, F32)
    bits = x.view(np.uint32)
    e = ((bits >> 23) & 0xFF).astype(np.int32) - 127
    m = ((bits & np.uint32(0x7FFFFF)) | np.uint32(0x3F800000)).view(F32)
    s = (m - F32(1.0)) / (m + F32(1.0))
    s2 = s * s
    p = ((s2 * F32(1.0 / 7.0) + F32(0.2)) * s2 + F32(1.0 / 3.0)) * s2 + F32(1.0)
    return e.astype(F32) * F32(0.69314718) + (F32(2.0) * s) * p


def tile_hit(gx, gy, A, B, C, op, tx, ty, width, height):
    """Exact tile culling (extension, see oracle.c): can alpha reach 1/255 anywhere on the tile's
    pixel-centre box?  All arguments are per-pair fp32 / int arrays."""
    with np.errstate(all="ignore"):
        x255 = F32(255.0) * op
        tau = np.where(x255 > 1, ln_approx(np.where(x255 > 1, x255, F32(2.0))) + F32(0.01), F32(-1.0)).astype(F32)
        xe = np.minimum(tx * TILE + TILE - 1, width - 1)
        ye = np.minimum(ty * TILE + TILE - 1, height - 1)
        X0, X1 = (tx * TILE).astype(F32) + F32(0.5), xe.astype(F32) + F32(0.5)
        Y0, Y1 = (ty * TILE).astype(F32) + F32(0.5), ye.astype(F32) + F32(0.5)
        dx0, dx1, dy0, dy1 = gx - X1, gx - X0, gy - Y1, gy - Y0
        inside = (dx0 <= 0) & (dx1 >= 0) & (dy0 <= 0) & (dy1 >= 0)

        def q(dx, dy):
            return F32(0.5) * ((A * dx) * dx + (C * dy) * dy) + (B * dx) * dy

        def clamp(t, lo, hi):
            return np.minimum(np.maximum(t, lo), hi)

        kx, ky = -(B / C), -(B / A)  # argmin of the quadratic along an edge: dy = kx*dx, dx = ky*dy
        best = q(dx0, clamp(kx * dx0, dy0, dy1))
        best = np.minimum(best, q(dx1, clamp(kx * dx1, dy0, dy1)))
        best = np.minimum(best, q(clamp(ky * dy0, dx0, dx1), dy0))
        best = np.minimum(best, q(clamp(ky * dy1, dx0, dx1), dy1))
    return (tau >= 0) & (inside | (best <= tau))


def isect_tiles(proj, width, height, opacities=None, cull=False):
    """Packed (visible-only, ascending Gaussian index) intersection list, sorted by
    (tile, depth bits) with a STABLE sort (= cub::DeviceRadixSort).  Returns dict with
    gaussian_ids[nnz], tiles_per_gauss[nnz], isect_ids[I] int64, flatten_ids[I] int32
    (index into the packed arrays), isect_offsets[th,tw] int32."""
    gids = np.nonzero(proj["radii"] > 0)[0].astype(np.int32)
    m2 = proj["means2d"][gids]
    rad = proj["radii"][gids]
    dep = proj["depths"][gids]
    x0, x1, y0, y1, tw, th = tile_bounds(m2, rad, width, height)
    tpg = ((y1 - y0) * (x1 - x0)).astype(np.int32)
    total = int(tpg.sum())
    flat = np.repeat(np.arange(len(gids), dtype=np.int32), tpg)
    start = np.cumsum(tpg) - tpg
    local = np.arange(total, dtype=np.int64) - np.repeat(start.astype(np.int64), tpg)
    bw = np.repeat((x1 - x0).astype(np.int64), tpg)
    ty = np.repeat(y0.astype(np.int64), tpg) + local // np.maximum(bw, 1)
    tx = np.repeat(x0.astype(np.int64), tpg) + local % np.maximum(bw, 1)
    tile_id = ty * tw + tx
    depth_bits = np.repeat(dep.view(np.int32).astype(np.int64), tpg)
    if cull:
        con = proj["conics"][gids][flat]
        hit = tile_hit(m2[flat, 0], m2[flat, 1], con[:, 0], con[:, 1], con[:, 2],
                       np.asarray(opacities, F32)[gids][flat], tx, ty, width, height)
        tile_id, depth_bits, flat = tile_id[hit], depth_bits[hit], flat[hit]
        tpg = np.bincount(flat, minlength=len(gids)).astype(np.int32)
        total = int(hit.sum())
    keys = (tile_id << 32) | depth_bits
    order = np.argsort(keys, kind="stable")
    isect_ids = keys[order]
    flatten_ids = flat[order]
    tiles = (isect_ids >> 32).astype(np.int64)
    offsets = np.searchsorted(tiles, np.arange(tw * th, dtype=np.int64), side="left").astype(np.int32)
    return dict(gaussian_ids=gids, tiles_per_gauss=tpg, isect_ids=isect_ids, flatten_ids=flatten_ids,
                isect_offsets=offsets.reshape(th, tw), tile_width=tw, tile_height=th, n_isects=total)


# --------------------------------------------------------------------------------------
# 3. compositing weights  (gsplat rasterize_to_pixels_fwd loop; §9.4)
# --------------------------------------------------------------------------------------
def _tile_weights(xy, conic, opac, ty, tx, width, height, dtype):
    """w[k, p] = alpha_k(p) * T_k(p) for the K Gaussians of one tile in sorted order and the 256
    pixels of tile (ty,tx) (p = 16*row + col).  Vectorised restatement of the sequential loop:
    skip if sigma<0 or alpha<1/255; stop (without compositing) at the first Gaussian whose
    T*(1-alpha) <= 1e-4.  Returns (w [K,256], T_final [256], n_used = #list entries the CTA would
    have touched before every pixel was done)."""
    f = dtype
    ii, jj = np.meshgrid(np.arange(TILE), np.arange(TILE), indexing="ij")
    py = (ty * TILE + ii).reshape(-1)
    px = (tx * TILE + jj).reshape(-1)
    inside = (py < height) & (px < width)
    pxf = px.astype(f) + f(0.5)
    pyf = py.astype(f) + f(0.5)
    dx = xy[:, 0:1].astype(f) - pxf[None, :]
    dy = xy[:, 1:2].astype(f) - pyf[None, :]
    cxx, cxy, cyy = (conic[:, i:i + 1].astype(f) for i in range(3))
    sigma = f(0.5) * ((cxx * dx) * dx + (cyy * dy) * dy) + (cxy * dx) * dy
    with np.errstate(over="ignore", under="ignore"):
        alpha = np.minimum(f(0.999), opac[:, None].astype(f) * np.exp(-sigma))
    valid = (sigma >= 0) & (alpha >= f(1.0 / 255.0)) & inside[None, :]
    factor = np.where(valid, f(1.0) - alpha, f(1.0))
    t_inc = np.multiply.accumulate(factor, axis=0, dtype=f)  # sequential products, like the kernel
    stop = valid & (t_inc <= f(1e-4))
    any_stop = stop.any(axis=0)
    k_stop = np.where(any_stop, stop.argmax(axis=0), xy.shape[0])
    karange = np.arange(xy.shape[0])[:, None]
    contrib = valid & (karange < k_stop[None, :])
    t_prev = np.concatenate([np.ones((1, 256), f), t_inc[:-1]], 0)
    w = np.where(contrib, alpha * t_prev, f(0.0))
    # transmittance after the last composited Gaussian
    last = np.where(contrib, karange, -1).max(axis=0)
    t_final = np.where(last >= 0, t_inc[np.maximum(last, 0), np.arange(256)], f(1.0))
    t_final = np.where(inside, t_final, f(1.0))
    done_at = np.where(inside, k_stop, 0)
    n_used = int(done_at.max()) if xy.shape[0] else 0
    return w, t_final, inside, n_used


def iter_tiles(proj, isect, opacities, width, height, dtype=F32):
    """Yield (ty, tx, packed_rows[K], w[K,256], T_final[256], inside[256]) for every non-empty tile."""
    gids = isect["gaussian_ids"]
    xy = proj["means2d"][gids]
    con = proj["conics"][gids]
    op = opacities[gids].astype(F32)
    off = isect["isect_offsets"].reshape(-1)
    total = isect["n_isects"]
    tw, th = isect["tile_width"], isect["tile_height"]
    flat = isect["flatten_ids"]
    for t in range(tw * th):
        s = int(off[t])
        e = int(off[t + 1]) if t + 1 < tw * th else total
        if e <= s:
            continue
        rows = flat[s:e]
        w, t_final, inside, n_used = _tile_weights(xy[rows], con[rows], op[rows], t // tw, t % tw, width, height, dtype)
        yield t // tw, t % tw, rows, w, t_final, inside, n_used


def _tile_pixels(ty, tx, width, height):
    ii, jj = np.meshgrid(np.arange(TILE), np.arange(TILE), indexing="ij")
    py = np.minimum(ty * TILE + ii.reshape(-1), height - 1)
    px = np.minimum(tx * TILE + jj.reshape(-1), width - 1)
    return py, px


# --------------------------------------------------------------------------------------
# 4. the hot path: per-view back-projection == d/d(colors) of <render(colors), F>
#    (backproject.py:115-151 through gsplat's rasterize_to_pixels_bwd v_colors; §9.5-9.6)
# --------------------------------------------------------------------------------------
def view_geometry(means, quats, scales, viewmat, K, width, height, opacities=None, cull=False, **kw):
    cov = quat_scale_to_covar(quats, scales)
    proj = project(means, cov, viewmat, K, width, height, **kw)
    isect = isect_tiles(proj, width, height, opacities, cull)
    return proj, isect


def backproject_view(means, quats, scales, opacities, viewmat, K, width, height, feats, dtype=F32,
                     stats=None, cull=False, **kw):
    """num_v[g,:] = sum_p w(g,p) F[p,:],  den_v[g] = sum_p w(g,p)  for one view.
    `feats` is [H,W,D] (any strides).  Returns (num_v [N,D] fp64, den_v [N] fp64)."""
    n = means.shape[0]
    d = feats.shape[2]
    proj, isect = view_geometry(means, quats, scales, viewmat, K, width, height, opacities, cull, **kw)
    num = np.zeros((n, d), np.float64)
    den = np.zeros(n, np.float64)
    gids = isect["gaussian_ids"]
    rows_nz = pairs = used = 0
    for ty, tx, rows, w, _tf, _ins, n_used in iter_tiles(proj, isect, opacities, width, height, dtype):
        py, px = _tile_pixels(ty, tx, width, height)
        ftile = feats[py, px, :].astype(np.float64)  # [256, D]; out-of-image pixels have w == 0
        nz = np.nonzero(w.any(axis=1))[0]
        if nz.size == 0:
            used += n_used
            continue
        wn = w[nz].astype(np.float64)
        g = gids[rows[nz]]
        np.add.at(num, g, wn @ ftile)
        np.add.at(den, g, wn.sum(axis=1))
        rows_nz += nz.size
        pairs += int((wn > 0).sum())
        used += n_used
    if stats is not None:
        stats.update(n_vis=int(gids.size), n_isects=int(isect["n_isects"]), rows_nonzero=rows_nz,
                     pairs=pairs, entries_walked=used)
    return num, den


def render_view(means, quats, scales, opacities, colors, viewmat, K, width, height, dtype=F32,
                backgrounds=None, **kw):
    """Forward D-channel render (segment.py:209-220 through rasterize_to_pixels_fwd; §9.4).
    Returns (render [H,W,D] fp64, alpha [H,W] fp64)."""
    d = colors.shape[1]
    proj, isect = view_geometry(means, quats, scales, viewmat, K, width, height, **kw)
    out = np.zeros((height, width, d), np.float64)
    alpha = np.zeros((height, width), np.float64)
    gids = isect["gaussian_ids"]
    for ty, tx, rows, w, t_final, inside, _ in iter_tiles(proj, isect, opacities, width, height, dtype):
        py, px = _tile_pixels(ty, tx, width, height)
        val = w.astype(np.float64).T @ colors[gids[rows]].astype(np.float64)  # [256, D]
        out[py[inside], px[inside]] = val[inside]
        alpha[py[inside], px[inside]] = 1.0 - t_final[inside].astype(np.float64)
    if backgrounds is not None:
        out += (1.0 - alpha)[..., None] * np.asarray(backgrounds, np.float64).reshape(1, 1, d)
    return out, alpha


def backproject(means, quats, scales, opacities, viewmats, K, width, height, feature_fn, d, dtype=F32):
    """Whole job: the loop of create_feature_field_lseg (backproject.py:62-63,74-165)."""
    n = means.shape[0]
    num = np.zeros((n, d), np.float64)
    den = np.full(n, 1e-12, np.float64)  # backproject.py:63
    for v in range(viewmats.shape[0]):
        nv, dv = backproject_view(means, quats, scales, opacities, viewmats[v], K, width, height,
                                  feature_fn(v), dtype)
        num += nv
        den += dv
    return num, den


def finalize(num, den):
    """backproject.py:166-169: f = num/den; f /= ||f||; NaN -> 0.  fp64 in, fp64 out."""
    with np.errstate(all="ignore"):
        f = num / den[:, None]
        f = f / np.linalg.norm(f, axis=-1, keepdims=True)
    f[np.isnan(f)] = 0.0
    return f


def ratio_backproject(views_num_den, width, height, d, eps=1e-12):
    """Per-view-ratio accumulation of affordance_transfer/demo_affordance_transfer.py:768-800:
    both losses are `.mean()`s, so grad = num_v/(H*W*D) and grad0[:,0] = den_v/(H*W*3) (3-channel ones render);
    features = sum_v grad / (grad0[:,0:1] + 1e-12), then L2-normalised rows (:800).
    views_num_den: iterable of per-view (num_v [N,D], den_v [N]) as returned by backproject_view.  fp64."""
    acc = None
    hw = float(width) * float(height)
    for num_v, den_v in views_num_den:
        g = num_v.astype(np.float64) / (hw * d)
        g0 = den_v.astype(np.float64) / (hw * 3.0)
        r = g / (g0[:, None] + eps)
        acc = r if acc is None else acc + r
    with np.errstate(all="ignore"):
        f = acc / np.linalg.norm(acc, axis=-1, keepdims=True)
    return acc, f  # rows never seen are NaN in f, as in the reference


def prune_mask(den_without_eps):
    """utils.py:236-257: a Gaussian survives iff its accumulated colour gradient is non-zero in
    at least one view, i.e. iff sum_v sum_p w > 0."""
    return den_without_eps > 0


# --------------------------------------------------------------------------------------
# 5. query side  (segment.py:52-58, 221-224)
# --------------------------------------------------------------------------------------
def _normalize(x, axis, eps=1e-12):
    return x / np.maximum(np.linalg.norm(x, axis=axis, keepdims=True), eps)


def mask3d(features, text, n_pos, threshold=None):
    score = _normalize(features.astype(np.float64), 1) @ _normalize(text.astype(np.float64), 1).T
    m = score[:, :n_pos].max(axis=1) > score[:, n_pos:].max(axis=1)
    if threshold is not None:
        m = m & (score[:, 0] > threshold)
    return m, score


def click_prompt(render_rgbd, viewmat, K, xy):
    """click_and_segment.py:254-275 for one clicked pixel: `render_rgbd` [H,W,D+1] is the RGB+D render
    (features + camera depth as last channel).  Returns (normalised prompt feature [D], world point [3])."""
    x, y = int(xy[0]), int(xy[1])
    out, Z = render_rgbd[y, x, :-1].astype(np.float64), float(render_rgbd[y, x, -1])
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    cam = np.array([(x - cx) / fx * Z, (y - cy) / fy * Z, Z, 1.0])
    world = np.linalg.inv(viewmat.astype(np.float64)) @ cam
    return _normalize(out[None], 1)[0], world[:3]


def mask2d(render, text, n_pos):
    r = _normalize(render.astype(np.float64), -1)
    score = r @ _normalize(text.astype(np.float64), 1).T
    return score[..., :n_pos].max(axis=2) > score[..., n_pos:].max(axis=2), score


# --------------------------------------------------------------------------------------
# 6. encoder-resolution feature maps  (backproject.py:110-112 bilinear, :245-249 nearest)
# --------------------------------------------------------------------------------------
def upsample(feats_low, height, width, mode="bilinear"):
    """torch.nn.functional.interpolate(x[1,D,h,w], size=(H,W), mode=mode) restated in fp32 numpy (align_corners=False:
    src = scale*(dst+0.5)-0.5 clamped at 0, scale = in/out; nearest: floor(dst*scale)).  feats_low: [h,w,D] -> [H,W,D].
    Checked against torch's own CPU kernel in tests/test_oracle.py."""
    f = feats_low.astype(F32, copy=False)
    h, w = f.shape[:2]

    def src(n_out, n_in):
        scale = F32(n_in) / F32(n_out)
        o = np.arange(n_out, dtype=F32)
        if mode == "nearest":
            i0 = np.minimum(np.floor(o * scale).astype(np.int64), n_in - 1)
            return i0, i0, np.zeros(n_out, F32)
        s = np.maximum(scale * (o + F32(0.5)) - F32(0.5), F32(0.0))
        i0 = np.minimum(s.astype(np.int64), n_in - 1)
        i1 = i0 + (i0 < n_in - 1)
        return i0, i1, (s - i0.astype(F32)).astype(F32)

    assert mode in ("bilinear", "nearest"), mode
    y0, y1, ly = src(height, h)
    x0, x1, lx = src(width, w)
    if mode == "nearest":
        return f[y0][:, x0]
    lx = lx[None, :, None]
    ly = ly[:, None, None]
    top = (F32(1.0) - lx) * f[y0][:, x0] + lx * f[y0][:, x1]
    bot = (F32(1.0) - lx) * f[y1][:, x0] + lx * f[y1][:, x1]
    return ((F32(1.0) - ly) * top + ly * bot).astype(F32)


# --------------------------------------------------------------------------------------
# 7. spherical-harmonics colours  (rasterization(sh_degree=3): backproject.py:88-100)
# --------------------------------------------------------------------------------------
def sh_basis(degree, dirs):
    """Real spherical harmonics up to `degree` from FIRST PRINCIPLES (associated Legendre functions with the
    Condon-Shortley phase, scipy.special.lpmv) -- deliberately not the hard-coded polynomial tables the CUDA kernel
    uses, so the two are independent.  3DGS / gsplat order: index l*l + l + m, sign convention
    Y_1 = (-C1 y, C1 z, -C1 x).  dirs [N,3] (normalised here).  Returns [N, (degree+1)^2] fp64."""
    from math import factorial, pi, sqrt

    from scipy.special import lpmv

    d = dirs.astype(np.float64)
    d = d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-30)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    phi = np.arctan2(y, x)
    ct = np.clip(z, -1.0, 1.0)
    out = np.zeros((d.shape[0], (degree + 1) ** 2))
    for l in range(degree + 1):
        for m in range(-l, l + 1):
            am = abs(m)
            k = sqrt((2 * l + 1) / (4 * pi) * factorial(l - am) / factorial(l + am))
            p = lpmv(am, l, ct)
            if m == 0:
                v = k * p
            elif m > 0:
                v = sqrt(2.0) * k * np.cos(am * phi) * p
            else:
                v = sqrt(2.0) * k * np.sin(am * phi) * p
            out[:, l * l + l + m] = v
    return out


def sh_colors(degree, means, coeffs, viewmat):
    """colour[g] = max(sum_k Y_k(dir_g) * coeffs[g,k,:] + 0.5, 0), dir_g = mean_g - camera position
    (gsplat-1.4.0 rendering.py: `dirs = means - camtoworlds[:, :3, 3]`, `clamp_min(colors + 0.5, 0)`).  fp64."""
    vm = viewmat.astype(np.float64)
    cam_pos = -(vm[:3, :3].T @ vm[:3, 3])
    basis = sh_basis(degree, means.astype(np.float64) - cam_pos[None])
    k = (degree + 1) ** 2
    col = np.einsum("nk,nkc->nc", basis, coeffs[:, :k].astype(np.float64))
    return np.maximum(col + 0.5, 0.0)
