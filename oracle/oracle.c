/* CPU ORACLE (test infrastructure, not product code) -- plain-C restatement of the
 * gradient-weighted feature back-projection path; the threaded twin of oracle/gsplat_oracle.py.
 *
 * PARITY UNPINNED: the arithmetic lives in the un-vendored wheel gsplat==1.4.0 (reference
 * requirements.txt:1); the reference has no tests / golden vectors for this path.
 * This file restates the published gsplat-1.4.0 algorithm and the reference's
 * driver math:
 *   rasterization(...) call sites ........ backproject.py:89-100,115-125,133-143
 *   num += grad, den += grad[:,0] ........ backproject.py:127-131,145-151
 *   forward feature render ............... segment.py:209-220
 * It is validated against hand-derived known answers and against gsplat_oracle.py (tests/test_oracle.py) and is the timed
 * "cpu_baseline" (kind "port") of bench.py.  Only tests/, __graft_entry__.smoke() and bench.py
 * may load it.
 *
 * Build:  make -C oracle        (gcc -O2 -fopenmp -ffp-contract=off, NO -ffast-math: projection
 *                                 and binning must round exactly like the numpy oracle and the
 *                                 -fmad=false CUDA kernel -- those stages are compared bit-exact)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16
#define NPIX 256

typedef struct OrcView {
    int n, W, H, tw, th, cull;
    int64_t n_vis, n_isects;
    /* unpacked [n] */
    int32_t *radii;
    float *means2d, *depths, *conics;
    const float *opac; /* borrowed */
    /* packed */
    int32_t *gaussian_ids; /* [n_vis] */
    int64_t *isect_ids;    /* [I] sorted */
    int32_t *flatten_ids;  /* [I] index into packed arrays */
    int32_t *offsets;      /* [th*tw] */
} OrcView;

/* quat (wxyz, un-normalised) + scale -> 6 unique entries of R S S^T R^T */
void orc_covar(int n, const float *quats, const float *scales, float *cov6) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        float w = quats[4 * i], x = quats[4 * i + 1], y = quats[4 * i + 2], z = quats[4 * i + 3];
        float n2 = ((w * w + x * x) + y * y) + z * z;
        float inv = 1.0f / sqrtf(n2);
        w *= inv; x *= inv; y *= inv; z *= inv;
        float x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z;
        float wx = w * x, wy = w * y, wz = w * z;
        float R[3][3] = {
            {1.0f - 2.0f * (y2 + z2), 2.0f * (xy - wz), 2.0f * (xz + wy)},
            {2.0f * (xy + wz), 1.0f - 2.0f * (x2 + z2), 2.0f * (yz - wx)},
            {2.0f * (xz - wy), 2.0f * (yz + wx), 1.0f - 2.0f * (x2 + y2)}};
        float M[3][3];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) M[r][c] = R[r][c] * scales[3 * i + c];
#define DOT(a, b) ((M[a][0] * M[b][0] + M[a][1] * M[b][1]) + M[a][2] * M[b][2])
        float *o = cov6 + 6 * (size_t)i;
        o[0] = DOT(0, 0); o[1] = DOT(0, 1); o[2] = DOT(0, 2);
        o[3] = DOT(1, 1); o[4] = DOT(1, 2); o[5] = DOT(2, 2);
#undef DOT
    }
}

/* EWA projection, one camera.  Same evaluation order as gsplat_oracle.project. */
static void project_all(OrcView *v, const float *means, const float *cov6, const float *V /*4x4*/,
                        const float *K /*3x3*/, float near_plane, float far_plane, float radius_clip,
                        float eps2d) {
    const float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const float Wf = (float)v->W, Hf = (float)v->H;
    const float tanx = (0.5f * Wf) / fx, tany = (0.5f * Hf) / fy;
    const float lim_xp = (Wf - cx) / fx + 0.3f * tanx, lim_xn = cx / fx + 0.3f * tanx;
    const float lim_yp = (Hf - cy) / fy + 0.3f * tany, lim_yn = cy / fy + 0.3f * tany;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < v->n; ++i) {
        const float mx = means[3 * (size_t)i], my = means[3 * (size_t)i + 1], mz = means[3 * (size_t)i + 2];
        const float *c = cov6 + 6 * (size_t)i;
        float p[3];
        for (int r = 0; r < 3; ++r) p[r] = ((V[4 * r] * mx + V[4 * r + 1] * my) + V[4 * r + 2] * mz) + V[4 * r + 3];
        const float S[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}};
        float T[3][3];
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 3; ++k)
                T[r][k] = (V[4 * r] * S[0][k] + V[4 * r + 1] * S[1][k]) + V[4 * r + 2] * S[2][k];
#define CC(a, b) ((T[a][0] * V[4 * b] + T[a][1] * V[4 * b + 1]) + T[a][2] * V[4 * b + 2])
        const float C00 = CC(0, 0), C01 = CC(0, 1), C02 = CC(0, 2), C11 = CC(1, 1), C12 = CC(1, 2), C22 = CC(2, 2);
#undef CC
        const float x = p[0], y = p[1], z = p[2];
        const float rz = 1.0f / z, rz2 = rz * rz;
        const float tx = z * fminf(lim_xp, fmaxf(-lim_xn, x * rz));
        const float ty = z * fminf(lim_yp, fmaxf(-lim_yn, y * rz));
        const float J00 = fx * rz, J11 = fy * rz;
        const float J02 = -((fx * tx) * rz2), J12 = -((fy * ty) * rz2);
        const float a0 = J00 * C00 + J02 * C02, a1 = J00 * C01 + J02 * C12, a2 = J00 * C02 + J02 * C22;
        const float b1 = J11 * C11 + J12 * C12, b2 = J11 * C12 + J12 * C22;
        const float s00 = a0 * J00 + a2 * J02, s01 = a1 * J11 + a2 * J12, s11 = b1 * J11 + b2 * J12;
        const float m2x = (fx * x) * rz + cx, m2y = (fy * y) * rz + cy;
        const float A = s00 + eps2d, Cc = s11 + eps2d;
        const float det = A * Cc - s01 * s01;
        const float inv_det = 1.0f / det;
        const float con_x = Cc * inv_det, con_y = -(s01 * inv_det), con_z = A * inv_det;
        const float b = 0.5f * (A + Cc);
        const float v1 = b + sqrtf(fmaxf(0.01f, b * b - det));
        float rad = ceilf(3.0f * sqrtf(v1));
        int ok = (z >= near_plane) && (z <= far_plane) && (det > 0.0f) && isfinite(rad);
        ok = ok && (rad > radius_clip);
        ok = ok && (m2x + rad > 0.0f) && (m2x - rad < Wf) && (m2y + rad > 0.0f) && (m2y - rad < Hf);
        ok = ok && isfinite(m2x) && isfinite(m2y) && isfinite(con_x) && isfinite(con_y) && isfinite(con_z);
        rad = ok ? fminf(rad, 16777216.0f) : 0.0f;
        v->radii[i] = (int32_t)rad;
        v->means2d[2 * (size_t)i] = m2x; v->means2d[2 * (size_t)i + 1] = m2y;
        v->depths[i] = z;
        v->conics[3 * (size_t)i] = con_x; v->conics[3 * (size_t)i + 1] = con_y; v->conics[3 * (size_t)i + 2] = con_z;
    }
}

static inline int clampi(float f, int hi) {
    if (!(f > 0.0f)) return 0;
    if (f >= (float)hi) return hi;
    return (int)f;
}

/* --- optional exact tile culling (an extension over gsplat-1.4.0, OFF by default) -------------------
 * A (Gaussian, tile) pair is dropped when alpha = op*exp(-sigma) < 1/255 on the whole pixel-centre
 * box of the tile, i.e. when min_box sigma > ln(255 op).  Dropped pairs have zero weight on every
 * pixel, so num/den are unchanged; the intersection list shrinks ~2x.  ln() is a fixed-order fp32
 * series (no libm) so that CPU and GPU agree bit for bit; +0.01 keeps the test conservative. */
static inline float orc_ln_approx(float x) { /* x > 0 */
    uint32_t bits; memcpy(&bits, &x, 4);
    const int e = (int)((bits >> 23) & 0xff) - 127;
    const uint32_t mb = (bits & 0x7fffffu) | 0x3f800000u;
    float m; memcpy(&m, &mb, 4);
    const float s = (m - 1.0f) / (m + 1.0f);
    const float s2 = s * s;
    const float p = ((s2 * (1.0f / 7.0f) + 0.2f) * s2 + (1.0f / 3.0f)) * s2 + 1.0f;
    return (float)e * 0.69314718f + (2.0f * s) * p;
}
static inline float orc_tau(float op) {
    const float x = 255.0f * op;
    return (x > 1.0f) ? orc_ln_approx(x) + 0.01f : -1.0f;
}
static inline float orc_q(float A, float B, float C, float dx, float dy) {
    return 0.5f * ((A * dx) * dx + (C * dy) * dy) + (B * dx) * dy;
}
static inline float orc_clampf(float t, float lo, float hi) { return fminf(fmaxf(t, lo), hi); }
static int tile_hit(float gx, float gy, float A, float B, float C, float tau, int tx, int ty, int W, int H) {
    if (tau < 0.0f) return 0;
    const int xe = (tx * TILE + TILE - 1 < W - 1) ? tx * TILE + TILE - 1 : W - 1;
    const int ye = (ty * TILE + TILE - 1 < H - 1) ? ty * TILE + TILE - 1 : H - 1;
    const float X0 = (float)(tx * TILE) + 0.5f, X1 = (float)xe + 0.5f;
    const float Y0 = (float)(ty * TILE) + 0.5f, Y1 = (float)ye + 0.5f;
    const float dx0 = gx - X1, dx1 = gx - X0, dy0 = gy - Y1, dy1 = gy - Y0;
    if (dx0 <= 0.0f && dx1 >= 0.0f && dy0 <= 0.0f && dy1 >= 0.0f) return 1;
    const float kx = -(B / C), ky = -(B / A); /* argmin of the quadratic along an edge: dy = kx*dx, dx = ky*dy */
    float best = orc_q(A, B, C, dx0, orc_clampf(kx * dx0, dy0, dy1));
    best = fminf(best, orc_q(A, B, C, dx1, orc_clampf(kx * dx1, dy0, dy1)));
    best = fminf(best, orc_q(A, B, C, orc_clampf(ky * dy0, dx0, dx1), dy0));
    best = fminf(best, orc_q(A, B, C, orc_clampf(ky * dy1, dx0, dx1), dy1));
    return best <= tau;
}

static void tile_rect(const OrcView *v, int g, int *x0, int *x1, int *y0, int *y1) {
    const float tr = (float)v->radii[g] / (float)TILE;
    const float txc = v->means2d[2 * (size_t)g] / (float)TILE, tyc = v->means2d[2 * (size_t)g + 1] / (float)TILE;
    *x0 = clampi(floorf(txc - tr), v->tw); *x1 = clampi(ceilf(txc + tr), v->tw);
    *y0 = clampi(floorf(tyc - tr), v->th); *y1 = clampi(ceilf(tyc + tr), v->th);
}

/* stable LSD radix sort on 64-bit keys with 32-bit payload, 16-bit digits (== cub::DeviceRadixSort order) */
static void radix_sort_pairs(int64_t *keys, int32_t *vals, int64_t n, int bits) {
    int64_t *k2 = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n ? n : 1));
    int32_t *v2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
    size_t *hist = (size_t *)malloc(sizeof(size_t) * 65536);
    for (int shift = 0; shift < bits; shift += 16) {
        memset(hist, 0, sizeof(size_t) * 65536);
        for (int64_t i = 0; i < n; ++i) hist[((uint64_t)keys[i] >> shift) & 0xffff]++;
        size_t run = 0;
        for (int d = 0; d < 65536; ++d) { size_t c = hist[d]; hist[d] = run; run += c; }
        for (int64_t i = 0; i < n; ++i) {
            size_t pos = hist[((uint64_t)keys[i] >> shift) & 0xffff]++;
            k2[pos] = keys[i]; v2[pos] = vals[i];
        }
        int64_t *tk = keys; keys = k2; k2 = tk;
        int32_t *tv = vals; vals = v2; v2 = tv;
    }
    /* after an even number of passes the data is back in the caller's arrays; otherwise copy */
    int passes = (bits + 15) / 16;
    if (passes & 1) {
        memcpy(k2, keys, sizeof(int64_t) * (size_t)n);
        memcpy(v2, vals, sizeof(int32_t) * (size_t)n);
        free(keys); free(vals);
    } else {
        free(k2); free(v2);
    }
    free(hist);
}

/* isect_tiles + sort + isect_offset_encode */
static void bin_and_sort(OrcView *v) {
    int64_t nvis = 0, total = 0;
    for (int g = 0; g < v->n; ++g) nvis += v->radii[g] > 0;
    v->gaussian_ids = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nvis ? nvis : 1));
    int64_t k = 0;
    for (int g = 0; g < v->n; ++g)
        if (v->radii[g] > 0) {
            int x0, x1, y0, y1;
            tile_rect(v, g, &x0, &x1, &y0, &y1);
            if (!v->cull) {
                total += (int64_t)(y1 - y0) * (x1 - x0);
            } else {
                const float tau = orc_tau(v->opac[g]);
                for (int i = y0; i < y1; ++i)
                    for (int j = x0; j < x1; ++j)
                        total += tile_hit(v->means2d[2 * (size_t)g], v->means2d[2 * (size_t)g + 1], v->conics[3 * (size_t)g],
                                          v->conics[3 * (size_t)g + 1], v->conics[3 * (size_t)g + 2], tau, j, i, v->W, v->H);
            }
            v->gaussian_ids[k++] = g;
        }
    v->n_vis = nvis; v->n_isects = total;
    v->isect_ids = (int64_t *)malloc(sizeof(int64_t) * (size_t)(total ? total : 1));
    v->flatten_ids = (int32_t *)malloc(sizeof(int32_t) * (size_t)(total ? total : 1));
    int64_t pos = 0;
    for (int64_t r = 0; r < nvis; ++r) {
        const int g = v->gaussian_ids[r];
        int x0, x1, y0, y1;
        tile_rect(v, g, &x0, &x1, &y0, &y1);
        int32_t dbits; memcpy(&dbits, &v->depths[g], 4);
        const float tau = orc_tau(v->opac[g]);
        for (int i = y0; i < y1; ++i)
            for (int j = x0; j < x1; ++j) {
                if (v->cull && !tile_hit(v->means2d[2 * (size_t)g], v->means2d[2 * (size_t)g + 1], v->conics[3 * (size_t)g],
                                         v->conics[3 * (size_t)g + 1], v->conics[3 * (size_t)g + 2], tau, j, i, v->W, v->H))
                    continue;
                v->isect_ids[pos] = ((int64_t)(i * v->tw + j) << 32) | (int64_t)(uint32_t)dbits;
                v->flatten_ids[pos] = (int32_t)r;
                ++pos;
            }
    }
    int tile_bits = 1; while ((1 << tile_bits) <= v->tw * v->th) ++tile_bits; /* floor(log2)+1 */
    radix_sort_pairs(v->isect_ids, v->flatten_ids, total, 32 + tile_bits);
    const int ntiles = v->tw * v->th;
    v->offsets = (int32_t *)malloc(sizeof(int32_t) * (size_t)ntiles);
    int64_t cur = 0;
    for (int t = 0; t < ntiles; ++t) {
        while (cur < total && (v->isect_ids[cur] >> 32) < t) ++cur;
        v->offsets[t] = (int32_t)cur;
    }
}

OrcView *orc_view_create(int n, const float *means, const float *quats, const float *scales, const float *opac,
                         const float *viewmat, const float *K, int W, int H, float near_plane, float far_plane,
                         float radius_clip, float eps2d, int cull) {
    OrcView *v = (OrcView *)calloc(1, sizeof(OrcView));
    v->cull = cull;
    v->n = n; v->W = W; v->H = H; v->tw = (W + TILE - 1) / TILE; v->th = (H + TILE - 1) / TILE;
    v->opac = opac;
    v->radii = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
    v->means2d = (float *)malloc(sizeof(float) * 2 * (size_t)(n ? n : 1));
    v->depths = (float *)malloc(sizeof(float) * (size_t)(n ? n : 1));
    v->conics = (float *)malloc(sizeof(float) * 3 * (size_t)(n ? n : 1));
    float *cov6 = (float *)malloc(sizeof(float) * 6 * (size_t)(n ? n : 1));
    orc_covar(n, quats, scales, cov6);
    project_all(v, means, cov6, viewmat, K, near_plane, far_plane, radius_clip, eps2d);
    free(cov6);
    bin_and_sort(v);
    return v;
}

void orc_view_destroy(OrcView *v) {
    if (!v) return;
    free(v->radii); free(v->means2d); free(v->depths); free(v->conics);
    free(v->gaussian_ids); free(v->isect_ids); free(v->flatten_ids); free(v->offsets);
    free(v);
}

void orc_view_counts(const OrcView *v, int64_t *n_vis, int64_t *n_isects, int32_t *tw, int32_t *th) {
    *n_vis = v->n_vis; *n_isects = v->n_isects; *tw = v->tw; *th = v->th;
}

void orc_view_export(const OrcView *v, int32_t *radii, float *means2d, float *depths, float *conics,
                     int32_t *gaussian_ids, int64_t *isect_ids, int32_t *flatten_ids, int32_t *offsets) {
    if (radii) memcpy(radii, v->radii, sizeof(int32_t) * (size_t)v->n);
    if (means2d) memcpy(means2d, v->means2d, sizeof(float) * 2 * (size_t)v->n);
    if (depths) memcpy(depths, v->depths, sizeof(float) * (size_t)v->n);
    if (conics) memcpy(conics, v->conics, sizeof(float) * 3 * (size_t)v->n);
    if (gaussian_ids) memcpy(gaussian_ids, v->gaussian_ids, sizeof(int32_t) * (size_t)v->n_vis);
    if (isect_ids) memcpy(isect_ids, v->isect_ids, sizeof(int64_t) * (size_t)v->n_isects);
    if (flatten_ids) memcpy(flatten_ids, v->flatten_ids, sizeof(int32_t) * (size_t)v->n_isects);
    if (offsets) memcpy(offsets, v->offsets, sizeof(int32_t) * (size_t)(v->tw * v->th));
}

/* one tile's sequential front-to-back walk.  For list entry k and pixel p:
 *   sigma<0 or alpha<1/255 -> skip;  T*(1-alpha)<=1e-4 -> pixel done, entry NOT composited.
 * `emit(k, w[256])` is called for every entry with at least one non-zero weight.  Returns the
 * number of entries walked before every pixel was done. */
typedef void (*emit_fn)(void *ctx, int g, const float *w);

static int walk_tile(const OrcView *v, int tile, float *T /*[256]*/, emit_fn emit, void *ctx) {
    const int ty = tile / v->tw, tx = tile % v->tw;
    const int64_t s = v->offsets[tile];
    const int64_t e = (tile + 1 < v->tw * v->th) ? v->offsets[tile + 1] : v->n_isects;
    unsigned char done[NPIX];
    float px[NPIX], py[NPIX], w[NPIX];
    int live = 0;
    for (int p = 0; p < NPIX; ++p) {
        const int yy = ty * TILE + p / TILE, xx = tx * TILE + p % TILE;
        px[p] = (float)xx + 0.5f; py[p] = (float)yy + 0.5f;
        done[p] = !(yy < v->H && xx < v->W);
        live += !done[p];
        T[p] = 1.0f;
    }
    int64_t k = s;
    for (; k < e && live > 0; ++k) {
        const int g = v->gaussian_ids[v->flatten_ids[k]];
        const float gx = v->means2d[2 * (size_t)g], gy = v->means2d[2 * (size_t)g + 1];
        const float cxx = v->conics[3 * (size_t)g], cxy = v->conics[3 * (size_t)g + 1], cyy = v->conics[3 * (size_t)g + 2];
        const float op = v->opac[g];
        int any = 0;
        for (int p = 0; p < NPIX; ++p) {
            w[p] = 0.0f;
            if (done[p]) continue;
            const float dx = gx - px[p], dy = gy - py[p];
            const float sigma = 0.5f * ((cxx * dx) * dx + (cyy * dy) * dy) + (cxy * dx) * dy;
            const float alpha = fminf(0.999f, op * expf(-sigma));
            if (sigma < 0.0f || alpha < 1.0f / 255.0f) continue;
            const float nT = T[p] * (1.0f - alpha);
            if (nT <= 1e-4f) { done[p] = 1; --live; continue; }
            w[p] = alpha * T[p];
            T[p] = nT;
            any = 1;
        }
        if (any && emit) emit(ctx, g, w);
    }
    return (int)(k - s);
}

typedef struct {
    const OrcView *v; const float *F; int64_t sH, sW, sD; int D, tile;
    double *num, *den; double *row; int64_t rows, pairs;
} BpCtx;

static void bp_emit(void *vctx, int g, const float *w) {
    BpCtx *c = (BpCtx *)vctx;
    const OrcView *v = c->v;
    const int ty = c->tile / v->tw, tx = c->tile % v->tw;
    double dsum = 0.0;
    memset(c->row, 0, sizeof(double) * (size_t)c->D);
    for (int p = 0; p < NPIX; ++p) {
        if (w[p] == 0.0f) continue;
        const int yy = ty * TILE + p / TILE, xx = tx * TILE + p % TILE;
        const float *f = c->F + yy * c->sH + xx * c->sW;
        const double wp = (double)w[p];
        for (int d = 0; d < c->D; ++d) c->row[d] += wp * (double)f[d * c->sD];
        dsum += wp;
        c->pairs++;
    }
    double *dst = c->num + (size_t)g * (size_t)c->D;
    for (int d = 0; d < c->D; ++d) {
#pragma omp atomic
        dst[d] += c->row[d];
    }
#pragma omp atomic
    c->den[g] += dsum;
    c->rows++;
}

/* backproject.py:127-151 for one view: num[g,:] += sum_p w F[p,:], den[g] += sum_p w  (fp64 accumulators).
 * stats[0..3] = rows with non-zero weight, contributing (pixel,Gaussian) pairs, entries walked, 0 */
void orc_view_backproject(const OrcView *v, const float *F, int64_t sH, int64_t sW, int64_t sD, int D,
                          double *num, double *den, int64_t *stats) {
    int64_t rows = 0, pairs = 0, walked = 0;
    const int ntiles = v->tw * v->th;
#pragma omp parallel reduction(+ : rows, pairs, walked)
    {
        double *row = (double *)malloc(sizeof(double) * (size_t)(D ? D : 1));
        float T[NPIX];
#pragma omp for schedule(dynamic, 1)
        for (int t = 0; t < ntiles; ++t) {
            BpCtx c = {v, F, sH, sW, sD, D, t, num, den, row, 0, 0};
            walked += walk_tile(v, t, T, bp_emit, &c);
            rows += c.rows; pairs += c.pairs;
        }
        free(row);
    }
    if (stats) { stats[0] = rows; stats[1] = pairs; stats[2] = walked; stats[3] = 0; }
}

typedef struct { const OrcView *v; const float *colors; int D, tile; double *acc; } RdCtx;

static void rd_emit(void *vctx, int g, const float *w) {
    RdCtx *c = (RdCtx *)vctx;
    const float *col = c->colors + (size_t)g * (size_t)c->D;
    for (int p = 0; p < NPIX; ++p) {
        if (w[p] == 0.0f) continue;
        double *a = c->acc + (size_t)p * (size_t)c->D;
        const double wp = (double)w[p];
        for (int d = 0; d < c->D; ++d) a[d] += wp * (double)col[d];
    }
}

/* segment.py:209-220: forward D-channel render.  out [H,W,D] fp64, alpha [H,W] fp64 */
void orc_view_render(const OrcView *v, const float *colors, int D, double *out, double *alpha) {
    const int ntiles = v->tw * v->th;
#pragma omp parallel
    {
        double *acc = (double *)malloc(sizeof(double) * NPIX * (size_t)(D ? D : 1));
        float T[NPIX];
#pragma omp for schedule(dynamic, 1)
        for (int t = 0; t < ntiles; ++t) {
            memset(acc, 0, sizeof(double) * NPIX * (size_t)D);
            RdCtx c = {v, colors, D, t, acc};
            walk_tile(v, t, T, rd_emit, &c);
            const int ty = t / v->tw, tx = t % v->tw;
            for (int p = 0; p < NPIX; ++p) {
                const int yy = ty * TILE + p / TILE, xx = tx * TILE + p % TILE;
                if (yy >= v->H || xx >= v->W) continue;
                memcpy(out + ((size_t)yy * v->W + xx) * (size_t)D, acc + (size_t)p * D, sizeof(double) * (size_t)D);
                alpha[(size_t)yy * v->W + xx] = 1.0 - (double)T[p];
            }
        }
        free(acc);
    }
}

/* ---- threshold-margin analysis (tests only) ---------------------------------------------------------
 * The compositing loop has two hard cut-offs: alpha < 1/255 -> skip, and T(1-alpha) <= 1e-4 -> stop.  A
 * GPU that evaluates exp() a few ulp differently can land a (pixel, Gaussian) pair on the other side of
 * one; parity tests must be able to tell such a flip from a real error.  For every Gaussian row this
 * walk records the SMALLEST relative distance to a threshold among the tests that can change the row:
 *   - the row's own tests, |alpha - 1/255| / (1/255) (when sigma >= 0) and |T' - 1e-4| / 1e-4
 *     (a flip changes w(g,p) itself);
 *   - every alpha test met EARLIER on the same pixel's walk (a flip there scales the transmittance of
 *     all later Gaussians by 1 - 1/255), counted only at pixels that carry at least `frac` of the
 *     row's total weight den_total[g], since a 0.4 % change of a smaller share cannot move the row.
 * A T-stop flip touches only the Gaussian it happens at: in either outcome the very next valid
 * Gaussian stops the pixel (T <~ 1e-4 already).
 * margin[] must be initialised by the caller (e.g. to +inf) and is min-updated, so views accumulate. */
void orc_view_margins(const OrcView *v, const double *den_total, double frac, float *margin) {
    const int ntiles = v->tw * v->th;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < ntiles; ++tile) {
        const int ty = tile / v->tw, tx = tile % v->tw;
        const int64_t s = v->offsets[tile];
        const int64_t e = (tile + 1 < ntiles) ? v->offsets[tile + 1] : v->n_isects;
        unsigned char done[NPIX];
        float px[NPIX], py[NPIX], T[NPIX], pm[NPIX];
        int live = 0;
        for (int p = 0; p < NPIX; ++p) {
            const int yy = ty * TILE + p / TILE, xx = tx * TILE + p % TILE;
            px[p] = (float)xx + 0.5f; py[p] = (float)yy + 0.5f;
            done[p] = !(yy < v->H && xx < v->W);
            live += !done[p];
            T[p] = 1.0f; pm[p] = INFINITY;
        }
        for (int64_t k = s; k < e && live > 0; ++k) {
            const int g = v->gaussian_ids[v->flatten_ids[k]];
            const float gx = v->means2d[2 * (size_t)g], gy = v->means2d[2 * (size_t)g + 1];
            const float cxx = v->conics[3 * (size_t)g], cxy = v->conics[3 * (size_t)g + 1], cyy = v->conics[3 * (size_t)g + 2];
            const float op = v->opac[g];
            float best = INFINITY;
            for (int p = 0; p < NPIX; ++p) {
                if (done[p]) continue;
                const float dx = gx - px[p], dy = gy - py[p];
                const float sigma = 0.5f * ((cxx * dx) * dx + (cyy * dy) * dy) + (cxy * dx) * dy;
                const float alpha = fminf(0.999f, op * expf(-sigma));
                float own = INFINITY;
                if (sigma >= 0.0f) own = fabsf(alpha - 1.0f / 255.0f) * 255.0f;
                const float upstream = pm[p];
                if (own < pm[p]) pm[p] = own; /* later Gaussians of this pixel see this alpha test */
                if (sigma < 0.0f || alpha < 1.0f / 255.0f) { if (own < best) best = own; continue; }
                const float nT = T[p] * (1.0f - alpha);
                const float mt = fabsf(nT - 1e-4f) * 1e4f;
                if (mt < own) own = mt;
                if (own < best) best = own;
                if (nT <= 1e-4f) { done[p] = 1; --live; continue; }
                const float w = alpha * T[p];
                T[p] = nT;
                if ((double)w >= frac * den_total[g] && upstream < best) best = upstream;
            }
            if (best < INFINITY) { /* atomic min: non-negative floats order like their bit patterns */
                uint32_t nb, *slot = (uint32_t *)(margin + g);
                memcpy(&nb, &best, 4);
                uint32_t cur = __atomic_load_n(slot, __ATOMIC_RELAXED);
                while (nb < cur && !__atomic_compare_exchange_n(slot, &cur, nb, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
            }
        }
    }
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU baseline sets its thread count explicitly */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
