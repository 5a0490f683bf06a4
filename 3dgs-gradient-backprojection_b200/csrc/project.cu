// Stage 1: scene packing, EWA projection, intersection emission.
//
// THIS FILE IS COMPILED WITH -fmad=false.  Every arithmetic statement below is one IEEE-fp32
// rounding in the same order as oracle/gsplat_oracle.py::project / oracle/oracle.c::project_all,
// so radii, tile rectangles and depth bits -- hence isect_ids / flatten_ids / isect_offsets --
// are BIT-EXACT against the oracle (tests/test_gpu_parity.py).  Semantics: gsplat-1.4.0
// fully_fused_projection_packed_fwd + isect_tiles, i.e. what every
// rasterization() call of backproject.py:89,115,133 does first.
//
// B200 notes: all three kernels are HBM-bound streaming kernels.  The scene is repacked ONCE
// into SoA float4/float4/float2 (40 B/Gaussian, 16-byte coalesced loads); projection reads those
// 40 B, writes an 8-byte count for every Gaussian and a 32-byte record for visible ones only.
#include "common.cuh"

namespace gwbp {

CamDev make_cam(const gwbp_camera &c) {
    CamDev d;
    for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 4; ++k) d.V[4 * r + k] = c.viewmat[4 * r + k];
    d.fx = c.K[0]; d.fy = c.K[4]; d.cx = c.K[2]; d.cy = c.K[5];
    d.W = c.width; d.H = c.height;
    d.tw = (c.width + kTile - 1) / kTile; d.th = (c.height + kTile - 1) / kTile;
    d.Wf = (float)c.width; d.Hf = (float)c.height;
    volatile float tanx = (0.5f * d.Wf) / d.fx, tany = (0.5f * d.Hf) / d.fy;
    volatile float t3x = 0.3f * tanx, t3y = 0.3f * tany;
    volatile float axp = (d.Wf - d.cx) / d.fx, axn = d.cx / d.fx;
    volatile float ayp = (d.Hf - d.cy) / d.fy, ayn = d.cy / d.fy;
    d.lim_xp = axp + t3x; d.lim_xn = axn + t3x;
    d.lim_yp = ayp + t3y; d.lim_yn = ayn + t3y;
    d.near_plane = c.near_plane; d.far_plane = c.far_plane;
    d.radius_clip = c.radius_clip; d.eps2d = c.eps2d;
    d.cull = 0;
    return d;
}

// ---------------------------------------------------------------------------------------------
// quat (wxyz, un-normalised) + scale -> world covariance; pack scene as SoA
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_scene_kernel(int64_t n, const float *__restrict__ means,
                                                         const float *__restrict__ quats,
                                                         const float *__restrict__ scales,
                                                         const float *__restrict__ opac, float4 *__restrict__ geo0,
                                                         float4 *__restrict__ geo1, float2 *__restrict__ geo2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 q = reinterpret_cast<const float4 *>(quats)[i];
    float w = q.x, x = q.y, y = q.z, z = q.w;
    const float n2 = ((w * w + x * x) + y * y) + z * z;
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(n2));
    w *= inv; x *= inv; y *= inv; z *= inv;
    const float x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z;
    const float wx = w * x, wy = w * y, wz = w * z;
    const float R[3][3] = {{1.0f - 2.0f * (y2 + z2), 2.0f * (xy - wz), 2.0f * (xz + wy)},
                           {2.0f * (xy + wz), 1.0f - 2.0f * (x2 + z2), 2.0f * (yz - wx)},
                           {2.0f * (xz - wy), 2.0f * (yz + wx), 1.0f - 2.0f * (x2 + y2)}};
    const float s[3] = {scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]};
    float M[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) M[r][c] = R[r][c] * s[c];
#define GWBP_DOT(a, b) ((M[a][0] * M[b][0] + M[a][1] * M[b][1]) + M[a][2] * M[b][2])
    geo0[i] = make_float4(means[3 * i], means[3 * i + 1], means[3 * i + 2], opac[i]);
    geo1[i] = make_float4(GWBP_DOT(0, 0), GWBP_DOT(0, 1), GWBP_DOT(0, 2), GWBP_DOT(1, 1));
    geo2[i] = make_float2(GWBP_DOT(1, 2), GWBP_DOT(2, 2));
#undef GWBP_DOT
}

int launch_pack_scene(int64_t n, const float *means, const float *quats, const float *scales,
                      const float *opac, void *geo, cudaStream_t st) {
    if (n == 0) return 0;
    float4 *g0 = (float4 *)geo;
    float4 *g1 = g0 + n;
    float2 *g2 = (float2 *)(g1 + n);
    pack_scene_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, means, quats, scales, opac, g0, g1, g2);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// tile rectangle of a projected Gaussian (gsplat isect_tiles)
// ---------------------------------------------------------------------------------------------
constexpr int kCoopTiles = 32;  // Gaussians covering more tiles than this are handled warp-wide

__device__ __forceinline__ int clamp_tile(float f, int hi) {
    if (!(f > 0.0f)) return 0;
    if (f >= (float)hi) return hi;
    return (int)f;
}

__device__ __forceinline__ void tile_rect(float m2x, float m2y, int radius, int tw, int th, int &x0, int &x1,
                                          int &y0, int &y1) {
    const float tr = (float)radius / (float)kTile;
    const float txc = m2x / (float)kTile, tyc = m2y / (float)kTile;
    x0 = clamp_tile(floorf(txc - tr), tw); x1 = clamp_tile(ceilf(txc + tr), tw);
    y0 = clamp_tile(floorf(tyc - tr), th); y1 = clamp_tile(ceilf(tyc + tr), th);
}

// ---------------------------------------------------------------------------------------------
// exact tile culling (GWBP_PREPARE_TILE_CULL; mirrors oracle.c tile_hit / gsplat_oracle.tile_hit)
// A (Gaussian, tile) pair is kept iff alpha = op*exp(-sigma) can reach 1/255 somewhere on the tile's
// pixel-centre box: min_box sigma <= ln(255 op) + 0.01.  Dropped pairs have zero weight on every
// pixel of the tile, so the accumulators are unchanged while the list to sort / walk shrinks ~40 %.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ln_approx(float x) {  // fixed-order fp32 series, x > 0
    const unsigned bits = __float_as_uint(x);
    const int e = (int)((bits >> 23) & 0xffu) - 127;
    const float m = __uint_as_float((bits & 0x7fffffu) | 0x3f800000u);
    const float s = __fdiv_rn(m - 1.0f, m + 1.0f);
    const float s2 = s * s;
    const float p = ((s2 * (1.0f / 7.0f) + 0.2f) * s2 + (1.0f / 3.0f)) * s2 + 1.0f;
    return (float)e * 0.69314718f + (2.0f * s) * p;
}
__device__ __forceinline__ float cull_tau(float op) {
    const float x = 255.0f * op;
    return (x > 1.0f) ? ln_approx(x) + 0.01f : -1.0f;
}
__device__ __forceinline__ float quad(float A, float B, float C, float dx, float dy) {
    return 0.5f * ((A * dx) * dx + (C * dy) * dy) + (B * dx) * dy;
}
__device__ __forceinline__ float clampf(float t, float lo, float hi) { return fminf(fmaxf(t, lo), hi); }
__device__ __forceinline__ bool tile_hit(float gx, float gy, float A, float B, float C, float tau, int tx, int ty,
                                         int W, int H) {
    if (tau < 0.0f) return false;
    const int xe = min(tx * kTile + kTile - 1, W - 1), ye = min(ty * kTile + kTile - 1, H - 1);
    const float X0 = (float)(tx * kTile) + 0.5f, X1 = (float)xe + 0.5f;
    const float Y0 = (float)(ty * kTile) + 0.5f, Y1 = (float)ye + 0.5f;
    const float dx0 = gx - X1, dx1 = gx - X0, dy0 = gy - Y1, dy1 = gy - Y0;
    if (dx0 <= 0.0f && dx1 >= 0.0f && dy0 <= 0.0f && dy1 >= 0.0f) return true;
    float best = quad(A, B, C, dx0, clampf(__fdiv_rn(-(B * dx0), C), dy0, dy1));
    best = fminf(best, quad(A, B, C, dx1, clampf(__fdiv_rn(-(B * dx1), C), dy0, dy1)));
    best = fminf(best, quad(A, B, C, clampf(__fdiv_rn(-(B * dy0), A), dx0, dx1), dy0));
    best = fminf(best, quad(A, B, C, clampf(__fdiv_rn(-(B * dy1), A), dx0, dx1), dy1));
    return best <= tau;
}

// ---------------------------------------------------------------------------------------------
// EWA projection: one thread per Gaussian
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) project_kernel(int64_t n, const float4 *__restrict__ geo0,
                                                      const float4 *__restrict__ geo1,
                                                      const float2 *__restrict__ geo2, CamDev cam,
                                                      unsigned long long *__restrict__ cnt,
                                                      float4 *__restrict__ rec) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == n) cnt[n] = 0ull;  // terminator so the exclusive scan yields the totals
    const bool in_range = i < n;
    const int64_t il = in_range ? i : 0;  // out-of-range lanes stay alive for the warp collectives below
    const float4 a = n ? geo0[il] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 b4 = n ? geo1[il] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float2 c2 = n ? geo2[il] : make_float2(0.f, 0.f);
    const float mx = a.x, my = a.y, mz = a.z;
    const float *V = cam.V;
    float p[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) p[r] = ((V[4 * r] * mx + V[4 * r + 1] * my) + V[4 * r + 2] * mz) + V[4 * r + 3];
    const float S[3][3] = {{b4.x, b4.y, b4.z}, {b4.y, b4.w, c2.x}, {b4.z, c2.x, c2.y}};
    float T[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) T[r][k] = (V[4 * r] * S[0][k] + V[4 * r + 1] * S[1][k]) + V[4 * r + 2] * S[2][k];
#define GWBP_CC(a_, b_) ((T[a_][0] * V[4 * b_] + T[a_][1] * V[4 * b_ + 1]) + T[a_][2] * V[4 * b_ + 2])
    const float C00 = GWBP_CC(0, 0), C01 = GWBP_CC(0, 1), C02 = GWBP_CC(0, 2);
    const float C11 = GWBP_CC(1, 1), C12 = GWBP_CC(1, 2), C22 = GWBP_CC(2, 2);
#undef GWBP_CC
    const float x = p[0], y = p[1], z = p[2];
    const float rz = __fdiv_rn(1.0f, z), rz2 = rz * rz;
    const float tx = z * fminf(cam.lim_xp, fmaxf(-cam.lim_xn, x * rz));
    const float ty = z * fminf(cam.lim_yp, fmaxf(-cam.lim_yn, y * rz));
    const float J00 = cam.fx * rz, J11 = cam.fy * rz;
    const float J02 = -((cam.fx * tx) * rz2), J12 = -((cam.fy * ty) * rz2);
    const float a0 = J00 * C00 + J02 * C02, a1 = J00 * C01 + J02 * C12, a2 = J00 * C02 + J02 * C22;
    const float b1 = J11 * C11 + J12 * C12, b2 = J11 * C12 + J12 * C22;
    const float s00 = a0 * J00 + a2 * J02, s01 = a1 * J11 + a2 * J12, s11 = b1 * J11 + b2 * J12;
    const float m2x = (cam.fx * x) * rz + cam.cx, m2y = (cam.fy * y) * rz + cam.cy;
    const float A = s00 + cam.eps2d, Cc = s11 + cam.eps2d;
    const float det = A * Cc - s01 * s01;
    const float inv_det = __fdiv_rn(1.0f, det);
    const float con_x = Cc * inv_det, con_y = -(s01 * inv_det), con_z = A * inv_det;
    const float bb = 0.5f * (A + Cc);
    const float v1 = bb + __fsqrt_rn(fmaxf(0.01f, bb * bb - det));
    float rad = ceilf(3.0f * __fsqrt_rn(v1));
    bool ok = in_range && (z >= cam.near_plane) && (z <= cam.far_plane) && (det > 0.0f) && isfinite(rad);
    ok = ok && (rad > cam.radius_clip);
    ok = ok && (m2x + rad > 0.0f) && (m2x - rad < cam.Wf) && (m2y + rad > 0.0f) && (m2y - rad < cam.Hf);
    ok = ok && isfinite(m2x) && isfinite(m2y) && isfinite(con_x) && isfinite(con_y) && isfinite(con_z);
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
    unsigned tiles = 0;
    if (ok) {
        const int radius = (int)fminf(rad, 16777216.0f);
        tile_rect(m2x, m2y, radius, cam.tw, cam.th, x0, x1, y0, y1);
        tiles = (unsigned)((y1 - y0) * (x1 - x0));
        rec[2 * i] = make_float4(m2x, m2y, a.w, z);
        rec[2 * i + 1] = make_float4(con_x, con_y, con_z, __int_as_float(radius));
    }
    if (cam.cull) {
        // count only the tiles the footprint can reach; big rectangles are counted by the whole warp
        const float tau = cull_tau(a.w);
        const int bw = x1 - x0;
        const bool big = ok && tiles > kCoopTiles;
        if (ok && !big) {
            unsigned hits = 0;
            for (unsigned k = 0; k < tiles; ++k)
                hits += tile_hit(m2x, m2y, con_x, con_y, con_z, tau, x0 + (int)(k % bw), y0 + (int)(k / bw), cam.W, cam.H);
            tiles = hits;
        }
        const int lane = threadIdx.x & 31;
        unsigned m = __ballot_sync(0xffffffffu, big);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float sx = __shfl_sync(0xffffffffu, m2x, src), sy = __shfl_sync(0xffffffffu, m2y, src);
            const float sA = __shfl_sync(0xffffffffu, con_x, src), sB = __shfl_sync(0xffffffffu, con_y, src);
            const float sC = __shfl_sync(0xffffffffu, con_z, src), st = __shfl_sync(0xffffffffu, tau, src);
            const int sx0 = __shfl_sync(0xffffffffu, x0, src), sy0 = __shfl_sync(0xffffffffu, y0, src);
            const int sbw = __shfl_sync(0xffffffffu, bw, src);
            const unsigned snt = __shfl_sync(0xffffffffu, tiles, src);
            unsigned hits = 0;
            for (unsigned k = lane; k < snt; k += 32)
                hits += tile_hit(sx, sy, sA, sB, sC, st, sx0 + (int)(k % sbw), sy0 + (int)(k / sbw), cam.W, cam.H);
#pragma unroll
            for (int o = 16; o; o >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o);
            if (lane == src) tiles = hits;
        }
    }
    if (in_range) cnt[i] = ok ? ((1ull << 32) | (unsigned long long)tiles) : 0ull;
}

int launch_project(int64_t n, const void *geo, const CamDev &cam, WsDev ws, cudaStream_t st) {
    const float4 *g0 = (const float4 *)geo;
    const float4 *g1 = g0 + n;
    const float2 *g2 = (const float2 *)(g1 + n);
    project_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(n, g0, g1, g2, cam, ws.cnt, ws.rec);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// emission: packed records + (tile|depth) keys in ascending-Gaussian, row-major-tile order
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) emit_kernel(int64_t n, CamDev cam, const unsigned long long *__restrict__ cnt,
                                                   const unsigned long long *__restrict__ scan,
                                                   const float4 *__restrict__ rec, float4 *__restrict__ grec,
                                                   int *__restrict__ radii, int *__restrict__ tpg,
                                                   long long *__restrict__ keys, int *__restrict__ vals,
                                                   int64_t cap) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool vis = false;
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0, pos = 0;
    long long base = 0, dbits = 0;
    float gx = 0.f, gy = 0.f, cA = 0.f, cB = 0.f, cC = 0.f, tau = -1.f;
    if (i < n && (cnt[i] >> 32)) {
        vis = true;
        const unsigned long long sc = scan[i];
        pos = (int)(sc >> 32);
        base = (long long)(sc & 0xffffffffull);
        const float4 r0 = rec[2 * i], r1 = rec[2 * i + 1];
        const int radius = __float_as_int(r1.w);
        tile_rect(r0.x, r0.y, radius, cam.tw, cam.th, x0, x1, y0, y1);
        grec[2 * (int64_t)pos] = make_float4(r0.x, r0.y, r0.z, __int_as_float((int)i));
        grec[2 * (int64_t)pos + 1] = make_float4(r1.x, r1.y, r1.z, r0.w);
        radii[pos] = radius;
        tpg[pos] = (int)(cnt[i] & 0xffffffffull);
        dbits = (long long)(unsigned)__float_as_int(r0.w);
        gx = r0.x; gy = r0.y; cA = r1.x; cB = r1.y; cC = r1.z;
        tau = cull_tau(r0.z);
    }
    const int bw = x1 - x0;
    const int ntiles = (y1 - y0) * bw;
    const bool big = vis && ntiles > kCoopTiles;
    if (vis && !big) {
        long long o = base;
        for (int k = 0; k < ntiles; ++k) {
            const int ty = y0 + k / bw, tx = x0 + k % bw;
            if (cam.cull && !tile_hit(gx, gy, cA, cB, cC, tau, tx, ty, cam.W, cam.H)) continue;
            if (o < cap) {
                keys[o] = ((long long)(ty * cam.tw + tx) << 32) | dbits;
                vals[o] = pos;
            }
            ++o;
        }
    }
    unsigned m = __ballot_sync(0xffffffffu, big);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int sx0 = __shfl_sync(0xffffffffu, x0, src), sy0 = __shfl_sync(0xffffffffu, y0, src);
        const int sbw = __shfl_sync(0xffffffffu, bw, src), snt = __shfl_sync(0xffffffffu, ntiles, src);
        const int spos = __shfl_sync(0xffffffffu, pos, src);
        long long sbase = __shfl_sync(0xffffffffu, base, src);
        const long long sd = __shfl_sync(0xffffffffu, dbits, src);
        const float sx = __shfl_sync(0xffffffffu, gx, src), sy = __shfl_sync(0xffffffffu, gy, src);
        const float sA = __shfl_sync(0xffffffffu, cA, src), sB = __shfl_sync(0xffffffffu, cB, src);
        const float sC = __shfl_sync(0xffffffffu, cC, src), st = __shfl_sync(0xffffffffu, tau, src);
        for (int k0 = 0; k0 < snt; k0 += 32) {  // ordered warp compaction keeps row-major tile order
            const int k = k0 + lane;
            const int ty = sy0 + k / sbw, tx = sx0 + k % sbw;
            const bool hit = (k < snt) && (!cam.cull || tile_hit(sx, sy, sA, sB, sC, st, tx, ty, cam.W, cam.H));
            const unsigned hm = __ballot_sync(0xffffffffu, hit);
            const long long o = sbase + __popc(hm & ((1u << lane) - 1u));
            if (hit && o < cap) {
                keys[o] = ((long long)(ty * cam.tw + tx) << 32) | sd;
                vals[o] = spos;
            }
            sbase += __popc(hm);
        }
    }
}

int launch_emit(int64_t n, const CamDev &cam, WsDev ws, int64_t cap, cudaStream_t st) {
    if (n == 0) return 0;
    emit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, cam, ws.cnt, ws.scan, ws.rec, ws.grec, ws.radii,
                                                           ws.tiles_per_gauss, ws.keys[0], ws.vals[0], cap);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace gwbp
