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
// EWA projection: one thread per Gaussian
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) project_kernel(int64_t n, const float4 *__restrict__ geo0,
                                                      const float4 *__restrict__ geo1,
                                                      const float2 *__restrict__ geo2, CamDev cam,
                                                      unsigned long long *__restrict__ cnt,
                                                      float4 *__restrict__ rec) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { cnt[n] = 0ull; return; }  // terminator so the exclusive scan yields the totals
    const float4 a = geo0[i];
    const float4 b4 = geo1[i];
    const float2 c2 = geo2[i];
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
    bool ok = (z >= cam.near_plane) && (z <= cam.far_plane) && (det > 0.0f) && isfinite(rad);
    ok = ok && (rad > cam.radius_clip);
    ok = ok && (m2x + rad > 0.0f) && (m2x - rad < cam.Wf) && (m2y + rad > 0.0f) && (m2y - rad < cam.Hf);
    ok = ok && isfinite(m2x) && isfinite(m2y) && isfinite(con_x) && isfinite(con_y) && isfinite(con_z);
    unsigned long long c = 0ull;
    if (ok) {
        const int radius = (int)fminf(rad, 16777216.0f);
        int x0, x1, y0, y1;
        tile_rect(m2x, m2y, radius, cam.tw, cam.th, x0, x1, y0, y1);
        const unsigned tiles = (unsigned)((y1 - y0) * (x1 - x0));
        c = (1ull << 32) | (unsigned long long)tiles;
        rec[2 * i] = make_float4(m2x, m2y, a.w, z);
        rec[2 * i + 1] = make_float4(con_x, con_y, con_z, __int_as_float(radius));
    }
    cnt[i] = c;
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
constexpr int kCoopTiles = 32;  // Gaussians covering more tiles than this are emitted warp-wide

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
        tpg[pos] = (y1 - y0) * (x1 - x0);
        dbits = (long long)(unsigned)__float_as_int(r0.w);
    }
    const int bw = x1 - x0;
    const int ntiles = (y1 - y0) * bw;
    const bool big = vis && ntiles > kCoopTiles;
    if (vis && !big) {
        for (int k = 0; k < ntiles; ++k) {
            const long long o = base + k;
            if (o < cap) {
                const int ty = y0 + k / bw, tx = x0 + k % bw;
                keys[o] = ((long long)(ty * cam.tw + tx) << 32) | dbits;
                vals[o] = pos;
            }
        }
    }
    unsigned m = __ballot_sync(0xffffffffu, big);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int sx0 = __shfl_sync(0xffffffffu, x0, src), sy0 = __shfl_sync(0xffffffffu, y0, src);
        const int sbw = __shfl_sync(0xffffffffu, bw, src), snt = __shfl_sync(0xffffffffu, ntiles, src);
        const int spos = __shfl_sync(0xffffffffu, pos, src);
        const long long sbase = __shfl_sync(0xffffffffu, base, src), sd = __shfl_sync(0xffffffffu, dbits, src);
        for (int k = lane; k < snt; k += 32) {
            const long long o = sbase + k;
            if (o < cap) {
                const int ty = sy0 + k / sbw, tx = sx0 + k % sbw;
                keys[o] = ((long long)(ty * cam.tw + tx) << 32) | sd;
                vals[o] = spos;
            }
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
