// Stage 1: scene packing, EWA projection, intersection emission.
//
// THIS FILE IS COMPILED WITH -fmad=false.  Every arithmetic statement below is one IEEE-fp32
// rounding in the same order as oracle/gsplat_oracle.py::project / oracle/oracle.c::project_all,
// so radii, tile rectangles and depth bits -- hence isect_ids / flatten_ids / isect_offsets --
// are BIT-EXACT against the oracle (tests/test_gpu_parity.py).  Semantics: gsplat-1.4.0
// fully_fused_projection_packed_fwd + isect_tiles, i.e. what every
// rasterization() call of backproject.py:89,115,133 does first.
//
// Binning is a two-stage sort (binning.cu): the ~n_vis visible Gaussians are sorted by depth bits
// first, intersections are emitted in that order, and a stable 13-bit sort on the tile id finishes;
// the result is identical to gsplat's 45-bit (tile|depth) sort of all intersections at ~1/5 the traffic.
//
// B200 notes: the scene is repacked ONCE into SoA float4/float4/float2 (40 B/Gaussian, 16-byte coalesced loads);
// projection reads those 40 B, writes an 8-byte count for every Gaussian and a 32-byte record for visible ones only.
// pack_scene / compact are HBM-bound streaming kernels; project_kernel is NOT: with every a*b+c split into two
// roundings (bit-exactness against the FMA-free oracle) and the exact tile test on ~7 candidate tiles per visible
// Gaussian it is instruction-issue-bound (ncu: issue slots 84 % busy, DRAM 20 %), see DESIGN.md.
#include "common.cuh"
#include "chain.cuh"

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
    d.super = 0;
    d.nsx = (d.tw + kSuperW - 1) / kSuperW;
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
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// tile rectangle of a projected Gaussian (gsplat isect_tiles)
// ---------------------------------------------------------------------------------------------
constexpr int kMaskTiles = 64;  // rectangles up to this many tiles are handled by one thread (hit mask = 1 word)

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
struct CullGauss {  // per-Gaussian constants of the tile test
    float gx, gy, A, B, C, kx, ky, tau;
};
__device__ __forceinline__ CullGauss cull_setup(float gx, float gy, float A, float B, float C, float op) {
    CullGauss g;
    g.gx = gx; g.gy = gy; g.A = A; g.B = B; g.C = C;
    g.kx = -__fdiv_rn(B, C);  // argmin over dy of the quadratic at fixed dx is kx*dx
    g.ky = -__fdiv_rn(B, A);
    const float x = 255.0f * op;
    g.tau = (x > 1.0f) ? ln_approx(x) + 0.01f : -1.0f;
    return g;
}
__device__ __forceinline__ float quad(const CullGauss &g, float dx, float dy) {
    return 0.5f * ((g.A * dx) * dx + (g.C * dy) * dy) + (g.B * dx) * dy;
}
__device__ __forceinline__ float clampf(float t, float lo, float hi) { return fminf(fmaxf(t, lo), hi); }
__device__ __forceinline__ bool tile_hit(const CullGauss &g, int tx, int ty, int W, int H) {
    if (g.tau < 0.0f) return false;
    const int xe = min(tx * kTile + kTile - 1, W - 1), ye = min(ty * kTile + kTile - 1, H - 1);
    const float X0 = (float)(tx * kTile) + 0.5f, X1 = (float)xe + 0.5f;
    const float Y0 = (float)(ty * kTile) + 0.5f, Y1 = (float)ye + 0.5f;
    const float dx0 = g.gx - X1, dx1 = g.gx - X0, dy0 = g.gy - Y1, dy1 = g.gy - Y0;
    if (dx0 <= 0.0f && dx1 >= 0.0f && dy0 <= 0.0f && dy1 >= 0.0f) return true;
    float best = quad(g, dx0, clampf(g.kx * dx0, dy0, dy1));
    best = fminf(best, quad(g, dx1, clampf(g.kx * dx1, dy0, dy1)));
    best = fminf(best, quad(g, clampf(g.ky * dy0, dx0, dx1), dy0));
    best = fminf(best, quad(g, clampf(g.ky * dy1, dx0, dx1), dy1));
    return best <= g.tau;
}
__device__ __forceinline__ CullGauss shfl_cull(const CullGauss &g, int src) {
    CullGauss r;
    r.gx = __shfl_sync(0xffffffffu, g.gx, src); r.gy = __shfl_sync(0xffffffffu, g.gy, src);
    r.A = __shfl_sync(0xffffffffu, g.A, src); r.B = __shfl_sync(0xffffffffu, g.B, src);
    r.C = __shfl_sync(0xffffffffu, g.C, src); r.kx = __shfl_sync(0xffffffffu, g.kx, src);
    r.ky = __shfl_sync(0xffffffffu, g.ky, src); r.tau = __shfl_sync(0xffffffffu, g.tau, src);
    return r;
}

// ---------------------------------------------------------------------------------------------
// Supertile binning (GWBP_PREPARE_SUPERTILE): a supertile is kSuperW x kSuperH = 8 x 4 tiles.  Per Gaussian and
// supertile ONE list entry carries the 32-bit mask of the supertile's tiles the Gaussian hits (bit = ly * 8 + lx).
// super_walk_small() enumerates the entries of a rectangle of <= 64 tiles from its row-major hit mask; the counting
// (projection kernel) and the emission use the same enumeration, so they cannot disagree.
// ---------------------------------------------------------------------------------------------
template <typename F>
__device__ __forceinline__ void super_walk_small(unsigned long long mk, int x0, int y0, int bw, int nsx, F f) {
    if (!mk) return;
    const unsigned long long rowmask = bw < 64 ? ((1ull << bw) - 1ull) : ~0ull;
    const int nrows = (63 - __clzll((long long)mk)) / bw + 1;  // rows of the rectangle that hold a set bit
    const int sy1 = (y0 + nrows - 1) / kSuperH, sx1 = (x0 + bw - 1) / kSuperW;
    for (int sy = y0 / kSuperH; sy <= sy1; ++sy)
        for (int sx = x0 / kSuperW; sx <= sx1; ++sx) {
            const int off = x0 - sx * kSuperW;  // column x0 relative to the supertile's first column
            unsigned sub = 0u;
#pragma unroll
            for (int ly = 0; ly < kSuperH; ++ly) {
                const int rr = sy * kSuperH + ly - y0;
                if (rr >= 0 && rr < nrows) {
                    const unsigned long long rowbits = (mk >> (rr * bw)) & rowmask;
                    const unsigned long long b8 = off >= 0 ? (rowbits << off) : (rowbits >> (-off));
                    sub |= (unsigned)(b8 & 0xffull) << (kSuperW * ly);
                }
            }
            if (sub) f(sy * nsx + sx, sub);
        }
}

// Number of entries super_walk_small() produces, without building the masks: per supertile ROW the union of the
// rectangle rows that fall into it, then the 8-column groups (aligned to the supertile grid) that hold a set bit.
__device__ __forceinline__ unsigned super_count_small(unsigned long long mk, int x0, int y0, int bw) {
    if (!mk) return 0u;
    const unsigned long long rowmask = bw < 64 ? ((1ull << bw) - 1ull) : ~0ull;
    const int off = x0 & (kSuperW - 1);
    unsigned cnt = 0u;
    int ty = y0;
    unsigned long long rest = mk;
    while (rest) {
        unsigned long long cols = 0ull;
        const int sy = ty / kSuperH;
        do {
            cols |= rest & rowmask;
            rest = bw < 64 ? rest >> bw : 0ull;
            ++ty;
        } while (rest && ty / kSuperH == sy);
        const unsigned long long lo = cols << off;
        const unsigned hi = off ? (unsigned)(cols >> (64 - off)) : 0u;  // columns that spill into a ninth group
        cnt += (unsigned)__popc(__vcmpne4((unsigned)lo, 0u) & 0x01010101u) +
               (unsigned)__popc(__vcmpne4((unsigned)(lo >> 32), 0u) & 0x01010101u) + (hi != 0u ? 1u : 0u);
    }
    return cnt;
}

// One rectangle of more than 64 tiles, walked supertile by supertile by the whole warp (lane = tile of the supertile).
// f(supertile id, mask) is called converged for every supertile of the rectangle's range, mask may be 0.
template <typename F>
__device__ __forceinline__ void super_walk_big(const CullGauss &sg, int x0, int x1, int y0, int y1, const CamDev &cam, F f) {
    const int lane = threadIdx.x & 31;
    const int lx = lane % kSuperW, ly = lane / kSuperW;
    for (int sy = y0 / kSuperH; sy <= (y1 - 1) / kSuperH; ++sy)
        for (int sx = x0 / kSuperW; sx <= (x1 - 1) / kSuperW; ++sx) {
            const int tx = sx * kSuperW + lx, ty = sy * kSuperH + ly;
            const bool inside = tx >= x0 && tx < x1 && ty >= y0 && ty < y1;
            const bool hit = inside && (!cam.cull || tile_hit(sg, tx, ty, cam.W, cam.H));
            f(sy * cam.nsx + sx, __ballot_sync(0xffffffffu, hit));
        }
}

// ---------------------------------------------------------------------------------------------
// EWA projection + tile test + ORDERED COMPACTION in one pass: one thread per Gaussian.
//
// Visible Gaussians are written straight to their packed position (ascending index = gsplat's packed=True order):
// the position is the exclusive prefix of the visible flags, obtained inside this kernel by a chained scan with
// decoupled look-back over the CTAs (status word per CTA: 2 flag bits + running count).  CTAs take their chunk of the
// scene from an atomic ticket, so a CTA's predecessors have always started and the look-back cannot dead-lock.
// This replaces project -> 64-bit scan -> compact and their per-Gaussian intermediates (counts, prefix, unpacked records,
// hit masks: 56 B written and read back per Gaussian).
// ---------------------------------------------------------------------------------------------
constexpr unsigned kErecBig = 0x80000000u;
constexpr int kFrontHdr = 4;  // front[0] unused, [1] = intersections, [2] = visible Gaussians, [3] = supertile entries, then status words

constexpr int kProjPerCta = 256;                  // Gaussians per CTA: warps 0-7, one thread each
constexpr int kProjThreads = kProjPerCta + 32;    // + warp 8: ticket and chained scan only
template <int kLook>
__global__ void __launch_bounds__(kProjThreads) project_pack_kernel(int64_t n, int nblocks, const float4 *__restrict__ geo0,
                                                           const float4 *__restrict__ geo1,
                                                           const float2 *__restrict__ geo2, CamDev cam,
                                                           unsigned long long *__restrict__ front,
                                                           float4 *__restrict__ grec, int *__restrict__ radii,
                                                           int *__restrict__ tpg, int *__restrict__ spg,
                                                           uint4 *__restrict__ erec,
                                                           unsigned *__restrict__ dkeys, unsigned *__restrict__ dvals,
                                                           unsigned long long *__restrict__ scan_n) {
    // The look-back costs a few L2 round trips (microseconds) while a CTA lives ~8 us: done by the compute warps it
    // would stretch every CTA's lifetime (measured: 0.26 -> 0.49 ms).  So warp 8 does nothing else, and its look-back
    // runs while warps 0-7 are busy with the tile tests, which need the visible flags but not the prefix.  The compute
    // warps never wait for each other (no CTA-wide barrier): they hand their counts to the scanner through named
    // barrier 1 (arrive only) and each picks up its base through its own barrier 2 + warp (with the scanner alone).
    // CTAs are chained in blockIdx order, as in CUB's device scan (1-D grids are dispatched in order).
    __shared__ unsigned s_wcount[8], s_wbase[8], s_done;
    __shared__ unsigned long long s_tiles, s_ents;
    const int lane = threadIdx.x & 31, wip = threadIdx.x >> 5;
    const bool scanner = wip == 8;
    const unsigned vb = blockIdx.x;
    if (threadIdx.x == kProjPerCta) {
        s_done = 0u;
        s_tiles = 0ull;
        s_ents = 0ull;
    }
    if (scanner) {
        asm volatile("bar.sync 1, %0;" ::"r"(kProjThreads) : "memory");  // the eight counts (and s_done / s_tiles) are in place
        unsigned agg = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) agg += s_wcount[w];
        if (lane == 0) chained_publish(front + kFrontHdr, vb, (unsigned long long)agg);
        const unsigned long long excl = chained_lookback<kLook>(front + kFrontHdr, vb, (unsigned long long)agg);
        if (lane == 0) {
            unsigned run = (unsigned)excl;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                s_wbase[w] = run;
                run += s_wcount[w];
            }
            if ((int)vb == nblocks - 1) {
                front[2] = excl + agg;
                *scan_n = (excl + agg) << kVisShift;  // the sort-free binning kernels read the visible count here
            }
        }
        __syncwarp();
#pragma unroll
        for (int w = 0; w < 8; ++w) asm volatile("bar.arrive %0, 64;" ::"r"(2 + w) : "memory");
        return;
    }
    const int64_t i = (int64_t)vb * kProjPerCta + threadIdx.x;
    const bool in_range = i < n;
    const int64_t il = in_range ? i : 0;  // out-of-range lanes stay alive for the warp collectives below
    const float4 a = geo0[il];
    const float4 b4 = geo1[il];
    const float2 c2 = geo2[il];
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
    // the CTA's visible count is known here, long before the tile tests are done: publish it now so that no successor
    // ever waits for it
    const unsigned vmask = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_wcount[wip] = (unsigned)__popc(vmask);
    __syncwarp();
    asm volatile("bar.arrive 1, %0;" ::"r"(kProjThreads) : "memory");
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0, radius = 0;
    unsigned tiles = 0;
    if (ok) {
        radius = (int)fminf(rad, 16777216.0f);
        tile_rect(m2x, m2y, radius, cam.tw, cam.th, x0, x1, y0, y1);
        tiles = (unsigned)((y1 - y0) * (x1 - x0));
    }
    const int bw = x1 - x0;
    const bool big = ok && tiles > kMaskTiles;
    // tile-hit mask of a rectangle of <= 64 tiles: bit k <-> k-th tile, row-major (all set without culling)
    unsigned long long mk = (ok && !big) ? (tiles >= 64 ? ~0ull : ((1ull << tiles) - 1ull)) : 0ull;
    if (cam.cull) {
        // Test every tile of the bounding rectangle ONCE and keep the result as a bit mask for the emission
        // pass (rectangles of <= 64 tiles; larger ones are counted and later re-tested by the whole warp).
        // Rectangle sizes vary from 1 to 64 tiles inside a warp, so the (Gaussian, tile) pairs of the 32
        // lanes are pooled and dealt out evenly: lane L tests pairs L, L+32, ... of the warp's list.
        __shared__ float4 s_cg[8][32][2];
        __shared__ int4 s_rc[8][32];        // x0, y0, bw, inclusive pair count
        __shared__ unsigned s_mask[8][32][2];
        const CullGauss cg = cull_setup(m2x, m2y, con_x, con_y, con_z, a.w);
        const int mine = (ok && !big) ? (int)tiles : 0;
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        s_cg[wip][lane][0] = make_float4(cg.gx, cg.gy, cg.A, cg.B);
        s_cg[wip][lane][1] = make_float4(cg.C, cg.kx, cg.ky, cg.tau);
        s_rc[wip][lane] = make_int4(x0, y0, bw, incl);
        s_mask[wip][lane][0] = 0u;
        s_mask[wip][lane][1] = 0u;
        __syncwarp();
        for (int pp = lane; pp < total; pp += 32) {
            int lo = 0, hi = 31;  // owner = first lane whose inclusive count exceeds pp
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const int mid = (lo + hi) >> 1;
                if (s_rc[wip][mid].w > pp) hi = mid; else lo = mid + 1;
            }
            const int4 rc = s_rc[wip][lo];
            const int first = lo ? s_rc[wip][lo - 1].w : 0;
            const int k = pp - first;
            const int row = (int)(((float)k + 0.5f) / (float)rc.z);  // k / bw for small non-negative ints
            const int col = k - row * rc.z;
            const float4 c0 = s_cg[wip][lo][0], c1 = s_cg[wip][lo][1];
            CullGauss g;
            g.gx = c0.x; g.gy = c0.y; g.A = c0.z; g.B = c0.w; g.C = c1.x; g.kx = c1.y; g.ky = c1.z; g.tau = c1.w;
            if (tile_hit(g, rc.x + col, rc.y + row, cam.W, cam.H)) atomicOr(&s_mask[wip][lo][k >> 5], 1u << (k & 31));
        }
        __syncwarp();
        if (ok && !big) {
            mk = ((unsigned long long)s_mask[wip][lane][1] << 32) | s_mask[wip][lane][0];
            tiles = (unsigned)__popcll(mk);
        }
        unsigned m = cam.super ? 0u : __ballot_sync(0xffffffffu, big);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const CullGauss sg = shfl_cull(cg, src);
            const int sx0 = __shfl_sync(0xffffffffu, x0, src), sy0 = __shfl_sync(0xffffffffu, y0, src);
            const int sbw = __shfl_sync(0xffffffffu, bw, src);
            const unsigned snt = __shfl_sync(0xffffffffu, tiles, src);
            unsigned hits = 0;
            for (unsigned k = lane; k < snt; k += 32)
                hits += tile_hit(sg, sx0 + (int)(k % sbw), sy0 + (int)(k / sbw), cam.W, cam.H);
#pragma unroll
            for (int o = 16; o; o >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o);
            if (lane == src) tiles = hits;
        }
    }
    unsigned sent = 0;  // (Gaussian, supertile) entries of this Gaussian
    if (cam.super) {
        if (ok && !big) sent = super_count_small(mk, x0, y0, bw);
        unsigned m = __ballot_sync(0xffffffffu, big);
        if (m) {
            const CullGauss cgb = cull_setup(m2x, m2y, con_x, con_y, con_z, a.w);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const CullGauss sg = shfl_cull(cgb, src);
                const int sx0 = __shfl_sync(0xffffffffu, x0, src), sx1 = __shfl_sync(0xffffffffu, x1, src);
                const int sy0 = __shfl_sync(0xffffffffu, y0, src), sy1 = __shfl_sync(0xffffffffu, y1, src);
                unsigned hits = 0, ents = 0;
                super_walk_big(sg, sx0, sx1, sy0, sy1, cam, [&](int, unsigned hm) {
                    hits += (unsigned)__popc(hm);
                    ents += hm != 0u;
                });
                if (lane == src) {
                    tiles = hits;
                    sent = ents;
                }
            }
        }
    }
    // ---- ordered compaction: CTA-local ranks + chained scan over the CTAs ----
    unsigned long long wt = ok ? (unsigned long long)tiles : 0ull;
#pragma unroll
    for (int o = 16; o; o >>= 1) wt += __shfl_xor_sync(0xffffffffu, wt, o);
    asm volatile("bar.sync %0, 64;" ::"r"(2 + wip) : "memory");  // this warp's base is in s_wbase (s_done / s_tiles are set)
    unsigned long long we = ok ? (unsigned long long)sent : 0ull;
#pragma unroll
    for (int o = 16; o; o >>= 1) we += __shfl_xor_sync(0xffffffffu, we, o);
    if (lane == 0) {  // the CTA's last warp adds the CTA's intersections (and supertile entries) to the view's totals
        if (wt) atomicAdd(&s_tiles, wt);
        if (we) atomicAdd(&s_ents, we);
        __threadfence_block();
        if (atomicAdd(&s_done, 1u) == 7u) {
            const unsigned long long bt = atomicAdd(&s_tiles, 0ull), be = atomicAdd(&s_ents, 0ull);
            if (bt) atomicAdd(front + 1, bt);
            if (be) atomicAdd(front + 3, be);
        }
    }
    if (ok) {
        const int pos = (int)(s_wbase[wip] + (unsigned)__popc(vmask & ((1u << lane) - 1u)));
        grec[2 * (int64_t)pos] = make_float4(m2x, m2y, a.w, __int_as_float((int)i));
        grec[2 * (int64_t)pos + 1] = make_float4(con_x, con_y, con_z, z);
        radii[pos] = radius;
        tpg[pos] = (int)tiles;
        if (cam.super) spg[pos] = (int)sent;
        erec[pos] = big ? make_uint4(0u, 0u, kErecBig, 0u)
                        : make_uint4((unsigned)mk, (unsigned)(mk >> 32),
                                     (unsigned)x0 | ((unsigned)y0 << 12) | ((unsigned)(bw - 1) << 24), 0u);
        dkeys[pos] = __float_as_uint(z);  // depth > 0: the bit pattern orders like the value
        dvals[pos] = (unsigned)pos;
    }
}

int launch_project_pack(int64_t n, const void *geo, const CamDev &cam, WsDev ws, cudaStream_t st) {
    const int nblocks = (int)((n + kProjPerCta - 1) / kProjPerCta);
    GWBP_CUDA_OK(cudaMemsetAsync(ws.front, 0, sizeof(unsigned long long) * (size_t)(kFrontHdr + nblocks), st));
    if (n == 0) return 0;
    const float4 *g0 = (const float4 *)geo;
    const float4 *g1 = g0 + n;
    const float2 *g2 = (const float2 *)(g1 + n);
    // look-back window: 64 predecessors per step (measured at config G: 32 / 64 / 128 / 256 -> 0.380 / 0.374 / 0.379 / 0.401 ms)
    project_pack_kernel<2><<<(unsigned)nblocks, kProjThreads, 0, st>>>(n, nblocks, g0, g1, g2, cam, ws.front, ws.grec, ws.radii,
                                                              ws.tiles_per_gauss, ws.spg, ws.erec, ws.dkeys[0], ws.dvals[0],
                                                              ws.scan + n);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

// per-Gaussian hit counts in DEPTH order (input of the scan that positions / balances the binning) and, for the
// sort-free binning, the 16-byte emission records gathered into depth order as well (erec_sorted may be NULL): the
// binning warps then stream them with coalesced loads instead of chasing order[] -> erec[] one group at a time
__global__ void __launch_bounds__(256) gather_counts_kernel(int64_t n_vis, const unsigned *__restrict__ order,
                                                            const int *__restrict__ tpg, unsigned *__restrict__ out,
                                                            const uint4 *__restrict__ erec, uint4 *__restrict__ erec_sorted) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_vis) {
        const unsigned pos = order[i];
        out[i] = (unsigned)tpg[pos];
        if (erec_sorted) erec_sorted[i] = erec[pos];
    }
    if (i == n_vis) out[i] = 0u;
}

int launch_gather_counts(int64_t n_vis, const unsigned *order, WsDev ws, bool gather_erec, cudaStream_t st,
                         bool super_counts) {
    gather_counts_kernel<<<(unsigned)((n_vis + 1 + 255) / 256), 256, 0, st>>>(
        n_vis, order, super_counts ? ws.spg : ws.tiles_per_gauss, ws.cnt2, ws.erec, gather_erec ? (uint4 *)ws.rec : nullptr);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// emission in DEPTH order: thread i takes the i-th nearest visible Gaussian and writes one
// (tile id, packed index) pair per tile it reaches.  A stable sort on the tile id alone then yields
// exactly the order of gsplat's 64-bit (tile | depth) sort (ties: ascending packed index).
// ---------------------------------------------------------------------------------------------
template <typename KT>
__global__ void __launch_bounds__(256) emit_kernel(int64_t n_vis, CamDev cam, const unsigned *__restrict__ order,
                                                   const unsigned *__restrict__ base2,
                                                   const float4 *__restrict__ grec, const int *__restrict__ radii,
                                                   const uint4 *__restrict__ erec,
                                                   KT *__restrict__ tkeys, int *__restrict__ tvals, int64_t cap) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool vis = i < n_vis;
    int pos = 0;
    long long base = 0;
    uint4 er = make_uint4(0u, 0u, 0u, 0u);
    if (vis) {
        pos = (int)order[i];
        base = base2[i];
        er = erec[pos];  // the only gather for rectangles of <= 64 tiles
    }
    const bool big = vis && (er.z & kErecBig);
    if (vis && !big) {
        long long o = base;
        unsigned long long mk = ((unsigned long long)er.y << 32) | er.x;
        const int x0 = (int)(er.z & 0xfffu), bw = (int)((er.z >> 24) & 0x3fu) + 1;
        unsigned tbase = (unsigned)((int)((er.z >> 12) & 0xfffu) * cam.tw + x0);
        const unsigned long long rowmask = (bw < 64) ? ((1ull << bw) - 1ull) : ~0ull;
        while (mk) {
            unsigned long long row = mk & rowmask;
            mk = (bw < 64) ? (mk >> bw) : 0ull;
            while (row) {
                const int k = __ffsll((long long)row) - 1;
                row &= row - 1;
                if (o < cap) { tkeys[o] = (KT)(tbase + (unsigned)k); tvals[o] = pos; }
                ++o;
            }
            tbase += (unsigned)cam.tw;
        }
    }
    unsigned m = __ballot_sync(0xffffffffu, big);
    if (m) {  // rare: rectangles of more than 64 tiles are re-tested and emitted by the whole warp
        int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
        CullGauss cg = {};
        if (big) {
            const float4 r0 = grec[2 * (int64_t)pos], r1 = grec[2 * (int64_t)pos + 1];
            tile_rect(r0.x, r0.y, radii[pos], cam.tw, cam.th, x0, x1, y0, y1);
            if (cam.cull) cg = cull_setup(r0.x, r0.y, r1.x, r1.y, r1.z, r0.z);
        }
        const int bw = x1 - x0, ntiles = (y1 - y0) * bw;
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const CullGauss sg = shfl_cull(cg, src);
            const int sx0 = __shfl_sync(0xffffffffu, x0, src), sy0 = __shfl_sync(0xffffffffu, y0, src);
            const int sbw = __shfl_sync(0xffffffffu, bw, src), snt = __shfl_sync(0xffffffffu, ntiles, src);
            const int spos = __shfl_sync(0xffffffffu, pos, src);
            long long sbase = __shfl_sync(0xffffffffu, base, src);
            for (int k0 = 0; k0 < snt; k0 += 32) {  // ordered warp compaction
                const int k = k0 + lane;
                const int ty = sy0 + k / sbw, tx = sx0 + k % sbw;
                const bool hit = (k < snt) && (!cam.cull || tile_hit(sg, tx, ty, cam.W, cam.H));
                const unsigned hm = __ballot_sync(0xffffffffu, hit);
                const long long o = sbase + __popc(hm & ((1u << lane) - 1u));
                if (hit && o < cap) { tkeys[o] = (KT)(ty * cam.tw + tx); tvals[o] = spos; }
                sbase += __popc(hm);
            }
        }
    }
}

int launch_emit(int64_t n_vis, const CamDev &cam, const unsigned *order, WsDev ws, int64_t cap, bool key16,
                cudaStream_t st) {
    if (n_vis == 0) return 0;
    const unsigned blocks = (unsigned)((n_vis + 255) / 256);
    if (key16)
        emit_kernel<unsigned short><<<blocks, 256, 0, st>>>(n_vis, cam, order, ws.base2, ws.grec, ws.radii, ws.erec,
                                                            (unsigned short *)ws.tkeys[0], ws.tvals[0], cap);
    else
        emit_kernel<unsigned><<<blocks, 256, 0, st>>>(n_vis, cam, order, ws.base2, ws.grec, ws.radii, ws.erec,
                                                      ws.tkeys[0], ws.tvals[0], cap);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Supertile emission in DEPTH order: thread i takes the i-th nearest visible Gaussian and writes one
// (supertile id, packed index | tile mask << 32) entry per supertile it reaches.  A stable sort on the supertile id
// (<= 8 bits up to 1920 x 1088: one radix pass) leaves every supertile's entries in (depth, packed index) order; the
// back-projection kernels pick the entries whose mask has their tile's bit.
// ---------------------------------------------------------------------------------------------
constexpr int kEmitStage = 1536;  // entries a CTA stages in shared memory (256 Gaussians x ~1.7 entries on average)
template <typename KT>
__global__ void __launch_bounds__(256) emit_super_kernel(int64_t n_vis, CamDev cam, const unsigned *__restrict__ order,
                                                         const unsigned *__restrict__ base2,
                                                         const float4 *__restrict__ grec, const int *__restrict__ radii,
                                                         const uint4 *__restrict__ erec, KT *__restrict__ skeys,
                                                         unsigned long long *__restrict__ svals, int64_t cap) {
    // A CTA's entries are one contiguous run of the output (base2 is the prefix in depth order).  Threads write theirs
    // at per-thread offsets -- 1- or 2-byte keys and 8-byte values, i.e. partial sectors all over the run -- so the
    // run is assembled in shared memory and copied out with coalesced stores (0.13 -> ~0.06 ms at config G).  A CTA whose
    // run does not fit (huge rectangles) writes directly.
    __shared__ unsigned long long s_val[kEmitStage];
    __shared__ KT s_key[kEmitStage];
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x;
    const int64_t i = i0 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool vis = i < n_vis;
    const long long cta_lo = base2[i0], cta_hi = base2[min(i0 + (int64_t)blockDim.x, n_vis)];
    const bool staged = cta_hi - cta_lo <= kEmitStage && cta_hi <= cap;
    int pos = 0;
    long long base = 0;
    uint4 er = make_uint4(0u, 0u, 0u, 0u);
    if (vis) {
        pos = (int)order[i];
        base = base2[i];
        er = erec[pos];
    }
    auto put = [&](long long o, int st, unsigned long long v) {
        if (staged) {
            s_key[o - cta_lo] = (KT)st;
            s_val[o - cta_lo] = v;
        } else if (o < cap) {
            skeys[o] = (KT)st;
            svals[o] = v;
        }
    };
    const bool big = vis && (er.z & kErecBig);
    if (vis && !big) {
        long long o = base;
        super_walk_small(((unsigned long long)er.y << 32) | er.x, (int)(er.z & 0xfffu), (int)((er.z >> 12) & 0xfffu),
                         (int)((er.z >> 24) & 0x3fu) + 1, cam.nsx, [&](int st, unsigned sub) {
                             put(o, st, (unsigned long long)(unsigned)pos | ((unsigned long long)sub << 32));
                             ++o;
                         });
    }
    unsigned m = __ballot_sync(0xffffffffu, big);
    if (m) {  // rare: rectangles of more than 64 tiles are re-tested and emitted by the whole warp
        int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
        CullGauss cg = {};
        if (big) {
            const float4 r0 = grec[2 * (int64_t)pos], r1 = grec[2 * (int64_t)pos + 1];
            tile_rect(r0.x, r0.y, radii[pos], cam.tw, cam.th, x0, x1, y0, y1);
            if (cam.cull) cg = cull_setup(r0.x, r0.y, r1.x, r1.y, r1.z, r0.z);
        }
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const CullGauss sg = shfl_cull(cg, src);
            const int sx0 = __shfl_sync(0xffffffffu, x0, src), sx1 = __shfl_sync(0xffffffffu, x1, src);
            const int sy0 = __shfl_sync(0xffffffffu, y0, src), sy1 = __shfl_sync(0xffffffffu, y1, src);
            const int spos = __shfl_sync(0xffffffffu, pos, src);
            long long sbase = __shfl_sync(0xffffffffu, base, src);
            super_walk_big(sg, sx0, sx1, sy0, sy1, cam, [&](int st, unsigned hm) {
                if (hm) {
                    if (lane == 0) put(sbase, st, (unsigned long long)(unsigned)spos | ((unsigned long long)hm << 32));
                    ++sbase;
                }
            });
        }
    }
    if (staged) {
        __syncthreads();
        const int cnt = (int)(cta_hi - cta_lo);
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
            skeys[cta_lo + k] = s_key[k];
            svals[cta_lo + k] = s_val[k];
        }
    }
}

int launch_emit_super(int64_t n_vis, const CamDev &cam, const unsigned *order, WsDev ws, int64_t cap, int key_bytes,
                      cudaStream_t st) {
    if (n_vis == 0) return 0;
    const unsigned blocks = (unsigned)((n_vis + 255) / 256);
    if (key_bytes == 1)
        emit_super_kernel<unsigned char><<<blocks, 256, 0, st>>>(n_vis, cam, order, ws.base2, ws.grec, ws.radii, ws.erec,
                                                                 (unsigned char *)ws.tkeys[0], ws.svals[0], cap);
    else if (key_bytes == 2)
        emit_super_kernel<unsigned short><<<blocks, 256, 0, st>>>(n_vis, cam, order, ws.base2, ws.grec, ws.radii, ws.erec,
                                                                  (unsigned short *)ws.tkeys[0], ws.svals[0], cap);
    else
        emit_super_kernel<unsigned><<<blocks, 256, 0, st>>>(n_vis, cam, order, ws.base2, ws.grec, ws.radii, ws.erec,
                                                            ws.tkeys[0], ws.svals[0], cap);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Tile binning without a sort (the default path; tiles <= kBinMaxTiles).
//
// After the depth sort, "sort the intersections by tile id, stably" is a COUNTING sort whose input never has
// to exist: the depth-ordered Gaussians are cut into `chunks` contiguous pieces of (nearly) equal intersection
// count, ONE WARP PER CHUNK, and
//   pass A  bin_count_kernel   : per-chunk histogram of tile hits (private shared-memory table, one counter/tile)
//   scans   bin_scan*_kernel   : exclusive prefix over chunks for every tile + exclusive prefix over tiles
//                                 (= isect_offsets) -- three small kernels over the [chunks x tiles] table
//   pass B  bin_scatter_kernel : every warp re-walks its chunk IN DEPTH ORDER, 32 intersections per step, and
//                                 writes each packed index straight to its final slot of flatten_ids
// replace emit + 2-pass radix sort + offsets (0.40 ms of a 2.8 ms view at config G) and their (tile, index)
// intermediate (2 x 100 MB).  The result is bit-identical to the stable radix sort: within a tile, entries keep
// emission order = (depth, packed index) order.
//
// Pass A needs no order: lane = Gaussian, every lane walks the set bits of its own hit mask and bumps the warp's
// table with shared-memory atomics.  Pass B needs emission order: the lanes first expand their Gaussians' hits into a
// small per-warp staging list (lane g writes its tiles at its exclusive offset), which is then consumed 32 entries
// per step; `match.any` on the tile id finds the lanes of a step that hit the same tile, the rank among them is the
// lane order, and the private running table (absolute positions, u32 per tile) is bumped once per tile by the last of
// them.  No atomics in pass B, so the output is deterministic.
// ---------------------------------------------------------------------------------------------
struct BinArgs {
    CamDev cam;
    const unsigned *order;   // depth order -> packed index
    const unsigned *base2;   // [n_vis + 1] exclusive prefix of the per-Gaussian hit counts in depth order
    const uint4 *erec;       // emission records IN DEPTH ORDER (gather_counts_kernel)
    const float4 *grec;
    const int *radii;
    const unsigned long long *scan;  // scan[n] holds the totals: visible count in the high bits
    long long n;
    unsigned *counts;        // [chunks_pad][tiles_pad]
    unsigned *segsum;        // [nseg][tiles_pad]
    unsigned *totals;        // [tiles_pad]
    int *offsets;            // [tiles + 1]
    int *flatten;            // [cap]
    long long cap;
    int chunks, tiles, tiles_pad, nseg;
    int wpc;                 // warps (= chunks) per CTA: as many private tables as fit in shared memory, <= kBinWarps
};

constexpr int kBinWarps = 8;     // warps (= chunks) per CTA, fewer when the per-warp tile table is large
constexpr int kBinSeg = 32;      // chunks per scan segment
constexpr int kBinStage = 1024;  // staging entries per warp in pass B (a group of 32 small rectangles has <= 2048 hits)

// first depth-ordered Gaussian whose first intersection index is >= target (32-ary search by the whole warp)
__device__ __forceinline__ long long bin_lower_bound(const unsigned *base2, long long n_vis, unsigned long long target) {
    const int lane = threadIdx.x & 31;
    long long lo = 0, hi = n_vis;  // answer in [lo, hi]; base2[n_vis] = total
    while (hi - lo > 0) {
        const long long span = hi - lo;
        const long long step = (span + 31) / 32;
        const long long probe = lo + (long long)lane * step;  // ascending in the lane index
        const bool ge = probe >= hi || (unsigned long long)base2[probe] >= target;
        const unsigned m = __ballot_sync(0xffffffffu, ge);   // monotone: 0..0 1..1
        const int firstge = m ? (__ffs(m) - 1) : 32;
        const long long nhi = firstge < 32 ? min(hi, lo + (long long)firstge * step) : hi;
        const long long nlo = firstge > 0 ? lo + (long long)(firstge - 1) * step + 1 : lo;
        if (firstge == 0) return lo;
        lo = min(nlo, nhi);
        hi = nhi;
    }
    return lo;
}

// chunk c owns the Gaussians whose FIRST intersection index lies in [c*q, (c+1)*q), q = ceil(I / chunks)
__device__ __forceinline__ void bin_chunk_range(const BinArgs &a, int chunk, long long &lo, long long &hi) {
    const long long n_vis = (long long)(a.scan[a.n] >> kVisShift);
    const unsigned long long total = a.base2[n_vis];
    const unsigned long long q = (total + a.chunks - 1) / a.chunks;
    if (q == 0) {  // no intersections at all
        lo = hi = 0;
        return;
    }
    lo = chunk == 0 ? 0 : bin_lower_bound(a.base2, n_vis, (unsigned long long)chunk * q);
    hi = chunk == a.chunks - 1 ? n_vis : bin_lower_bound(a.base2, n_vis, (unsigned long long)(chunk + 1) * q);
}

// tile id of bit b of a small rectangle's hit mask; code = x0 | y0 << 12 | (bw - 1) << 24
__device__ __forceinline__ int bin_tile_of_bit(unsigned code, int b, int tw) {
    const int bw = (int)((code >> 24) & 0x3fu) + 1;
    const int row = (b * ((65536 + bw - 1) / bw)) >> 16;  // == b / bw for b < 64, bw <= 64
    return ((int)((code >> 12) & 0xfffu) + row) * tw + (int)(code & 0xfffu) + (b - row * bw);
}

// one rectangle of more than 64 tiles, re-tested and emitted by the whole warp, 32 tiles per step in row-major order
template <typename F>
__device__ __forceinline__ void bin_walk_big(const BinArgs &a, int src, int pos, F f) {
    const int lane = threadIdx.x & 31;
    int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
    CullGauss cg = {};
    if (lane == src) {
        const float4 r0 = a.grec[2 * (long long)pos], r1 = a.grec[2 * (long long)pos + 1];
        tile_rect(r0.x, r0.y, a.radii[pos], a.cam.tw, a.cam.th, x0, x1, y0, y1);
        if (a.cam.cull) cg = cull_setup(r0.x, r0.y, r1.x, r1.y, r1.z, r0.z);
    }
    const CullGauss sg = shfl_cull(cg, src);
    const int sx0 = __shfl_sync(0xffffffffu, x0, src), sy0 = __shfl_sync(0xffffffffu, y0, src);
    const int sbw = __shfl_sync(0xffffffffu, x1 - x0, src);
    const int snt = __shfl_sync(0xffffffffu, (y1 - y0) * (x1 - x0), src);
    const int spos = __shfl_sync(0xffffffffu, pos, src);
    for (int k0 = 0; k0 < snt; k0 += 32) {
        const int k = k0 + lane;
        const int ty = sy0 + k / sbw, tx = sx0 + k % sbw;
        const bool hit = (k < snt) && (!a.cam.cull || tile_hit(sg, tx, ty, a.cam.W, a.cam.H));
        f(ty * a.cam.tw + tx, spos, hit);
    }
}

// Walks one chunk and calls f(tile, packed_index, active) for every intersection.
//   kOrdered = false (pass A): any order, f is called divergently (lane = Gaussian);
//   kOrdered = true  (pass B): emission order, f is called with the whole warp converged, 32 consecutive
//                              intersections per call; `stage` = kBinStage words of shared memory of this warp.
// Shared by the counting and the scattering pass so that they cannot disagree about which hits exist.
template <bool kOrdered, typename F>
__device__ __forceinline__ void bin_walk_chunk(const BinArgs &a, long long lo, long long hi, unsigned *stage, F f) {
    const int lane = threadIdx.x & 31;
    // records and packed indices are streamed in depth order (coalesced), two groups ahead of their use, so the
    // warp's serial path never waits for DRAM
    int pos_n[2] = {0, 0};
    uint4 er_n[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
#pragma unroll
    for (int k = 0; k < 2; ++k)
        if (lo + 32 * k + lane < hi) {
            pos_n[k] = (int)a.order[lo + 32 * k + lane];
            er_n[k] = a.erec[lo + 32 * k + lane];
        }
    for (long long g0 = lo; g0 < hi; g0 += 32) {
        const bool vis = g0 + lane < hi;
        const int pos = pos_n[0];
        const uint4 er = er_n[0];
        pos_n[0] = pos_n[1];
        er_n[0] = er_n[1];
        pos_n[1] = 0;
        er_n[1] = make_uint4(0u, 0u, 0u, 0u);
        if (g0 + 64 + lane < hi) {
            pos_n[1] = (int)a.order[g0 + 64 + lane];
            er_n[1] = a.erec[g0 + 64 + lane];
        }
        const bool big = vis && (er.z & kErecBig);
        unsigned bigmask = __ballot_sync(0xffffffffu, big);
        if (!kOrdered) {
            if (vis && !big) {
                unsigned long long mk = ((unsigned long long)er.y << 32) | er.x;
                while (mk) {
                    const int b = __ffsll((long long)mk) - 1;
                    mk &= mk - 1;
                    f(bin_tile_of_bit(er.z, b, a.cam.tw), pos, true);
                }
            }
            __syncwarp();  // reconverge before the warp-cooperative part
            while (bigmask) {
                const int src = __ffs(bigmask) - 1;
                bigmask &= bigmask - 1;
                bin_walk_big(a, src, pos, f);
            }
            continue;
        }
        const int cnt = (vis && !big) ? __popc(er.x) + __popc(er.y) : 0;
        int first = 0;  // lanes [first, stop) form a run of small rectangles, lane `stop` (if < 32) is a big one
        while (first < 32) {
            const int stop = bigmask ? (__ffs(bigmask) - 1) : 32;
            // ---- run of small rectangles: inclusive prefix of their hit counts over the run's lanes
            const bool in_run = lane >= first && lane < stop;
            int incl = in_run ? cnt : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const int excl = incl - (in_run ? cnt : 0);
            for (int w0 = 0; w0 < total; w0 += kBinStage) {  // one window unless the group has > kBinStage hits
                if (in_run && cnt) {
                    unsigned long long mk = ((unsigned long long)er.y << 32) | er.x;
                    int e = excl - w0;
                    while (mk && e < kBinStage) {
                        const int b = __ffsll((long long)mk) - 1;
                        mk &= mk - 1;
                        if (e >= 0) stage[e] = (unsigned)bin_tile_of_bit(er.z, b, a.cam.tw) | ((unsigned)lane << 16);
                        ++e;
                    }
                }
                __syncwarp();
                const int lim = min(total - w0, kBinStage);
                for (int t0 = 0; t0 < lim; t0 += 32) {
                    const bool act = t0 + lane < lim;
                    const unsigned v = act ? stage[t0 + lane] : 0u;
                    const int opos = __shfl_sync(0xffffffffu, pos, (int)(v >> 16));
                    f((int)(v & 0xffffu), opos, act);
                }
                __syncwarp();
            }
            if (stop >= 32) break;
            bin_walk_big(a, stop, pos, f);
            bigmask &= bigmask - 1;
            first = stop + 1;
        }
    }
}

__global__ void __launch_bounds__(32 * kBinWarps) bin_count_kernel(const BinArgs a) {
    extern __shared__ unsigned bin_tab[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.x * a.wpc + warp;
    unsigned *tab = bin_tab + (size_t)warp * a.tiles_pad;
    for (int t = lane; t < a.tiles_pad; t += 32) tab[t] = 0u;
    __syncwarp();
    if (chunk < a.chunks) {
        long long lo, hi;
        bin_chunk_range(a, chunk, lo, hi);
        bin_walk_chunk<false>(a, lo, hi, nullptr, [&](int tile, int, bool act) {
            if (act) atomicAdd(&tab[tile], 1u);  // lanes may hit the same tile at the same time
        });
    }
    __syncwarp();
    unsigned *dst = a.counts + (size_t)chunk * a.tiles_pad;  // rows chunks..chunks_pad-1 are written as zeros
    for (int t = lane; t < a.tiles_pad; t += 32) dst[t] = tab[t];
}

// counts[c][t] -> exclusive prefix over the 32 chunks of its segment (in place); segsum[s][t] = segment total
__global__ void __launch_bounds__(256) bin_scan1_kernel(const BinArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
    if (t >= a.tiles_pad) return;
    unsigned *col = a.counts + (size_t)s * kBinSeg * a.tiles_pad + t;
    unsigned v[kBinSeg];
#pragma unroll
    for (int j = 0; j < kBinSeg; ++j) v[j] = col[(size_t)j * a.tiles_pad];
    unsigned acc = 0;
#pragma unroll
    for (int j = 0; j < kBinSeg; ++j) {
        col[(size_t)j * a.tiles_pad] = acc;
        acc += v[j];
    }
    a.segsum[(size_t)s * a.tiles_pad + t] = acc;
}

// segsum[s][t] -> exclusive prefix over segments (in place); totals[t] = intersections of tile t
__global__ void __launch_bounds__(256) bin_scan2_kernel(const BinArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.tiles_pad) return;
    unsigned acc = 0;
    for (int s = 0; s < a.nseg; ++s) {
        const unsigned v = a.segsum[(size_t)s * a.tiles_pad + t];
        a.segsum[(size_t)s * a.tiles_pad + t] = acc;
        acc += v;
    }
    a.totals[t] = acc;
}

// offsets[t] = exclusive prefix of totals over tiles (= gsplat isect_offsets), offsets[tiles] = n_isects.  One CTA.
__global__ void __launch_bounds__(1024) bin_scan3_kernel(const BinArgs a) {
    __shared__ unsigned warp_sums[32];
    __shared__ unsigned carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0u;
    __syncthreads();
    for (int t0 = 0; t0 < a.tiles; t0 += 1024) {
        const int t = t0 + tid;
        const unsigned v = t < a.tiles ? a.totals[t] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            warp_sums[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const unsigned base = carry + (warp ? warp_sums[warp - 1] : 0u);
        if (t < a.tiles) a.offsets[t] = (int)(base + incl - v);
        __syncthreads();
        if (tid == 1023) carry = base + incl;
        __syncthreads();
    }
    if (tid == 0) a.offsets[a.tiles] = (int)carry;
}

__global__ void __launch_bounds__(32 * kBinWarps) bin_scatter_kernel(const BinArgs a) {
    extern __shared__ unsigned bin_tab[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.x * a.wpc + warp;
    if (chunk >= a.chunks) return;
    long long lo, hi;
    bin_chunk_range(a, chunk, lo, hi);
    if (lo >= hi) return;
    unsigned *tab = bin_tab + (size_t)warp * a.tiles_pad;
    unsigned *stage = bin_tab + (size_t)a.wpc * a.tiles_pad + (size_t)warp * kBinStage;
    const unsigned *cbase = a.counts + (size_t)chunk * a.tiles_pad;
    const unsigned *sbase = a.segsum + (size_t)(chunk / kBinSeg) * a.tiles_pad;
    for (int t = lane; t < a.tiles; t += 32) tab[t] = (unsigned)a.offsets[t] + sbase[t] + cbase[t];  // absolute slots
    __syncwarp();
    bin_walk_chunk<true>(a, lo, hi, stage, [&](int tile, int pos, bool act) {
        const unsigned key = act ? (unsigned)tile : (0x80000000u | (unsigned)lane);  // inactive lanes match nobody
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int rank = __popc(peers & ((1u << lane) - 1u)), npeers = __popc(peers);
        unsigned slot = 0;
        if (act) slot = tab[tile];
        __syncwarp();
        if (act && rank == npeers - 1) tab[tile] = slot + (unsigned)npeers;
        __syncwarp();
        if (act && (long long)(slot + rank) < a.cap) a.flatten[slot + rank] = pos;
    });
}

// geometry of the binning tables for an image of n_tiles tiles (host only; also sizes the workspace)
size_t bin_table_bytes(int n_tiles, int *chunks_pad, int *tiles_pad, int *nseg, int *chunks, int *wpc) {
    const int tp = (n_tiles + 31) & ~31;
    const size_t tab = (size_t)tp * sizeof(unsigned);
    int w = (int)((size_t)(200 * 1024) / (tab + kBinStage * sizeof(unsigned)));
    if (w > kBinWarps) w = kBinWarps;
    if (w < 1) w = 1;
    const int ch = num_sms() * w;  // one CTA per SM
    const int ns = (ch + kBinSeg - 1) / kBinSeg;
    if (chunks) *chunks = ch;
    if (chunks_pad) *chunks_pad = (ns * kBinSeg + w - 1) / w * w;  // whole CTAs; rows beyond `chunks` hold zeros
    if (tiles_pad) *tiles_pad = tp;
    if (nseg) *nseg = ns;
    if (wpc) *wpc = w;
    return tab;
}

bool bin_fast_supported(int n_tiles) { return n_tiles >= 1 && n_tiles <= kBinMaxTiles; }

int launch_bin(int64_t n, const CamDev &cam, const unsigned *order, WsDev ws, int64_t cap, cudaStream_t st) {
    BinArgs a;
    a.cam = cam;
    a.order = order; a.base2 = ws.base2; a.erec = (const uint4 *)ws.rec; a.grec = ws.grec; a.radii = ws.radii; a.scan = ws.scan;
    a.n = n;
    a.counts = ws.bin_counts; a.segsum = ws.bin_seg; a.totals = ws.bin_tot;
    a.offsets = ws.offsets; a.flatten = ws.tvals[0]; a.cap = cap;
    a.tiles = cam.tw * cam.th;
    int chunks_pad = 0;
    const size_t tab = bin_table_bytes(a.tiles, &chunks_pad, &a.tiles_pad, &a.nseg, &a.chunks, &a.wpc);
    const size_t smem_a = tab * a.wpc, smem_b = (tab + kBinStage * sizeof(unsigned)) * a.wpc;
    GWBP_CUDA_OK(cudaFuncSetAttribute(bin_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    GWBP_CUDA_OK(cudaFuncSetAttribute(bin_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    const unsigned ctas = (unsigned)(chunks_pad / a.wpc);
    bin_count_kernel<<<ctas, 32 * a.wpc, smem_a, st>>>(a);
    count_launches(1);
    bin_scan1_kernel<<<dim3((unsigned)((a.tiles_pad + 255) / 256), (unsigned)a.nseg), 256, 0, st>>>(a);
    count_launches(1);
    bin_scan2_kernel<<<(unsigned)((a.tiles_pad + 255) / 256), 256, 0, st>>>(a);
    count_launches(1);
    bin_scan3_kernel<<<1, 1024, 0, st>>>(a);
    count_launches(1);
    bin_scatter_kernel<<<ctas, 32 * a.wpc, smem_b, st>>>(a);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace gwbp
