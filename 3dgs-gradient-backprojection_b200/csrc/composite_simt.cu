// Stage 3 (CUDA-core "truth" kernels): per-tile front-to-back compositing fused with
//   * the back-projection  num[g,:] += sum_p w(g,p) F[p,:],  den[g] += sum_p w(g,p)
//     == gsplat rasterize_to_pixels_bwd v_colors for v_render = F, plus the reference's
//     accumulation (backproject.py:127-131,145-151), and
//   * the forward D-channel render (segment.py:209-220).
//
// These kernels work for any D and any feature-map strides, accumulate in fp32 and are the
// on-device reference the tcgen05 kernel (backproject_tc.cu) is validated against.  One CTA =
// one 16x16 tile = 256 threads; a thread is a pixel while weights are generated and a feature
// column while they are contracted.  Weights are generated ONCE per (tile, Gaussian, pixel)
// -- the reference regenerates them 35x per view (3 forward + 17 backward launches at D=512).
#include "common.cuh"

namespace gwbp {

template <int NC, int BG>
__global__ void __launch_bounds__(256) bp_simt_kernel(TileCtx t, const float *__restrict__ F, int64_t sH,
                                                      int64_t sW, int64_t sD, int d, float *__restrict__ num,
                                                      float *__restrict__ den, long long *__restrict__ stats) {
    __shared__ __align__(16) float Ws[BG][kTilePix];
    __shared__ float4 sg0[BG], sg1[BG];
    __shared__ float rowsum[BG];

    const int tile = blockIdx.x;
    const int ty = tile / t.tw, tx = tile % t.tw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = t.offsets[tile], e = t.offsets[tile + 1];
    const int yy = ty * kTile + tid / kTile, xx = tx * kTile + tid % kTile;
    const float px = (float)xx + 0.5f, py = (float)yy + 0.5f;
    bool done = !(yy < t.H && xx < t.W);
    float T = 1.0f;
    long long walked = 0, rows = 0;

    for (int b = s; b < e; b += BG) {
        if (__syncthreads_count(!done) == 0) break;
        const int nb = min(BG, e - b);
        if (tid < nb) {
            const int id = t.flatten[b + tid];
            sg0[tid] = t.grec[2 * (int64_t)id];
            sg1[tid] = t.grec[2 * (int64_t)id + 1];
        }
        __syncthreads();
        // thread = pixel: weights of this batch
#pragma unroll 4
        for (int k = 0; k < nb; ++k) {
            const float4 g0 = sg0[k], g1 = sg1[k];
            Ws[k][tid] = composite_step(g0.x, g0.y, g0.z, g1.x, g1.y, g1.z, px, py, T, done);
        }
        for (int k = nb; k < BG; ++k) Ws[k][tid] = 0.0f;
        __syncthreads();
        // den contribution of each row (also the "row is non-zero" flag: every weight is >= 0)
        for (int k = warp; k < BG; k += 8) {
            float v = 0.0f;
#pragma unroll
            for (int j = 0; j < kTilePix / 32; ++j) v += Ws[k][lane + 32 * j];
#pragma unroll
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) rowsum[k] = v;
        }
        __syncthreads();
        unsigned live = 0;
#pragma unroll
        for (int k = 0; k < BG; ++k) live |= (rowsum[k] > 0.0f ? 1u : 0u) << k;
        walked += nb;
        rows += __popc(live);
        if (live) {
            // thread = feature column(s): acc[k][c] = sum_p Ws[k][p] * F[p][col_c]
            float acc[BG][NC];
#pragma unroll
            for (int k = 0; k < BG; ++k)
#pragma unroll
                for (int c = 0; c < NC; ++c) acc[k][c] = 0.0f;
            for (int p4 = 0; p4 < kTilePix / 4; ++p4) {
                const int y4 = ty * kTile + p4 / 4, x4 = tx * kTile + (p4 % 4) * 4;
                float f[4][NC];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const int col = tid + 256 * c;
                        const bool ok = (y4 < t.H) && (x4 + i < t.W) && (col < d);
                        f[i][c] = ok ? __ldg(F + y4 * sH + (x4 + i) * sW + col * sD) : 0.0f;
                    }
#pragma unroll
                for (int k = 0; k < BG; ++k) {
                    if (live >> k & 1u) {
                        const float4 w4 = *reinterpret_cast<const float4 *>(&Ws[k][4 * p4]);
#pragma unroll
                        for (int c = 0; c < NC; ++c)
                            acc[k][c] += w4.x * f[0][c] + w4.y * f[1][c] + w4.z * f[2][c] + w4.w * f[3][c];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < BG; ++k) {
                if (live >> k & 1u) {
                    const int64_t gid = __float_as_int(sg0[k].w);
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const int col = tid + 256 * c;
                        if (col < d) atomicAdd(num + gid * d + col, acc[k][c]);
                    }
                    if (tid == 0) atomicAdd(den + gid, rowsum[k]);
                }
            }
        }
    }
    if (stats && tid == 0) {
        atomicAdd((unsigned long long *)&stats[0], (unsigned long long)rows);
        atomicAdd((unsigned long long *)&stats[1], (unsigned long long)walked);
    }
}

int launch_backproject_simt(const TileCtx &t, const float *F, int64_t sH, int64_t sW, int64_t sD, int d,
                            float *num, float *den, long long *stats, cudaStream_t st) {
    const int tiles = t.tw * t.th;
    if (tiles == 0 || d == 0) return 0;
    const int nc = (d + 255) / 256;
    switch (nc) {
        case 1: bp_simt_kernel<1, 32><<<tiles, 256, 0, st>>>(t, F, sH, sW, sD, d, num, den, stats); break;
        case 2: bp_simt_kernel<2, 32><<<tiles, 256, 0, st>>>(t, F, sH, sW, sD, d, num, den, stats); break;
        case 3: bp_simt_kernel<3, 16><<<tiles, 256, 0, st>>>(t, F, sH, sW, sD, d, num, den, stats); break;
        case 4: bp_simt_kernel<4, 16><<<tiles, 256, 0, st>>>(t, F, sH, sW, sD, d, num, den, stats); break;
        default: set_error("SIMT back-projection supports D <= 1024 (got %d)", d); return -1;
    }
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// forward render: thread = pixel, channels in chunks of CH (weights regenerated per chunk)
// ---------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256) render_simt_kernel(TileCtx t, const float *__restrict__ colors,
                                                          int64_t cstride, int d, const float *__restrict__ bg,
                                                          float *__restrict__ render, float *__restrict__ alpha) {
    constexpr int BG = 64;
    __shared__ float4 sg0[BG], sg1[BG];
    __shared__ __align__(16) float scol[BG][CH];
    const int tile = blockIdx.x;
    const int ty = tile / t.tw, tx = tile % t.tw;
    const int tid = threadIdx.x;
    const int s = t.offsets[tile], e = t.offsets[tile + 1];
    const int yy = ty * kTile + tid / kTile, xx = tx * kTile + tid % kTile;
    const bool inside = (yy < t.H && xx < t.W);
    const float px = (float)xx + 0.5f, py = (float)yy + 0.5f;

    for (int c0 = 0; c0 < d; c0 += CH) {
        bool done = !inside;
        float T = 1.0f;
        float acc[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = 0.0f;
        for (int b = s; b < e; b += BG) {
            if (__syncthreads_count(!done) == 0) break;
            const int nb = min(BG, e - b);
            if (tid < nb) {
                const int id = t.flatten[b + tid];
                sg0[tid] = t.grec[2 * (int64_t)id];
                sg1[tid] = t.grec[2 * (int64_t)id + 1];
            }
            __syncthreads();
            for (int idx = tid; idx < nb * CH; idx += 256) {
                const int k = idx / CH, c = idx % CH;
                const int64_t gid = __float_as_int(sg0[k].w);
                scol[k][c] = (c0 + c < d) ? __ldg(colors + gid * cstride + c0 + c) : 0.0f;
            }
            __syncthreads();
            for (int k = 0; k < nb; ++k) {
                const float4 g0 = sg0[k], g1 = sg1[k];
                const float w = composite_step(g0.x, g0.y, g0.z, g1.x, g1.y, g1.z, px, py, T, done);
                if (w > 0.0f) {
#pragma unroll
                    for (int c = 0; c < CH; c += 4) {
                        const float4 v = *reinterpret_cast<const float4 *>(&scol[k][c]);
                        acc[c] += w * v.x; acc[c + 1] += w * v.y; acc[c + 2] += w * v.z; acc[c + 3] += w * v.w;
                    }
                }
            }
        }
        __syncthreads();  // every warp is past the batch loop before smem is reused by the next chunk
        if (inside) {
            float *o = render + ((int64_t)yy * t.W + xx) * d + c0;
#pragma unroll
            for (int c = 0; c < CH; ++c)
                if (c0 + c < d) o[c] = acc[c] + (bg ? T * bg[c0 + c] : 0.0f);
            if (c0 == 0 && alpha) alpha[(int64_t)yy * t.W + xx] = 1.0f - T;
        }
    }
}

int launch_render_simt(const TileCtx &t, const float *colors, int64_t cstride, int d, const float *bg,
                       float *render, float *alpha, cudaStream_t st) {
    const int tiles = t.tw * t.th;
    if (tiles == 0 || d == 0) return 0;
    if (d <= 4)
        render_simt_kernel<4><<<tiles, 256, 0, st>>>(t, colors, cstride, d, bg, render, alpha);
    else
        render_simt_kernel<32><<<tiles, 256, 0, st>>>(t, colors, cstride, d, bg, render, alpha);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// probe render: the same composite at a handful of pixels (one CTA per probe pixel).
// alpha is evaluated for 128 Gaussians in parallel, the transmittance chain T is walked by one thread
// (it is a 128-step scalar recurrence), then every thread accumulates its channels over the batch.
// Same arithmetic and channel summation order as render_simt_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kProbeBatch = 128, kProbeRegs = 8;  // up to 256*8 = 2048 channels

__global__ void __launch_bounds__(256) render_pixels_kernel(TileCtx t, const float *__restrict__ colors,
                                                            int64_t cstride, int d, const float *__restrict__ extra,
                                                            const int *__restrict__ xy, float *__restrict__ out,
                                                            float *__restrict__ alpha_out) {
    __shared__ float s_a[kProbeBatch], s_w[kProbeBatch];
    __shared__ int s_gid[kProbeBatch];
    __shared__ float s_T, s_extra;
    __shared__ int s_done;
    const int tid = threadIdx.x, probe = blockIdx.x;
    const int xx = xy[2 * probe], yy = xy[2 * probe + 1];
    const int od = d + (extra ? 1 : 0);
    float acc[kProbeRegs];
#pragma unroll
    for (int j = 0; j < kProbeRegs; ++j) acc[j] = 0.0f;
    const bool inside = xx >= 0 && yy >= 0 && xx < t.W && yy < t.H;
    if (tid == 0) { s_T = 1.0f; s_extra = 0.0f; s_done = inside ? 0 : 1; }
    __syncthreads();
    if (inside) {
        const int tile = (yy / kTile) * t.tw + xx / kTile;
        const int s = t.offsets[tile], e = t.offsets[tile + 1];
        const float px = (float)xx + 0.5f, py = (float)yy + 0.5f;
        for (int b = s; b < e && !s_done; b += kProbeBatch) {
            const int nb = min(kProbeBatch, e - b);
            if (tid < nb) {
                const int id = t.flatten[b + tid];
                const float4 g0 = t.grec[2 * (int64_t)id], g1 = t.grec[2 * (int64_t)id + 1];
                const float dx = g0.x - px, dy = g0.y - py;
                const float sigma = pair_sigma(dx, dy, 0.5f * g1.x, g1.y, 0.5f * g1.z);
                const float a = pair_alpha(g0.z, sigma);
                s_a[tid] = (sigma >= 0.0f && a >= kAlphaMin) ? a : 0.0f;
                s_gid[tid] = __float_as_int(g0.w);
            }
            __syncthreads();
            if (tid == 0) {
                float T = s_T, ex = s_extra;
                bool done = false;
                for (int k = 0; k < nb; ++k) {
                    float w = 0.0f;
                    const float a = s_a[k];
                    if (!done && a > 0.0f) {
                        const float nT = __fmul_rn(T, __fsub_rn(1.0f, a));
                        if (nT <= kTMin) done = true;
                        else { w = __fmul_rn(a, T); T = nT; }
                    }
                    s_w[k] = w;
                    if (extra && w > 0.0f) ex += w * extra[s_gid[k]];
                }
                s_T = T; s_extra = ex; s_done = done ? 1 : 0;
            }
            __syncthreads();
            for (int k = 0; k < nb; ++k) {
                const float w = s_w[k];
                if (w > 0.0f) {
                    const float *row = colors + (int64_t)s_gid[k] * cstride;
#pragma unroll
                    for (int j = 0; j < kProbeRegs; ++j) {
                        const int c = tid + 256 * j;
                        if (c < d) acc[j] += w * __ldg(row + c);
                    }
                }
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int j = 0; j < kProbeRegs; ++j) {
        const int c = tid + 256 * j;
        if (c < d) out[(int64_t)probe * od + c] = acc[j];
    }
    if (tid == 0) {
        if (extra) out[(int64_t)probe * od + d] = s_extra;
        if (alpha_out) alpha_out[probe] = 1.0f - s_T;
    }
}

int launch_render_pixels(const TileCtx &t, const float *colors, int64_t cstride, int d, const float *extra,
                         const int *xy, int k, float *out, float *alpha, cudaStream_t st) {
    if (k == 0) return 0;
    GWBP_REQUIRE(d <= 256 * kProbeRegs, "render_pixels supports D <= %d (got %d)", 256 * kProbeRegs, d);
    render_pixels_kernel<<<k, 256, 0, st>>>(t, colors, cstride, d, extra, xy, out, alpha);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace gwbp
