// Dense [rows, D] passes of the path, each fused into ONE HBM-bound kernel:
//   finalize : f = num/den; f /= ||f||; NaN -> 0          (backproject.py:166-169: 3 passes + temporaries)
//   mask     : normalise(x) @ normalise(text).T, max(pos) > max(neg) [, score0 > thr]
//              (segment.py:52-58 for Gaussians, segment.py:221-224 for rendered pixels)
// One warp per row, float4 loads when D % 4 == 0.
#include "common.cuh"

namespace gwbp {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// rows of up to 32*kRegs floats are held in registers: ONE read and one write of HBM per element
constexpr int kRegs = 32;
// VEC (template): float4 registers per lane on the vectorised paths -- 4 covers D <= 512, 8 covers D <= 1024

template <int VEC>
__global__ void __launch_bounds__(256) finalize_kernel(const float *__restrict__ num, const float *__restrict__ den,
                                                       float *__restrict__ out, int64_t n, int d) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float dn = den[row];
    const float *src = num + row * d;
    float *dst = out + row * d;
    if constexpr (VEC > 0) {  // launcher guarantees d % 4 == 0 and d <= 128 * VEC
        // 16-byte loads, the row held in registers, and ONE division + one rsqrt-style reciprocal per row: the
        // element-wise IEEE divisions of the first version made this pass instruction-bound (1.1 TB/s).
        // x * (1/den) and q * (1/norm) differ from x/den, q/norm by <= 1 ulp each; NaN/inf propagate identically
        // (den = 0 -> inf/NaN -> row of NaN -> 0, backproject.py:169).
        const float rd = 1.0f / dn;
        const int n4 = d >> 2;
        float4 q[VEC];
        float ss = 0.0f;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const int k = lane + 32 * j;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < n4) {
                v = __ldg(reinterpret_cast<const float4 *>(src) + k);
                v.x *= rd; v.y *= rd; v.z *= rd; v.w *= rd;
            }
            q[j] = v;
            ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        ss = warp_sum(ss);
        const float rn = 1.0f / sqrtf(ss);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const int k = lane + 32 * j;
            if (k < n4) {
                float4 v = make_float4(q[j].x * rn, q[j].y * rn, q[j].z * rn, q[j].w * rn);
                v.x = isnan(v.x) ? 0.0f : v.x; v.y = isnan(v.y) ? 0.0f : v.y;
                v.z = isnan(v.z) ? 0.0f : v.z; v.w = isnan(v.w) ? 0.0f : v.w;
                reinterpret_cast<float4 *>(dst)[k] = v;
            }
        }
        return;
    } else if (d <= 32 * kRegs) {
        float q[kRegs];
        float ss = 0.0f;
#pragma unroll
        for (int j = 0; j < kRegs; ++j) {
            const int k = lane + 32 * j;
            q[j] = (k < d) ? src[k] / dn : 0.0f;
            ss += q[j] * q[j];
        }
        ss = warp_sum(ss);
        const float nrm = sqrtf(ss);
#pragma unroll
        for (int j = 0; j < kRegs; ++j) {
            const int k = lane + 32 * j;
            if (k < d) {
                float v = q[j] / nrm;
                if (isnan(v)) v = 0.0f;
                dst[k] = v;
            }
        }
        return;
    } else {
        float ss = 0.0f;
        for (int j = lane; j < d; j += 32) {
            const float q = src[j] / dn;
            ss += q * q;
        }
        ss = warp_sum(ss);
        const float nrm = sqrtf(ss);
        for (int j = lane; j < d; j += 32) {
            float v = (src[j] / dn) / nrm;
            if (isnan(v)) v = 0.0f;
            dst[j] = v;
        }
    }
}

int launch_finalize(const float *num, const float *den, float *out, int64_t n, int d, cudaStream_t st) {
    if (n == 0 || d == 0) return 0;
    const unsigned blocks = (unsigned)((n + 7) / 8);
    const bool vec = (d & 3) == 0 && (((uintptr_t)num | (uintptr_t)out) & 15) == 0;
    if (vec && d <= 512)
        finalize_kernel<4><<<blocks, 256, 0, st>>>(num, den, out, n, d);
    else if (vec && d <= 1024)
        finalize_kernel<8><<<blocks, 256, 0, st>>>(num, den, out, n, d);
    else
        finalize_kernel<0><<<blocks, 256, 0, st>>>(num, den, out, n, d);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Closing step of a view-sharded job, fused: sparse reduce-scatter over NVLink PEER MEMORY + finalise in ONE kernel.
//
// Every rank holds full (num[N,D], den[N]) accumulators but its views only touched a fraction of the rows (config G,
// 8 ranks x 20 views: 12 % per rank -- the rows a view reaches are the un-occluded surface, which is why the reference
// needs prune_by_gradients at all).  A dense reduce-scatter moves all 11.9 GB regardless.  Here rank r owns rows
// [lo, lo + rows); one warp per owned row reads the W den values (4 bytes per peer), and pulls a peer's 2 KB num row
// through its NVLink mapping ONLY if that peer touched the row (den_peer > eps); the partial rows are added in rank
// order (deterministic, unlike a ring), divided by the summed den, normalised (backproject.py:166-169) and written
// once.  Rows nobody touched cost W den reads and one zero row.  Pointers come from cudaIpcOpenMemHandle
// (gwbp_ipc_*), W <= 8 ranks of one node.
// ---------------------------------------------------------------------------------------------------------------
struct PeerPtrs {
    const float *num[kMaxPeers];
    const float *den[kMaxPeers];
};

__device__ __forceinline__ float4 ld_peer4(const float4 *p) {  // L2-only: peer rows are read once
    return __ldcg(p);
}

template <int VEC>
__global__ void __launch_bounds__(256) peer_reduce_finalize_kernel(PeerPtrs p, int world, int64_t lo, int64_t rows, int d,
                                                                   float eps, float *__restrict__ out_feat,
                                                                   float *__restrict__ out_num, float *__restrict__ out_den) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int64_t g = lo + row;
    const float dr = lane < world ? __ldcg(p.den[lane] + g) : 0.0f;
    const unsigned touched = __ballot_sync(0xffffffffu, lane < world && dr > eps);
    // the reference's 1e-12 initialiser is counted once (dist.py): den_0 + sum_{r>0} (den_r - eps), rank order
    float total = __shfl_sync(0xffffffffu, dr, 0);
    for (int r = 1; r < world; ++r) total += __shfl_sync(0xffffffffu, dr, r) - eps;
    if (lane == 0 && out_den) out_den[row] = total;
    const int n4 = d >> 2;
    float4 acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned m = touched;
    while (m) {  // two peers' rows in flight per step, added in rank order
        const int r0 = __ffs(m) - 1;
        m &= m - 1;
        const int r1 = m ? __ffs(m) - 1 : -1;
        if (r1 >= 0) m &= m - 1;
        const float4 *s0 = reinterpret_cast<const float4 *>(p.num[r0] + g * d);
        const float4 *s1 = r1 >= 0 ? reinterpret_cast<const float4 *>(p.num[r1] + g * d) : nullptr;
        float4 a[VEC], b[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const int k = lane + 32 * j;
            a[j] = k < n4 ? ld_peer4(s0 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const int k = lane + 32 * j;
            b[j] = (s1 && k < n4) ? ld_peer4(s1 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            acc[j].x += a[j].x; acc[j].y += a[j].y; acc[j].z += a[j].z; acc[j].w += a[j].w;
            acc[j].x += b[j].x; acc[j].y += b[j].y; acc[j].z += b[j].z; acc[j].w += b[j].w;
        }
    }
    if (out_num) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const int k = lane + 32 * j;
            if (k < n4) reinterpret_cast<float4 *>(out_num + row * d)[k] = acc[j];
        }
    }
    if (!out_feat) return;
    // finalise exactly as finalize_kernel<VEC> (one reciprocal per row; NaN -> 0)
    const float rd = 1.0f / total;
    float ss = 0.0f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        acc[j].x *= rd; acc[j].y *= rd; acc[j].z *= rd; acc[j].w *= rd;
        ss += acc[j].x * acc[j].x + acc[j].y * acc[j].y + acc[j].z * acc[j].z + acc[j].w * acc[j].w;
    }
    ss = warp_sum(ss);
    const float rn = 1.0f / sqrtf(ss);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int k = lane + 32 * j;
        if (k < n4) {
            float4 v = make_float4(acc[j].x * rn, acc[j].y * rn, acc[j].z * rn, acc[j].w * rn);
            v.x = isnan(v.x) ? 0.0f : v.x; v.y = isnan(v.y) ? 0.0f : v.y;
            v.z = isnan(v.z) ? 0.0f : v.z; v.w = isnan(v.w) ? 0.0f : v.w;
            reinterpret_cast<float4 *>(out_feat + row * d)[k] = v;
        }
    }
}

bool peer_reduce_supported(int world, int d) { return world >= 1 && world <= kMaxPeers && (d & 3) == 0 && d >= 4 && d <= 1024; }

int launch_peer_reduce_finalize(const float *const *num_ptrs, const float *const *den_ptrs, int world, int64_t lo,
                                int64_t rows, int d, float eps, float *out_feat, float *out_num, float *out_den,
                                cudaStream_t st) {
    if (rows == 0) return 0;
    PeerPtrs p = {};
    for (int r = 0; r < world; ++r) {
        p.num[r] = num_ptrs[r];
        p.den[r] = den_ptrs[r];
    }
    const unsigned blocks = (unsigned)((rows + 7) / 8);
    if (d <= 512)
        peer_reduce_finalize_kernel<4><<<blocks, 256, 0, st>>>(p, world, lo, rows, d, eps, out_feat, out_num, out_den);
    else
        peer_reduce_finalize_kernel<8><<<blocks, 256, 0, st>>>(p, world, lo, rows, d, eps, out_feat, out_num, out_den);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

// Per-view-ratio accumulation over the Gaussians one view saw (packed records, depth order): one warp per
// record.  Rows the view never touched hold zeros in (num_v, den_v) and contribute 0/(0+eps) = 0 in the
// reference's dense expression, so skipping them is exact.
__global__ void __launch_bounds__(256) ratio_accumulate_kernel(const float4 *__restrict__ grec, int64_t n_vis,
                                                               float *__restrict__ num_v, float *__restrict__ den_v,
                                                               float *__restrict__ acc, float *__restrict__ den_acc, int d,
                                                               float num_scale, float den_scale, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n_vis) return;
    const int64_t gid = __float_as_int(grec[2 * i].w);
    const float dv = den_v[gid];
    if (dv == 0.0f) return;  // warp-uniform
    const float denom = den_scale * dv + eps;
    float *src = num_v + gid * d, *dst = acc + gid * d;
    for (int c = lane; c < d; c += 32) {
        dst[c] += (num_scale * src[c]) / denom;
        src[c] = 0.0f;
    }
    __syncwarp();
    if (lane == 0) {
        den_v[gid] = 0.0f;
        if (den_acc) den_acc[gid] += dv;
    }
}

int launch_ratio_accumulate(const float4 *grec, int64_t n_vis, float *num_v, float *den_v, float *acc, float *den_acc,
                            int d, float num_scale, float den_scale, float eps, cudaStream_t st) {
    if (n_vis == 0) return 0;
    ratio_accumulate_kernel<<<(unsigned)((n_vis + 7) / 8), 256, 0, st>>>(grec, n_vis, num_v, den_v, acc, den_acc,
                                                                         d, num_scale, den_scale, eps);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

// text [p, d] is staged (normalised) in shared memory; p*d*4 bytes must fit (checked on the host)
template <int VEC>
__global__ void __launch_bounds__(256) mask_kernel(const float *__restrict__ x, int64_t rows, int d,
                                                   const float *__restrict__ text, int p, int npos, float thr,
                                                   int use_thr, uint8_t *__restrict__ mask,
                                                   float *__restrict__ score) {
    extern __shared__ __align__(16) float stext[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int j = warp; j < p; j += nwarps) {
        float ss = 0.0f;
        for (int k = lane; k < d; k += 32) { const float v = text[(int64_t)j * d + k]; ss += v * v; }
        ss = warp_sum(ss);
        const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        for (int k = lane; k < d; k += 32) stext[j * d + k] = text[(int64_t)j * d + k] * inv;
    }
    __syncthreads();
    for (int64_t row = (int64_t)blockIdx.x * nwarps + warp; row < rows; row += (int64_t)gridDim.x * nwarps) {
        const float *src = x + row * d;
        float best_pos = -INFINITY, best_neg = -INFINITY, s0 = 0.0f;
        if constexpr (VEC > 0) {  // launcher guarantees d % 4 == 0 and d <= 128 * VEC
            // 16-byte global and shared loads; the row norm is applied to the P dot products, not to the D elements
            // (the scalar version spent 85 % of its issue slots on per-element work: 2.9 TB/s)
            const int n4 = d >> 2;
            const float4 *t4 = reinterpret_cast<const float4 *>(stext);
            float4 q[VEC];
            float ss = 0.0f;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const int k = lane + 32 * j;
                q[j] = (k < n4) ? __ldg(reinterpret_cast<const float4 *>(src) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                ss += q[j].x * q[j].x + q[j].y * q[j].y + q[j].z * q[j].z + q[j].w * q[j].w;
            }
            ss = warp_sum(ss);
            const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
            for (int jp = 0; jp < p; ++jp) {
                float dot = 0.0f;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const int k = lane + 32 * j;
                    if (k < n4) {
                        const float4 tv = t4[jp * n4 + k];
                        dot += q[j].x * tv.x + q[j].y * tv.y + q[j].z * tv.z + q[j].w * tv.w;
                    }
                }
                dot = warp_sum(dot) * inv;
                if (jp == 0) s0 = dot;
                if (jp < npos) best_pos = fmaxf(best_pos, dot); else best_neg = fmaxf(best_neg, dot);
                if (score && lane == 0) score[row * p + jp] = dot;
            }
        } else if (d <= 32 * kRegs) {  // one HBM read of the row: keep it in registers
            float q[kRegs];
            float ss = 0.0f;
#pragma unroll
            for (int j = 0; j < kRegs; ++j) {
                const int k = lane + 32 * j;
                q[j] = (k < d) ? src[k] : 0.0f;
                ss += q[j] * q[j];
            }
            ss = warp_sum(ss);
            const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
            for (int jp = 0; jp < p; ++jp) {
                float dot = 0.0f;
#pragma unroll
                for (int j = 0; j < kRegs; ++j) {
                    const int k = lane + 32 * j;
                    if (k < d) dot += (q[j] * inv) * stext[jp * d + k];
                }
                dot = warp_sum(dot);
                if (jp == 0) s0 = dot;
                if (jp < npos) best_pos = fmaxf(best_pos, dot); else best_neg = fmaxf(best_neg, dot);
                if (score && lane == 0) score[row * p + jp] = dot;
            }
        } else {
            float ss = 0.0f;
            for (int k = lane; k < d; k += 32) { const float v = src[k]; ss += v * v; }
            ss = warp_sum(ss);
            const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
            for (int j = 0; j < p; ++j) {
                float dot = 0.0f;
                for (int k = lane; k < d; k += 32) dot += (src[k] * inv) * stext[j * d + k];
                dot = warp_sum(dot);
                if (j == 0) s0 = dot;
                if (j < npos) best_pos = fmaxf(best_pos, dot); else best_neg = fmaxf(best_neg, dot);
                if (score && lane == 0) score[row * p + j] = dot;
            }
        }
        if (lane == 0) {
            bool m = best_pos > best_neg;
            if (use_thr) m = m && (s0 > thr);
            mask[row] = m ? 1 : 0;
        }
    }
}

int launch_mask(const float *x, int64_t rows, int d, const float *text, int p, int npos, float thr, int use_thr,
                uint8_t *mask, float *score, cudaStream_t st) {
    if (rows == 0) return 0;
    const size_t smem = (size_t)p * d * sizeof(float);
    GWBP_REQUIRE(p >= 1 && npos >= 1 && npos <= p, "mask: need 1 <= npos <= p (npos=%d p=%d)", npos, p);
    GWBP_REQUIRE(smem <= 200 * 1024, "mask: %d prompts x %d dims do not fit in shared memory", p, d);
    if (smem > 48 * 1024) {
        GWBP_CUDA_OK(cudaFuncSetAttribute(mask_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GWBP_CUDA_OK(cudaFuncSetAttribute(mask_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GWBP_CUDA_OK(cudaFuncSetAttribute(mask_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int64_t blocks = (rows + 7) / 8;
    if (blocks > num_sms() * 16) blocks = num_sms() * 16;
    const bool vec = (d & 3) == 0 && ((uintptr_t)x & 15) == 0;
    if (vec && d <= 512)
        mask_kernel<4><<<(unsigned)blocks, 256, smem, st>>>(x, rows, d, text, p, npos, thr, use_thr, mask, score);
    else if (vec && d <= 1024)
        mask_kernel<8><<<(unsigned)blocks, 256, smem, st>>>(x, rows, d, text, p, npos, thr, use_thr, mask, score);
    else
        mask_kernel<0><<<(unsigned)blocks, 256, smem, st>>>(x, rows, d, text, p, npos, thr, use_thr, mask, score);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace gwbp
