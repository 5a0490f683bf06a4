// Chained scan with decoupled look-back: status words and the warp-wide look-back shared by the projection kernel
// (ordered compaction), the exclusive scan of the per-Gaussian hit counts and -- per digit -- the radix sort passes.
#pragma once
#include <cuda_runtime.h>

namespace gwbp {

constexpr unsigned long long kDescAgg = 1ull << 62, kDescIncl = 2ull << 62, kDescVal = (1ull << 62) - 1ull;
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Chained scan with decoupled look-back over the CTAs (ticket order).  chained_publish() makes a CTA's aggregate
// visible as soon as it is known; chained_lookback(), called later by ONE WARP of the CTA, walks back over the
// predecessors' status words -- 256 per step (8 per lane, nearest first), because all CTAs of a wave publish at about the
// same time and the nearest word that already holds an inclusive prefix is typically a whole wave (~600 CTAs) away --
// publishes the CTA's inclusive prefix and returns the exclusive one (all lanes).
__device__ __forceinline__ void chained_publish(unsigned long long *desc, unsigned vb, unsigned long long agg) {
    st_relaxed(desc + vb, (vb == 0 ? kDescIncl : kDescAgg) | agg);
}
template <int kLookPerLane>
__device__ __forceinline__ unsigned long long chained_lookback(unsigned long long *desc, unsigned vb, unsigned long long agg) {
    const int lane = threadIdx.x & 31;
    unsigned long long excl = 0ull;
    if (vb > 0) {
        long long j = (long long)vb - 1;
        while (true) {
            // word u*32 + lane of the window = predecessor j - (u*32 + lane): every load instruction reads 256 contiguous
            // bytes (8 sectors).  [With 8 consecutive words per lane a window cost 256 sector requests and the ~600
            // resident look-back warps saturated the L2 request rate: ~4 us per round trip.]
            unsigned long long dsc[kLookPerLane];
#pragma unroll
            for (int u = 0; u < kLookPerLane; ++u) {  // all loads of a step are in flight together
                const long long idx = j - (long long)(u * 32 + lane);
                dsc[u] = idx >= 0 ? ld_relaxed(desc + idx) : kDescIncl;
            }
            while (true) {  // rare: a predecessor has started (dispatch order) but not published yet
                bool ready = true;
#pragma unroll
                for (int u = 0; u < kLookPerLane; ++u) ready = ready && (dsc[u] >> 62) != 0ull;
                if (__all_sync(0xffffffffu, ready)) break;
                __nanosleep(64);
#pragma unroll
                for (int u = 0; u < kLookPerLane; ++u)
                    if ((dsc[u] >> 62) == 0ull) dsc[u] = ld_relaxed(desc + (j - (long long)(u * 32 + lane)));
            }
            unsigned long long v = 0ull;
            bool found = false;
#pragma unroll
            for (int u = 0; u < kLookPerLane; ++u) {  // nearest group of 32 first
                if (!found) {
                    const unsigned imask = __ballot_sync(0xffffffffu, (dsc[u] >> 62) == 2ull);
                    const int first = __ffs(imask) - 1;  // nearest word of this group that holds an inclusive prefix
                    if (first < 0 || lane <= first) v += dsc[u] & kDescVal;
                    found = first >= 0;
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            excl += v;
            if (found) break;
            j -= 32 * kLookPerLane;
        }
        if (lane == 0) st_relaxed(desc + vb, kDescIncl | (excl + agg));
    }
    return excl;
}

// 32-bit status words (2 flag bits + 30-bit value) for the per-digit chains of the radix sort
constexpr unsigned kStAgg = 1u << 30, kStIncl = 2u << 30, kStVal = (1u << 30) - 1u;
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned *p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

}  // namespace gwbp
