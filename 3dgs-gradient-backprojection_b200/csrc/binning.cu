// Stage 2: prefix scan, the two stable radix sorts and per-tile range finding.  Integer work, bit-exact against the
// oracle.  Semantics: gsplat-1.4.0 isect_tiles / radix sort / isect_offset_encode -- the sorted
// (tile, depth, packed index) order is the same, but instead of one 45-bit sort over all I intersections the n_vis
// visible Gaussians are depth-sorted first (4 passes x 8 B) and, after emission in depth order, a stable sort on the
// <= 13 tile bits (2 passes) or the <= 8 supertile bits (1 pass) finishes the job.
//
// Everything here is hand-written (no CUB):
//   radix_hist_kernel      digit histograms of ALL passes in one read of the keys (shared-memory bumps, one flush per CTA)
//   radix_scan_kernel      exclusive scan of each pass's 256 bins = where every digit's run starts
//   radix_onesweep_kernel  one stable LSD pass: a CTA ranks its tile of keys per digit (match.any inside a warp, running
//                          per-warp counters, prefix over the warps), learns how many keys with the same digit the
//                          EARLIER tiles hold from a chained scan with decoupled look-back (one status word per tile
//                          and digit, 16 predecessors fetched per round trip), reorders the tile in shared memory so
//                          that every digit's keys leave as one contiguous run, and scatters keys + values
//   scan_counts_kernel     exclusive scan of the per-Gaussian hit counts (chained scan, one pass over the data)
//   offsets_kernel         per-tile ranges of the sorted list
// At the sizes of a view (4 M depth keys, 7-17 M list entries) a pass moves ~64 MB (10 us of HBM time) but takes ~40 us:
// ncu (profiles/r02_radix_onesweep_summary.txt) shows the MIO pipe as the limiter -- match.any, shuffles and the
// shared-memory counter chains of the ranking (short-scoreboard stalls 12 of 26 stall cycles per issue) -- then the
// barrier in front of the scatter while 8 of the 16 warps look back.  Measured against the CUB onesweep it replaces
// (same box, config G): depth sort 0.192 -> 0.188 ms, supertile pass + its histogram 0.078 -> 0.073 ms.
#include <stdlib.h>

#include "common.cuh"
#include "chain.cuh"

namespace gwbp {

namespace {

constexpr int kRadixBits = 8, kRadixBins = 1 << kRadixBits;
constexpr int kSortThreads = 512, kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 8, kSortTile = kSortThreads * kSortItems;  // keys per CTA and pass
constexpr int kMaxPasses = 4;

template <typename K>
__device__ __forceinline__ unsigned digit_of(K key, int shift, unsigned mask) {
    return ((unsigned)key >> shift) & mask;
}

// ---- histograms of every pass: one read of the keys ----
template <typename K>
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const K *__restrict__ keys, int64_t n, int passes, int bits,
                                                                  unsigned *__restrict__ ghist) {
    __shared__ unsigned s_hist[kMaxPasses][kRadixBins];
    for (int i = threadIdx.x; i < kMaxPasses * kRadixBins; i += kSortThreads) (&s_hist[0][0])[i] = 0u;
    __syncthreads();
    for (int64_t base = (int64_t)blockIdx.x * kSortTile; base < n; base += (int64_t)gridDim.x * kSortTile) {
#pragma unroll
        for (int k = 0; k < kSortItems; ++k) {
            const int64_t i = base + k * kSortThreads + threadIdx.x;
            if (i < n) {
                const K key = keys[i];
                // plain shared-memory bumps: lanes that share a digit (the high bytes of depth keys) serialise inside the
                // atomic unit, which is cheaper than finding them with match.any first
                for (int p = 0; p < passes; ++p) {
                    const int nb = min(kRadixBits, bits - p * kRadixBits);
                    atomicAdd(&s_hist[p][digit_of(key, p * kRadixBits, (1u << nb) - 1u)], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadixBins; i += kSortThreads) {
        const unsigned v = (&s_hist[0][0])[i];
        if (v) atomicAdd(ghist + i, v);
    }
}

// ---- per pass: exclusive scan of the 256 bins ----
// offsets_out (single-pass sorts): the bin offsets ARE the per-list ranges of the sorted array -- offsets_out[d] for
// d < n_lists, offsets_out[n_lists] = n -- so no range-finding pass over the sorted keys is needed.
__global__ void __launch_bounds__(kRadixBins) radix_scan_kernel(const unsigned *__restrict__ ghist, unsigned *__restrict__ gofs,
                                                                int *__restrict__ offsets_out, int n_lists, unsigned n) {
    __shared__ unsigned s_w[kRadixBins / 32];
    const int d = threadIdx.x, lane = d & 31, w = d >> 5;
    const unsigned h = ghist[blockIdx.x * kRadixBins + d];
    unsigned incl = h;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    unsigned off = incl - h;
    for (int k = 0; k < w; ++k) off += s_w[k];
    gofs[blockIdx.x * kRadixBins + d] = off;
    if (offsets_out && blockIdx.x == 0) {
        if (d < n_lists) offsets_out[d] = (int)off;
        if (d == 0) offsets_out[n_lists] = (int)n;
    }
}

// ---- one stable LSD pass ----
// Tile order = memory order: warp w owns keys [w * 256, (w + 1) * 256) of the tile, item k of lane l is key k * 32 + l
// of that piece, so "earlier" means (earlier warp) or (earlier item) or (same item, lower lane).
template <typename K, typename V, int kSortLook>
__global__ void __launch_bounds__(kSortThreads, 3) radix_onesweep_kernel(const K *__restrict__ kin, K *__restrict__ kout,
                                                                      const V *__restrict__ vin, V *__restrict__ vout,
                                                                      int64_t n, int shift, int nbits,
                                                                      const unsigned *__restrict__ gofs,
                                                                      unsigned *__restrict__ status,
                                                                      const int *__restrict__ gather_src,
                                                                      unsigned *__restrict__ gather_dst) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    V *s_v = reinterpret_cast<V *>(s_dyn);
    K *s_k = reinterpret_cast<K *>(s_dyn + sizeof(V) * kSortTile);
    __shared__ unsigned s_whist[kSortWarps][kRadixBins];  // per-warp digit counts -> keys of earlier warps with the digit
    __shared__ unsigned s_tofs[kRadixBins];               // first slot of the digit inside the reordered tile
    __shared__ unsigned s_gbase[kRadixBins];              // global slot of the digit's first key of this tile, minus s_tofs
    __shared__ unsigned s_wsum[kRadixBins / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tile = blockIdx.x;
    const int64_t base = (int64_t)tile * kSortTile;
    const int nv = (int)min((int64_t)kSortTile, n - base);  // keys of this tile
    const unsigned mask = (1u << nbits) - 1u;
    for (int i = tid; i < kSortWarps * kRadixBins; i += kSortThreads) (&s_whist[0][0])[i] = 0u;
    K key[kSortItems];
    unsigned dig[kSortItems];
    const int first = warp * 32 * kSortItems + lane;
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const int p = first + k * 32;
        key[k] = p < nv ? kin[base + p] : (K)0;
        dig[k] = p < nv ? digit_of(key[k], shift, mask) : (unsigned)kRadixBins;  // padding matches only padding
    }
    __syncthreads();
    // rank inside the warp: keys of the same digit in earlier items (running counter) + lower lanes of this item.
    // The eight match.any are independent and issued back to back (their latency was 26 % of the kernel's stall samples
    // when each one sat in front of the counter update that needs it); only the counter updates form a chain.
    unsigned rank[kSortItems], peers[kSortItems];
    unsigned *wh = s_whist[warp];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) peers[k] = __match_any_sync(0xffffffffu, dig[k]);
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const unsigned d = dig[k];
        const int leader = __ffs(peers[k]) - 1;
        unsigned old = 0u;
        if (lane == leader && d < (unsigned)kRadixBins) {
            old = wh[d];
            wh[d] = old + (unsigned)__popc(peers[k]);
        }
        rank[k] = __shfl_sync(0xffffffffu, old, leader) + (unsigned)__popc(peers[k] & lt);
        __syncwarp();  // the next item's leader may be another lane: order its read after this write
    }
    __syncthreads();
    unsigned *mine = status + (size_t)tile * kRadixBins + tid;
    unsigned total = 0u;
    if (tid < kRadixBins) {
        // prefix over the warps for digit `tid`, then over the digits for the tile-local layout
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const unsigned t = s_whist[w][tid];
            s_whist[w][tid] = total;
            total += t;
        }
        st_relaxed_u32(mine, (tile == 0 ? kStIncl : kStAgg) | total);  // visible to the successors as early as possible
        unsigned incl = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_wsum[warp] = incl;
        s_tofs[tid] = incl - total;  // completed below with the earlier warps' sums
    }
    __syncthreads();
    if (tid < kRadixBins) {
        unsigned off = s_tofs[tid];
        for (int w = 0; w < warp; ++w) off += s_wsum[w];
        s_tofs[tid] = off;
    }
    __syncthreads();
    // reorder the tile in shared memory: every digit's keys become one run, in tile order
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const int p = first + k * 32;
        if (p < nv) {
            const unsigned slot = s_tofs[dig[k]] + wh[dig[k]] + rank[k];
            s_k[slot] = key[k];
            s_v[slot] = vin[base + p];
        }
    }
    if (tid < kRadixBins) {
        // keys with digit `tid` in all earlier tiles: walk back over the predecessors' status words, kSortLook per step,
        // up to the nearest one that already knows its inclusive prefix
        unsigned excl = 0u;
        if (tile > 0) {
            long long j = (long long)tile - 1;
            bool found = false;
            while (!found) {
                unsigned v[kSortLook];
#pragma unroll
                for (int k = 0; k < kSortLook; ++k)
                    v[k] = j - k >= 0 ? ld_relaxed_u32(status + (size_t)(j - k) * kRadixBins + tid) : kStIncl;
#pragma unroll
                for (int k = 0; k < kSortLook; ++k) {
                    if (!found) {
                        while ((v[k] >> 30) == 0u) {  // started (dispatch order) but not published yet
                            __nanosleep(32);
                            v[k] = ld_relaxed_u32(status + (size_t)(j - k) * kRadixBins + tid);
                        }
                        excl += v[k] & kStVal;
                        found = (v[k] >> 30) == 2u;
                    }
                }
                j -= kSortLook;
            }
            st_relaxed_u32(mine, kStIncl | (excl + total));
        }
        s_gbase[tid] = gofs[tid] + excl - s_tofs[tid];
    }
    __syncthreads();
    // scatter: consecutive threads hold consecutive slots of the reordered tile = consecutive global addresses per digit
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const int p = k * kSortThreads + tid;
        if (p < nv) {
            const K kk = s_k[p];
            const unsigned dst = s_gbase[digit_of(kk, shift, mask)] + (unsigned)p;
            const V vv = s_v[p];
            kout[dst] = kk;
            vout[dst] = vv;
            // last pass of the depth sort: the per-Gaussian entry counts follow their Gaussian into depth order (input
            // of the scan that positions the emission) instead of a separate gather kernel
            if (gather_src) gather_dst[dst] = (unsigned)gather_src[(size_t)vv];
        }
    }
    if (gather_src && tile == 0 && tid == 0) gather_dst[n] = 0u;  // terminator: the scan runs over n + 1 counters
}

struct SortTmp {
    unsigned *ghist, *gofs, *status;
    size_t status_words;  // per pass
};

size_t sort_tmp_need(int64_t n, int passes) {
    const size_t tiles = (size_t)((n + kSortTile - 1) / kSortTile);
    return sizeof(unsigned) * (2 * kMaxPasses * kRadixBins + (size_t)passes * tiles * kRadixBins);
}

template <typename K, typename V>
int sort_pairs(K *k0, K *k1, V *v0, V *v1, int64_t n, int bits, WsDev ws, int *sorted_buf, cudaStream_t st,
               const int *gather_src = nullptr, unsigned *gather_dst = nullptr, int *offsets_out = nullptr, int n_lists = 0) {
    *sorted_buf = 0;
    if (n == 0 || bits <= 0) {
        if (gather_src) GWBP_CUDA_OK(cudaMemsetAsync(gather_dst, 0, sizeof(unsigned), st));
        if (offsets_out) GWBP_CUDA_OK(cudaMemsetAsync(offsets_out, 0, sizeof(int) * (size_t)(n_lists + 1), st));
        return 0;
    }
    GWBP_REQUIRE(n < (1ll << 30), "radix sort: %lld keys exceed the 2^30 limit of the 30-bit chain counters", (long long)n);
    GWBP_REQUIRE(bits <= (int)(8 * sizeof(K)) && bits <= kMaxPasses * kRadixBits, "radix sort: bad key width %d", bits);
    const int passes = (bits + kRadixBits - 1) / kRadixBits;
    const size_t need = sort_tmp_need(n, passes);
    GWBP_REQUIRE(need <= ws.sort_tmp_bytes, "sort scratch too small: %zu > %zu", need, ws.sort_tmp_bytes);
    const unsigned tiles = (unsigned)((n + kSortTile - 1) / kSortTile);
    unsigned *ghist = (unsigned *)ws.sort_tmp, *gofs = ghist + kMaxPasses * kRadixBins, *status = gofs + kMaxPasses * kRadixBins;
    GWBP_CUDA_OK(cudaMemsetAsync(ws.sort_tmp, 0, need, st));  // histograms + every pass's status words
    const unsigned hist_blocks = (unsigned)min((long long)tiles, (long long)num_sms() * 4);
    radix_hist_kernel<K><<<hist_blocks, kSortThreads, 0, st>>>(k0, n, passes, bits, ghist);
    GWBP_REQUIRE(offsets_out == nullptr || (passes == 1 && n_lists <= kRadixBins), "list offsets need a single-pass sort");
    radix_scan_kernel<<<passes, kRadixBins, 0, st>>>(ghist, gofs, offsets_out, n_lists, (unsigned)n);
    constexpr int smem = (int)(sizeof(K) + sizeof(V)) * kSortTile;
    int look = 16;  // predecessors per look-back step
#ifdef GWBP_EXPERIMENTS
    if (const char *e = getenv("GWBP_SORT_LOOK")) look = atoi(e);
#endif
    K *kb[2] = {k0, k1};
    V *vb[2] = {v0, v1};
    auto run = [&](auto kern) -> int {
        static bool attr_set[64] = {false};
        int dev = 0;
        GWBP_CUDA_OK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            GWBP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
        for (int p = 0; p < passes; ++p) {
            const int nb = bits - p * kRadixBits < kRadixBits ? bits - p * kRadixBits : kRadixBits;
            const bool lastp = p == passes - 1;
            kern<<<tiles, kSortThreads, smem, st>>>(kb[p & 1], kb[(p + 1) & 1], vb[p & 1], vb[(p + 1) & 1], n, p * kRadixBits, nb,
                                                    gofs + p * kRadixBins, status + (size_t)p * tiles * kRadixBins,
                                                    lastp ? gather_src : nullptr, lastp ? gather_dst : nullptr);
        }
        return 0;
    };
    int rc;
    if (look == 8) rc = run(radix_onesweep_kernel<K, V, 8>);
    else if (look == 32) rc = run(radix_onesweep_kernel<K, V, 32>);
    else rc = run(radix_onesweep_kernel<K, V, 16>);
    if (rc) return rc;
    count_launches(2 + passes);
    GWBP_CUDA_OK(cudaGetLastError());
    *sorted_buf = passes & 1;
    return 0;
}

// ---- exclusive scan of u32 counters (chained scan, one pass): 256 threads x 8 consecutive items per CTA ----
constexpr int kScanItems = 8, kScanTile = 256 * kScanItems;
__global__ void __launch_bounds__(256) scan_counts_kernel(const unsigned *__restrict__ in, unsigned *__restrict__ out, int64_t n,
                                                          unsigned long long *__restrict__ desc) {
    __shared__ unsigned s_w[8];
    __shared__ unsigned long long s_excl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tile = blockIdx.x;
    const int64_t i0 = (int64_t)tile * kScanTile + tid * kScanItems;
    unsigned v[kScanItems];
    if (i0 + kScanItems <= n) {
        const uint4 a = *reinterpret_cast<const uint4 *>(in + i0), b = *reinterpret_cast<const uint4 *>(in + i0 + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) v[k] = i0 + k < n ? in[i0 + k] : 0u;
    }
    unsigned sum = 0u;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) sum += v[k];
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned agg = 0u;
#pragma unroll
        for (int w = 0; w < 8; ++w) agg += s_w[w];
        if (lane == 0) chained_publish(desc, tile, (unsigned long long)agg);
        const unsigned long long excl = chained_lookback<2>(desc, tile, (unsigned long long)agg);
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    unsigned run = (unsigned)s_excl + incl - sum;
    for (int w = 0; w < warp; ++w) run += s_w[w];
    unsigned o[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        o[k] = run;
        run += v[k];
    }
    if (i0 + kScanItems <= n) {
        *reinterpret_cast<uint4 *>(out + i0) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4 *>(out + i0 + 4) = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k)
            if (i0 + k < n) out[i0 + k] = o[k];
    }
}

}  // namespace

// Host-only (no CUDA call, so gwbp_workspace_layout works without a GPU): scratch of the sorts and the scan -- the
// histograms / bin offsets (8 KB), one 1 KB row of status words per 4096-key tile and pass, one 8-byte word per scan tile.
size_t binning_tmp_bytes(int64_t n, int64_t cap) {
    const size_t a = sort_tmp_need(n + 1, 4);                    // depth sort: 32-bit keys
    const size_t b = sort_tmp_need(cap, 4);                      // tile / supertile sort (<= 32 key bits)
    const size_t c = sizeof(unsigned long long) * (size_t)((n + 1 + kScanTile) / kScanTile + 1);
    size_t m = a > b ? a : b;
    m = m > c ? m : c;
    return (m + 4096 + 255) & ~(size_t)255;
}

// stage 1: visible Gaussians by depth bits (stable: ties keep ascending packed index)
// counts_src (optional): per-Gaussian entry counts in PACKED order; the last pass also writes them in depth order to
// ws.cnt2 (+ terminator), which replaces gather_counts_kernel
int launch_depth_sort(int64_t n_vis, WsDev ws, int *sorted_buf, cudaStream_t st, const int *counts_src) {
    return sort_pairs(ws.dkeys[0], ws.dkeys[1], ws.dvals[0], ws.dvals[1], n_vis, 32, ws, sorted_buf, st, counts_src, ws.cnt2);
}

// exclusive offsets of the per-Gaussian tile counts in depth order (n_vis + 1 entries)
int launch_scan_counts(int64_t n_vis, WsDev ws, cudaStream_t st) {
    const int64_t n = n_vis + 1;
    const unsigned tiles = (unsigned)((n + kScanTile - 1) / kScanTile);
    GWBP_REQUIRE(sizeof(unsigned long long) * (size_t)tiles <= ws.sort_tmp_bytes, "scan scratch too small");
    GWBP_CUDA_OK(cudaMemsetAsync(ws.sort_tmp, 0, sizeof(unsigned long long) * (size_t)tiles, st));
    scan_counts_kernel<<<tiles, 256, 0, st>>>(ws.cnt2, ws.base2, n, (unsigned long long *)ws.sort_tmp);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

// stage 2: intersections (already in depth order) by tile id only -- stable, floor(log2(tiles))+1 bits
int launch_tile_sort(int64_t n_isects, int tile_bits, WsDev ws, bool key16, int *sorted_buf, cudaStream_t st) {
    if (key16)
        return sort_pairs((unsigned short *)ws.tkeys[0], (unsigned short *)ws.tkeys[1], ws.tvals[0], ws.tvals[1], n_isects,
                          tile_bits, ws, sorted_buf, st);
    return sort_pairs(ws.tkeys[0], ws.tkeys[1], ws.tvals[0], ws.tvals[1], n_isects, tile_bits, ws, sorted_buf, st);
}

// supertile lists: stable sort of the depth-ordered (supertile id, packed index | tile mask << 32) entries by supertile id
// n_lists > 0 (single-pass sorts only): ws.offsets is filled from the pass's bin offsets, no range-finding kernel needed
int launch_super_sort(int64_t n_entries, int bits, WsDev ws, int key_bytes, int *sorted_buf, cudaStream_t st, int n_lists) {
    if (key_bytes == 1)
        return sort_pairs((unsigned char *)ws.tkeys[0], (unsigned char *)ws.tkeys[1], ws.svals[0], ws.svals[1], n_entries,
                          bits, ws, sorted_buf, st, nullptr, nullptr, n_lists > 0 ? ws.offsets : nullptr, n_lists);
    if (key_bytes == 2)
        return sort_pairs((unsigned short *)ws.tkeys[0], (unsigned short *)ws.tkeys[1], ws.svals[0], ws.svals[1], n_entries,
                          bits, ws, sorted_buf, st);
    return sort_pairs(ws.tkeys[0], ws.tkeys[1], ws.svals[0], ws.svals[1], n_entries, bits, ws, sorted_buf, st);
}

// offsets[t] = first sorted position whose tile id >= t; offsets[n_tiles] = n_isects.  Each thread scans kPer
// consecutive keys (the list is sorted, so tile boundaries are where neighbours differ).
template <typename KT>
__global__ void __launch_bounds__(256) offsets_kernel(int64_t n_isects, int n_tiles, const KT *__restrict__ keys,
                                                      int *__restrict__ offsets) {
    constexpr int kPer = 16 / sizeof(KT) * 2;  // two 16-byte loads per thread
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kPer;
    if (n_isects == 0) {
        for (int64_t t = i0; t < i0 + kPer; ++t)
            if (t <= n_tiles) offsets[t] = 0;
        return;
    }
    if (i0 >= n_isects) return;
    KT k[kPer];
    if (i0 + kPer <= n_isects) {
        const uint4 *src = reinterpret_cast<const uint4 *>(keys + i0);  // 16-byte aligned: i0 % kPer == 0
        *reinterpret_cast<uint4 *>(&k[0]) = __ldg(src);
        *reinterpret_cast<uint4 *>(&k[kPer / 2]) = __ldg(src + 1);
    } else {
#pragma unroll
        for (int j = 0; j < kPer; ++j) k[j] = i0 + j < n_isects ? keys[i0 + j] : (KT)0;
    }
    int prev = i0 ? (int)keys[i0 - 1] : -1;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int64_t i = i0 + j;
        if (i >= n_isects) break;
        const int cur = (int)k[j];
        for (int t = prev + 1; t <= cur; ++t) offsets[t] = (int)i;
        prev = cur;
        if (i == n_isects - 1)
            for (int t = cur + 1; t <= n_tiles; ++t) offsets[t] = (int)n_isects;
    }
}

int launch_offsets(int64_t n_isects, int n_tiles, const void *keys, bool key16, int *offsets, cudaStream_t st, bool key8) {
    const int64_t work = n_isects > 0 ? n_isects : (int64_t)n_tiles + 1;
    const int per = key8 ? 32 : key16 ? 16 : 8;  // keys per thread (offsets_kernel::kPer)
    const unsigned blocks = (unsigned)((work + 256 * per - 1) / (256 * per));
    if (key8)
        offsets_kernel<<<blocks, 256, 0, st>>>(n_isects, n_tiles, (const unsigned char *)keys, offsets);
    else if (key16)
        offsets_kernel<<<blocks, 256, 0, st>>>(n_isects, n_tiles, (const unsigned short *)keys, offsets);
    else
        offsets_kernel<<<blocks, 256, 0, st>>>(n_isects, n_tiles, (const unsigned *)keys, offsets);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace gwbp
