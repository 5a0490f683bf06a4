// Stage 2: prefix scans, the two stable radix sorts and per-tile range finding.  Integer work,
// bit-exact against the oracle.  Semantics: gsplat-1.4.0 isect_tiles / radix sort /
// isect_offset_encode -- the sorted (tile, depth, packed index) order is the same,
// but instead of one 45-bit sort over all I intersections (6 onesweep passes x 12 B x 2) the n_vis
// visible Gaussians are depth-sorted first (4 passes x 8 B) and, after emission in depth order, a
// stable sort on the <= 13 tile bits (2 passes x 8 B) finishes the job.
//
// Scan and sort are CUB device primitives (part of the CUDA toolkit, like cuBLAS for a plain GEMM).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace gwbp {

// Host-only (no CUDA call, so gwbp_workspace_layout works without a GPU): a generous bound on
// CUB's scratch.  With DoubleBuffer the sort needs O(#blocks) scratch, not O(n); the real
// requirement is queried at launch and checked against this bound.
size_t binning_tmp_bytes(int64_t n, int64_t cap) {
    size_t m = (size_t)(8u << 20) + (size_t)cap / 2 + (size_t)n / 4;
    return (m + 255) & ~(size_t)255;
}

template <typename K, typename V>
static int sort_pairs(K *k0, K *k1, V *v0, V *v1, int64_t n, int bits, WsDev ws, int *sorted_buf, cudaStream_t st) {
    *sorted_buf = 0;
    if (n == 0) return 0;
    cub::DoubleBuffer<K> k(k0, k1);
    cub::DoubleBuffer<V> v(v0, v1);
    size_t need = 0;
    GWBP_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, need, k, v, (long long)n, 0, bits, st));
    GWBP_REQUIRE(need <= ws.cub_tmp_bytes, "sort scratch too small: %zu > %zu", need, ws.cub_tmp_bytes);
    size_t b = ws.cub_tmp_bytes;
    GWBP_CUDA_OK(cub::DeviceRadixSort::SortPairs(ws.cub_tmp, b, k, v, (long long)n, 0, bits, st));
    count_launches(2 + (bits + 7) / 8);  // histogram + exclusive-sum + one onesweep pass per 8-bit digit
    *sorted_buf = k.selector;
    if (v.selector != k.selector) {
        set_error("radix sort returned mismatched key/value buffers");
        return -1;
    }
    return 0;
}

// stage 1: visible Gaussians by depth bits (stable: ties keep ascending packed index)
int launch_depth_sort(int64_t n_vis, WsDev ws, int *sorted_buf, cudaStream_t st) {
    return sort_pairs(ws.dkeys[0], ws.dkeys[1], ws.dvals[0], ws.dvals[1], n_vis, 32, ws, sorted_buf, st);
}

// exclusive offsets of the per-Gaussian tile counts in depth order (n_vis + 1 entries)
int launch_scan_counts(int64_t n_vis, WsDev ws, cudaStream_t st) {
    size_t need = 0;
    GWBP_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, need, ws.cnt2, ws.base2, (long long)(n_vis + 1), st));
    GWBP_REQUIRE(need <= ws.cub_tmp_bytes, "scan scratch too small: %zu > %zu", need, ws.cub_tmp_bytes);
    size_t b = ws.cub_tmp_bytes;
    GWBP_CUDA_OK(cub::DeviceScan::ExclusiveSum(ws.cub_tmp, b, ws.cnt2, ws.base2, (long long)(n_vis + 1), st));
    count_launches(2);
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
int launch_super_sort(int64_t n_entries, int bits, WsDev ws, int key_bytes, int *sorted_buf, cudaStream_t st) {
    if (key_bytes == 1)
        return sort_pairs((unsigned char *)ws.tkeys[0], (unsigned char *)ws.tkeys[1], ws.svals[0], ws.svals[1], n_entries,
                          bits, ws, sorted_buf, st);
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
