// Stage 2: prefix scan of per-Gaussian (visible, tiles) counts, stable radix sort of
// (tile | depth-bits) keys, per-tile range finding.  Integer work, bit-exact against the oracle.
// Semantics: gsplat-1.4.0 isect_tiles / cub radix sort / isect_offset_encode.
//
// The scan and the sort are CUB device primitives (part of the CUDA toolkit, like cuBLAS for a
// plain GEMM); they are HBM-bound passes over 8 B (scan) and 12 B (sort) records.  Only the bits
// that can differ are sorted: 32 depth bits + floor(log2(tiles))+1 tile bits.
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

int launch_scan(int64_t n, WsDev ws, cudaStream_t st) {
    size_t need = 0;
    GWBP_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, need, ws.cnt, ws.scan, (long long)(n + 1), st));
    GWBP_REQUIRE(need <= ws.cub_tmp_bytes, "scan scratch too small: %zu > %zu", need, ws.cub_tmp_bytes);
    size_t b = ws.cub_tmp_bytes;
    GWBP_CUDA_OK(cub::DeviceScan::ExclusiveSum(ws.cub_tmp, b, ws.cnt, ws.scan, (long long)(n + 1), st));
    return 0;
}

int launch_sort(int64_t n_isects, int tile_bits, WsDev ws, int *sorted_buf, cudaStream_t st) {
    *sorted_buf = 0;
    if (n_isects == 0) return 0;
    cub::DoubleBuffer<long long> k(ws.keys[0], ws.keys[1]);
    cub::DoubleBuffer<int> v(ws.vals[0], ws.vals[1]);
    size_t need = 0;
    GWBP_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, need, k, v, (long long)n_isects, 0, 32 + tile_bits, st));
    GWBP_REQUIRE(need <= ws.cub_tmp_bytes, "sort scratch too small: %zu > %zu", need, ws.cub_tmp_bytes);
    size_t b = ws.cub_tmp_bytes;
    GWBP_CUDA_OK(cub::DeviceRadixSort::SortPairs(ws.cub_tmp, b, k, v, (long long)n_isects, 0, 32 + tile_bits, st));
    *sorted_buf = k.selector;
    if (v.selector != k.selector) {
        set_error("radix sort returned mismatched key/value buffers");
        return -1;
    }
    return 0;
}

__global__ void __launch_bounds__(256) offsets_kernel(int64_t n_isects, int n_tiles,
                                                      const long long *__restrict__ keys, int *__restrict__ offsets) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n_isects == 0) {
        if (i <= n_tiles) offsets[i] = 0;
        return;
    }
    if (i >= n_isects) return;
    const int cur = (int)(keys[i] >> 32);
    const int prev = i ? (int)(keys[i - 1] >> 32) : -1;
    for (int t = prev + 1; t <= cur; ++t) offsets[t] = (int)i;
    if (i == n_isects - 1)
        for (int t = cur + 1; t <= n_tiles; ++t) offsets[t] = (int)n_isects;
}

int launch_offsets(int64_t n_isects, int n_tiles, const long long *keys, int *offsets, cudaStream_t st) {
    const int64_t work = n_isects > 0 ? n_isects : (int64_t)n_tiles + 1;
    offsets_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>(n_isects, n_tiles, keys, offsets);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace gwbp
