// B200 probe: (A) validates the UMMA descriptor / canonical-layout conventions used by
// backproject_tc.cu against a host GEMM, (B) measures scattered fp32 row-accumulation throughput
// for the candidate epilogue forms, (C) measures bulk-copy (TMA) streaming bandwidth from an
// L2-resident and an HBM-resident region.  Build: see tools/build_probe.sh.  Not product code.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../3dgs-gradient-backprojection_b200/csrc/tc_common.cuh"

using namespace gwbp::tc;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

// ------------------------------------------------------------------------------------------------
// (A) one CTA: D[128 x N] = A[128 x K] * B[N x K]^T with A, B bf16 MN-major SWIZZLE_NONE in smem
// ------------------------------------------------------------------------------------------------
template <int N, int K>
__global__ void __launch_bounds__(128) umma_probe(const float *__restrict__ A, const float *__restrict__ B,
                                                  float *__restrict__ D, int swap_lbo_sbo, int two_pass) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    constexpr uint32_t SBO_A = 128, LBO_A = 16 * 128;      // 16 M-groups of 8
    constexpr uint32_t SBO_B = 128, LBO_B = (N / 8) * 128;  // N/8 N-groups of 8
    uint8_t *sA = smem;                                     // K/8 * LBO_A bytes
    uint8_t *sB = smem + (K / 8) * LBO_A;
    const int tid = threadIdx.x, warp = tid >> 5;
    // element (m,k) -> (m/8)*SBO + (k/8)*LBO + (k%8)*16 + (m%8)*2
    for (int i = tid; i < 128 * K; i += 128) {
        const int m = i % 128, k = i / 128;
        *reinterpret_cast<__nv_bfloat16 *>(sA + (m / 8) * SBO_A + (k / 8) * LBO_A + (k % 8) * 16 + (m % 8) * 2) =
            __float2bfloat16(A[m * K + k]);
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i % N, k = i / N;
        *reinterpret_cast<__nv_bfloat16 *>(sB + (n / 8) * SBO_B + (k / 8) * LBO_B + (k % 8) * 16 + (n % 8) * 2) =
            __float2bfloat16(B[n * K + k]);
    }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base));
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        mbar_init_fence();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, N, true, true);
        for (int pass = 0; pass <= two_pass; ++pass)
            for (int ks = 0; ks < K / 16; ++ks) {
                const uint32_t a0 = smem_u32(sA) + ks * 2 * LBO_A, b0 = smem_u32(sB) + ks * 2 * LBO_B;
                const uint64_t da = swap_lbo_sbo ? umma_smem_desc(a0, SBO_A, LBO_A) : umma_smem_desc(a0, LBO_A, SBO_A);
                const uint64_t db = swap_lbo_sbo ? umma_smem_desc(b0, SBO_B, LBO_B) : umma_smem_desc(b0, LBO_B, SBO_B);
                umma_bf16(tm, da, db, idesc, (ks > 0 || pass > 0) ? 1u : 0u);
            }
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    for (int c = 0; c < N; c += 32) {
        float v[32];
        tmem_ld32(tm + ((uint32_t)(32 * warp) << 16) + c, v);
        for (int j = 0; j < 32; ++j)
            if (c + j < N) D[(32 * warp + (tid & 31)) * N + c + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

template <int N, int K>
static void run_umma_probe() {
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    srand(1);
    auto rb = []() {  // bf16-exact values
        float f = (float)(rand() % 255 - 127) / 64.0f;
        return f;
    };
    for (auto &x : A) x = rb();
    for (auto &x : B) x = rb();
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4));
    CK(cudaMalloc(&dB, B.size() * 4));
    CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)(K / 8) * 16 * 128 + (size_t)(K / 8) * (N / 8) * 128;
    CK(cudaFuncSetAttribute(umma_probe<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int two = 0; two < 2; ++two)
        for (int swap = 0; swap < 1; ++swap) {  // swapped LBO/SBO faults (verified on B200): the cute convention is right
            CK(cudaMemset(dD, 0, D.size() * 4));
            umma_probe<N, K><<<1, 128, smem>>>(dA, dB, dD, swap, two);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("[umma N=%d K=%d swap=%d two=%d] kernel error: %s\n", N, K, swap, two, cudaGetErrorString(e));
                exit(2);
            }
            CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
            double maxerr = 0, maxref = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    double r = 0;
                    for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[n * K + k];
                    r *= (two ? 2.0 : 1.0);
                    maxerr = fmax(maxerr, fabs(r - D[m * N + n]));
                    maxref = fmax(maxref, fabs(r));
                }
            printf("[umma N=%d K=%d lbo/sbo-swapped=%d accumulate-twice=%d] max|err|=%.4g (max|ref|=%.4g) %s\n", N, K,
                   swap, two, maxerr, maxref, maxerr < 1e-3 * maxref ? "MATCH" : "mismatch");
        }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

// (A2) same GEMM with B K-major (the forward render's W operand: element (n,k) = 8 consecutive k per 16-byte
// core-matrix row): validates which of LBO/SBO is the K-direction stride for a K-major SWIZZLE_NONE operand.
template <int N, int K>
__global__ void __launch_bounds__(128) umma_probe_kmajor_b(const float *__restrict__ A, const float *__restrict__ B,
                                                           float *__restrict__ D, int swap_lbo_sbo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    constexpr uint32_t SBO_A = 128, LBO_A = 16 * 128;
    constexpr uint32_t NSTR_B = 128, KSTR_B = (N / 8) * 128;  // N-group stride / K-group stride of B's core matrices
    uint8_t *sA = smem;
    uint8_t *sB = smem + (K / 8) * LBO_A;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * K; i += 128) {
        const int m = i % 128, k = i / 128;
        *reinterpret_cast<__nv_bfloat16 *>(sA + (m / 8) * SBO_A + (k / 8) * LBO_A + (k % 8) * 16 + (m % 8) * 2) =
            __float2bfloat16(A[m * K + k]);
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i % N, k = i / N;
        *reinterpret_cast<__nv_bfloat16 *>(sB + (n / 8) * NSTR_B + (k / 8) * KSTR_B + (n % 8) * 16 + (k % 8) * 2) =
            __float2bfloat16(B[n * K + k]);
    }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base));
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        mbar_init_fence();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, N, true, false);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t a0 = smem_u32(sA) + ks * 2 * LBO_A, b0 = smem_u32(sB) + ks * 2 * KSTR_B;
            const uint64_t da = umma_smem_desc(a0, LBO_A, SBO_A);
            const uint64_t db = swap_lbo_sbo ? umma_smem_desc(b0, NSTR_B, KSTR_B) : umma_smem_desc(b0, KSTR_B, NSTR_B);
            umma_bf16(tm, da, db, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    for (int c = 0; c < N; c += 32) {
        float v[32];
        tmem_ld32(tm + ((uint32_t)(32 * warp) << 16) + c, v);
        for (int j = 0; j < 32; ++j)
            if (c + j < N) D[(32 * warp + (tid & 31)) * N + c + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

template <int N, int K>
static void run_umma_probe_kmajor_b(int swap) {
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    srand(2);
    for (auto &x : A) x = (float)(rand() % 255 - 127) / 64.0f;
    for (auto &x : B) x = (float)(rand() % 255 - 127) / 64.0f;
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4));
    CK(cudaMalloc(&dB, B.size() * 4));
    CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, D.size() * 4));
    const size_t smem = (size_t)(K / 8) * 16 * 128 + (size_t)(K / 8) * (N / 8) * 128;
    CK(cudaFuncSetAttribute(umma_probe_kmajor_b<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_probe_kmajor_b<N, K><<<1, 128, smem>>>(dA, dB, dD, swap);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("[umma B K-major N=%d K=%d kdir-stride-in-%s] kernel error: %s\n", N, K, swap ? "SBO" : "LBO",
               cudaGetErrorString(e));
        exit(2);
    }
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double r = 0;
            for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[n * K + k];
            maxerr = fmax(maxerr, fabs(r - D[m * N + n]));
            maxref = fmax(maxref, fabs(r));
        }
    printf("[umma B K-major N=%d K=%d kdir-stride-in-%s] max|err|=%.4g (max|ref|=%.4g) %s\n", N, K,
           swap ? "SBO" : "LBO", maxerr, maxref, maxerr < 1e-3 * maxref ? "MATCH" : "mismatch");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

// ------------------------------------------------------------------------------------------------
// (B) scattered accumulation of 2 KB fp32 rows into a table much larger than L2
// ------------------------------------------------------------------------------------------------
constexpr int ROW = 512;

__global__ void __launch_bounds__(256) acc_warp_row_v4(float *tab, const int *rows, int nrows) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int r = warp; r < nrows; r += nw) {
        float *dst = tab + (size_t)rows[r] * ROW;
#pragma unroll
        for (int j = 0; j < ROW / 128; ++j) red_add_v4(dst + j * 128 + lane * 4, 1.f, 2.f, 3.f, 4.f);
    }
}
__global__ void __launch_bounds__(256) acc_warp_row_scalar(float *tab, const int *rows, int nrows) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int r = warp; r < nrows; r += nw) {
        float *dst = tab + (size_t)rows[r] * ROW;
#pragma unroll
        for (int j = 0; j < ROW / 32; ++j) atomicAdd(dst + j * 32 + lane, 1.f);
    }
}
__global__ void __launch_bounds__(256) acc_lane_row_v4(float *tab, const int *rows, int nrows) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int r = t; r < nrows; r += nt) {
        float *dst = tab + (size_t)rows[r] * ROW;
#pragma unroll 8
        for (int j = 0; j < ROW / 4; ++j) red_add_v4(dst + j * 4, 1.f, 2.f, 3.f, 4.f);
    }
}
// what tcgen05.ld.16x256b hands out: a quad of lanes holds 8 consecutive columns (32 B) of one row, 8 rows per warp
__global__ void __launch_bounds__(256) acc_quad_sector_v2(float *tab, const int *rows, int nrows) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31, sub = lane >> 2, q = lane & 3;
    for (int r0 = warp * 8; r0 < nrows; r0 += nw * 8) {
        const int r = r0 + sub;
        if (r >= nrows) continue;
        float *dst = tab + (size_t)rows[r] * ROW + 2 * q;
#pragma unroll 8
        for (int j = 0; j < ROW / 8; ++j)
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst + 8 * j), "f"(1.f), "f"(2.f) : "memory");
    }
}
// 4 lanes x 16 B = 64-byte row pieces, 8 rows per instruction (the current epilogue after a smem transpose)
__global__ void __launch_bounds__(256) acc_8rows_64B(float *tab, const int *rows, int nrows) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31, sub = lane >> 2, q = lane & 3;
    for (int r0 = warp * 8; r0 < nrows; r0 += nw * 8) {
        const int r = r0 + sub;
        if (r >= nrows) continue;
        float *dst = tab + (size_t)rows[r] * ROW + 4 * q;
#pragma unroll 8
        for (int j = 0; j < ROW / 16; ++j) red_add_v4(dst + 16 * j, 1.f, 2.f, 3.f, 4.f);
    }
}
template <int BYTES>
__global__ void __launch_bounds__(256) acc_bulk(float *tab, const int *rows, int nrows) {
    extern __shared__ __align__(128) uint8_t sm[];
    float *stage = reinterpret_cast<float *>(sm);  // one 2 KB row per warp
    const int warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    float *mine = stage + warp_in_cta * ROW;
    for (int j = lane; j < ROW; j += 32) mine[j] = 1.0f;
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        int pending = 0;
        for (int r = warp; r < nrows; r += nw) {
            float *dst = tab + (size_t)rows[r] * ROW;
#pragma unroll
            for (int o = 0; o < ROW * 4; o += BYTES)
                bulk_reduce_add_f32(reinterpret_cast<uint8_t *>(dst) + o, smem_u32(mine) + o, BYTES);
            bulk_commit();
            if (++pending == 8) { bulk_wait_read<4>(); pending = 4; }
        }
        bulk_wait_all<0>();
    }
}

static void run_acc_probe() {
    const size_t table_rows = 4u << 20;  // 4M rows x 2 KB = 8.6 GB >> L2
    const int nrows = 1 << 20;           // 2.1 GB of payload per launch
    float *tab;
    int *rows;
    CK(cudaMalloc(&tab, table_rows * ROW * 4));
    CK(cudaMemset(tab, 0, table_rows * ROW * 4));
    std::vector<int> h(nrows);
    srand(7);
    for (auto &x : h) x = (int)(((size_t)rand() * 65536u + rand()) % table_rows);
    CK(cudaMalloc(&rows, nrows * 4));
    CK(cudaMemcpy(rows, h.data(), nrows * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8;
    auto time_it = [&](const char *name, auto launch) {
        launch();
        CK(cudaDeviceSynchronize());
        float best = 1e9;
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            best = fminf(best, ms);
        }
        printf("[acc %-28s] %.3f ms  payload %.0f GB/s\n", name, best, (double)nrows * ROW * 4 / best / 1e6);
    };
    time_it("warp-per-row red.v4 (512B/instr)", [&] { acc_warp_row_v4<<<grid, 256>>>(tab, rows, nrows); });
    time_it("warp-per-row red.f32 (128B/instr)", [&] { acc_warp_row_scalar<<<grid, 256>>>(tab, rows, nrows); });
    time_it("lane-per-row red.v4 (scattered)", [&] { acc_lane_row_v4<<<grid, 256>>>(tab, rows, nrows); });
    time_it("quad-per-sector red.v2 (8 rows/instr)", [&] { acc_quad_sector_v2<<<grid, 256>>>(tab, rows, nrows); });
    time_it("8 rows x 64B red.v4 / instr", [&] { acc_8rows_64B<<<grid, 256>>>(tab, rows, nrows); });
    CK(cudaFuncSetAttribute(acc_bulk<2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * ROW * 4));
    time_it("bulk reduce 2048B/row", [&] { acc_bulk<2048><<<grid, 256, 8 * ROW * 4>>>(tab, rows, nrows); });
    time_it("bulk reduce 8x256B/row", [&] { acc_bulk<256><<<grid, 256, 8 * ROW * 4>>>(tab, rows, nrows); });
    time_it("bulk reduce 16x128B/row", [&] { acc_bulk<128><<<grid, 256, 8 * ROW * 4>>>(tab, rows, nrows); });
    // locality: every row hit 4x in a row (neighbouring tiles of one Gaussian)
    for (int i = 0; i < nrows; ++i) h[i] = h[(i / 4) * 4];
    CK(cudaMemcpy(rows, h.data(), nrows * 4, cudaMemcpyHostToDevice));
    time_it("warp-per-row red.v4, 4x reuse", [&] { acc_warp_row_v4<<<grid, 256>>>(tab, rows, nrows); });
    time_it("bulk reduce 2048B, 4x reuse", [&] { acc_bulk<2048><<<grid, 256, 8 * ROW * 4>>>(tab, rows, nrows); });
    cudaFree(tab); cudaFree(rows);
}

// ------------------------------------------------------------------------------------------------
// (C) bulk-copy streaming: each CTA pulls 16 KB blocks into a 4-deep smem ring
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) stream_probe(const uint8_t *src, size_t region_bytes, int blocks_per_cta,
                                                   unsigned long long *sink) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t full[4];
    constexpr uint32_t BLK = 16384;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&full[i]), 1);
        mbar_init_fence();
        const size_t nblk = region_bytes / BLK;
        size_t cur = ((size_t)blockIdx.x * 7919u) % nblk;
        for (int i = 0; i < blocks_per_cta + 4; ++i) {
            if (i >= 4) mbar_wait(smem_u32(&full[i & 3]), ((i - 4) >> 2) & 1);
            if (i < blocks_per_cta) {
                mbar_arrive_expect_tx(smem_u32(&full[i & 3]), BLK);
                bulk_g2s(smem_u32(sm) + (i & 3) * BLK, src + cur * BLK, BLK, smem_u32(&full[i & 3]));
                cur = (cur + gridDim.x) % nblk;
            }
        }
        sink[blockIdx.x] = sm[0];
    }
}

static void run_stream_probe() {
    const size_t big = 4ull << 30;
    uint8_t *buf;
    unsigned long long *sink;
    CK(cudaMalloc(&buf, big));
    CK(cudaMemset(buf, 1, big));
    CK(cudaMalloc(&sink, 4096 * 8));
    CK(cudaFuncSetAttribute(stream_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (size_t region : {(size_t)48 << 20, (size_t)96 << 20, big}) {
        for (int ctas_per_sm : {1, 2}) {
            const int grid = 148 * ctas_per_sm, per = 2048;
            stream_probe<<<grid, 32, 65536>>>(buf, region, per, sink);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            stream_probe<<<grid, 32, 65536>>>(buf, region, per, sink);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("[stream region=%5zu MB ctas/SM=%d] %.3f ms  %.0f GB/s  (%.1f B/clk/SM @1.9GHz)\n", region >> 20,
                   ctas_per_sm, ms, (double)grid * per * 16384 / ms / 1e6,
                   (double)grid * per * 16384 / ms / 1e6 / 148 / 1.9);
        }
    }
    cudaFree(buf); cudaFree(sink);
}

int main(int argc, char **argv) {
    const char *what = argc > 1 ? argv[1] : "all";
    if (!strcmp(what, "all") || !strcmp(what, "umma")) {
        run_umma_probe<256, 64>();
        run_umma_probe<64, 32>();
        run_umma_probe<16, 16>();
        run_umma_probe<112, 32>();
    }
    if (!strcmp(what, "kmajor")) {  // run each convention in its own process: the wrong one may fault
        const int swap = argc > 2 ? atoi(argv[2]) : 0;
        run_umma_probe_kmajor_b<256, 64>(swap);
        run_umma_probe_kmajor_b<64, 32>(swap);
    }
    if (!strcmp(what, "all") || !strcmp(what, "acc")) run_acc_probe();
    if (!strcmp(what, "all") || !strcmp(what, "stream")) run_stream_probe();
    return 0;
}
