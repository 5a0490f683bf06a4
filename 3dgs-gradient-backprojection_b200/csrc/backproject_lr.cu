// Stage 3 for ENCODER-RESOLUTION feature maps: the back-projection against the map the
// reference actually has before `F.interpolate` (backproject.py:108-112: [512, 240, 240] bilinear; :236-249:
// 64 x 64 DINOv2 tokens, nearest) WITHOUT ever building the [H, W, D] map or its packed copy (2 x 2.2 GB at config G).
//
//      num[g, :] += sum_p w(g,p) * (U F_low)[p, :]  =  sum_q ( sum_p w(g,p) U[p,q] ) F_low[q, :]
//
// U = the interpolation operator (4 taps per pixel, torch align_corners=False arithmetic).  Per (tile, batch of 128
// Gaussians) the kernel runs TWO chained tcgen05 GEMMs instead of one big one:
//      GEMM1   W'[128 g x 64 q] = W[128 g x 256 px] . U[256 px x 64 q]     (q = the tile's window of the low-res map:
//                                                                          8 source rows x 8 texels)
//      GEMM2   acc[128 g x D]  = W'[128 g x 64 q]  . F_low[64 q x D]       (F_low window fetched by TMA tensor maps)
// i.e. the weights are DOWN-sampled on the tensor cores (the adjoint of the up-sample) and contracted with the
// L2-resident low-resolution map: ~1/5 of the tensor work of bp_tc_kernel, 1/4 of its shared-memory operand traffic,
// and the weight block W is consumed by one short sweep, so the ALU warps never wait for a second column-chunk sweep.
//
// Persistent CTA, 1 per SM, 736 threads, warp-specialised:
//   warps 0-7   ALU       : exactly bp_tc_kernel's weight generation (thread = pixel, sequential T, bf16 hi/lo W^T)
//   warps 8-11  epilogue  : exactly bp_tc_kernel's (tcgen05.ld -> smem transpose -> red.global.add.v4.f32 rows)
//   warps 12-15 converter : tcgen05.ld W' (fp32, lane = Gaussian) -> bf16 hi/lo -> A operand of GEMM2 in smem
//   warp 16     producer  : cp.async.bulk.tensor.3d (TMA tensor map) of the F_low window, K-step by K-step
//   warp 17     MMA       : one elected lane issues both GEMMs
//   warps 18-20 U writers : regenerate the 16-pixel-row slices of U (4 KB each, hi/lo) per batch, one ring slot per warp
//                           (a single writer made the GEMM1 sweep wait ~600 cycles per slice: profiles/r02_bp_lr_*)
// Split-bf16 everywhere (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM): ~2^-16 per contraction.
//
// Measured alternative (git history, "separable down-sampling"): contracting x on the tensor cores against a static 1 KB
// operand (16 MMAs of N = 16 per batch) and applying the y taps in the converters removes the U writers, but the 16 tiny
// MMAs still take ~6 k cycles to issue (~125 cycles each, size-independent) and the converters' 16 dependent TMEM
// loads ~9 k: 1.17 ms against 1.08 ms for this version (profiles/r02_lr_trace_separable.txt).
//
// TMEM (512 columns): acc buffers at 0 / 192 (192 columns each), W' buffers at 384 / 448 (64 columns each).
// The packed low-res map (flow_pack_kernel) is [y][channel group][x][8 channels] bf16, hi and lo: a TMA box of
// {8 texels x 8 channels, 24 groups, 2 rows} lands in shared memory directly in UMMA core-matrix order.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "lister.cuh"
#include "tc_common.cuh"

namespace gwbp {

using namespace tc;

namespace {

constexpr int MB = 128;            // Gaussians per batch = UMMA M
constexpr int KSL = 16;            // pixels per K-slice of GEMM1 (one tile row) = one UMMA K step
constexpr int QY = 8, QX = 8;      // low-res window of a tile: source rows x texels (one core matrix per row)
constexpr int NQ = QY * QX;        // 64 = GEMM1 N = GEMM2 K
constexpr int NC2 = 192;           // feature columns per GEMM2 chunk (2 x 192 + 2 x 64 = all 512 TMEM columns)
constexpr int NG2 = NC2 / 8;       // channel groups per chunk = TMA box extent
constexpr int K2STEPS = NQ / 16;   // 4
constexpr int RING = 3;            // batches in flight between ALU and epilogue
constexpr uint32_t A_SBO = 128, A_LBO = (MB / 8) * 128;    // W^T and W'^T: MN-major, 16 row-groups of 8 Gaussians per K-group
constexpr int W_PART_BYTES = (kTilePix / 8) * A_LBO;       // 64 KB per hi / lo part
constexpr int A2_PART_BYTES = (NQ / 8) * A_LBO;            // 16 KB per hi / lo part
constexpr uint32_t U_LBO = (NQ / 8) * 128;                 // U slice: [16 px x 64 q], MN-major
constexpr int U_PART_BYTES = KSL * NQ * 2;                 // 2 KB
constexpr int U_SLOT_BYTES = 2 * U_PART_BYTES;
constexpr int NUSLOT = 3;
constexpr uint32_t F_LBO = NG2 * 128;                      // F_low stage: [16 q x 192 cols], MN-major as TMA writes it
constexpr int F_PART_BYTES = KSL * NC2 * 2;                // 6 KB
constexpr int F_STAGE_BYTES = 2 * F_PART_BYTES;
constexpr int NFSTAGE = 3;
constexpr int EPI_COLS = 32, EPI_ROWS = 16, EPI_PITCH = EPI_COLS * 4 + 16;
constexpr int TM_ACC = 0, TM_D1 = 2 * NC2;                 // TMEM column bases

constexpr int kEpiWarp0 = 8, kCvtWarp0 = 12, kProducerWarp = 16, kMmaWarp = 17, kUWarp0 = 18, kListerWarp0 = 18 + NUSLOT,
              kThreads = (20 + NUSLOT) * 32;  // + two lister warps

struct RowInfo {
    int gid[MB];
    float den[MB];
    int exit_flag, pad[3];
};

struct Smem {
    static constexpr int w_hi = 0;
    static constexpr int w_lo = W_PART_BYTES;
    static constexpr int a2_hi = 2 * W_PART_BYTES;
    static constexpr int a2_lo = a2_hi + A2_PART_BYTES;
    static constexpr int uring = a2_lo + A2_PART_BYTES;
    static constexpr int fring = uring + NUSLOT * U_SLOT_BYTES;
    static constexpr int stage_out = fring + NFSTAGE * F_STAGE_BYTES;
    static constexpr int gbuf = stage_out + 4 * EPI_ROWS * EPI_PITCH;
    static constexpr int rows = gbuf + MB * 24;
    static constexpr int ctrl = rows + RING * (int)sizeof(RowInfo);
    static constexpr int lring = ctrl + 64;  // lst::Ring: id batches from the lister warp
    static constexpr int bars = lring + lst::NLISTERS * (int)sizeof(lst::Ring);
    static constexpr int w_full = 0, w_free = 8, u_full = 16, u_empty = u_full + NUSLOT, f_full = u_empty + NUSLOT,
                         f_empty = f_full + NFSTAGE, d1_full = f_empty + NFSTAGE, d1_empty = d1_full + 2,
                         a2_full = d1_empty + 2, a2_empty = a2_full + 1, acc_full = a2_empty + 1, acc_empty = acc_full + 2,
                         rows_ready = acc_empty + 2, rows_free = rows_ready + RING, ctrl_full = rows_free + RING,
                         ctrl_empty = ctrl_full + RING, l_full = ctrl_empty + RING, l_free = l_full + lst::NL,
                         nbars = l_full + lst::NLISTERS * 2 * lst::NL;  // per lister: full[NL], free[NL]
    static constexpr int tmem_slot = bars + nbars * 8;
    static constexpr int total = tmem_slot + 16;
};
static_assert(Smem::total + 256 <= 232448, "shared memory budget (227 KB) exceeded");
static_assert(NUSLOT == 3, "the MMA warp tracks three U slots");
static_assert(Smem::fring % 128 == 0 && Smem::uring % 128 == 0 && Smem::a2_hi % 128 == 0, "operand alignment");

__device__ __forceinline__ int bar_red_popc_alu(bool pred) {
    int cnt;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %1, 0;\n\t"
        "bar.red.popc.u32 %0, 1, 256, p;\n\t}"
        : "=r"(cnt)
        : "r"((int)pred)
        : "memory");
    return cnt;
}
__device__ __forceinline__ void bar_sync_alu() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 32 lanes x 16 columns of fp32 (see tmem_ld32)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// TMA: 3-D tensor-map box -> shared memory, completion counted on an mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            dst_smem),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

// torch area_pixel_compute_source_index (align_corners=False) / nearest: the same arithmetic as fpack_lowres_*_kernel
// and oracle/gsplat_oracle.py::upsample
struct SrcIdx {
    int i0, i1;
    float l;
};
__host__ __device__ __forceinline__ SrcIdx src_index(int o, float scale, int n_src, int nearest) {
    SrcIdx s;
    if (nearest) {
        s.i0 = s.i1 = min((int)floorf((float)o * scale), n_src - 1);
        s.l = 0.0f;
    } else {
        const float f = fmaxf(scale * ((float)o + 0.5f) - 0.5f, 0.0f);
        s.i0 = min((int)f, n_src - 1);
        s.i1 = s.i0 + (s.i0 < n_src - 1 ? 1 : 0);
        s.l = f - (float)s.i0;
    }
    return s;
}

constexpr int kBand = 4;
using lst::unit_to_tile;

struct LrArgs {
    TileCtx t;
    float *num, *den;
    int d, dp, nchunks, nunits;
    int sh, sw, nearest;      // low-res map size, interpolation mode
    float scale_y, scale_x;   // sh / H, sw / W as fp32 quotients (torch's area_pixel_compute_scale)
    int *unit_counter;
    long long *stats;
    int debug;  // GWBP_LR_DEBUG (experiment builds only): 1 = skip the accumulator reductions (timing only)
    unsigned long long *trace;  // gwbp_debug_set_trace: CTA 0 records (role, event, batch, chunk, clock64); nullptr = off
};

__global__ void __launch_bounds__(kThreads, 1) bp_lr_kernel(const LrArgs a, const __grid_constant__ CUtensorMap map_hi,
                                                            const __grid_constant__ CUtensorMap map_lo) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto bar = [&](int i) -> uint32_t { return sbase + Smem::bars + 8 * i; };
    RowInfo *rows = reinterpret_cast<RowInfo *>(smem + Smem::rows);
    volatile int *ctrl = reinterpret_cast<volatile int *>(smem + Smem::ctrl);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + Smem::tmem_slot);

    if (tid == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(bar(Smem::w_full + i), 1); mbar_init(bar(Smem::w_free + i), 1); }
        for (int i = 0; i < NUSLOT; ++i) { mbar_init(bar(Smem::u_full + i), 1); mbar_init(bar(Smem::u_empty + i), 1); }
        for (int i = 0; i < NFSTAGE; ++i) { mbar_init(bar(Smem::f_full + i), 1); mbar_init(bar(Smem::f_empty + i), 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(Smem::d1_full + i), 1);
            mbar_init(bar(Smem::d1_empty + i), 4);
            mbar_init(bar(Smem::acc_full + i), 1);
            mbar_init(bar(Smem::acc_empty + i), 4);
        }
        mbar_init(bar(Smem::a2_full), 4);
        mbar_init(bar(Smem::a2_empty), 1);
        for (int i = 0; i < RING; ++i) {
            mbar_init(bar(Smem::rows_ready + i), 8);
            mbar_init(bar(Smem::rows_free + i), 4);
            mbar_init(bar(Smem::ctrl_full + i), 1);
            mbar_init(bar(Smem::ctrl_empty + i), 4);  // producer, MMA, U writer 0, converters (warp 12)
        }
        for (int r = 0; r < lst::NLISTERS; ++r) {
            for (int i = 0; i < lst::NL; ++i) {
                mbar_init(bar(Smem::l_full + 2 * lst::NL * r + i), 1);
                mbar_init(bar(Smem::l_free + 2 * lst::NL * r + i), 8);
            }
            reinterpret_cast<lst::Ring *>(smem + Smem::lring)[r].abort_unit = -1;
        }
        mbar_init_fence();
    }
    if (warp == kMmaWarp) tmem_alloc<512>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // optional event trace (debug only): roles 0 = ALU warp 0, 1 = converter warp 12, 2 = epilogue warp 8, 3 = MMA
    int trace_n = 0;
    auto trace = [&](int role, int ev, int q, int c) {
        if (a.trace != nullptr && blockIdx.x == 0 && lane == 0 && trace_n < 4096) {
            unsigned long long *p = a.trace + ((size_t)role * 4096 + trace_n) * 2;
            p[0] = ((unsigned long long)ev << 48) | ((unsigned long long)(c & 0xffff) << 32) | (unsigned)q;
            p[1] = (unsigned long long)clock64();
            ++trace_n;
        }
    };

    if (warp < 8) {
        // ======================================= ALU =========================================
        // identical to bp_tc_kernel's ALU role (same pair arithmetic, same W^T layout, same den butterfly)
        float4 *gbuf = reinterpret_cast<float4 *>(smem + Smem::gbuf);
        lst::Ring *rings = reinterpret_cast<lst::Ring *>(smem + Smem::lring);
        // lr = the ring (lister) the current tile comes from; qlc / alive_c = batches consumed from it / it still has
        // tiles; qlo / alive_o = the same for the other ring (scalars, swapped at a ring switch: no local-memory arrays)
        int q = 0, lr = 0, qlc = 0, qlo = 0;
        bool alive_c = true, alive_o = true;
        long long walked = 0;
        const uint32_t wslab = (uint32_t)(tid >> 3) * A_LBO + (uint32_t)(tid & 7) * 16;
        // The next batch to process, as handed over by the lister warp (which owns the work queue and runs a few
        // batches ahead, across tile boundaries): its tile, size, whether it closes the tile, and (threads 0..127) the
        // record of its row `tid`, loaded while the previous batch is processed.
        int nu = -1, nn = 0, nlast = 1;
        float4 r0 = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)), r1 = make_float4(0.f, 0.f, 0.f, 0.f);
        auto fetch = [&]() {  // next batch of ring lr
            lst::Ring *ring = rings + lr;
            const int ls = qlc % lst::NL;
            mbar_wait(bar(Smem::l_full + 2 * lst::NL * lr + ls), (qlc / lst::NL) & 1);
            nu = ring->unit[ls]; nn = ring->n[ls]; nlast = ring->last[ls];
            r0 = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
            r1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tid < MB && tid < nn) {
                const int id = ring->ids[ls][tid];
                r0 = a.t.grec[2 * (int64_t)id];
                r1 = a.t.grec[2 * (int64_t)id + 1];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(Smem::l_free + 2 * lst::NL * lr + ls));
            ++qlc;
        };
        // first batch of the next tile: the listers take alternate tiles, so it comes from the other ring while that
        // lister still has tiles (a lister that has published its exit marker is never read again)
        auto fetch_next_tile = [&]() {
            while (true) {
                if (alive_o) {
                    lr ^= 1;
                    const int tq = qlc; qlc = qlo; qlo = tq;
                    const bool ta = alive_c; alive_c = alive_o; alive_o = ta;
                } else if (!alive_c) {
                    nu = -1; nn = 0; nlast = 1;
                    return;
                }
                fetch();
                if (nu >= 0) return;
                alive_c = false;
            }
        };
        lr = 1;
        fetch_next_tile();  // ring 0 first
        int unit = -2;
        float2 npx = make_float2(0.f, 0.f), npy = make_float2(0.f, 0.f);
        bool done = true;
        float T = 1.0f;
        while (nu >= 0) {
            if (nu != unit) {  // first batch of a new tile
                unit = nu;
                const int tile = unit_to_tile(unit, a.t.tw, a.t.th, kBand);
                const int ty = tile / a.t.tw, tx = tile % a.t.tw;
                const int yy = ty * kTile + (tid >> 4), xx = tx * kTile + (tid & 15);
                const float px = (float)xx + 0.5f, py = (float)yy + 0.5f;
                npx = make_float2(-px, -px);
                npy = make_float2(-py, -py);
                done = !(yy < a.t.H && xx < a.t.W);
                T = 1.0f;
            }
            const int n_cur = nn;
            const bool last_cur = nlast != 0;
            if (n_cur == 0) {  // empty closing batch of a tile
                fetch_next_tile();
                continue;
            }
            {
                if (bar_red_popc_alu(!done) == 0) {
                    // every pixel of the tile is finished: tell the lister and skip the tile's remaining batches
                    if (!last_cur) {
                        if (tid == 0) rings[lr].abort_unit = unit;
                        do {
                            fetch();
                        } while (!nlast);
                    }
                    fetch_next_tile();
                    continue;
                }
                if (warp == 0) trace(0, 0, q, 0);
                const int slot = q % RING;
                if (q >= RING) mbar_wait(bar(Smem::rows_free + slot), ((q / RING) - 1) & 1);
                if (tid < MB) {
                    float *gp = reinterpret_cast<float *>(gbuf + 3 * (tid >> 1)) + (tid & 1);
                    gp[0] = r0.x; gp[2] = r0.y;
                    gp[4] = 0.5f * r1.x; gp[6] = r1.y;
                    gp[8] = 0.5f * r1.z; gp[10] = r0.z;
                    rows[slot].gid[tid] = __float_as_int(r0.w);
                    rows[slot].den[tid] = 0.0f;
                    if (tid == 0) rows[slot].exit_flag = 0;
                }
                if (tid == 0) {
                    if (q >= RING) mbar_wait(bar(Smem::ctrl_empty + slot), ((q / RING) - 1) & 1);
                    ctrl[slot] = unit;
                    mbar_arrive(bar(Smem::ctrl_full + slot));
                }
                bar_sync_alu();
                // the next batch (of this tile or the next one): its record loads are in flight during this batch
                if (last_cur) fetch_next_tile(); else fetch();
                if (q >= 1) mbar_wait(bar(Smem::w_free + warp), (q - 1) & 1);
                if (warp == 0) trace(0, 1, q, 0);
                walked += n_cur;
                if (__all_sync(0xffffffffu, done)) {
                    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 4
                    for (int j = 0; j < MB / 8; ++j) {
                        *reinterpret_cast<uint4 *>(smem + Smem::w_hi + j * A_SBO + wslab) = z;
                        *reinterpret_cast<uint4 *>(smem + Smem::w_lo + j * A_SBO + wslab) = z;
                    }
                } else {
#pragma unroll 2
                    for (int j = 0; j < MB / 16; ++j) {
                        // every pixel of this warp finished INSIDE this batch (typically the tile's last one), or the
                        // batch holds no more than 16 j Gaussians (the tail of a list that ends before the tile
                        // saturates): the rest of the slab is zero, written without evaluating a single pair
                        if (j > 0 && (16 * j >= n_cur || __all_sync(0xffffffffu, done))) {
                            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                            for (int jj = 2 * j; jj < MB / 8; ++jj) {
                                *reinterpret_cast<uint4 *>(smem + Smem::w_hi + jj * A_SBO + wslab) = z;
                                *reinterpret_cast<uint4 *>(smem + Smem::w_lo + jj * A_SBO + wslab) = z;
                            }
                            break;
                        }
                        float w[16];
#pragma unroll
                        for (int i2 = 0; i2 < 8; ++i2) {
                            const float4 q0 = gbuf[3 * (8 * j + i2)], q1 = gbuf[3 * (8 * j + i2) + 1],
                                         q2 = gbuf[3 * (8 * j + i2) + 2];
                            const float2 dx = add2_rn(make_float2(q0.x, q0.y), npx);
                            const float2 dy = add2_rn(make_float2(q0.z, q0.w), npy);
                            const float2 sg = pair_sigma2(dx, dy, make_float2(q1.x, q1.y), make_float2(q1.z, q1.w),
                                                          make_float2(q2.x, q2.y));
                            const float2 ex = mul2_rn(sg, make_float2(-kLog2e, -kLog2e));
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const float sigma = h ? sg.y : sg.x;
                                const float alpha = fminf(kAlphaMax, __fmul_rn(h ? q2.w : q2.z, ex2_approx(h ? ex.y : ex.x)));
                                const float nT = __fmul_rn(T, __fsub_rn(1.0f, alpha));
                                const bool valid = !done && sigma >= 0.0f && alpha >= kAlphaMin;
                                const bool stop = valid && nT <= kTMin;
                                const bool take = valid && !stop;
                                w[2 * i2 + h] = take ? __fmul_rn(alpha, T) : 0.0f;
                                T = take ? nT : T;
                                done = done || stop;
                            }
                        }
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint4 hi, lo;
                            split_bf16x2(w[8 * h + 0], w[8 * h + 1], hi.x, lo.x);
                            split_bf16x2(w[8 * h + 2], w[8 * h + 3], hi.y, lo.y);
                            split_bf16x2(w[8 * h + 4], w[8 * h + 5], hi.z, lo.z);
                            split_bf16x2(w[8 * h + 6], w[8 * h + 7], hi.w, lo.w);
                            const uint32_t off = (uint32_t)(2 * j + h) * A_SBO + wslab;
                            *reinterpret_cast<uint4 *>(smem + Smem::w_hi + off) = hi;
                            *reinterpret_cast<uint4 *>(smem + Smem::w_lo + off) = lo;
                        }
                        const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2;
                        float v8[8], v4[4], v2[2], v1;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float send = b16 ? w[i] : w[i + 8];
                            v8[i] = (b16 ? w[i + 8] : w[i]) + __shfl_xor_sync(0xffffffffu, send, 16);
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float send = b8 ? v8[i] : v8[i + 4];
                            v4[i] = (b8 ? v8[i + 4] : v8[i]) + __shfl_xor_sync(0xffffffffu, send, 8);
                        }
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const float send = b4 ? v4[i] : v4[i + 2];
                            v2[i] = (b4 ? v4[i + 2] : v4[i]) + __shfl_xor_sync(0xffffffffu, send, 4);
                        }
                        {
                            const float send = b2 ? v2[0] : v2[1];
                            v1 = (b2 ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, send, 2);
                        }
                        v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
                        if ((lane & 1) == 0 && v1 > 0.0f) {
                            const int gi = (b16 ? 8 : 0) + (b8 ? 4 : 0) + (b4 ? 2 : 0) + (b2 ? 1 : 0);
                            atomicAdd(&rows[slot].den[16 * j + gi], v1);
                        }
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar(Smem::w_full + warp));
                    mbar_arrive(bar(Smem::rows_ready + slot));
                }
                if (warp == 0) trace(0, 2, q, 0);
                ++q;
            }
        }
        {   // exit sentinel for the other roles
            const int slot = q % RING;
            if (q >= RING) mbar_wait(bar(Smem::rows_free + slot), ((q / RING) - 1) & 1);
            if (tid == 0) {
                rows[slot].exit_flag = 1;
                if (q >= RING) mbar_wait(bar(Smem::ctrl_empty + slot), ((q / RING) - 1) & 1);
                ctrl[slot] = -1;
                mbar_arrive(bar(Smem::ctrl_full + slot));
            }
            bar_sync_alu();
            if (lane == 0) mbar_arrive(bar(Smem::rows_ready + slot));
        }
        if (a.stats && tid == 0) atomicAdd((unsigned long long *)&a.stats[1], (unsigned long long)walked);
    } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
        // ===================================== epilogue ======================================
        // bp_tc_kernel's epilogue with 192-column accumulator buffers
        const int quarter = warp & 3;
        const uint32_t lane_base = (uint32_t)(32 * quarter) << 16;
        const int r = 32 * quarter + lane;
        uint8_t *wstage = smem + Smem::stage_out + (EPI_ROWS * quarter) * EPI_PITCH;
        uint8_t *srow = wstage + (lane & 15) * EPI_PITCH;
        const int rsub = lane >> 3, piece = lane & 7;
        long long live_rows = 0;
        for (int q = 0;; ++q) {
            const int slot = q % RING;
            mbar_wait(bar(Smem::rows_ready + slot), (q / RING) & 1);
            if (rows[slot].exit_flag) break;
            const int gid = rows[slot].gid[r];
            const float dn = rows[slot].den[r];
            const bool live = (gid >= 0) && (dn > 0.0f);
            const unsigned live_mask = __ballot_sync(0xffffffffu, live);
            int64_t grow[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) grow[i] = (int64_t)__shfl_sync(0xffffffffu, gid, 4 * i + rsub) * a.d;
            for (int c = 0; c < a.nchunks; ++c) {
                const int u = q * a.nchunks + c, ab = u & 1;
                const int ncols = min(NC2, a.dp - c * NC2);
                mbar_wait(bar(Smem::acc_full + ab), (u >> 1) & 1);
                tc_fence_after();
                if (quarter == 0) trace(2, 0, q, c);
                for (int c0 = 0; c0 < ncols; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem + lane_base + (uint32_t)(TM_ACC + ab * NC2 + c0), v);
                    const int col = c * NC2 + c0 + 4 * piece;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        __syncwarp();
                        if (live && (lane >> 4) == half) {
#pragma unroll
                            for (int i = 0; i < 32; i += 4)
                                *reinterpret_cast<float4 *>(srow + 4 * i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        }
                        __syncwarp();
                        if (col < a.d) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int row = 16 * half + 4 * i + rsub;
                                if (live_mask >> row & 1u) {
                                    const float4 x = *reinterpret_cast<const float4 *>(wstage + (4 * i + rsub) * EPI_PITCH + 16 * piece);
                                    if (!(a.debug & 1)) red_add_v4(a.num + grow[4 * half + i] + col, x.x, x.y, x.z, x.w);
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(Smem::acc_empty + ab));
                if (quarter == 0) trace(2, 1, q, c);
            }
            if (live) {
                atomicAdd(a.den + gid, dn);
                ++live_rows;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(Smem::rows_free + slot));
        }
        if (a.stats) {
#pragma unroll
            for (int o = 16; o; o >>= 1) live_rows += __shfl_xor_sync(0xffffffffu, live_rows, o);
            if (lane == 0) atomicAdd((unsigned long long *)&a.stats[0], (unsigned long long)live_rows);
        }
    } else if (warp >= kCvtWarp0 && warp < kCvtWarp0 + 4) {
        // ===================================== converter =====================================
        // W' (fp32 in TMEM, lane = Gaussian row) -> bf16 hi/lo, MN-major A operand of GEMM2 (same layout as W^T)
        const int quarter = warp & 3;
        const uint32_t lane_base = (uint32_t)(32 * quarter) << 16;
        const int g = 32 * quarter + lane;
        const uint32_t goff = (uint32_t)(g >> 3) * A_SBO + (uint32_t)(g & 7) * 2;
        for (int q = 0;; ++q) {
            const int slot = q % RING;
            mbar_wait(bar(Smem::ctrl_full + slot), (q / RING) & 1);
            const int unit = ctrl[slot];
            __syncwarp();
            if (quarter == 0 && lane == 0) mbar_arrive(bar(Smem::ctrl_empty + slot));
            if (unit < 0) break;
            const int db = q & 1;
            mbar_wait(bar(Smem::d1_full + db), (q >> 1) & 1);
            tc_fence_after();
            if (quarter == 0) trace(1, 0, q, 0);
            if (q >= 1) mbar_wait(bar(Smem::a2_empty), (q - 1) & 1);  // GEMM2 of the previous batch has read A2
            if (quarter == 0) trace(1, 1, q, 0);
#pragma unroll 1
            for (int k0 = 0; k0 < NQ; k0 += 16) {
                float t[16];
                tmem_ld16(tmem + lane_base + (uint32_t)(TM_D1 + db * NQ + k0), t);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int k = k0 + i;
                    const __nv_bfloat16 h = __float2bfloat16_rn(t[i]);
                    const __nv_bfloat16 l = __float2bfloat16_rn(t[i] - __bfloat162float(h));
                    const uint32_t off = (uint32_t)(k >> 3) * A_LBO + (uint32_t)(k & 7) * 16 + goff;
                    *reinterpret_cast<__nv_bfloat16 *>(smem + Smem::a2_hi + off) = h;
                    *reinterpret_cast<__nv_bfloat16 *>(smem + Smem::a2_lo + off) = l;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(Smem::d1_empty + db));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(Smem::a2_full));
            if (quarter == 0) trace(1, 2, q, 0);
        }
    } else if (warp == kProducerWarp) {
        // ===================================== producer ======================================
        if (lane == 0) {
            int stage = 0, use = 0;
            for (int q = 0;; ++q) {
                const int slot = q % RING;
                mbar_wait(bar(Smem::ctrl_full + slot), (q / RING) & 1);
                const int unit = ctrl[slot];
                mbar_arrive(bar(Smem::ctrl_empty + slot));
                if (unit < 0) break;
                const int tile = unit_to_tile(unit, a.t.tw, a.t.th, kBand);
                const int ty = tile / a.t.tw, tx = tile % a.t.tw;
                const int ylo = src_index(min(ty * kTile, a.t.H - 1), a.scale_y, a.sh, a.nearest).i0;
                const int xlo = src_index(min(tx * kTile, a.t.W - 1), a.scale_x, a.sw, a.nearest).i0;
                for (int c = 0; c < a.nchunks; ++c) {
                    for (int k2 = 0; k2 < K2STEPS; ++k2) {
                        if (use >= 1) mbar_wait(bar(Smem::f_empty + stage), (use - 1) & 1);
                        mbar_arrive_expect_tx(bar(Smem::f_full + stage), F_STAGE_BYTES);
                        const uint32_t dst = sbase + Smem::fring + stage * F_STAGE_BYTES;
                        tma_load_3d(dst, &map_hi, xlo * 8, c * NG2, ylo + 2 * k2, bar(Smem::f_full + stage));
                        tma_load_3d(dst + F_PART_BYTES, &map_lo, xlo * 8, c * NG2, ylo + 2 * k2, bar(Smem::f_full + stage));
                        if (++stage == NFSTAGE) { stage = 0; ++use; }
                    }
                }
            }
        }
    } else if (warp >= kUWarp0 && warp < kUWarp0 + NUSLOT) {
        // ===================================== U writers =====================================
        // slice ks = the interpolation weights of the tile's pixel row ks onto the 8 x 8 window: [16 px x 64 q], bf16
        // hi/lo, MN-major (a pixel's 8 consecutive q = one 16-byte core-matrix row, q = 8 * source row + texel).
        // Writer j owns ring slot j and the slices ks = j, j + NUSLOT, ...
        const int us = warp - kUWarp0;
        int uuse = 0;
        for (int q = 0;; ++q) {
            const int slot = q % RING;
            mbar_wait(bar(Smem::ctrl_full + slot), (q / RING) & 1);
            const int unit = ctrl[slot];
            __syncwarp();
            if (us == 0 && lane == 0) mbar_arrive(bar(Smem::ctrl_empty + slot));
            if (unit < 0) break;
            const int tile = unit_to_tile(unit, a.t.tw, a.t.th, kBand);
            const int ty = tile / a.t.tw, tx = tile % a.t.tw;
            const int ylo = src_index(min(ty * kTile, a.t.H - 1), a.scale_y, a.sh, a.nearest).i0;
            const int xlo = src_index(min(tx * kTile, a.t.W - 1), a.scale_x, a.sw, a.nearest).i0;
            // this lane's four (pixel, source row) items: item = lane + 32 i -> p = item % 16, row = item / 16
            const int p = lane & 15;
            const SrcIdx sx = src_index(min(tx * kTile + p, a.t.W - 1), a.scale_x, a.sw, a.nearest);
            const int x0 = sx.i0 - xlo, x1 = sx.i1 - xlo;
            float wx[QX];
#pragma unroll
            for (int j = 0; j < QX; ++j) wx[j] = (j == x0 ? 1.0f - sx.l : 0.0f) + (j == x1 ? sx.l : 0.0f);
            uint8_t *blk = smem + Smem::uring + us * U_SLOT_BYTES;
            for (int ks = us; ks < kTilePix / KSL; ks += NUSLOT, ++uuse) {
                const SrcIdx sy = src_index(min(ty * kTile + ks, a.t.H - 1), a.scale_y, a.sh, a.nearest);
                const int y0 = sy.i0 - ylo, y1 = sy.i1 - ylo;
                if (uuse >= 1) mbar_wait(bar(Smem::u_empty + us), (uuse - 1) & 1);
#pragma unroll
                for (int i = 0; i < QY / 2; ++i) {
                    const int row = (lane >> 4) + 2 * i;  // source row of the window: 0..7
                    const float wy = (row == y0 ? 1.0f - sy.l : 0.0f) + (row == y1 ? sy.l : 0.0f);
                    uint4 hi, lo;
                    split_bf16x2(wy * wx[0], wy * wx[1], hi.x, lo.x);
                    split_bf16x2(wy * wx[2], wy * wx[3], hi.y, lo.y);
                    split_bf16x2(wy * wx[4], wy * wx[5], hi.z, lo.z);
                    split_bf16x2(wy * wx[6], wy * wx[7], hi.w, lo.w);
                    const uint32_t off = (uint32_t)(p >> 3) * U_LBO + (uint32_t)row * 128 + (uint32_t)(p & 7) * 16;
                    *reinterpret_cast<uint4 *>(blk + off) = hi;
                    *reinterpret_cast<uint4 *>(blk + U_PART_BYTES + off) = lo;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(Smem::u_full + us));
            }
        }
    } else if (warp >= kListerWarp0) {
        // ====================================== listers ======================================
        const int r = warp - kListerWarp0;
        lst::run_lister(a.t, a.unit_counter, a.nunits, kBand, reinterpret_cast<lst::Ring *>(smem + Smem::lring) + r,
                        bar(Smem::l_full + 2 * lst::NL * r));
    } else if (warp == kMmaWarp) {
        // ======================================= MMA =========================================
        int ucnt[NUSLOT] = {0, 0, 0};  // uses of each U ring slot (slice ks lives in slot ks % NUSLOT, written by writer ks % NUSLOT)
        int fs = 0, fuse = 0;
        const uint64_t a_hi0 = umma_smem_desc(sbase + Smem::w_hi, A_LBO, A_SBO);
        const uint64_t a_lo0 = umma_smem_desc(sbase + Smem::w_lo, A_LBO, A_SBO);
        const uint64_t a2_hi0 = umma_smem_desc(sbase + Smem::a2_hi, A_LBO, A_SBO);
        const uint64_t a2_lo0 = umma_smem_desc(sbase + Smem::a2_lo, A_LBO, A_SBO);
        constexpr uint64_t kAStep = (2 * A_LBO) >> 4;  // start-address advance per 16-element K-slice (W and W')
        constexpr uint32_t idesc1 = umma_idesc_bf16(MB, NQ, true, true);
        for (int q = 0;; ++q) {
            const int slot = q % RING;
            mbar_wait(bar(Smem::ctrl_full + slot), (q / RING) & 1);
            const int unit = ctrl[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(Smem::ctrl_empty + slot));
            if (unit < 0) break;
            // ---- GEMM1: W' = W . U into D1[q & 1]
            const int db = q & 1;
            if (q >= 2) mbar_wait(bar(Smem::d1_empty + db), ((q >> 1) - 1) & 1);
            tc_fence_after();
            trace(3, 0, q, 0);
            const uint32_t d1 = tmem + (uint32_t)(TM_D1 + db * NQ);
#pragma unroll 1
            for (int ks = 0; ks < kTilePix / KSL; ++ks) {
                if ((ks & 1) == 0) mbar_wait(bar(Smem::w_full + (ks >> 1)), q & 1);
                const int us = ks % NUSLOT;
                const int uuse = us == 0 ? ucnt[0] : us == 1 ? ucnt[1] : ucnt[2];
                mbar_wait(bar(Smem::u_full + us), uuse & 1);
                tc_fence_after();
                const uint64_t a_hi = a_hi0 + (uint64_t)ks * kAStep, a_lo = a_lo0 + (uint64_t)ks * kAStep;
                const uint32_t ub = sbase + Smem::uring + us * U_SLOT_BYTES;
                const uint64_t u_hi = umma_smem_desc(ub, U_LBO, 128), u_lo = umma_smem_desc(ub + U_PART_BYTES, U_LBO, 128);
                if (elect_one()) {
                    umma_bf16(d1, a_hi, u_hi, idesc1, ks > 0 ? 1u : 0u);
                    umma_bf16(d1, a_hi, u_lo, idesc1, 1u);
                    umma_bf16(d1, a_lo, u_hi, idesc1, 1u);
                    umma_commit(bar(Smem::u_empty + us));
                    if (ks & 1) umma_commit(bar(Smem::w_free + (ks >> 1)));
                }
                __syncwarp();
                if (us == 0) ++ucnt[0]; else if (us == 1) ++ucnt[1]; else ++ucnt[2];
            }
            if (elect_one()) umma_commit(bar(Smem::d1_full + db));
            __syncwarp();
            trace(3, 1, q, 0);
            // ---- GEMM2: acc = W' . F_low, one 192-column chunk at a time
            mbar_wait(bar(Smem::a2_full), q & 1);
            tc_fence_after();
            trace(3, 2, q, 0);
            for (int c = 0; c < a.nchunks; ++c) {
                const int u = q * a.nchunks + c, ab = u & 1;
                const int ncols = min(NC2, a.dp - c * NC2);
                const uint32_t idesc2 = umma_idesc_bf16(MB, ncols, true, true);
                if (u >= 2) mbar_wait(bar(Smem::acc_empty + ab), ((u >> 1) - 1) & 1);
                tc_fence_after();
                const uint32_t d2 = tmem + (uint32_t)(TM_ACC + ab * NC2);
#pragma unroll 1
                for (int k2 = 0; k2 < K2STEPS; ++k2) {
                    mbar_wait(bar(Smem::f_full + fs), fuse & 1);
                    tc_fence_after();
                    const uint64_t x_hi = a2_hi0 + (uint64_t)k2 * kAStep, x_lo = a2_lo0 + (uint64_t)k2 * kAStep;
                    const uint32_t fb = sbase + Smem::fring + fs * F_STAGE_BYTES;
                    const uint64_t f_hi = umma_smem_desc(fb, F_LBO, 128), f_lo = umma_smem_desc(fb + F_PART_BYTES, F_LBO, 128);
                    if (elect_one()) {
                        umma_bf16(d2, x_hi, f_hi, idesc2, k2 > 0 ? 1u : 0u);
                        umma_bf16(d2, x_hi, f_lo, idesc2, 1u);
                        umma_bf16(d2, x_lo, f_hi, idesc2, 1u);
                        umma_commit(bar(Smem::f_empty + fs));
                    }
                    __syncwarp();
                    if (++fs == NFSTAGE) { fs = 0; ++fuse; }
                }
                if (elect_one()) {
                    umma_commit(bar(Smem::acc_full + ab));
                    if (c == a.nchunks - 1) umma_commit(bar(Smem::a2_empty));
                }
                __syncwarp();
                trace(3, 3, q, c);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc<512>(tmem);
}

// low-res map fp32 [h, w, D] (element strides) -> bf16 hi / lo, [y][channel group][x][8 channels]: a TMA box of 8
// texels x 8 channels is then one 128-byte UMMA core matrix.  One thread per (y, group, x): 8 strided loads
// (coalesced over x for the reference's planar [D, h, w] encoder output), two 16-byte stores.
__global__ void __launch_bounds__(256) flow_pack_kernel(const float *__restrict__ S, int sh, int sw, int64_t ssh, int64_t ssw,
                                                        int64_t ssd, int d, int ngp, uint4 *__restrict__ hi_out,
                                                        uint4 *__restrict__ lo_out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)sh * ngp * sw;
    if (idx >= total) return;
    const int x = (int)(idx % sw);
    const int ng = (int)((idx / sw) % ngp);
    const int y = (int)(idx / ((int64_t)sw * ngp));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = 8 * ng + j;
        v[j] = c < d ? __ldg(S + y * ssh + x * ssw + c * ssd) : 0.0f;
    }
    uint4 hi, lo;
    split_bf16x2(v[0], v[1], hi.x, lo.x);
    split_bf16x2(v[2], v[3], hi.y, lo.y);
    split_bf16x2(v[4], v[5], hi.z, lo.z);
    split_bf16x2(v[6], v[7], hi.w, lo.w);
    hi_out[idx] = hi;
    lo_out[idx] = lo;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

// Can the adjoint kernel take this geometry?  Every tile's window of the low-res map must fit 8 source rows x 8 texels
// (true for any up-sampling factor >= ~2.5, e.g. 240 -> 840 x 1297 and 64 tokens -> anything larger than 160); otherwise the caller falls back to the fused-upsample re-layout + bp_tc_kernel.
bool lr_supported(int W, int H, int sh, int sw, int d, int nearest) {
    if (!(d >= 16 && d <= 2048 && d % 4 == 0) || sh < 1 || sw < 1 || W < 1 || H < 1) return false;
    const float scale_y = (float)sh / (float)H, scale_x = (float)sw / (float)W;
    const int tw = (W + kTile - 1) / kTile, th = (H + kTile - 1) / kTile;
    for (int ty = 0; ty < th; ++ty) {
        const int lo = src_index(min(ty * kTile, H - 1), scale_y, sh, nearest).i0;
        const int hi = src_index(min(ty * kTile + kTile - 1, H - 1), scale_y, sh, nearest).i1;
        if (hi - lo + 1 > QY) return false;
    }
    for (int tx = 0; tx < tw; ++tx) {
        const int lo = src_index(min(tx * kTile, W - 1), scale_x, sw, nearest).i0;
        const int hi = src_index(min(tx * kTile + kTile - 1, W - 1), scale_x, sw, nearest).i1;
        if (hi - lo + 1 > QX) return false;
    }
    return encode_tiled_fn() != nullptr;
}

size_t lr_scratch_bytes(int sh, int sw, int d) {
    const size_t part = (size_t)sh * sw * round_up(d, 16) * 2;
    return 2 * ((part + 127) & ~(size_t)127);
}

// bf16 hi/lo copy of the low-res map in the layout the TMA boxes expect (2 x 59 MB at 240 x 240 x 512, ~39 us): depends
// only on the map, so the host side may run it on a second stream next to the geometry pipeline of the same view
int launch_lr_pack(const float *S, int sh, int sw, int64_t ssh, int64_t ssw, int64_t ssd, int d, void *scratch,
                   cudaStream_t st) {
    GWBP_REQUIRE(((uintptr_t)scratch & 127) == 0, "scratch must be 128-byte aligned");
    const int dp = round_up(d, 16), ngp = dp / 8;
    const size_t part = (((size_t)sh * sw * dp * 2) + 127) & ~(size_t)127;
    uint4 *hi = (uint4 *)scratch, *lo = (uint4 *)((uint8_t *)scratch + part);
    const int64_t total = (int64_t)sh * ngp * sw;
    flow_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(S, sh, sw, ssh, ssw, ssd, d, ngp, hi, lo);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_backproject_lr(const TileCtx &t, const float *S, int sh, int sw, int64_t ssh, int64_t ssw, int64_t ssd,
                          int nearest, int d, float *num, float *den, void *scratch, bool packed, long long *stats,
                          cudaStream_t st) {
    const int ntiles = t.tw * t.th;
    if (ntiles == 0) return 0;
    GWBP_REQUIRE(((uintptr_t)scratch & 127) == 0, "scratch must be 128-byte aligned");
    GWBP_REQUIRE(((uintptr_t)num & 15) == 0, "num must be 16-byte aligned");
    const int dp = round_up(d, 16), ngp = dp / 8;
    const size_t part = (((size_t)sh * sw * dp * 2) + 127) & ~(size_t)127;
    uint4 *hi = (uint4 *)scratch, *lo = (uint4 *)((uint8_t *)scratch + part);
    if (!packed)
        if (int rc = launch_lr_pack(S, sh, sw, ssh, ssw, ssd, d, scratch, st)) return rc;

    EncodeTiledFn enc = encode_tiled_fn();
    GWBP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap maps[2];
    const cuuint64_t dims[3] = {(cuuint64_t)sw * 8, (cuuint64_t)ngp, (cuuint64_t)sh};
    const cuuint64_t strides[2] = {(cuuint64_t)sw * 16, (cuuint64_t)ngp * sw * 16};  // bytes, dims 1 and 2
    const cuuint32_t box[3] = {64, (cuuint32_t)NG2, 2};
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int k = 0; k < 2; ++k) {
        const CUresult r = enc(&maps[k], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, k ? (void *)lo : (void *)hi, dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        GWBP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for a %d x %d x %d map", (int)r, sh, sw, dp);
    }
    LrArgs a;
    a.t = t;
    a.num = num; a.den = den;
    a.d = d; a.dp = dp; a.nchunks = (dp + NC2 - 1) / NC2; a.nunits = ntiles;
    a.sh = sh; a.sw = sw; a.nearest = nearest;
    a.scale_y = (float)sh / (float)t.H; a.scale_x = (float)sw / (float)t.W;
    a.unit_counter = (int *)t.scratch;
    a.stats = stats;
    a.debug = 0;
    a.trace = (unsigned long long *)tc_trace_buffer();
#ifdef GWBP_EXPERIMENTS
    static const int dbg = getenv("GWBP_LR_DEBUG") ? atoi(getenv("GWBP_LR_DEBUG")) : 0;
    a.debug = dbg;
#endif
    GWBP_CUDA_OK(cudaMemsetAsync(a.unit_counter, 0, sizeof(int), st));
    GWBP_CUDA_OK(cudaFuncSetAttribute(bp_lr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::total));
    const int grid = a.nunits < num_sms() ? a.nunits : num_sms();
    bp_lr_kernel<<<grid, kThreads, Smem::total, st>>>(a, maps[0], maps[1]);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace gwbp
