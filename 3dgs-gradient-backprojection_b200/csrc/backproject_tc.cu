// Stage 3, tensor-core path: per-tile compositing fused with the back-projection contraction
//      num[g, :] += sum_p w(g,p) F[p, :]          (== gsplat rasterize_to_pixels_bwd v_colors with
//      den[g]    += sum_p w(g,p)                      v_render = F, backproject.py:127-131,145-151)
// as a [128 Gaussians x 256 pixels] x [256 pixels x D] GEMM per (tile, Gaussian batch) on the
// 5th-gen tensor cores (tcgen05.mma, cta_group::1, kind::f16), fp32 accumulation in TMEM.
//
// Precision: both operands are split into bf16 hi + lo (x ~= hi + lo, 16 mantissa bits) and three
// MMAs are issued per K-step (hi*hi + hi*lo + lo*hi), so the contraction carries ~2^-16 relative
// error -- the 1e-4 parity bar is not reachable with plain bf16 operands (8 bits).
//
// Work unit = one tile.  D is processed in chunks of <= 256 columns: for every batch of 128 Gaussians the
// weights are generated ONCE and the MMA loops over the chunks, alternating between the two
// [128 x 256] fp32 accumulators in the 512 TMEM columns, so the epilogue of chunk c overlaps the MMAs
// of chunk c+1 (or of the next batch's chunk 0).
// Data flow of one persistent CTA (1 per SM, 480 threads, warp-specialised):
//   warps 0-7   ALU      : thread = pixel.  Walk the tile's depth-sorted list 128 Gaussians at a time,
//                          generate w = alpha*T (sequential T per pixel), write W^T as bf16 hi/lo
//                          straight into the UMMA canonical layout (MN-major, SWIZZLE_NONE) in smem.
//   warp 13     producer : streams the pre-packed bf16 hi/lo feature tile through a 5-stage smem ring
//                          with cp.async.bulk (TMA engine), one 16-pixel K-slice x <=256 columns per stage.
//   warp 14     MMA      : one thread issues tcgen05.mma; accumulators [128 x <=256] fp32 are double
//                          buffered in the 512 TMEM columns so the epilogue of one batch overlaps the
//                          MMAs of the next.
//   warps 8-11  epilogue : tcgen05.ld the accumulator (lane = Gaussian row), transpose 16-row x 32-column pieces
//                          through padded smem and reduce them into num[N,D] with row-contiguous
//                          red.global.add.v4.f32 (measured 2.6 TB/s of payload on scattered 2 KB rows vs
//                          0.6 TB/s for per-lane rows -- profiles/r01_probe.txt).
// The W buffer (128 KB) is single: warp w re-fills its 32-pixel slab for batch q+1 as soon as the MMA
// of batch q's LAST column chunk has consumed it (per-warp mbarriers), so generation and MMA overlap.
//
// The feature map is re-laid-out once per view by fpack_planar_kernel / fpack_kernel (fp32 [H,W,D], any strides
// -> bf16 hi/lo, tile-major, already in UMMA core-matrix order) so the producer needs no tensor map.
#include <stdlib.h>

#include "common.cuh"
#include "lister.cuh"
#include "tc_common.cuh"

namespace gwbp {

using namespace tc;

namespace {

constexpr int MB = 128;             // Gaussians per batch = UMMA M
constexpr int NCMAX = 256;          // columns per work unit = UMMA N (max)
constexpr int KSL = 16;             // pixels per K-slice (= one tile row) = one UMMA K step for bf16
constexpr int NSTAGE = 5;           // feature ring depth (16 KB stages): bytes in flight bound the MMA rate
constexpr int STAGE_BYTES = NCMAX * KSL * 2 * 2;  // hi + lo = 16 KB
constexpr int RING = 3;             // batches in flight between ALU and epilogue
constexpr uint32_t A_SBO = 128, A_LBO = (MB / 8) * 128;  // W^T: 16 row-groups of 8 Gaussians per K-group
constexpr int W_PART_BYTES = (kTilePix / 8) * A_LBO;     // 64 KB per hi / lo part
constexpr int EPI_COLS = 32;                             // columns per epilogue piece (128 B per row)
constexpr int EPI_ROWS = 16;                             // rows staged per round and warp (half of the warp's 32)
constexpr int EPI_PITCH = EPI_COLS * 4 + 16;             // padded row pitch (144 B): conflict-free 16-byte stores

constexpr int kEpiWarp0 = 8, kListerWarp0 = 12, kProducerWarp = 13, kMmaWarp = 14, kListerWarp1 = 15, kThreads = 512;

struct RowInfo {
    int gid[MB];
    float den[MB];
    int exit_flag, pad[3];
};

struct Smem {
    // offsets into dynamic shared memory
    static constexpr int w_hi = 0;
    static constexpr int w_lo = W_PART_BYTES;
    static constexpr int fring = 2 * W_PART_BYTES;
    static constexpr int stage_out = fring + NSTAGE * STAGE_BYTES;   // epilogue staging: 4 warps x EPI_ROWS x EPI_PITCH
    static constexpr int gbuf = stage_out + 4 * EPI_ROWS * EPI_PITCH;  // 64 Gaussian pairs x 3 float4
    static constexpr int rows = gbuf + MB * 24;
    static constexpr int ctrl = rows + RING * (int)sizeof(RowInfo);  // int[RING]
    static constexpr int lring = ctrl + 64;                           // lst::Ring: id batches from the lister warp
    static constexpr int bars = lring + lst::NLISTERS * (int)sizeof(lst::Ring);
    // barrier indices
    static constexpr int w_full = 0, w_free = 8, f_full = 16, f_empty = f_full + NSTAGE,
                         acc_full = f_empty + NSTAGE, acc_empty = acc_full + 2, rows_ready = acc_empty + 2,
                         rows_free = rows_ready + RING, ctrl_full = rows_free + RING, ctrl_empty = ctrl_full + RING,
                         l_full = ctrl_empty + RING, l_free = l_full + lst::NL,
                         nbars = l_full + lst::NLISTERS * 2 * lst::NL;  // per lister: full[NL], free[NL]
    static constexpr int tmem_slot = bars + nbars * 8;
    static constexpr int total = tmem_slot + 16;
};
static_assert(sizeof(lst::Ring) % 8 == 0, "mbarriers behind the ring stay 8-byte aligned");
static_assert(Smem::total + 256 <= 232448, "shared memory budget (227 KB) exceeded");

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

// Optional event trace (debug/profiling only: gwbp_debug_set_trace).  CTA 0 records
// (role, event, batch, chunk, clock64) tuples: roles 0/1 = ALU warp 0/7, 2 = epilogue warp 0, 3 = MMA.
constexpr int kTraceRoles = 4, kTraceCap = 4096;
constexpr int kBand = 4;   // tile rows per band of the visiting order (lst::unit_to_tile); measured at config G:
                           // 2 -> 1.063, 4 -> 1.056, 8 -> 1.082, 16 -> 1.092 ms, whole image -> 1.096
using lst::unit_to_tile;

struct TcArgs {
    unsigned long long *trace;  // [kTraceRoles][kTraceCap][2] or nullptr
    TileCtx t;
    const uint8_t *fpack;
    float *num, *den;
    int d, dp, nchunks, nunits;
    int debug;  // GWBP_TC_DEBUG env (experiments only): 1 = skip the accumulator reductions, 2 = every CTA re-reads one
                // packed tile (features always L2-resident: wrong results, timing only)
    int band;   // tile rows per band of the visiting order (kBand; GWBP_TC_BAND overrides for experiments)
    int *unit_counter;
    long long *stats;
};

// -------------------------------------------------------------------------------------------------
// work unit = tile: every CTA pulls tiles from a global counter
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) bp_tc_kernel(const TcArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto bar = [&](int i) -> uint32_t { return sbase + Smem::bars + 8 * i; };
    RowInfo *rows = reinterpret_cast<RowInfo *>(smem + Smem::rows);
    volatile int *ctrl = reinterpret_cast<volatile int *>(smem + Smem::ctrl);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + Smem::tmem_slot);

    if (tid == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(bar(Smem::w_full + i), 1); mbar_init(bar(Smem::w_free + i), 1); }
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar(Smem::f_full + i), 1); mbar_init(bar(Smem::f_empty + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar(Smem::acc_full + i), 1); mbar_init(bar(Smem::acc_empty + i), 4); }
        for (int i = 0; i < RING; ++i) {
            mbar_init(bar(Smem::rows_ready + i), 8);
            mbar_init(bar(Smem::rows_free + i), 4);
            mbar_init(bar(Smem::ctrl_full + i), 1);
            mbar_init(bar(Smem::ctrl_empty + i), 2);
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
    int trace_n = 0;
    auto trace = [&](int role, int ev, int q, int c) {
        if (a.trace != nullptr && blockIdx.x == 0 && lane == 0 && trace_n < kTraceCap) {
            unsigned long long *p = a.trace + ((size_t)role * kTraceCap + trace_n) * 2;
            p[0] = ((unsigned long long)ev << 48) | ((unsigned long long)(c & 0xffff) << 32) | (unsigned)q;
            p[1] = (unsigned long long)clock64();
            ++trace_n;
        }
    };

    if (warp < 8) {
        // ======================================= ALU =========================================
        float4 *gbuf = reinterpret_cast<float4 *>(smem + Smem::gbuf);
        lst::Ring *rings = reinterpret_cast<lst::Ring *>(smem + Smem::lring);
        // lr = the ring (lister) the current tile comes from; qlc / alive_c = batches consumed from it / it still has
        // tiles; qlo / alive_o = the same for the other ring (scalars, swapped at a ring switch: no local-memory arrays)
        int q = 0, lr = 0, qlc = 0, qlo = 0;
        bool alive_c = true, alive_o = true;
        long long walked = 0;
        const uint32_t wslab = (uint32_t)(tid >> 3) * A_LBO + (uint32_t)(tid & 7) * 16;  // this pixel's K-row
        // The next batch to process, as handed over by the lister warp: its tile (work unit), size, whether it closes
        // the tile, and (threads 0..127) the record of its row `tid`, loaded while the previous batch is processed.
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
                const int tile = unit_to_tile(unit, a.t.tw, a.t.th, a.band);
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
                if (bar_red_popc_alu(!done) == 0) {  // also: every warp is done reading gbuf of batch q-1
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
                if (warp == 0 || warp == 7) trace(warp ? 1 : 0, 0, q, 0);
                const int slot = q % RING;
                if (q >= RING) mbar_wait(bar(Smem::rows_free + slot), ((q / RING) - 1) & 1);
                if (tid < MB) {
                    // records of Gaussians 2k, 2k+1 interleaved so that the ALU loop loads packed operand pairs:
                    // float4 (gx0,gx1,gy0,gy1), (hxx0,hxx1,cxy0,cxy1), (hyy0,hyy1,op0,op1), h** = half conic (exact)
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
                if (warp == 0 || warp == 7) trace(warp ? 1 : 0, 1, q, 0);
                walked += n_cur;
                if (__all_sync(0xffffffffu, done)) {
                    // this warp's 32 pixels are finished: its slab of W is all zero
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
                        // 16 Gaussians per step: alpha evaluation is independent across Gaussians (ILP);
                        // only the T update is a serial chain
                        float w[16];
#pragma unroll
                        for (int i2 = 0; i2 < 8; ++i2) {
                            // sigma of two Gaussians at once on the packed fp32 pipe (FADD2/FMUL2/FFMA2): same
                            // roundings as the scalar pair_sigma() of the CUDA-core kernels, half the instructions
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
                        // den: per-Gaussian sum over the warp's 32 pixels.  Butterfly transpose-reduce:
                        // 16 values x 32 lanes -> lane L ends with the full sum of Gaussian (L >> 1).
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
                if (warp == 0 || warp == 7) trace(warp ? 1 : 0, 2, q, 0);
                ++q;
            }
        }
        // exit sentinel for the other roles
        {
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
        // TMEM lane = Gaussian row, but a reduction wants one ROW contiguous per instruction (measured:
        // 2.6 TB/s coalesced vs 0.6 TB/s for per-lane rows, profiles/r01_probe.txt).  So each warp stages
        // 16 of its rows x 32 columns in padded smem and re-reads them row-wise: 8 lanes x 16 B = one 128-byte
        // row piece, 4 rows per `red.global.add.v4.f32` instruction.
        const int quarter = warp & 3;
        const uint32_t lane_base = (uint32_t)(32 * quarter) << 16;
        const int r = 32 * quarter + lane;
        uint8_t *wstage = smem + Smem::stage_out + (EPI_ROWS * quarter) * EPI_PITCH;  // this warp's 16 staging rows
        uint8_t *srow = wstage + (lane & 15) * EPI_PITCH;
        const int rsub = lane >> 3, piece = lane & 7;  // row-in-group (4 rows per instruction) / 16-byte piece of the 128-byte row
        long long live_rows = 0;
        for (int q = 0;; ++q) {
            const int slot = q % RING;
            mbar_wait(bar(Smem::rows_ready + slot), (q / RING) & 1);
            if (rows[slot].exit_flag) break;
            const int gid = rows[slot].gid[r];
            const float dn = rows[slot].den[r];
            const bool live = (gid >= 0) && (dn > 0.0f);
            const unsigned live_mask = __ballot_sync(0xffffffffu, live);
            // the 8 rows this lane will reduce (rows 4*i + rsub) and their accumulator rows
            int64_t grow[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) grow[i] = (int64_t)__shfl_sync(0xffffffffu, gid, 4 * i + rsub) * a.d;
            for (int c = 0; c < a.nchunks; ++c) {
                const int u = q * a.nchunks + c, ab = u & 1;
                const int ncols = min(NCMAX, a.dp - c * NCMAX);   // padded columns of this chunk
                mbar_wait(bar(Smem::acc_full + ab), (u >> 1) & 1);
                tc_fence_after();
                if (quarter == 0) trace(2, 0, q, c);
                for (int c0 = 0; c0 < ncols; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem + lane_base + (uint32_t)(ab * NCMAX + c0), v);
                    const int col = c * NCMAX + c0 + 4 * piece;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {  // lanes 0-15, then lanes 16-31 stage their rows
                        __syncwarp();  // previous round fully read before it is overwritten
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
    } else if (warp == kListerWarp0 || warp == kListerWarp1) {
        // ====================================== listers ======================================
        const int r = warp == kListerWarp1;
        lst::run_lister(a.t, a.unit_counter, a.nunits, a.band, reinterpret_cast<lst::Ring *>(smem + Smem::lring) + r,
                        bar(Smem::l_full + 2 * lst::NL * r));
    } else if (warp == kProducerWarp) {
        // ===================================== producer ======================================
        if (lane == 0) {
            int stage = 0, use = 0;
            const int64_t tile_bytes = (int64_t)kTilePix * a.dp * 4;
            for (int q = 0;; ++q) {
                const int slot = q % RING;
                mbar_wait(bar(Smem::ctrl_full + slot), (q / RING) & 1);
                const int unit = ctrl[slot];
                mbar_arrive(bar(Smem::ctrl_empty + slot));
                if (unit < 0) break;
                const uint8_t *tbase = a.fpack + ((a.debug & 2) ? blockIdx.x : unit_to_tile(unit, a.t.tw, a.t.th, a.band)) * tile_bytes;
                for (int c = 0; c < a.nchunks; ++c) {
                    const int ncols = min(NCMAX, a.dp - c * NCMAX);
                    const uint32_t bytes = (uint32_t)ncols * KSL * 4;  // hi + lo
                    const uint8_t *cbase = tbase + (int64_t)c * NCMAX * kTilePix * 4;
                    for (int ks = 0; ks < kTilePix / KSL; ++ks) {
                        if (use >= 1) mbar_wait(bar(Smem::f_empty + stage), (use - 1) & 1);
                        mbar_arrive_expect_tx(bar(Smem::f_full + stage), bytes);
                        bulk_g2s(sbase + Smem::fring + stage * STAGE_BYTES, cbase + (int64_t)ks * bytes, bytes,
                                 bar(Smem::f_full + stage));
                        if (++stage == NSTAGE) { stage = 0; ++use; }
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ======================================= MMA =========================================
        // The whole warp runs this loop converged and every operand below is warp-uniform, so the
        // tcgen05 instructions are issued straight from uniform registers by one elected lane.
        // (Issuing from a divergent `if (lane == 0)` makes ptxas wrap each UTCHMMA/UTCBAR in an
        // ELECT/BRA.U.ANY serialisation loop: ~630 cycles of issue overhead per 384-cycle K-slice.)
        int stage = 0, use = 0;
        const uint64_t a_hi0 = umma_smem_desc(sbase + Smem::w_hi, A_LBO, A_SBO);
        const uint64_t a_lo0 = umma_smem_desc(sbase + Smem::w_lo, A_LBO, A_SBO);
        constexpr uint64_t kAStep = (2 * A_LBO) >> 4;  // start-address field advance per 16-pixel K-slice
        for (int q = 0;; ++q) {
            const int slot = q % RING;
            mbar_wait(bar(Smem::ctrl_full + slot), (q / RING) & 1);
            const int unit = ctrl[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(Smem::ctrl_empty + slot));
            if (unit < 0) break;
            for (int c = 0; c < a.nchunks; ++c) {
                const int u = q * a.nchunks + c, ab = u & 1;
                const int ncols = min(NCMAX, a.dp - c * NCMAX);
                const uint32_t idesc = umma_idesc_bf16(MB, ncols, true, true);
                const uint32_t b_lbo = (uint32_t)(ncols / 8) * 128, b_part = (uint32_t)ncols * KSL * 2;
                const uint64_t b_hi0 = umma_smem_desc(sbase + Smem::fring, b_lbo, 128);
                const uint64_t b_lo0 = umma_smem_desc(sbase + Smem::fring + b_part, b_lbo, 128);
                if (u >= 2) mbar_wait(bar(Smem::acc_empty + ab), ((u >> 1) - 1) & 1);
                tc_fence_after();
                trace(3, 0, q, c);
                const uint32_t d_tmem = tmem + (uint32_t)(ab * NCMAX);
                const bool last = (c == a.nchunks - 1);
#pragma unroll 1
                for (int ks = 0; ks < kTilePix / KSL; ++ks) {
                    if (c == 0 && (ks & 1) == 0) mbar_wait(bar(Smem::w_full + (ks >> 1)), q & 1);
                    mbar_wait(bar(Smem::f_full + stage), use & 1);
                    tc_fence_after();
                    const uint64_t a_hi = a_hi0 + (uint64_t)ks * kAStep, a_lo = a_lo0 + (uint64_t)ks * kAStep;
                    const uint64_t boff = (uint64_t)((stage * STAGE_BYTES) >> 4);
                    if (elect_one()) {
                        umma_bf16(d_tmem, a_hi, b_hi0 + boff, idesc, ks > 0 ? 1u : 0u);
                        umma_bf16(d_tmem, a_hi, b_lo0 + boff, idesc, 1u);
                        umma_bf16(d_tmem, a_lo, b_hi0 + boff, idesc, 1u);
                        umma_commit(bar(Smem::f_empty + stage));
                        if (last && (ks & 1)) umma_commit(bar(Smem::w_free + (ks >> 1)));
                    }
                    __syncwarp();
                    if (++stage == NSTAGE) { stage = 0; ++use; }
                }
                if (elect_one()) umma_commit(bar(Smem::acc_full + ab));
                __syncwarp();
                trace(3, 1, q, c);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc<512>(tmem);
}

// -------------------------------------------------------------------------------------------------
// feature re-layout: fp32 [H,W,D] (element strides sH,sW,sD) -> bf16 hi/lo, tile-major, UMMA
// core-matrix order.  Per (tile, chunk c, K-slice ks) one contiguous block:
//   [hi: (p/8)*LBO + (n/8)*128 + (p%8)*16 + (n%8)*2][lo: same]   LBO = (ncols/8)*128, p in 0..15
// One CTA per (tile, chunk); HBM-bound: reads 4 B and writes 4 B per feature element.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fpack_kernel(const float *__restrict__ F, int64_t sH, int64_t sW, int64_t sD,
                                                    int W, int H, int tw, int d, int dp, int nchunks,
                                                    uint8_t *__restrict__ out) {
    __shared__ __align__(16) float slab[KSL][NCMAX + 4];
    const int tile = blockIdx.x / nchunks, c = blockIdx.x % nchunks;
    const int ty = tile / tw, tx = tile % tw;
    const int ncols = min(NCMAX, dp - c * NCMAX);
    const int t = threadIdx.x;
    uint8_t *cbase = out + (int64_t)tile * kTilePix * dp * 4 + (int64_t)c * NCMAX * kTilePix * 4;
    const uint32_t lbo = (uint32_t)(ncols / 8) * 128, part = (uint32_t)ncols * KSL * 2;
    const int col = c * NCMAX + t;
    for (int ks = 0; ks < kTilePix / KSL; ++ks) {
        const int y = ty * kTile + ks;
        if (t < ncols) {
#pragma unroll
            for (int p = 0; p < KSL; ++p) {
                const int x = tx * kTile + p;
                const bool ok = (y < H) && (x < W) && (col < d);
                slab[p][t] = ok ? __ldg(F + y * sH + x * sW + col * sD) : 0.0f;
            }
        }
        __syncthreads();
        // item = (pixel p, 8-column group ng): one 16-byte row of a core matrix, hi and lo
        for (int item = t; item < KSL * (ncols / 8); item += 256) {
            const int p = item % KSL, ng = item / KSL;
            const float4 f0 = *reinterpret_cast<const float4 *>(&slab[p][8 * ng]);
            const float4 f1 = *reinterpret_cast<const float4 *>(&slab[p][8 * ng + 4]);
            uint4 hi, lo;
            split_bf16x2(f0.x, f0.y, hi.x, lo.x);
            split_bf16x2(f0.z, f0.w, hi.y, lo.y);
            split_bf16x2(f1.x, f1.y, hi.z, lo.z);
            split_bf16x2(f1.z, f1.w, hi.w, lo.w);
            const uint32_t off = (uint32_t)(p / 8) * lbo + (uint32_t)ng * 128 + (uint32_t)(p % 8) * 16;
            uint8_t *blk = cbase + (int64_t)ks * part * 2;
            *reinterpret_cast<uint4 *>(blk + off) = hi;
            *reinterpret_cast<uint4 *>(blk + part + off) = lo;
        }
        __syncthreads();
    }
}


// Planar input ([D,H,W] buffer exposed as a permuted [H,W,D] view: sW == 1, the reference's layout,
// backproject.py:110-113).  lane = pixel: a CTA takes 256 consecutive pixels of one image row x 64 channels.  A lane
// loads its pixel's 64 channel values (each warp load = 128 contiguous bytes of one channel row, the CTA's 8 warps
// 1 KB: DRAM page locality, the planes are megabytes apart), splits them and writes 16-byte core-matrix rows; 8
// consecutive lanes write 128 contiguous bytes, the eight 8-channel groups of a thread 1 KB.  No shared memory: the
// transposition is free because a lane gathers its own pixel's channels.  64 independent loads in flight per thread;
// measured at config G (tools/pack_bench.py): 0.739 ms = 6.04 TB/s read + write, 94 % of the measured copy peak
// (the previous version staged [channel][pixel] slabs in shared memory: 0.863 ms; 128 px x 32 ch: 0.805;
// 256 x 32: 0.777; 512 x 32: 0.766; 128 x 16: 0.794).
template <int kPix, int kCols>
__global__ void __launch_bounds__(kPix) fpack_planar_kernel(const float *__restrict__ F, int64_t sH, int64_t sD,
                                                               int W, int H, int tw, int d, int dp,
                                                               uint8_t *__restrict__ out) {
    const int x = blockIdx.x * kPix + threadIdx.x, y = blockIdx.y, col0 = blockIdx.z * kCols;
    const int tx = x >> 4, p = x & 15;
    if (tx >= tw) return;
    const int c = col0 / NCMAX, ncols = min(NCMAX, dp - c * NCMAX), n0 = col0 - c * NCMAX;
    const int ngroups = min(kCols, ncols - n0) / 8;  // 8-channel groups of this block that exist in the chunk
    const bool ok = x < W;
    const float *src = F + (int64_t)y * sH + x + (int64_t)col0 * sD;
    float v[kCols];
#pragma unroll
    for (int i = 0; i < kCols; ++i) v[i] = (ok && col0 + i < d) ? __ldcs(src + (int64_t)i * sD) : 0.0f;
    const uint32_t lbo = (uint32_t)(ncols / 8) * 128, part = (uint32_t)ncols * KSL * 2;
    uint8_t *blk = out + (int64_t)((y >> 4) * tw + tx) * kTilePix * dp * 4 + (int64_t)c * NCMAX * kTilePix * 4 +
                   (int64_t)(y & 15) * part * 2 + (uint32_t)(p >> 3) * lbo + (uint32_t)(n0 / 8) * 128 + (uint32_t)(p & 7) * 16;
#pragma unroll
    for (int g = 0; g < kCols / 8; ++g) {
        if (g < ngroups) {
            uint4 hi, lo;
            split_bf16x2(v[8 * g + 0], v[8 * g + 1], hi.x, lo.x);
            split_bf16x2(v[8 * g + 2], v[8 * g + 3], hi.y, lo.y);
            split_bf16x2(v[8 * g + 4], v[8 * g + 5], hi.z, lo.z);
            split_bf16x2(v[8 * g + 6], v[8 * g + 7], hi.w, lo.w);
            *reinterpret_cast<uint4 *>(blk + g * 128) = hi;
            *reinterpret_cast<uint4 *>(blk + part + g * 128) = lo;
        }
    }
}

// Fused upsample + re-layout.  The reference materialises
//   F = interpolate(encoder_out[1,D,h,w], size=(H,W), mode="bilinear")      (backproject.py:110-112, 2.2 GB/view)
// (mode="nearest" for the DINOv2 tokens, backproject.py:245-249) only to contract it once.  Here the
// encoder-resolution map (118 MB at 240x240x512, L2-resident) is sampled on the fly with PyTorch's
// align_corners=False arithmetic and written straight into the packed bf16 hi/lo layout.
__global__ void __launch_bounds__(256) fpack_lowres_kernel(const float *__restrict__ S, int sh, int sw, int64_t ssh,
                                                           int64_t ssw, int64_t ssd, int nearest, int W, int H, int tw,
                                                           int d, int dp, int nchunks, uint8_t *__restrict__ out) {
    __shared__ float slab[NCMAX][33];
    const int span = blockIdx.x, y = blockIdx.y, c = blockIdx.z;
    const int ncols = min(NCMAX, dp - c * NCMAX);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int x = span * 32 + lane;
    const bool xok = x < W;
    // source coordinates (torch area_pixel_compute_source_index, align_corners=False / nearest)
    const float scale_y = (float)sh / (float)H, scale_x = (float)sw / (float)W;
    int y0, y1, x0, x1;
    float ly, lx;
    if (nearest) {
        y0 = y1 = min((int)floorf((float)y * scale_y), sh - 1);
        x0 = x1 = min((int)floorf((float)x * scale_x), sw - 1);
        ly = lx = 0.0f;
    } else {
        const float fy = fmaxf(scale_y * ((float)y + 0.5f) - 0.5f, 0.0f);
        const float fx = fmaxf(scale_x * ((float)x + 0.5f) - 0.5f, 0.0f);
        y0 = min((int)fy, sh - 1); x0 = min((int)fx, sw - 1);
        y1 = y0 + (y0 < sh - 1 ? 1 : 0); x1 = x0 + (x0 < sw - 1 ? 1 : 0);
        ly = fy - (float)y0; lx = fx - (float)x0;
    }
    const int64_t o00 = y0 * ssh + x0 * ssw, o01 = y0 * ssh + x1 * ssw, o10 = y1 * ssh + x0 * ssw, o11 = y1 * ssh + x1 * ssw;
    // 4 channels x 4 taps = 16 independent (L1/L2-resident) loads in flight per lane
    for (int n0 = warp; n0 < ncols; n0 += 32) {
        float t00[4], t01[4], t10[4], t11[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int n = n0 + 8 * u, col = c * NCMAX + n;
            const bool okc = xok && n < ncols && col < d;
            const float *p = S + (okc ? col : 0) * ssd;
            t00[u] = okc ? __ldg(p + o00) : 0.0f;
            t01[u] = okc ? __ldg(p + o01) : 0.0f;
            t10[u] = okc ? __ldg(p + o10) : 0.0f;
            t11[u] = okc ? __ldg(p + o11) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float top = (1.0f - lx) * t00[u] + lx * t01[u];
            const float bot = (1.0f - lx) * t10[u] + lx * t11[u];
            if (n0 + 8 * u < ncols) slab[n0 + 8 * u][lane] = (1.0f - ly) * top + ly * bot;
        }
    }
    __syncthreads();
    const int ty = y / kTile, ks = y % kTile;
    const uint32_t lbo = (uint32_t)(ncols / 8) * 128, part = (uint32_t)ncols * KSL * 2;
    for (int item = t; item < 32 * (ncols / 8); item += 256) {
        const int px = item & 31, ng = item >> 5;
        const int tx = span * 2 + (px >> 4), p = px & 15;
        if (tx >= tw) continue;
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = slab[8 * ng + i][px];
        uint4 hi, lo;
        split_bf16x2(f[0], f[1], hi.x, lo.x);
        split_bf16x2(f[2], f[3], hi.y, lo.y);
        split_bf16x2(f[4], f[5], hi.z, lo.z);
        split_bf16x2(f[6], f[7], hi.w, lo.w);
        const int tile = ty * tw + tx;
        uint8_t *blk = out + (int64_t)tile * kTilePix * dp * 4 + (int64_t)c * NCMAX * kTilePix * 4 + (int64_t)ks * part * 2;
        const uint32_t off = (uint32_t)(p / 8) * lbo + (uint32_t)ng * 128 + (uint32_t)(p % 8) * 16;
        *reinterpret_cast<uint4 *>(blk + off) = hi;
        *reinterpret_cast<uint4 *>(blk + part + off) = lo;
    }
}

// Faster variant for UPsampling (the reference's case: 240x240 -> 1297x840): one CTA = one tile row (16 image rows)
// x 32 pixels x 64 channels.  The low-resolution window the CTA needs (a few rows x columns per channel) is staged
// in shared memory once; a thread owns one pixel column and 8 channels, keeps the two horizontally interpolated
// source rows it is between in registers and walks down the 16 output rows, so a source texel is fetched ~5x per
// CTA instead of once per output element (the first version issued 4 global loads per element: 1.36 ms at config G).
constexpr int kLowCh = 64, kLowSpan = 32, kLowMaxWin = 160;  // window texels per channel (+1 pad): 64*161*4 = 41 KB
__global__ void __launch_bounds__(256) fpack_lowres_tile_kernel(const float *__restrict__ S, int sh, int sw, int64_t ssh,
                                                                int64_t ssw, int64_t ssd, int nearest, int W, int H,
                                                                int tw, int d, int dp, uint8_t *__restrict__ out) {
    extern __shared__ float win[];  // [kLowCh][chs], chs = qy*qx | 1 (odd: conflict-free across channels)
    const int span = blockIdx.x, ty = blockIdx.y, cb = blockIdx.z * kLowCh;
    const int t = threadIdx.x, px = t & 31, g = t >> 5;
    const float scale_y = (float)sh / (float)H, scale_x = (float)sw / (float)W;
    auto src = [&](int o, float scale, int n_src, int &i0, int &i1, float &l) {
        if (nearest) {
            i0 = i1 = min((int)floorf((float)o * scale), n_src - 1);
            l = 0.0f;
        } else {  // torch area_pixel_compute_source_index, align_corners=False
            const float f = fmaxf(scale * ((float)o + 0.5f) - 0.5f, 0.0f);
            i0 = min((int)f, n_src - 1);
            i1 = i0 + (i0 < n_src - 1 ? 1 : 0);
            l = f - (float)i0;
        }
    };
    // window of source rows / columns this CTA touches (uniform)
    int ylo, yhi, xlo, xhi, tmp;
    float ftmp;
    src(ty * kTile, scale_y, sh, ylo, tmp, ftmp);
    src(min(ty * kTile + kTile - 1, H - 1), scale_y, sh, tmp, yhi, ftmp);
    src(span * kLowSpan, scale_x, sw, xlo, tmp, ftmp);
    src(min(span * kLowSpan + kLowSpan - 1, W - 1), scale_x, sw, tmp, xhi, ftmp);
    const int qy = yhi - ylo + 1, qx = xhi - xlo + 1, chs = (qy * qx) | 1;
    for (int idx = t; idx < kLowCh * qy * qx; idx += 256) {
        const int ch = idx / (qy * qx), r = idx - ch * (qy * qx);
        const int yy = r / qx, xx = r - yy * qx;
        const int col = cb + ch;
        win[ch * chs + r] = col < d ? __ldg(S + col * ssd + (ylo + yy) * ssh + (xlo + xx) * ssw) : 0.0f;
    }
    __syncthreads();
    const int x = span * kLowSpan + px;
    const int tx = span * 2 + (px >> 4), p = px & 15;
    if (tx >= tw || cb + 8 * g >= dp) return;
    int x0, x1;
    float lx;
    src(min(x, W - 1), scale_x, sw, x0, x1, lx);
    x0 -= xlo; x1 -= xlo;
    const float *wch = win + (8 * g) * chs;
    const int c = cb / NCMAX, ncols = min(NCMAX, dp - c * NCMAX), ng = (cb - c * NCMAX) / 8 + g;
    const uint32_t lbo = (uint32_t)(ncols / 8) * 128, part = (uint32_t)ncols * KSL * 2;
    uint8_t *tbase = out + (int64_t)(ty * tw + tx) * kTilePix * dp * 4 + (int64_t)c * NCMAX * kTilePix * 4 +
                     (uint32_t)(p / 8) * lbo + (uint32_t)ng * 128 + (uint32_t)(p % 8) * 16;
    float ha[8], hb[8];
    int cur = -1;
    for (int ks = 0; ks < kTile; ++ks) {
        const int y = ty * kTile + ks;
        uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
        if (y < H && x < W) {
            int y0, y1;
            float ly;
            src(y, scale_y, sh, y0, y1, ly);
            if (y0 != cur) {  // uniform: a new pair of source rows, interpolated horizontally once
                cur = y0;
                const int ra = (y0 - ylo) * qx, rb = (y1 - ylo) * qx;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float *w = wch + j * chs;
                    ha[j] = (1.0f - lx) * w[ra + x0] + lx * w[ra + x1];
                    hb[j] = (1.0f - lx) * w[rb + x0] + lx * w[rb + x1];
                }
            }
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (1.0f - ly) * ha[j] + ly * hb[j];
            split_bf16x2(v[0], v[1], hi.x, lo.x);
            split_bf16x2(v[2], v[3], hi.y, lo.y);
            split_bf16x2(v[4], v[5], hi.z, lo.z);
            split_bf16x2(v[6], v[7], hi.w, lo.w);
        }
        uint8_t *blk = tbase + (int64_t)ks * part * 2;
        *reinterpret_cast<uint4 *>(blk) = hi;
        *reinterpret_cast<uint4 *>(blk + part) = lo;
    }
}

}  // namespace

static unsigned long long *g_trace = nullptr;
void tc_set_trace(void *buf, size_t bytes) {
    g_trace = (buf && bytes >= (size_t)kTraceRoles * kTraceCap * 16) ? (unsigned long long *)buf : nullptr;
}
void *tc_trace_buffer() { return g_trace; }

static int round_up(int x, int m) { return (x + m - 1) / m * m; }

// 16-byte row pieces for the bulk reduction need D % 4 == 0; tiny D is not worth a GEMM
bool tc_supported(int d) { return d >= 16 && d <= 2048 && d % 4 == 0; }

size_t fpack_bytes(int W, int H, int d) {
    const size_t tiles = (size_t)((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile);
    return tiles * kTilePix * (size_t)round_up(d, 16) * 4;
}

// feature re-layout only (may run on a side stream, concurrently with projection / binning of the same view)
int launch_fpack(int W, int H, const float *F, int64_t sH, int64_t sW, int64_t sD, int d, void *fpack, cudaStream_t st) {
    const int tw = (W + kTile - 1) / kTile, th = (H + kTile - 1) / kTile, ntiles = tw * th;
    if (ntiles == 0) return 0;
    GWBP_REQUIRE(((uintptr_t)fpack & 127) == 0, "fpack must be 128-byte aligned");
    const int dp = round_up(d, 16), nchunks = (dp + NCMAX - 1) / NCMAX;
    if (sW == 1 && sD != 1 && H <= 65535) {  // (grid.y = H)
        // rows of the last tile row beyond H are never written by the planar kernel: they are only
        // ever multiplied by zero weights, but must not hold NaN/Inf bit patterns -> clear once per view
        if (H % kTile)
            GWBP_CUDA_OK(cudaMemsetAsync((uint8_t *)fpack + (size_t)(th - 1) * tw * kTilePix * dp * 4, 0,
                                         (size_t)tw * kTilePix * dp * 4, st));
        constexpr int kPix = 256, kCols = 64;
        dim3 grid((tw * kTile + kPix - 1) / kPix, H, (dp + kCols - 1) / kCols);
        fpack_planar_kernel<kPix, kCols><<<grid, kPix, 0, st>>>(F, sH, sD, W, H, tw, d, dp, (uint8_t *)fpack);
        count_launches(1);
    } else {
        fpack_kernel<<<ntiles * nchunks, 256, 0, st>>>(F, sH, sW, sD, W, H, tw, d, dp, nchunks, (uint8_t *)fpack);
        count_launches(1);
    }
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_fpack_lowres(int W, int H, const float *S, int sh, int sw, int64_t ssh, int64_t ssw, int64_t ssd, int nearest,
                        int d, void *fpack, cudaStream_t st) {
    const int tw = (W + kTile - 1) / kTile, th = (H + kTile - 1) / kTile;
    if (tw * th == 0) return 0;
    GWBP_REQUIRE(((uintptr_t)fpack & 127) == 0, "fpack must be 128-byte aligned");
    GWBP_REQUIRE(sh >= 1 && sw >= 1, "low-resolution map must be at least 1x1");
    const int dp = round_up(d, 16), nchunks = (dp + NCMAX - 1) / NCMAX;
    // window bound of the tile variant: rows/columns of the source one CTA touches
    const int qy = (int)((double)kTile * sh / H) + 3, qx = (int)((double)kLowSpan * sw / W) + 3;
    if (qy * qx <= kLowMaxWin) {
        const size_t smem = (size_t)kLowCh * ((qy * qx) | 1) * sizeof(float);
        GWBP_CUDA_OK(cudaFuncSetAttribute(fpack_lowres_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((W + kLowSpan - 1) / kLowSpan, th, (dp + kLowCh - 1) / kLowCh);
        fpack_lowres_tile_kernel<<<grid, 256, smem, st>>>(S, sh, sw, ssh, ssw, ssd, nearest, W, H, tw, d, dp,
                                                          (uint8_t *)fpack);
        count_launches(1);
        GWBP_CUDA_OK(cudaGetLastError());
        return 0;
    }
    // downsampling / huge windows: the per-element variant
    if (H % kTile)
        GWBP_CUDA_OK(cudaMemsetAsync((uint8_t *)fpack + (size_t)(th - 1) * tw * kTilePix * dp * 4, 0,
                                     (size_t)tw * kTilePix * dp * 4, st));
    dim3 grid((W + 31) / 32, H, nchunks);
    fpack_lowres_kernel<<<grid, 256, 0, st>>>(S, sh, sw, ssh, ssw, ssd, nearest, W, H, tw, d, dp, nchunks, (uint8_t *)fpack);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_backproject_tc(const TileCtx &t, const float *F, int64_t sH, int64_t sW, int64_t sD, int d, float *num,
                          float *den, void *fpack, bool fpack_ready, long long *stats, cudaStream_t st) {
    const int ntiles = t.tw * t.th;
    if (ntiles == 0) return 0;
    GWBP_REQUIRE(((uintptr_t)fpack & 127) == 0, "fpack must be 128-byte aligned");
    GWBP_REQUIRE(((uintptr_t)num & 15) == 0, "num must be 16-byte aligned");
    const int dp = round_up(d, 16), nchunks = (dp + NCMAX - 1) / NCMAX;
    if (!fpack_ready)
        if (int rc = launch_fpack(t.W, t.H, F, sH, sW, sD, d, fpack, st)) return rc;

    TcArgs a;
    a.trace = g_trace;
    a.t = t;
    a.fpack = (const uint8_t *)fpack;
    a.num = num; a.den = den;
    a.d = d; a.dp = dp; a.nchunks = nchunks; a.nunits = ntiles;
    a.unit_counter = (int *)t.scratch;
    a.stats = stats;
    a.debug = 0;
    a.band = kBand;
#ifdef GWBP_EXPERIMENTS  // result-altering timing knobs exist only in experiment builds (never in lib/libgwbp.so)
    static const int dbg = getenv("GWBP_TC_DEBUG") ? atoi(getenv("GWBP_TC_DEBUG")) : 0;
    a.debug = dbg;
    static const int band_env = getenv("GWBP_TC_BAND") ? atoi(getenv("GWBP_TC_BAND")) : kBand;
    a.band = band_env > 0 ? band_env : kBand;
#endif
    GWBP_CUDA_OK(cudaMemsetAsync(a.unit_counter, 0, sizeof(int), st));
    // per-device function attribute: set on every launch (microseconds) so a process that drives several
    // devices never launches with the default 48 KB limit
    GWBP_CUDA_OK(cudaFuncSetAttribute(bp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::total));
    const int grid = a.nunits < num_sms() ? a.nunits : num_sms();
    bp_tc_kernel<<<grid, kThreads, Smem::total, st>>>(a);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace gwbp
