// Forward D-channel render on the 5th-gen tensor cores (segment.py:209-220: `rasterization(colors=features)`;
// gsplat rasterize_to_pixels_fwd semantics):
//      render[p, :] = sum_g w(g,p) X[g, :]        w = alpha*T, front-to-back
// computed per tile as the transposed GEMM  render^T[c, p] = sum_g X^T[c, g] * W[g, p]:
//      A = X^T  [M = 128 features x K = 16 Gaussians]   MN-major (a Gaussian's 8 consecutive features = 16 B)
//      B = W    [K = 16 Gaussians x N = 256 pixels]     K-major  (a pixel's 8 consecutive Gaussians = 16 B)
//      D        [128 features (TMEM lanes) x 256 pixels (TMEM columns)] fp32, double buffered in the 512 columns
// The [128 x 256] accumulator pair fills the TMEM, so a D-channel render takes D/128 passes over a tile's
// Gaussians.  The weights are GENERATED once per tile, in pass 0, which also records them (64 KB per 64-Gaussian
// batch, per-CTA scratch in the dead part of the workspace, L2-resident); the passes of the other feature chunks copy
// them back instead of recomputing them (the chunked CUDA-core kernel -- composite_simt.cu, and gsplat itself --
// regenerates them for every 32 channels).  The epilogue writes 128 contiguous bytes per pixel without a transpose
// (lane = feature).  With two accumulator buffers the epilogue's stores of one (tile, chunk) unit (the whole [H,W,D]
// output must reach DRAM: 2.2 GB at config G) drain while the next unit is being contracted.  A 256-feature unit
// halves the passes but serialises those stores behind the MMAs: measured 2.8 ms at config G
// (profiles/r01_render_tc.txt).
//
// Both operands are split into bf16 hi + lo and three MMAs are issued per K-step (hi*hi + hi*lo + lo*hi),
// as in backproject_tc.cu: ~2^-16 relative error in the contraction.
//
// One persistent CTA per SM, warp-specialised (544 threads):
//   warps 0-7   ALU      : thread = pixel; walks the tile's depth-sorted list 64 Gaussians at a time and writes
//                          W (bf16 hi/lo) straight into the UMMA layout, double buffered (pass 0: generated and
//                          recorded; later passes: replayed from the record).
//   warps 8-11  loaders  : gather the batch's X rows (fp32, 512 B pieces), split to bf16 hi/lo and store them
//                          in UMMA layout into an 8-stage ring (one 16-Gaussian K-step per stage).
//   warps 12-15 epilogue : at the end of a (tile, chunk) unit, tcgen05.ld the accumulators and store
//                          render (+ T*background); alpha = 1 - T is written by the ALU threads.
//   warp 16     MMA      : one elected lane issues tcgen05.mma.
// The ALU publishes one control entry per batch (Gaussian ids) and one per unit end (final transmittances).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gwbp {

using namespace tc;

namespace {

constexpr int GB = 64;        // Gaussians per batch
constexpr int KST = 16;       // Gaussians per UMMA K-step
constexpr int MC = 128;       // feature columns per work unit = UMMA M
constexpr int XSTAGES = 8;    // X^T ring depth (one K-step per stage): two batches
constexpr int ERING = 4;      // control entries in flight
constexpr uint32_t W_KSTR = (kTilePix / 8) * 128;  // W: stride between core matrices along K (8 Gaussians) = 4 KB
constexpr uint32_t W_PART = (GB / 8) * W_KSTR;     // 32 KB per hi / lo part
constexpr uint32_t W_BUF = 2 * W_PART;
constexpr uint32_t X_KSTR = (MC / 8) * 128;        // X^T: stride between core matrices along K = 2 KB
constexpr uint32_t X_PART = (KST / 8) * X_KSTR;    // 4 KB per hi / lo part
constexpr uint32_t X_STAGE = 2 * X_PART;

constexpr int kLoadWarp0 = 8, kEpiWarp0 = 12, kMmaWarp = 16, kThreads = 17 * 32;

struct Entry {
    int unit;      // -1 = exit
    int nb;        // Gaussians in this batch; 0 = end-of-unit marker
    int first;     // first batch of its unit (accumulators are overwritten, not accumulated)
    int nbatches;  // marker only: batches the unit had (0 = empty tile, accumulators untouched)
    int gid[GB];
};

struct Smem {
    static constexpr int w = 0;                                  // 2 buffers x (hi | lo)
    static constexpr int x = 2 * (int)W_BUF;                     // XSTAGES x (hi | lo)
    static constexpr int gbuf = x + XSTAGES * (int)X_STAGE;      // GB x 2 float4
    static constexpr int ent = gbuf + GB * 32;
    static constexpr int tfin = ent + ERING * (int)sizeof(Entry);  // ERING x 256 floats
    static constexpr int tend = tfin + ERING * kTilePix * 4;     // 256 floats: pass 0's final transmittances
    static constexpr int misc = tend + kTilePix * 4;             // work-queue slot
    static constexpr int bars = misc + 16;
    static constexpr int ent_full = 0, ent_empty = ent_full + ERING, w_full = ent_empty + ERING, w_free = w_full + 2,
                         x_full = w_free + 2, x_empty = x_full + XSTAGES, acc_full = x_empty + XSTAGES,
                         acc_empty = acc_full + 2, nbars = acc_empty + 2;
    static constexpr int tmem_slot = bars + nbars * 8;
    static constexpr int total = tmem_slot + 16;
};
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

// Debug cycle accounting: tick(cat) charges the time since the previous tick to category `cat`.
struct Prof {
    long long t[6] = {0, 0, 0, 0, 0, 0};
    long long last = 0;
    bool on = false;
    __device__ __forceinline__ void start(bool enable) { on = enable; if (on) last = clock64(); }
    __device__ __forceinline__ void tick(int cat) {
        if (on) { const long long now = clock64(); t[cat] += now - last; last = now; }
    }
    __device__ __forceinline__ void flush(unsigned long long *dst) {
        if (on)
            for (int i = 0; i < 6; ++i) atomicAdd(dst + i, (unsigned long long)t[i]);
    }
};

__device__ __forceinline__ void st_global_pred(float *p, float v, bool pred) {  // no branch around the store
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.u32 q, %2, 0;\n\t"
        "@q st.global.f32 [%0], %1;\n\t}" ::"l"(p),
        "f"(v), "r"((int)pred)
        : "memory");
}

struct RenderArgs {
    TileCtx t;
    const float *colors;
    int64_t cstride;
    const float *bg;
    float *render, *alpha;
    int d, dp, nchunks, ntiles;
    int *unit_counter;
    uint4 *wsave;  // weight cache: per CTA `cap` batches x (8 Gaussian groups x hi/lo x 256 threads) uint4 ...
    int *gsave;    // ... and the batches' Gaussian ids (GB ints each)
    int cap;       // batches per tile the cache holds (0 = regenerate the weights for every feature chunk)
    unsigned long long *prof;  // debug: per-role cycle accounting, [4 roles][8 categories] (gwbp_debug_set_trace)
    int debug;  // GWBP_RENDER_DEBUG (experiments only): 1 = no X loads, 2 = no render stores, 8 = no weight cache
};

// kCache = false compiles the recording / replay code away (single-chunk renders, or no scratch to cache in)
template <bool kCache>
__global__ void __launch_bounds__(kThreads, 1) render_tc_kernel(const RenderArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto bar = [&](int i) -> uint32_t { return sbase + Smem::bars + 8 * i; };
    Entry *ent = reinterpret_cast<Entry *>(smem + Smem::ent);
    float *tfin = reinterpret_cast<float *>(smem + Smem::tfin);
    volatile int *s_unit = reinterpret_cast<volatile int *>(smem + Smem::misc);  // work-queue broadcast, 2 slots
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + Smem::tmem_slot);

    if (tid == 0) {
        for (int i = 0; i < ERING; ++i) { mbar_init(bar(Smem::ent_full + i), 1); mbar_init(bar(Smem::ent_empty + i), 9); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar(Smem::w_full + i), 8); mbar_init(bar(Smem::w_free + i), 1); }
        for (int i = 0; i < XSTAGES; ++i) { mbar_init(bar(Smem::x_full + i), 4); mbar_init(bar(Smem::x_empty + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar(Smem::acc_full + i), 1); mbar_init(bar(Smem::acc_empty + i), 4); }
        mbar_init_fence();
    }
    if (warp == kMmaWarp) tmem_alloc<512>(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 8) {
        // ======================================= ALU =========================================
        float4 *gbuf = reinterpret_cast<float4 *>(smem + Smem::gbuf);
        int q = 0, e = 0;
        const uint32_t pslab = (uint32_t)(tid >> 3) * 128 + (uint32_t)(tid & 7) * 16;  // this pixel's core-matrix row
        auto wait_entry_slot = [&](int slot) {
            if (e >= ERING) mbar_wait(bar(Smem::ent_empty + slot), ((e / ERING) - 1) & 1);
        };
        Prof pf;  // ALU: 0 other, 1 popc barrier, 2 wait entry slot, 3 wait w_free, 4 generate + store W
        pf.start(a.prof != nullptr && tid == 0);
        // Work queue: TILES.  The NEXT tile id is requested at the start of the current one (tid 0 keeps the atomic's
        // result in a register) so its round trip is hidden behind the tile's work.
        //
        // Weight cache.  The [128 feature x 256 pixel] accumulator pair fills the TMEM, so a D-channel render takes
        // D/128 passes over the tile's Gaussians.  The weights do not depend on the feature chunk: pass 0 generates
        // them (the kernel's critical path: ~26 instructions per pixel and Gaussian) and also streams every thread's
        // 16 core-matrix rows per batch to a per-CTA scratch in the dead part of the workspace (coalesced, L2
        // resident: 64 KB per batch); passes 1.. copy them back into the W buffers instead of recomputing them.  A
        // tile that walks more than `cap` batches falls back to regeneration for all its passes.
        // (the cache addresses and the per-pixel final transmittance of pass 0 are kept out of registers: the weight
        // generation loop below needs all of them for its 16-wide instruction-level parallelism)
        auto wrow_of = [&](int batch) { return a.wsave + ((size_t)blockIdx.x * a.cap + batch) * (16 * 256) + tid; };
        auto gsave_of = [&](int batch) { return a.gsave + ((size_t)blockIdx.x * a.cap + batch) * GB; };
        float *const tend = reinterpret_cast<float *>(smem + Smem::tend);
        if (tid == 0) s_unit[1] = atomicAdd(a.unit_counter, 1);
        bar_sync_alu();
        int tile = s_unit[1];
        for (int useq = 0; tile < a.ntiles; ++useq) {
            int next_tile = 0;
            if (tid == 0) next_tile = atomicAdd(a.unit_counter, 1);
            const int ty = tile / a.t.tw, tx = tile % a.t.tw;
            const int s = a.t.offsets[tile], eend = a.t.offsets[tile + 1];
            const int yy = ty * kTile + (tid >> 4), xx = tx * kTile + (tid & 15);
            const bool inside = yy < a.t.H && xx < a.t.W;
            const float px = (float)xx + 0.5f, py = (float)yy + 0.5f;
            const float2 npx = make_float2(-px, -px), npy = make_float2(-py, -py);
            bool cached = kCache && a.cap > 0;
            int nb0 = 0;       // batches pass 0 walked
            for (int chunk = 0; chunk < a.nchunks; ++chunk) {
                const int unit = tile * a.nchunks + chunk;
                int nbatches = 0;
                float T = 1.0f;
                if (!kCache || chunk == 0 || !cached) {
                    // ---------------- generate (and, in pass 0, record) ----------------
                    bool done = !inside;
                    // Records are fetched two-deep: the list entry (Gaussian index) of batch b+2 and the record of
                    // batch b+1 are requested while batch b is processed, so no load waits on another load in the loop.
                    float4 r0 = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)), r1 = make_float4(0.f, 0.f, 0.f, 0.f);
                    int idn = -1;
                    if (tid < GB) {
                        if (s + tid < eend) {
                            const int id = a.t.flatten[s + tid];
                            r0 = a.t.grec[2 * (int64_t)id];
                            r1 = a.t.grec[2 * (int64_t)id + 1];
                        }
                        if (s + GB + tid < eend) idn = a.t.flatten[s + GB + tid];
                    }
                    for (int b = s; b < eend; b += GB) {
                        pf.tick(0);
                        if (bar_red_popc_alu(!done) == 0) break;  // also: every warp is done reading gbuf of the previous batch
                        pf.tick(1);
                        const int slot = e % ERING;
                        wait_entry_slot(slot);
                        pf.tick(2);
                        if (kCache && chunk == 0 && nbatches >= a.cap) cached = false;  // uniform: the tile outgrew the cache
                        const bool save = kCache && chunk == 0 && cached;
                        if (tid < GB) {
                            // pair-interleaved records (see bp_tc_kernel): (gx0,gx1,gy0,gy1), (hxx0,hxx1,cxy0,cxy1), (hyy0,hyy1,op0,op1)
                            float *gp = reinterpret_cast<float *>(gbuf + 3 * (tid >> 1)) + (tid & 1);
                            gp[0] = r0.x; gp[2] = r0.y;
                            gp[4] = 0.5f * r1.x; gp[6] = r1.y;
                            gp[8] = 0.5f * r1.z; gp[10] = r0.z;
                            ent[slot].gid[tid] = __float_as_int(r0.w);
                            if (save) gsave_of(nbatches)[tid] = __float_as_int(r0.w);
                        }
                        if (tid == 0) {
                            ent[slot].unit = unit;
                            ent[slot].nb = min(GB, eend - b);
                            ent[slot].first = (nbatches == 0);
                            ent[slot].nbatches = 0;
                        }
                        bar_sync_alu();
                        if (tid == 0) mbar_arrive(bar(Smem::ent_full + slot));
                        ++e;
                        r0 = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
                        r1 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (tid < GB) {
                            if (idn >= 0) {
                                r0 = a.t.grec[2 * (int64_t)idn];
                                r1 = a.t.grec[2 * (int64_t)idn + 1];
                            }
                            idn = (b + 2 * GB + tid < eend) ? a.t.flatten[b + 2 * GB + tid] : -1;
                        }
                        const int buf = q & 1;
                        pf.tick(0);
                        if (q >= 2) mbar_wait(bar(Smem::w_free + buf), ((q >> 1) - 1) & 1);
                        pf.tick(3);
                        uint8_t *whi = smem + Smem::w + buf * W_BUF + pslab, *wlo = whi + W_PART;
                        if (__all_sync(0xffffffffu, done)) {
                            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                            for (int kg = 0; kg < GB / 8; ++kg) {
                                *reinterpret_cast<uint4 *>(whi + kg * W_KSTR) = z;
                                *reinterpret_cast<uint4 *>(wlo + kg * W_KSTR) = z;
                                if (save) { uint4 *wrow = wrow_of(nbatches); wrow[(2 * kg) * 256] = z; wrow[(2 * kg + 1) * 256] = z; }
                            }
                        } else {
#pragma unroll 1
                            for (int j = 0; j < GB / 16; ++j) {
                                float w[16];
#pragma unroll
                                for (int i2 = 0; i2 < 8; ++i2) {
                                    // same pair arithmetic as every other kernel (common.cuh), two Gaussians per packed op
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
                                    *reinterpret_cast<uint4 *>(whi + (2 * j + h) * W_KSTR) = hi;
                                    *reinterpret_cast<uint4 *>(wlo + (2 * j + h) * W_KSTR) = lo;
                                    if (save) { uint4 *wrow = wrow_of(nbatches); wrow[(2 * (2 * j + h)) * 256] = hi; wrow[(2 * (2 * j + h) + 1) * 256] = lo; }
                                }
                            }
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar(Smem::w_full + buf));
                        pf.tick(4);
                        ++q;
                        ++nbatches;
                    }
                    if (kCache && chunk == 0) { nb0 = nbatches; tend[tid] = T; }
                } else if (kCache) {
                    // ---------------- replay the recorded weights ----------------
                    T = tend[tid];
                    for (int bi = 0; bi < nb0; ++bi) {
                        pf.tick(0);
                        const int slot = e % ERING;
                        wait_entry_slot(slot);
                        pf.tick(2);
                        if (tid < GB) ent[slot].gid[tid] = gsave_of(bi)[tid];
                        if (tid == 0) {
                            ent[slot].unit = unit;
                            ent[slot].nb = min(GB, eend - (s + bi * GB));
                            ent[slot].first = (bi == 0);
                            ent[slot].nbatches = 0;
                        }
                        bar_sync_alu();
                        if (tid == 0) mbar_arrive(bar(Smem::ent_full + slot));
                        ++e;
                        const uint4 *wrow = wrow_of(bi);
                        uint4 v[8];  // two halves of 8 rows: 32 registers in flight
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = __ldcg(wrow + i * 256);
                        const int buf = q & 1;
                        pf.tick(0);
                        if (q >= 2) mbar_wait(bar(Smem::w_free + buf), ((q >> 1) - 1) & 1);
                        pf.tick(3);
                        uint8_t *whi = smem + Smem::w + buf * W_BUF + pslab, *wlo = whi + W_PART;
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int kg = 4 * half + k;
                                *reinterpret_cast<uint4 *>(whi + kg * W_KSTR) = v[2 * k];
                                *reinterpret_cast<uint4 *>(wlo + kg * W_KSTR) = v[2 * k + 1];
                            }
                            if (half == 0) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = __ldcg(wrow + (8 + i) * 256);
                            }
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar(Smem::w_full + buf));
                        pf.tick(4);
                        ++q;
                    }
                    nbatches = nb0;
                }
                // end-of-unit marker: final transmittances for the epilogue (background term), alpha straight out
                {
                    const int slot = e % ERING;
                    wait_entry_slot(slot);
                    tfin[slot * kTilePix + tid] = T;
                    if (chunk == 0 && inside && a.alpha) a.alpha[(int64_t)yy * a.t.W + xx] = 1.0f - T;
                    if (tid == 0) {
                        ent[slot].unit = unit;
                        ent[slot].nb = 0;
                        ent[slot].first = 0;
                        ent[slot].nbatches = nbatches;
                    }
                    const bool last = chunk == a.nchunks - 1;
                    if (last && tid == 0) s_unit[useq & 1] = next_tile;
                    bar_sync_alu();
                    if (tid == 0) mbar_arrive(bar(Smem::ent_full + slot));
                    ++e;
                    if (last) tile = s_unit[useq & 1];  // this slot is rewritten two tiles later, i.e. after another barrier
                }
            }
        }
        {
            const int slot = e % ERING;
            wait_entry_slot(slot);
            if (tid == 0) {
                ent[slot].unit = -1;
                ent[slot].nb = 0;
                mbar_arrive(bar(Smem::ent_full + slot));
            }
        }
        pf.tick(0);
        pf.flush(a.prof);
    } else if (warp >= kLoadWarp0 && warp < kLoadWarp0 + 4) {
        // ===================================== X loaders =====================================
        // item = (Gaussian row r of the K-step, 8-feature group cg): 32 B of fp32 in, one 16-byte core-matrix
        // row of hi and of lo out.  Lane -> (r % 8, cg % 4): a quarter-warp stores 128 contiguous bytes.
        const int lw = warp - kLoadWarp0;
        const int rl = (lw & 1) * 8 + (lane & 7);         // row within the K-step
        const int cg0 = (lw >> 1) * 8 + (lane >> 3);      // feature groups cg0 + 4*i, i = 0..1
        const uint32_t xoff = (uint32_t)(lw & 1) * X_KSTR + (uint32_t)(lane & 7) * 16;
        int xs = 0;
        Prof pf;  // loader: 0 other, 1 wait entry, 2 issue loads, 3 wait x_empty, 4 (wait data +) split + store
        pf.start(a.prof != nullptr && warp == kLoadWarp0 && lane == 0);
        for (int e = 0;; ++e) {
            const int slot = e % ERING;
            pf.tick(0);
            mbar_wait(bar(Smem::ent_full + slot), (e / ERING) & 1);
            pf.tick(1);
            const int unit = ent[slot].unit, nb = ent[slot].nb;
            int gid[GB / KST];
#pragma unroll
            for (int ks = 0; ks < GB / KST; ++ks)
                gid[ks] = (unit >= 0 && KST * ks + rl < nb) ? ent[slot].gid[KST * ks + rl] : -1;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(Smem::ent_empty + slot));
            if (unit < 0) break;
            if (nb == 0) continue;
            const int chunk = unit % a.nchunks, cbase = chunk * MC;
            const int ksteps = (nb + KST - 1) / KST;
            // all of the batch's loads in flight at once (16 x 16 B per thread), then store K-step by K-step
            float4 v[GB / KST][2][2];
#pragma unroll
            for (int ks = 0; ks < GB / KST; ++ks) {
                const int g = gid[ks];
                const float *row = a.colors + (int64_t)(g < 0 ? 0 : g) * a.cstride + cbase;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int col = 8 * (cg0 + 4 * i);
                    const bool ok = g >= 0 && !(a.debug & 1);
                    v[ks][i][0] = (ok && cbase + col < a.d) ? __ldg(reinterpret_cast<const float4 *>(row + col))
                                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[ks][i][1] = (ok && cbase + col + 4 < a.d) ? __ldg(reinterpret_cast<const float4 *>(row + col + 4))
                                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            pf.tick(2);
#pragma unroll
            for (int ks = 0; ks < GB / KST; ++ks) {
                if (ks >= ksteps) break;  // warp-uniform
                const int stage = xs % XSTAGES;
                if (xs >= XSTAGES) mbar_wait(bar(Smem::x_empty + stage), ((xs / XSTAGES) - 1) & 1);
                pf.tick(3);
                uint8_t *xhi = smem + Smem::x + stage * X_STAGE + xoff;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int cg = cg0 + 4 * i;
                    uint4 hi, lo;
                    split_bf16x2(v[ks][i][0].x, v[ks][i][0].y, hi.x, lo.x);
                    split_bf16x2(v[ks][i][0].z, v[ks][i][0].w, hi.y, lo.y);
                    split_bf16x2(v[ks][i][1].x, v[ks][i][1].y, hi.z, lo.z);
                    split_bf16x2(v[ks][i][1].z, v[ks][i][1].w, hi.w, lo.w);
                    *reinterpret_cast<uint4 *>(xhi + cg * 128) = hi;
                    *reinterpret_cast<uint4 *>(xhi + X_PART + cg * 128) = lo;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(Smem::x_full + stage));
                pf.tick(4);
                ++xs;
            }
        }
        pf.flush(a.prof + 8);
    } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
        // ===================================== epilogue ======================================
        const int quarter = warp & 3;
        const uint32_t lane_base = (uint32_t)(32 * quarter) << 16;
        int nacc = 0;
        Prof pf;  // epilogue: 0 other, 1 wait entry, 2 wait acc_full, 3 tcgen05.ld + stores
        pf.start(a.prof != nullptr && warp == kEpiWarp0 && lane == 0);
        for (int e = 0;; ++e) {
            const int slot = e % ERING;
            pf.tick(0);
            mbar_wait(bar(Smem::ent_full + slot), (e / ERING) & 1);
            pf.tick(1);
            const int unit = ent[slot].unit, nb = ent[slot].nb, nbatches = ent[slot].nbatches;
            if (unit >= 0 && nb == 0) {
                const int tile = unit / a.nchunks, chunk = unit - tile * a.nchunks;
                const int ty = tile / a.t.tw, tx = tile % a.t.tw;
                const int cbase = chunk * MC, ab = nacc & 1;
                const float *tf = tfin + slot * kTilePix;
                if (nbatches > 0) {
                    mbar_wait(bar(Smem::acc_full + ab), (nacc >> 1) & 1);
                    tc_fence_after();
                }
                pf.tick(2);
                // lane = feature c: one store instruction writes 128 contiguous bytes of one pixel.  Everything but
                // `cok` is warp-uniform, and the transmittances are fetched up front, so the 32 stores of a
                // group issue back to back (a per-store branch + smem load costs ~150 cycles each: measured).
                const int c = cbase + 32 * quarter + lane;
                const bool cok = c < a.d && !(a.debug & 2);
                const bool has_bg = a.bg != nullptr;
                const float bgc = (has_bg && c < a.d) ? a.bg[c] : 0.0f;
                const int nx = min(kTile, a.t.W - tx * kTile);
                for (int p0 = 0; p0 < kTilePix; p0 += 32) {
                    float v[32];
                    if (nbatches > 0) {
                        tmem_ld32(tmem + lane_base + (uint32_t)(ab * kTilePix + p0), v);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = 0.0f;
                    }
                    if (has_bg) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 t4 = *reinterpret_cast<const float4 *>(tf + p0 + i);
                            v[i] = fmaf(t4.x, bgc, v[i]);
                            v[i + 1] = fmaf(t4.y, bgc, v[i + 1]);
                            v[i + 2] = fmaf(t4.z, bgc, v[i + 2]);
                            v[i + 3] = fmaf(t4.w, bgc, v[i + 3]);
                        }
                    }
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int yy = ty * kTile + (p0 >> 4) + half;
                        if (yy >= a.t.H) break;
                        float *pp = a.render + ((int64_t)yy * a.t.W + tx * kTile) * a.d + c;
                        if (nx == kTile) {  // interior tile: straight-line predicated stores
#pragma unroll
                            for (int j = 0; j < kTile; ++j, pp += a.d) st_global_pred(pp, v[16 * half + j], cok);
                        } else {
#pragma unroll
                            for (int j = 0; j < kTile; ++j, pp += a.d) st_global_pred(pp, v[16 * half + j], cok && j < nx);
                        }
                    }
                }
                if (nbatches > 0) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(Smem::acc_empty + ab));
                    ++nacc;
                }
                pf.tick(3);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(Smem::ent_empty + slot));
            if (unit < 0) break;
        }
        pf.flush(a.prof + 16);
    } else if (warp == kMmaWarp) {
        // ======================================= MMA =========================================
        // whole warp converged, operands warp-uniform, one elected lane issues (see backproject_tc.cu)
        constexpr uint32_t idesc = umma_idesc_bf16(128, kTilePix, true, false);
        int q = 0, xs = 0, nacc = 0;
        Prof pf;  // MMA: 0 other, 1 wait entry, 2 wait acc_empty, 3 wait w_full, 4 wait x_full, 5 issue
        pf.start(a.prof != nullptr && lane == 0);
        for (int e = 0;; ++e) {
            const int slot = e % ERING;
            pf.tick(0);
            mbar_wait(bar(Smem::ent_full + slot), (e / ERING) & 1);
            pf.tick(1);
            const int unit = ent[slot].unit, nb = ent[slot].nb, first = ent[slot].first, nbatches = ent[slot].nbatches;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(Smem::ent_empty + slot));
            if (unit < 0) break;
            if (nb == 0) {
                if (nbatches > 0) {
                    if (elect_one()) umma_commit(bar(Smem::acc_full + (nacc & 1)));
                    __syncwarp();
                    ++nacc;
                }
                continue;
            }
            const int ab = nacc & 1;
            pf.tick(0);
            if (first && nacc >= 2) mbar_wait(bar(Smem::acc_empty + ab), ((nacc >> 1) - 1) & 1);
            pf.tick(2);
            const int buf = q & 1;
            mbar_wait(bar(Smem::w_full + buf), (q >> 1) & 1);
            tc_fence_after();
            pf.tick(3);
            const int ksteps = (nb + KST - 1) / KST;
            const uint64_t b_hi0 = umma_smem_desc(sbase + Smem::w + buf * W_BUF, W_KSTR, 128);
            const uint64_t b_lo0 = umma_smem_desc(sbase + Smem::w + buf * W_BUF + W_PART, W_KSTR, 128);
            constexpr uint64_t kBStep = (2 * W_KSTR) >> 4;
#pragma unroll 1
            for (int ks = 0; ks < ksteps; ++ks) {
                const int stage = xs % XSTAGES;
                mbar_wait(bar(Smem::x_full + stage), (xs / XSTAGES) & 1);
                tc_fence_after();
                pf.tick(4);
                const uint64_t a_hi0 = umma_smem_desc(sbase + Smem::x + stage * X_STAGE, X_KSTR, 128);
                const uint64_t a_lo0 = umma_smem_desc(sbase + Smem::x + stage * X_STAGE + X_PART, X_KSTR, 128);
                const uint64_t b_hi = b_hi0 + (uint64_t)ks * kBStep, b_lo = b_lo0 + (uint64_t)ks * kBStep;
                const uint32_t acc = (first && ks == 0) ? 0u : 1u;
                if (elect_one()) {
                    const uint32_t d_tmem = tmem + (uint32_t)(ab * kTilePix);
                    umma_bf16(d_tmem, a_hi0, b_hi, idesc, acc);
                    umma_bf16(d_tmem, a_hi0, b_lo, idesc, 1u);
                    umma_bf16(d_tmem, a_lo0, b_hi, idesc, 1u);
                    umma_commit(bar(Smem::x_empty + stage));
                    if (ks == ksteps - 1) umma_commit(bar(Smem::w_free + buf));
                }
                __syncwarp();
                pf.tick(5);
                ++xs;
            }
            ++q;
        }
        pf.flush(a.prof + 24);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc<512>(tmem);
}

}  // namespace

bool render_tc_supported(const float *colors, int64_t cstride, int d) {
    return d >= 32 && d <= 4096 && d % 4 == 0 && cstride % 4 == 0 && ((uintptr_t)colors & 15) == 0;
}

int launch_render_tc(const TileCtx &t, const float *colors, int64_t cstride, int d, const float *bg, float *render,
                     float *alpha, cudaStream_t st) {
    const int ntiles = t.tw * t.th;
    if (ntiles == 0 || d == 0) return 0;
    GWBP_REQUIRE(render_tc_supported(colors, cstride, d),
                 "tcgen05 render needs 32 <= D <= 4096, D %% 4 == 0 and 16-byte aligned rows (D=%d stride=%lld)", d,
                 (long long)cstride);
    RenderArgs a;
    a.t = t;
    a.colors = colors; a.cstride = cstride; a.bg = bg;
    a.render = render; a.alpha = alpha;
    a.d = d;
    a.dp = (d + MC - 1) / MC * MC;
    a.nchunks = (a.dp + MC - 1) / MC;
    a.ntiles = ntiles;
    a.unit_counter = (int *)t.scratch;
    int dbg = 0;
#ifdef GWBP_EXPERIMENTS  // result-altering timing knobs exist only in experiment builds (never in lib/libgwbp.so)
    static const int dbg_env = getenv("GWBP_RENDER_DEBUG") ? atoi(getenv("GWBP_RENDER_DEBUG")) : 0;
    dbg = dbg_env;
#endif
    a.debug = dbg;
    a.prof = (unsigned long long *)tc_trace_buffer();
    GWBP_CUDA_OK(cudaMemsetAsync(a.unit_counter, 0, sizeof(int), st));
    GWBP_CUDA_OK(cudaFuncSetAttribute(render_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::total));
    GWBP_CUDA_OK(cudaFuncSetAttribute(render_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::total));
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    // weight cache in the dead part of the workspace: per CTA and batch 64 KB of weights + GB ids
    constexpr size_t kBatchBytes = 16 * 256 * sizeof(uint4) + GB * sizeof(int);
    size_t cap = (a.nchunks > 1 && t.dead && !(dbg & 8)) ? t.dead_bytes / ((size_t)grid * kBatchBytes) : 0;
    if (cap > 64) cap = 64;
    a.cap = (int)cap;
    a.wsave = (uint4 *)t.dead;
    a.gsave = (int *)((char *)t.dead + (size_t)grid * cap * 16 * 256 * sizeof(uint4));
    if (a.cap > 0)
        render_tc_kernel<true><<<grid, kThreads, Smem::total, st>>>(a);
    else
        render_tc_kernel<false><<<grid, kThreads, Smem::total, st>>>(a);
    count_launches(1);
    GWBP_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace gwbp
