// Shared declarations for libgwbp.so (sm_100a only).  See include/gwbp.h for the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gwbp.h"

namespace gwbp {

constexpr int kTile = GWBP_TILE;
constexpr int kTilePix = kTile * kTile;
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kAlphaMax = 0.999f;
constexpr float kTMin = 1e-4f;
// per-Gaussian (visible, tile hits) counters travel through ONE 64-bit scan: tile hits in the low 33 bits, the
// visible flag above them, so a view can hold up to 8.5e9 intersections before the sum spills into the visible count
// (any total above the workspace capacity, which is < 2^31, is reported as "capacity exceeded")
constexpr int kVisShift = 33;
constexpr unsigned long long kTileCountMask = (1ull << kVisShift) - 1ull;
constexpr int kBinMaxTiles = 12288;  // sort-free tile binning up to this many tiles (48 KB table per warp); CUB sort above
// SMs of the CURRENT device (cudaDevAttrMultiProcessorCount, cached per device; 148 on a full B200): persistent
// grids and per-CTA scratch partitions are sized with it, so MIG slices / other sm_100a SKUs are not mis-subscribed.
int num_sms();

// thread-local error string behind gwbp_last_error()
void set_error(const char *fmt, ...);

#define GWBP_CUDA_OK(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            ::gwbp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                              __FILE__, __LINE__);                                      \
            return (int)_e;                                                             \
        }                                                                               \
    } while (0)

#define GWBP_REQUIRE(cond, ...)             \
    do {                                    \
        if (!(cond)) {                      \
            ::gwbp::set_error(__VA_ARGS__); \
            return -1;                      \
        }                                   \
    } while (0)

// Camera constants precomputed on the host in the oracle's evaluation order (fp32, no FMA).
struct CamDev {
    float V[12];  // rows 0..2 of the world->camera matrix
    float fx, fy, cx, cy;
    float Wf, Hf;
    float lim_xp, lim_xn, lim_yp, lim_yn;
    float near_plane, far_plane, radius_clip, eps2d;
    int W, H, tw, th;
    int cull;  // exact tile culling (GWBP_PREPARE_TILE_CULL)
    int super; // supertile binning (GWBP_PREPARE_SUPERTILE): also count the (Gaussian, supertile) entries
    int nsx;   // supertiles per row
};
constexpr int kSuperW = GWBP_SUPER_W, kSuperH = GWBP_SUPER_H;  // tiles per supertile: 8 x 4 = one 32-bit tile mask

// Typed view of the caller's workspace.
struct WsDev {
    unsigned long long *cnt, *scan, *mask;
    uint4 *erec;
    float4 *rec, *grec;
    int *radii, *tiles_per_gauss;
    unsigned *dkeys[2], *dvals[2];  // depth sort of the visible Gaussians
    unsigned *cnt2, *base2;         // tile counts / exclusive offsets in depth order
    unsigned *tkeys[2];             // tile id per intersection
    int *tvals[2];                  // packed Gaussian index per intersection (flatten_ids)
    int *offsets;
    long long *stats;
    unsigned *bin_counts, *bin_seg, *bin_tot;  // sort-free tile binning tables (project.cu)
    int *spg;                                  // supertile entries per packed Gaussian
    unsigned long long *svals[2];              // (packed index | tile mask << 32) entries, sort double buffer
    unsigned long long *front;                 // ticket, totals and chained-scan status words of project_pack_kernel
    void *sort_tmp;
    size_t sort_tmp_bytes;
};

inline WsDev ws_view(void *base, const gwbp_ws_layout &L) {
    char *b = (char *)base;
    WsDev w;
    w.cnt = (unsigned long long *)(b + L.cnt);
    w.scan = (unsigned long long *)(b + L.scan);
    w.mask = (unsigned long long *)(b + L.mask);
    w.erec = (uint4 *)(b + L.erec);
    w.rec = (float4 *)(b + L.rec);
    w.grec = (float4 *)(b + L.grec);
    w.radii = (int *)(b + L.radii);
    w.tiles_per_gauss = (int *)(b + L.tiles_per_gauss);
    w.dkeys[0] = (unsigned *)(b + L.dkeys0); w.dkeys[1] = (unsigned *)(b + L.dkeys1);
    w.dvals[0] = (unsigned *)(b + L.dvals0); w.dvals[1] = (unsigned *)(b + L.dvals1);
    w.cnt2 = (unsigned *)(b + L.cnt2);
    w.base2 = (unsigned *)(b + L.base2);
    w.tkeys[0] = (unsigned *)(b + L.tkeys0); w.tkeys[1] = (unsigned *)(b + L.tkeys1);
    w.tvals[0] = (int *)(b + L.tvals0); w.tvals[1] = (int *)(b + L.tvals1);
    w.offsets = (int *)(b + L.offsets);
    w.stats = (long long *)(b + L.stats);
    w.bin_counts = (unsigned *)(b + L.bin_counts);
    w.bin_seg = (unsigned *)(b + L.bin_seg);
    w.bin_tot = (unsigned *)(b + L.bin_tot);
    w.spg = (int *)(b + L.spg);
    w.svals[0] = (unsigned long long *)(b + L.tvals0);  // tvals0 and tvals1 are adjacent: 8 * cap bytes
    w.svals[1] = (unsigned long long *)(b + L.svals);
    w.front = (unsigned long long *)(b + L.front);
    w.sort_tmp = (void *)(b + L.sort_tmp);
    w.sort_tmp_bytes = L.sort_tmp_bytes;
    return w;
}

CamDev make_cam(const gwbp_camera &c);

// ---- stage launchers (one per .cu) -------------------------------------------------------
int launch_pack_scene(int64_t n, const float *means, const float *quats, const float *scales,
                      const float *opac, void *geo, cudaStream_t st);
// projection + tile test + ordered compaction in one kernel; totals land in ws.front[1] (intersections), [2] (visible)
int launch_project_pack(int64_t n, const void *geo, const CamDev &cam, WsDev ws, cudaStream_t st);
int launch_gather_counts(int64_t n_vis, const unsigned *order, WsDev ws, bool gather_erec, cudaStream_t st,
                         bool super_counts = false);
// supertile binning: (supertile id, packed index | tile mask << 32) entries in depth order, stable sort on the id, ranges
int launch_emit_super(int64_t n_vis, const CamDev &cam, const unsigned *order, WsDev ws, int64_t cap, int key_bytes,
                      cudaStream_t st);
int launch_super_sort(int64_t n_entries, int bits, WsDev ws, int key_bytes, int *sorted_buf, cudaStream_t st, int n_lists = 0);
int launch_emit(int64_t n_vis, const CamDev &cam, const unsigned *order, WsDev ws, int64_t cap, bool key16, cudaStream_t st);
size_t binning_tmp_bytes(int64_t n, int64_t cap);
int launch_depth_sort(int64_t n_vis, WsDev ws, int *sorted_buf, cudaStream_t st, const int *counts_src = nullptr);
int launch_scan_counts(int64_t n_vis, WsDev ws, cudaStream_t st);
// key16: tile ids are stored as uint16 in the tkeys buffers (tiles <= 65536): 25 % less sort traffic
int launch_tile_sort(int64_t n_isects, int tile_bits, WsDev ws, bool key16, int *sorted_buf, cudaStream_t st);
int launch_offsets(int64_t n_isects, int n_tiles, const void *tkeys, bool key16, int *offsets, cudaStream_t st,
                   bool key8 = false);
// sort-free stable tile binning (project.cu): counting pass, three scans, ordered scatter -> offsets + flatten (tvals[0])
bool bin_fast_supported(int n_tiles);
size_t bin_table_bytes(int n_tiles, int *chunks_pad, int *tiles_pad, int *nseg, int *chunks, int *wpc);
int launch_bin(int64_t n, const CamDev &cam, const unsigned *order, WsDev ws, int64_t cap, cudaStream_t st);
// kernels launched by this library so far (process-wide; bench.py reports the difference over its timed region)
void count_launches(int k);

struct TileCtx {
    const float4 *grec;
    const int *flatten;
    const int *offsets;          // per-tile ranges over flatten, or (sents != nullptr) per-supertile ranges over sents
    const uint2 *sents;          // supertile lists: (packed index, mask of the supertile's 8 x 4 tiles), depth-ordered
    int nsx;                     // supertiles per row
    long long *scratch;  // 16 device int64 in the workspace (work-queue counters)
    void *dead;          // workspace regions that are dead once the view is prepared (counts, scans, unpacked
    size_t dead_bytes;   // records, hit masks): scratch for the kernels that run on a prepared view
    int W, H, tw, th;
};

int launch_backproject_simt(const TileCtx &t, const float *F, int64_t sH, int64_t sW, int64_t sD, int d,
                            float *num, float *den, long long *stats, cudaStream_t st);
int launch_render_simt(const TileCtx &t, const float *colors, int64_t cstride, int d, const float *bg,
                       float *render, float *alpha, cudaStream_t st);
int launch_render_pixels(const TileCtx &t, const float *colors, int64_t cstride, int d, const float *extra,
                         const int *xy, int k, float *out, float *alpha, cudaStream_t st);
int launch_ratio_accumulate(const float4 *grec, int64_t n_vis, float *num_v, float *den_v, float *acc, float *den_acc,
                            int d, float num_scale, float den_scale, float eps, cudaStream_t st);
int launch_sh_colors(int64_t n, int degree, const float *means, const float *coeffs, int64_t sN, int64_t sK, int64_t sC,
                     const float *cam_pos_host, float *out, cudaStream_t st);
int launch_finalize(const float *num, const float *den, float *out, int64_t n, int d, cudaStream_t st);
// sparse reduce-scatter over peer memory fused with the finalise (finalize.cu): rows [lo, lo + rows) of the field
constexpr int kMaxPeers = GWBP_MAX_PEERS;
bool peer_reduce_supported(int world, int d);
int launch_peer_reduce_finalize(const float *const *num_ptrs, const float *const *den_ptrs, int world, int64_t lo,
                                int64_t rows, int d, float eps, float *out_feat, float *out_num, float *out_den,
                                cudaStream_t st);
int launch_mask(const float *x, int64_t rows, int d, const float *text, int p, int npos, float thr,
                int use_thr, uint8_t *mask, float *score, cudaStream_t st);

// tcgen05 path (backproject_tc.cu)
size_t fpack_bytes(int W, int H, int d);
bool tc_supported(int d);
void tc_set_trace(void *buf, size_t bytes);
void *tc_trace_buffer();  // debug buffer set by gwbp_debug_set_trace (nullptr = off)
int launch_fpack(int W, int H, const float *F, int64_t sH, int64_t sW, int64_t sD, int d, void *fpack, cudaStream_t st);
int launch_fpack_lowres(int W, int H, const float *S, int sh, int sw, int64_t ssh, int64_t ssw, int64_t ssd, int nearest,
                        int d, void *fpack, cudaStream_t st);
int launch_backproject_tc(const TileCtx &t, const float *F, int64_t sH, int64_t sW, int64_t sD, int d,
                          float *num, float *den, void *fpack, bool fpack_ready, long long *stats, cudaStream_t st);

// encoder-resolution maps: down-sampled weights x low-res map, two chained tcgen05 GEMMs (backproject_lr.cu)
bool lr_supported(int W, int H, int sh, int sw, int d, int nearest);
size_t lr_scratch_bytes(int sh, int sw, int d);
int launch_lr_pack(const float *S, int sh, int sw, int64_t ssh, int64_t ssw, int64_t ssd, int d, void *scratch,
                   cudaStream_t st);
int launch_backproject_lr(const TileCtx &t, const float *S, int sh, int sw, int64_t ssh, int64_t ssw, int64_t ssd,
                          int nearest, int d, float *num, float *den, void *scratch, bool packed, long long *stats,
                          cudaStream_t st);

// tcgen05 forward render (render_tc.cu)
bool render_tc_supported(const float *colors, int64_t cstride, int d);
int launch_render_tc(const TileCtx &t, const float *colors, int64_t cstride, int d, const float *bg, float *render,
                     float *alpha, cudaStream_t st);

// ---- device helpers ------------------------------------------------------------------------
// One (pixel, Gaussian) pair, gsplat-1.4.0 rasterize_to_pixels_{fwd,bwd} arithmetic:
//     sigma = 0.5f * (conic.x*dx*dx + conic.z*dy*dy) + conic.y*dx*dy;   alpha = min(0.999f, opac * __expf(-sigma));
// The roundings are spelled out (the contraction nvcc applies to that expression under its default
// -fmad=true, and __expf(x) = ex2.approx(x * log2(e))) so that EVERY kernel of this library -- CUDA-core and
// tcgen05, forward and backward -- evaluates a pair bit-identically and a pair sitting on a threshold
// (alpha = 1/255, T(1-alpha) = 1e-4) falls on the same side everywhere.
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_approx(float x) {  // MUFU.EX2
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// hxx = 0.5f*conic.x, hyy = 0.5f*conic.z: halving is exact, so fma(hxx*dx, dx, (hyy*dy)*dy) == 0.5f * s bit for bit.
// The cross term is folded into the last FMA: sigma = fma(cxy*dx, dy, 0.5f*s) (one of the two contractions a
// compiler may pick for the expression above; ptxas picks it for packed operands whatever the PTX says, see below).
__device__ __forceinline__ float pair_sigma(float dx, float dy, float hxx, float cxy, float hyy) {
    const float s = __fmaf_rn(__fmul_rn(hxx, dx), dx, __fmul_rn(__fmul_rn(hyy, dy), dy));
    return __fmaf_rn(__fmul_rn(cxy, dx), dy, s);
}
__device__ __forceinline__ float pair_alpha(float op, float sigma) {
    return fminf(kAlphaMax, __fmul_rn(op, ex2_approx(__fmul_rn(sigma, -kLog2e))));
}
// Packed fp32 (sm_100 FADD2 / FMUL2 / FFMA2): two pairs per instruction, inline PTX with explicit .rn.
// NB: ptxas 12.9 fuses a mul.rn.f32x2 feeding an add.rn.f32x2 into ONE FFMA2 (even under -fmad=false; the scalar
// .rn forms are left alone), so pair_sigma() is defined with that FMA on the scalar side too and no packed
// mul->add pair is ever written here.
__device__ __forceinline__ float2 add2_rn(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rr;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rr, ra, rb;\n\tmov.b64 {%0, %1}, rr;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 mul2_rn(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rr;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rr, ra, rb;\n\tmov.b64 {%0, %1}, rr;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2_rn(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc, rr;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rr, ra, rb, rc;\n\tmov.b64 {%0, %1}, rr;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
// pair_sigma() for two Gaussians at once: identical roundings, half the instructions
__device__ __forceinline__ float2 pair_sigma2(float2 dx, float2 dy, float2 hxx, float2 cxy, float2 hyy) {
    const float2 s = fma2_rn(mul2_rn(hxx, dx), dx, mul2_rn(mul2_rn(hyy, dy), dy));
    return fma2_rn(mul2_rn(cxy, dx), dy, s);
}
// Returns the weight alpha*T (0 if skipped) and updates T / done.
__device__ __forceinline__ float composite_step(float gx, float gy, float op, float cxx, float cxy,
                                                float cyy, float px, float py, float &T, bool &done) {
    const float dx = gx - px, dy = gy - py;
    const float sigma = pair_sigma(dx, dy, 0.5f * cxx, cxy, 0.5f * cyy);
    const float alpha = pair_alpha(op, sigma);
    float w = 0.0f;
    if (!done && sigma >= 0.0f && alpha >= kAlphaMin) {
        const float nT = __fmul_rn(T, __fsub_rn(1.0f, alpha));
        if (nT <= kTMin) {
            done = true;
        } else {
            w = __fmul_rn(alpha, T);
            T = nT;
        }
    }
    return w;
}

}  // namespace gwbp
