// extern "C" surface of libgwbp.so (see include/gwbp.h for the contract and the reference
// call sites each entry point replaces).
#include <cuda.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace gwbp {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cache[dev] = v;
    }
    return cache[dev];
}

static unsigned long long g_launches = 0;  // kernels launched by this library (all threads; relaxed counter)
void count_launches(int k) { __atomic_fetch_add(&g_launches, (unsigned long long)k, __ATOMIC_RELAXED); }

// ---- optional per-stage timing (gwbp_profile_*): CUDA events recorded on the caller's stream at stage boundaries ----
enum { kEvProject0 = 0, kEvProject1, kEvCounts, kEvCompact, kEvDepthSort, kEvBin, kEvPack0, kEvPack1, kEvBp0, kEvBp1, kNumEv };
static bool g_prof_on = false;
static cudaEvent_t g_ev[kNumEv];
static bool g_ev_ready = false, g_ev_set[kNumEv] = {false};
static void prof_mark(int i, cudaStream_t st) {
    if (!g_prof_on) return;
    if (!g_ev_ready) {
        for (int k = 0; k < kNumEv; ++k) cudaEventCreate(&g_ev[k]);
        g_ev_ready = true;
    }
    cudaEventRecord(g_ev[i], st);
    g_ev_set[i] = true;
}

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

static int tile_bits_for(int n_tiles) {
    int b = 1;
    while ((1 << b) <= n_tiles) ++b;  // floor(log2(n_tiles)) + 1, as gsplat
    return b;
}

static int check_cam(const gwbp_camera *c) {
    GWBP_REQUIRE(c != nullptr, "camera is NULL");
    GWBP_REQUIRE(c->width > 0 && c->height > 0, "width/height must be positive (got %d x %d)", c->width, c->height);
    GWBP_REQUIRE(c->width <= 65536 && c->height <= 65536, "image too large (%d x %d)", c->width, c->height);
    GWBP_REQUIRE(c->K[0] > 0.0f && c->K[4] > 0.0f, "focal lengths must be positive");
    return 0;
}

static TileCtx tile_ctx(const gwbp_camera *cam, const void *ws, const gwbp_ws_layout &L, const gwbp_view_info *info) {
    WsDev w = ws_view(const_cast<void *>(ws), L);
    TileCtx t;
    t.grec = w.grec;
    t.flatten = w.tvals[info->sorted_buf];
    t.offsets = w.offsets;
    t.sents = info->list_kind == 1 ? (const uint2 *)w.svals[info->sorted_buf] : nullptr;
    t.nsx = info->super_w;
    t.scratch = w.stats;
    t.dead = (char *)const_cast<void *>(ws) + L.cnt;
    t.dead_bytes = L.grec - L.cnt;
    t.W = cam->width; t.H = cam->height;
    t.tw = info->tile_w; t.th = info->tile_h;
    return t;
}

}  // namespace gwbp

using namespace gwbp;

extern "C" {

int gwbp_abi_version(void) { return GWBP_ABI_VERSION; }

const char *gwbp_last_error(void) { return g_err; }

int gwbp_workspace_layout(int64_t n, int32_t width, int32_t height, int64_t cap, gwbp_ws_layout *L) {
    GWBP_REQUIRE(L != nullptr, "layout pointer is NULL");
    GWBP_REQUIRE(n >= 0 && n < (1ll << 31) - 1, "n out of range (%lld)", (long long)n);
    GWBP_REQUIRE(width > 0 && height > 0, "width/height must be positive");
    GWBP_REQUIRE(cap >= 0 && cap < (1ll << 31), "cap_isects out of range (%lld)", (long long)cap);
    const int64_t tiles = (int64_t)((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
    const int64_t n1 = n + 1, c1 = cap > 0 ? cap : 1;
    size_t o = 0;
    memset(L, 0, sizeof(*L));
    L->cnt = o; o = align_up(o + sizeof(unsigned long long) * n1);
    L->scan = o; o = align_up(o + sizeof(unsigned long long) * n1);
    L->rec = o; o = align_up(o + sizeof(float4) * 2 * n1);
    L->mask = o; o = align_up(o + sizeof(unsigned long long) * n1);
    L->grec = o; o = align_up(o + sizeof(float4) * 2 * n1);
    L->erec = o; o = align_up(o + sizeof(uint4) * n1);
    L->radii = o; o = align_up(o + sizeof(int) * n1);
    L->tiles_per_gauss = o; o = align_up(o + sizeof(int) * n1);
    L->dkeys0 = o; o = align_up(o + sizeof(unsigned) * n1);
    L->dkeys1 = o; o = align_up(o + sizeof(unsigned) * n1);
    L->dvals0 = o; o = align_up(o + sizeof(unsigned) * n1);
    L->dvals1 = o; o = align_up(o + sizeof(unsigned) * n1);
    L->cnt2 = o; o = align_up(o + sizeof(unsigned) * n1);
    L->base2 = o; o = align_up(o + sizeof(unsigned) * n1);
    L->tkeys0 = o; o = align_up(o + sizeof(unsigned) * c1);
    L->tkeys1 = o; o = align_up(o + sizeof(unsigned) * c1);
    L->tvals0 = o; o = align_up(o + sizeof(int) * c1);
    L->tvals1 = o; o = align_up(o + sizeof(int) * c1);
    L->offsets = o; o = align_up(o + sizeof(int) * (tiles + 1));
    L->stats = o; o = align_up(o + sizeof(long long) * 16);
    {   // sort-free tile binning tables (only used when the image has <= kBinMaxTiles tiles)
        int chunks_pad = 0, tiles_pad = 0, nseg = 0;
        const bool fast = bin_fast_supported((int)tiles);
        if (fast) bin_table_bytes((int)tiles, &chunks_pad, &tiles_pad, &nseg, nullptr, nullptr);
        L->bin_counts = o; o = align_up(o + sizeof(unsigned) * (size_t)chunks_pad * tiles_pad + 16);
        L->bin_seg = o; o = align_up(o + sizeof(unsigned) * (size_t)nseg * tiles_pad + 16);
        L->bin_tot = o; o = align_up(o + sizeof(unsigned) * (size_t)tiles_pad + 16);
    }
    L->spg = o; o = align_up(o + sizeof(int) * n1);
    L->svals = o; o = align_up(o + sizeof(unsigned long long) * c1);
    L->front = o; o = align_up(o + sizeof(unsigned long long) * (size_t)(4 + (n + 255) / 256));
    L->sort_tmp_bytes = binning_tmp_bytes(n, c1);
    L->sort_tmp = o; o = align_up(o + L->sort_tmp_bytes);
    L->total = o;
    return 0;
}

int gwbp_pack_scene(int64_t n, const float *means, const float *quats, const float *scales,
                    const float *opacities, void *geo, void *stream) {
    GWBP_REQUIRE(n >= 0, "n must be >= 0");
    if (n == 0) return 0;
    GWBP_REQUIRE(means && quats && scales && opacities && geo, "pack_scene: NULL pointer");
    GWBP_REQUIRE(((uintptr_t)quats & 15) == 0 && ((uintptr_t)geo & 15) == 0, "quats/geo must be 16-byte aligned");
    return launch_pack_scene(n, means, quats, scales, opacities, geo, (cudaStream_t)stream);
}

int gwbp_view_prepare(const gwbp_scene *scene, const gwbp_camera *cam, void *ws, size_t ws_bytes, int64_t cap,
                      int32_t flags, void *stream, gwbp_view_info *info) {
    GWBP_REQUIRE(scene && ws && info, "view_prepare: NULL pointer");
    if (int rc = check_cam(cam)) return rc;
    gwbp_ws_layout L;
    if (int rc = gwbp_workspace_layout(scene->n, cam->width, cam->height, cap, &L)) return rc;
    GWBP_REQUIRE(ws_bytes >= L.total, "workspace too small: %zu < %zu", ws_bytes, L.total);
    GWBP_REQUIRE(scene->n == 0 || scene->geo, "scene.geo is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    WsDev w = ws_view(ws, L);
    CamDev cd = make_cam(*cam);
    cd.cull = (flags & GWBP_PREPARE_TILE_CULL) ? 1 : 0;
    cd.super = (flags & GWBP_PREPARE_SUPERTILE) ? 1 : 0;
    cd.nsx = (cd.tw + kSuperW - 1) / kSuperW;
    const int nsy = (cd.th + kSuperH - 1) / kSuperH, n_super = cd.nsx * nsy;
    const int64_t n = scene->n;
    memset(info, 0, sizeof(*info));
    info->tile_w = cd.tw; info->tile_h = cd.th;
    info->cap_isects = cap;

    prof_mark(kEvProject0, st);
    if (int rc = launch_project_pack(n, scene->geo, cd, w, st)) return rc;  // projection + tile test + ordered compaction
    prof_mark(kEvProject1, st);
    unsigned long long totals[3] = {0ull, 0ull, 0ull};  // intersections, visible Gaussians, supertile entries (64-bit each)
    GWBP_CUDA_OK(cudaMemcpyAsync(totals, w.front + 1, sizeof(totals), cudaMemcpyDeviceToHost, st));
    GWBP_CUDA_OK(cudaStreamSynchronize(st));
    prof_mark(kEvCounts, st);
    info->n_vis = (int64_t)totals[1];
    info->n_isects = totals[0] > (unsigned long long)INT64_MAX ? INT64_MAX : (int64_t)totals[0];
    if (info->n_isects > cap) {
        set_error("intersection capacity exceeded: need %lld, workspace sized for %lld",
                  (long long)info->n_isects, (long long)cap);
        return -2;
    }
    info->n_entries = cd.super ? (int64_t)totals[2] : info->n_isects;
    prof_mark(kEvCompact, st);
    int dsel = 0;
    // the per-Gaussian entry counts ride along with the last pass of the depth sort (the counting-bin path gathers its
    // emission records as well and keeps its own gather kernel)
    const bool counting = !cd.super && bin_fast_supported(cd.tw * cd.th) && (flags & GWBP_PREPARE_COUNTING_BIN);
    const int *counts_src = counting ? nullptr : cd.super ? w.spg : w.tiles_per_gauss;
    if (int rc = launch_depth_sort(info->n_vis, w, &dsel, st, counts_src)) return rc;
    prof_mark(kEvDepthSort, st);
    const unsigned *order = w.dvals[dsel];
    const int n_tiles = cd.tw * cd.th;
    if (cd.super) {
        // supertile lists: one entry per (Gaussian, 8 x 4-tile supertile) in depth order, ONE stable radix pass on the
        // supertile id while there are <= 256 of them, ranges per supertile in `offsets`
        info->list_kind = 1;
        info->super_w = cd.nsx; info->super_h = nsy;
        if (int rc = launch_scan_counts(info->n_vis, w, st)) return rc;
        const int kb = n_super <= 256 ? 1 : n_super <= 65536 ? 2 : 4;
        info->tile_key_bytes = kb;
        if (int rc = launch_emit_super(info->n_vis, cd, order, w, cap, kb, st)) return rc;
        int sorted = 0, sbits = 1;
        while ((1 << sbits) < n_super) ++sbits;
        // <= 256 supertiles: the single pass's bin offsets are the per-supertile ranges
        if (int rc = launch_super_sort(info->n_entries, sbits, w, kb, &sorted, st, kb == 1 ? n_super : 0)) return rc;
        info->sorted_buf = sorted;
        const int rc = kb == 1 ? 0 : launch_offsets(info->n_entries, n_super, w.tkeys[sorted], kb == 2, w.offsets, st, false);
        prof_mark(kEvBin, st);
        return rc;
    }
    if (bin_fast_supported(n_tiles) && (flags & GWBP_PREPARE_COUNTING_BIN)) {
        // hand-written stable counting sort fused with the emission: no (tile, index) intermediate, no radix sort;
        // the depth-ordered prefix of the per-Gaussian hit counts cuts the list into chunks of equal work
        if (int rc = launch_gather_counts(info->n_vis, order, w, true, st)) return rc;
        if (int rc = launch_scan_counts(info->n_vis, w, st)) return rc;
        info->tile_key_bytes = 0;
        info->sorted_buf = 0;
        const int rc = launch_bin(n, cd, order, w, cap, st);
        prof_mark(kEvBin, st);
        return rc;
    }
    // default: emit (tile, index) pairs in depth order, stable radix sort on the <= 13 tile bits, range finding
    if (int rc = launch_scan_counts(info->n_vis, w, st)) return rc;
    const bool key16 = n_tiles <= 65536;
    info->tile_key_bytes = key16 ? 2 : 4;
    if (int rc = launch_emit(info->n_vis, cd, order, w, cap, key16, st)) return rc;
    int sorted = 0;
    const int tb = tile_bits_for(n_tiles);
    if (int rc = launch_tile_sort(info->n_isects, key16 && tb > 16 ? 16 : tb, w, key16, &sorted, st)) return rc;
    info->sorted_buf = sorted;
    const int rc = launch_offsets(info->n_isects, n_tiles, w.tkeys[sorted], key16, w.offsets, st);
    prof_mark(kEvBin, st);
    return rc;
}

int gwbp_profile_enable(int on) {
    g_prof_on = on != 0;
    for (int k = 0; k < kNumEv; ++k) g_ev_set[k] = false;
    return 0;
}

int gwbp_profile_read(float *ms_host, int n) {
    static const int pairs[GWBP_PROFILE_STAGES][2] = {{kEvProject0, kEvProject1}, {kEvProject1, kEvCounts},
                                                      {kEvCounts, kEvCompact},   {kEvCompact, kEvDepthSort},
                                                      {kEvDepthSort, kEvBin},    {kEvPack0, kEvPack1},
                                                      {kEvBp0, kEvBp1}};
    GWBP_REQUIRE(ms_host && n >= GWBP_PROFILE_STAGES, "profile_read: need room for %d floats", GWBP_PROFILE_STAGES);
    for (int i = 0; i < GWBP_PROFILE_STAGES; ++i) {
        ms_host[i] = -1.0f;
        const int a = pairs[i][0], b = pairs[i][1];
        if (!g_ev_ready || !g_ev_set[a] || !g_ev_set[b]) continue;
        GWBP_CUDA_OK(cudaEventSynchronize(g_ev[b]));
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, g_ev[a], g_ev[b]) == cudaSuccess) ms_host[i] = ms;
    }
    return 0;
}

unsigned long long gwbp_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int gwbp_debug_set_trace(void *buf, size_t bytes) {
    tc_set_trace(buf, bytes);
    return 0;
}

size_t gwbp_fpack_bytes(int32_t width, int32_t height, int32_t d) {
    if (width <= 0 || height <= 0 || !tc_supported(d)) return 0;
    return fpack_bytes(width, height, d);
}

int gwbp_pack_features(int32_t width, int32_t height, const float *F, int64_t sH, int64_t sW, int64_t sD, int32_t d,
                       void *fpack, void *stream) {
    GWBP_REQUIRE(width > 0 && height > 0, "pack_features: bad image size");
    GWBP_REQUIRE(tc_supported(d), "pack_features: tcgen05 path does not support D=%d", d);
    GWBP_REQUIRE(F && fpack, "pack_features: NULL pointer");
    prof_mark(kEvPack0, (cudaStream_t)stream);
    const int rc = launch_fpack(width, height, F, sH, sW, sD, d, fpack, (cudaStream_t)stream);
    prof_mark(kEvPack1, (cudaStream_t)stream);
    return rc;
}

int gwbp_pack_features_lowres(int32_t width, int32_t height, const float *S, int32_t src_h, int32_t src_w, int64_t sH,
                              int64_t sW, int64_t sD, int32_t nearest, int32_t d, void *fpack, void *stream) {
    GWBP_REQUIRE(width > 0 && height > 0, "pack_features_lowres: bad image size");
    GWBP_REQUIRE(tc_supported(d), "pack_features_lowres: tcgen05 path does not support D=%d", d);
    GWBP_REQUIRE(S && fpack, "pack_features_lowres: NULL pointer");
    prof_mark(kEvPack0, (cudaStream_t)stream);
    const int rc = launch_fpack_lowres(width, height, S, src_h, src_w, sH, sW, sD, nearest, d, fpack, (cudaStream_t)stream);
    prof_mark(kEvPack1, (cudaStream_t)stream);
    return rc;
}

int gwbp_backproject_view(const gwbp_scene *scene, const gwbp_camera *cam, const void *ws,
                          const gwbp_view_info *info, const float *F, int64_t sH, int64_t sW, int64_t sD, int32_t d,
                          float *num, float *den, int32_t kernel, void *fpack, int64_t *stats, void *stream) {
    GWBP_REQUIRE(scene && info, "backproject_view: NULL pointer");
    if (scene->n == 0 || info->n_isects == 0) return 0;
    GWBP_REQUIRE(ws && F && num && den, "backproject_view: NULL pointer");
    if (int rc = check_cam(cam)) return rc;
    GWBP_REQUIRE(d >= 1, "feature dimension must be >= 1 (got %d)", d);
    gwbp_ws_layout L;
    if (int rc = gwbp_workspace_layout(scene->n, cam->width, cam->height, info->cap_isects, &L)) return rc;
    const TileCtx t = tile_ctx(cam, ws, L, info);
    cudaStream_t st = (cudaStream_t)stream;
    if (info->n_isects == 0) return 0;
    int k = kernel & 0xff;
    if (k == GWBP_KERNEL_AUTO) k = (tc_supported(d) && fpack) ? GWBP_KERNEL_TC : GWBP_KERNEL_SIMT;
    if (k == GWBP_KERNEL_TC) {
        GWBP_REQUIRE(tc_supported(d), "tcgen05 back-projection does not support D=%d", d);
        GWBP_REQUIRE(fpack != nullptr, "tcgen05 back-projection needs the fpack buffer (gwbp_fpack_bytes)");
        prof_mark(kEvBp0, st);
        const int rc = launch_backproject_tc(t, F, sH, sW, sD, d, num, den, fpack, (kernel & GWBP_KERNEL_FPACK_READY) != 0,
                                             (long long *)stats, st);
        prof_mark(kEvBp1, st);
        return rc;
    }
    GWBP_REQUIRE(k == GWBP_KERNEL_SIMT, "unknown kernel id %d", kernel);
    GWBP_REQUIRE(info->list_kind == 0, "the CUDA-core kernel needs per-tile lists (view prepared with GWBP_PREPARE_SUPERTILE)");
    prof_mark(kEvBp0, st);
    const int rc = launch_backproject_simt(t, F, sH, sW, sD, d, num, den, (long long *)stats, st);
    prof_mark(kEvBp1, st);
    return rc;
}

int gwbp_lowres_adjoint_supported(int32_t width, int32_t height, int32_t src_h, int32_t src_w, int32_t d, int32_t nearest) {
    return lr_supported(width, height, src_h, src_w, d, nearest) ? 1 : 0;
}

int gwbp_pack_lowres_adjoint(const float *S, int32_t src_h, int32_t src_w, int64_t sH, int64_t sW, int64_t sD, int32_t d,
                             void *fpack, void *stream) {
    GWBP_REQUIRE(S && fpack, "pack_lowres_adjoint: NULL pointer");
    GWBP_REQUIRE(src_h >= 1 && src_w >= 1 && tc_supported(d), "pack_lowres_adjoint: bad shape (%d x %d x %d)", src_h, src_w, d);
    prof_mark(kEvPack0, (cudaStream_t)stream);
    const int rc = launch_lr_pack(S, src_h, src_w, sH, sW, sD, d, fpack, (cudaStream_t)stream);
    prof_mark(kEvPack1, (cudaStream_t)stream);
    return rc;
}

int gwbp_backproject_view_lowres(const gwbp_scene *scene, const gwbp_camera *cam, const void *ws,
                                 const gwbp_view_info *info, const float *S, int32_t src_h, int32_t src_w, int64_t sH,
                                 int64_t sW, int64_t sD, int32_t nearest, int32_t d, float *num, float *den, void *fpack,
                                 int64_t *stats, void *stream) {
    GWBP_REQUIRE(scene && info, "backproject_view_lowres: NULL pointer");
    if (scene->n == 0 || info->n_isects == 0) return 0;
    GWBP_REQUIRE(ws && S && num && den && fpack, "backproject_view_lowres: NULL pointer");
    if (int rc = check_cam(cam)) return rc;
    GWBP_REQUIRE(src_h >= 1 && src_w >= 1, "low-resolution map must be at least 1x1");
    GWBP_REQUIRE(tc_supported(d), "backproject_view_lowres: tcgen05 path does not support D=%d", d);
    gwbp_ws_layout L;
    if (int rc = gwbp_workspace_layout(scene->n, cam->width, cam->height, info->cap_isects, &L)) return rc;
    const TileCtx t = tile_ctx(cam, ws, L, info);
    cudaStream_t st = (cudaStream_t)stream;
    const bool packed = (nearest & GWBP_LOWRES_PACKED) != 0;
    nearest &= 1;
    if (lr_supported(cam->width, cam->height, src_h, src_w, d, nearest)) {
        // adjoint path: weights down-sampled on the tensor cores, contracted with the low-res map itself
        if (!packed) {
            prof_mark(kEvPack0, st);
            prof_mark(kEvPack1, st);
        }
        prof_mark(kEvBp0, st);
        const int rc = launch_backproject_lr(t, S, src_h, src_w, sH, sW, sD, nearest, d, num, den, fpack, packed,
                                             (long long *)stats, st);
        prof_mark(kEvBp1, st);
        return rc;
    }
    GWBP_REQUIRE(!packed, "GWBP_LOWRES_PACKED: this geometry is not covered by the adjoint kernel (gwbp_lowres_adjoint_supported)");
    // windows too large for the adjoint kernel (down-sampling or mild up-sampling): fused upsample + re-layout, then the
    // full-resolution contraction
    prof_mark(kEvPack0, st);
    if (int rc = launch_fpack_lowres(cam->width, cam->height, S, src_h, src_w, sH, sW, sD, nearest, d, fpack, st)) return rc;
    prof_mark(kEvPack1, st);
    prof_mark(kEvBp0, st);
    const int rc = launch_backproject_tc(t, nullptr, 0, 0, 0, d, num, den, fpack, true, (long long *)stats, st);
    prof_mark(kEvBp1, st);
    return rc;
}

int gwbp_render_view(const gwbp_scene *scene, const gwbp_camera *cam, const void *ws, const gwbp_view_info *info,
                     const float *colors, int64_t color_stride, int32_t d, const float *background, float *render,
                     float *alpha, int32_t kernel, void *stream) {
    GWBP_REQUIRE(scene && info, "render_view: NULL pointer");
    if (scene->n == 0 || info->n_isects == 0) return 0;
    GWBP_REQUIRE(ws && colors && render, "render_view: NULL pointer");
    if (int rc = check_cam(cam)) return rc;
    GWBP_REQUIRE(d >= 1, "channel count must be >= 1 (got %d)", d);
    GWBP_REQUIRE(color_stride >= d, "color_stride (%lld) < d (%d)", (long long)color_stride, d);
    GWBP_REQUIRE(info->list_kind == 0, "render_view needs per-tile lists (view prepared with GWBP_PREPARE_SUPERTILE)");
    gwbp_ws_layout L;
    if (int rc = gwbp_workspace_layout(scene->n, cam->width, cam->height, info->cap_isects, &L)) return rc;
    const TileCtx t = tile_ctx(cam, ws, L, info);
    int k = kernel & 0xff;
    if (k == GWBP_KERNEL_AUTO)
        k = (d >= 64 && render_tc_supported(colors, color_stride, d)) ? GWBP_KERNEL_TC : GWBP_KERNEL_SIMT;
    if (k == GWBP_KERNEL_TC) return launch_render_tc(t, colors, color_stride, d, background, render, alpha, (cudaStream_t)stream);
    GWBP_REQUIRE(k == GWBP_KERNEL_SIMT, "unknown kernel id %d", kernel);
    return launch_render_simt(t, colors, color_stride, d, background, render, alpha, (cudaStream_t)stream);
}

int gwbp_render_pixels(const gwbp_scene *scene, const gwbp_camera *cam, const void *ws, const gwbp_view_info *info,
                       const float *colors, int64_t color_stride, int32_t d, const float *extra, const int32_t *xy,
                       int32_t k, float *out, float *alpha, void *stream) {
    GWBP_REQUIRE(scene && info, "render_pixels: NULL pointer");
    GWBP_REQUIRE(k >= 0 && d >= 1, "render_pixels: bad shape (k=%d d=%d)", k, d);
    if (k == 0) return 0;
    GWBP_REQUIRE(xy && out, "render_pixels: NULL pointer");
    const int od = d + (extra ? 1 : 0);
    if (scene->n == 0 || info->n_isects == 0) {
        GWBP_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)k * od, (cudaStream_t)stream));
        if (alpha) GWBP_CUDA_OK(cudaMemsetAsync(alpha, 0, sizeof(float) * (size_t)k, (cudaStream_t)stream));
        return 0;
    }
    GWBP_REQUIRE(ws && colors, "render_pixels: NULL pointer");
    GWBP_REQUIRE(info->list_kind == 0, "render_pixels needs per-tile lists (view prepared with GWBP_PREPARE_SUPERTILE)");
    if (int rc = check_cam(cam)) return rc;
    GWBP_REQUIRE(color_stride >= d, "color_stride (%lld) < d (%d)", (long long)color_stride, d);
    gwbp_ws_layout L;
    if (int rc = gwbp_workspace_layout(scene->n, cam->width, cam->height, info->cap_isects, &L)) return rc;
    const TileCtx t = tile_ctx(cam, ws, L, info);
    return launch_render_pixels(t, colors, color_stride, d, extra, xy, k, out, alpha, (cudaStream_t)stream);
}

int gwbp_ratio_accumulate(const gwbp_scene *scene, const gwbp_camera *cam, const void *ws, const gwbp_view_info *info,
                          float *num_v, float *den_v, float *acc, float *den_acc, int32_t d, float num_scale,
                          float den_scale, float eps, void *stream) {
    GWBP_REQUIRE(scene && info, "ratio_accumulate: NULL pointer");
    GWBP_REQUIRE(d >= 1, "ratio_accumulate: bad shape (d=%d)", d);
    if (scene->n == 0 || info->n_vis == 0) return 0;
    GWBP_REQUIRE(ws && num_v && den_v && acc, "ratio_accumulate: NULL pointer");
    if (int rc = check_cam(cam)) return rc;
    gwbp_ws_layout L;
    if (int rc = gwbp_workspace_layout(scene->n, cam->width, cam->height, info->cap_isects, &L)) return rc;
    WsDev w = ws_view(const_cast<void *>(ws), L);
    return launch_ratio_accumulate(w.grec, info->n_vis, num_v, den_v, acc, den_acc, d, num_scale, den_scale, eps,
                                   (cudaStream_t)stream);
}

int gwbp_sh_colors(int64_t n, int32_t degree, const float *means, const float *coeffs, int64_t sN, int64_t sK, int64_t sC,
                   const float *cam_pos_host, float *out, void *stream) {
    GWBP_REQUIRE(n >= 0, "sh_colors: n must be >= 0");
    GWBP_REQUIRE(degree >= 0 && degree <= 4, "sh_colors: sh_degree must be 0..4 (got %d)", degree);
    if (n == 0) return 0;
    GWBP_REQUIRE(means && coeffs && cam_pos_host && out, "sh_colors: NULL pointer");
    return launch_sh_colors(n, degree, means, coeffs, sN, sK, sC, cam_pos_host, out, (cudaStream_t)stream);
}

int gwbp_finalize(const float *num, const float *den, float *out, int64_t n, int32_t d, void *stream) {
    GWBP_REQUIRE(n >= 0 && d >= 1, "finalize: bad shape");
    if (n == 0) return 0;
    GWBP_REQUIRE(num && den && out, "finalize: NULL pointer");
    return launch_finalize(num, den, out, n, d, (cudaStream_t)stream);
}

// ---- peer memory (CUDA IPC) + the fused closing step ----
int gwbp_ipc_export(const void *ptr, void *handle_out, int64_t *offset_out) {
    GWBP_REQUIRE(ptr && handle_out && offset_out, "ipc_export: NULL pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == GWBP_IPC_HANDLE_BYTES, "IPC handle size");
    typedef CUresult (*RangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);
    static RangeFn range = nullptr;
    if (!range) {
        void *fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            range = (RangeFn)fp;
    }
    GWBP_REQUIRE(range != nullptr, "cuMemGetAddressRange is not available from this driver");
    CUdeviceptr base = 0;
    size_t size = 0;
    const CUresult r = range(&base, &size, (CUdeviceptr)ptr);
    GWBP_REQUIRE(r == CUDA_SUCCESS, "cuMemGetAddressRange failed (%d): not a device allocation?", (int)r);
    cudaIpcMemHandle_t h;
    GWBP_CUDA_OK(cudaIpcGetMemHandle(&h, (void *)base));  // fails for VMM memory (expandable segments): caller falls back
    memcpy(handle_out, &h, sizeof(h));
    *offset_out = (int64_t)((CUdeviceptr)ptr - base);
    return 0;
}

int gwbp_ipc_open(const void *handle, void **base_out) {
    GWBP_REQUIRE(handle && base_out, "ipc_open: NULL pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    GWBP_CUDA_OK(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int gwbp_ipc_close(void *base) {
    if (!base) return 0;
    GWBP_CUDA_OK(cudaIpcCloseMemHandle(base));
    return 0;
}

int gwbp_peer_reduce_supported(int32_t world, int32_t d) { return peer_reduce_supported(world, d) ? 1 : 0; }

int gwbp_peer_reduce_finalize(const void *const *num_ptrs, const void *const *den_ptrs, int32_t world, int64_t lo,
                              int64_t rows, int32_t d, float eps, float *out_feat, float *out_num, float *out_den,
                              void *stream) {
    GWBP_REQUIRE(peer_reduce_supported(world, d), "peer_reduce_finalize: unsupported shape (world=%d, d=%d)", world, d);
    GWBP_REQUIRE(lo >= 0 && rows >= 0, "peer_reduce_finalize: bad row range");
    if (rows == 0) return 0;
    GWBP_REQUIRE(num_ptrs && den_ptrs, "peer_reduce_finalize: NULL pointer table");
    for (int r = 0; r < world; ++r) {
        GWBP_REQUIRE(num_ptrs[r] && den_ptrs[r], "peer_reduce_finalize: NULL pointer for rank %d", r);
        GWBP_REQUIRE(((uintptr_t)num_ptrs[r] & 15) == 0, "peer_reduce_finalize: num of rank %d is not 16-byte aligned", r);
    }
    GWBP_REQUIRE((((uintptr_t)out_feat | (uintptr_t)out_num) & 15) == 0, "peer_reduce_finalize: outputs must be 16-byte aligned");
    return launch_peer_reduce_finalize((const float *const *)num_ptrs, (const float *const *)den_ptrs, world, lo, rows, d, eps,
                                       out_feat, out_num, out_den, (cudaStream_t)stream);
}

int gwbp_mask3d(const float *x, int64_t rows, int32_t d, const float *text, int32_t p, int32_t npos, float threshold,
                int32_t use_threshold, uint8_t *mask, float *score, void *stream) {
    GWBP_REQUIRE(rows >= 0 && d >= 1, "mask3d: bad shape");
    if (rows == 0) return 0;
    GWBP_REQUIRE(x && text && mask, "mask3d: NULL pointer");
    return launch_mask(x, rows, d, text, p, npos, threshold, use_threshold, mask, score, (cudaStream_t)stream);
}

int gwbp_mask2d(const float *render, int64_t pixels, int32_t d, const float *text, int32_t p, int32_t npos,
                uint8_t *mask, void *stream) {
    GWBP_REQUIRE(pixels >= 0 && d >= 1, "mask2d: bad shape");
    if (pixels == 0) return 0;
    GWBP_REQUIRE(render && text && mask, "mask2d: NULL pointer");
    return launch_mask(render, pixels, d, text, p, npos, 0.0f, 0, mask, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
