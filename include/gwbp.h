/* gwbp -- gradient-weighted feature back-projection for 3D Gaussian splats, B200 (sm_100a).
 *
 * C ABI of lib/libgwbp.so.  This is the drop-in boundary for ONE path of
 * JojiJoseph/3dgs-gradient-backprojection: everything `create_feature_field_lseg`
 * (backproject.py:25-172) asks of `gsplat.rasterization` + autograd, and the forward feature
 * render / cosine query of segment.py:26-61,209-224.  Plain pointers and sizes only; every
 * pointer is a DEVICE pointer unless the name says `_host`.  The library never allocates
 * caller-visible memory: the caller sizes one workspace with gwbp_workspace_layout() and owns it.
 *
 * Return value of every int function: 0 = ok, <0 = bad argument / capacity, >0 = cudaError_t.
 * gwbp_last_error() returns a thread-local description of the last failure.
 * A handle-free design: all per-view state lives in the caller's workspace, so one workspace
 * per (device, stream); functions are not thread-safe on the same workspace.
 *
 * Reference interface each entry point replaces (file:line in the reference repo; gsplat-1.4.0
 * internals per SURVEY.md §9):
 *   gwbp_pack_scene ........ quat/scale -> covariance part of fully_fused_projection
 *                            (called inside rasterization(): backproject.py:89,115,133)
 *   gwbp_view_prepare ...... fully_fused_projection(packed) + isect_tiles + radix sort +
 *                            isect_offset_encode of ONE rasterization() call
 *   gwbp_backproject_view .. rasterize_to_pixels backward w.r.t. colors driven by
 *                            `(render*feats).sum().backward()` and `render.sum().backward()`
 *                            (backproject.py:127-131,145-151) INCLUDING the accumulation
 *                            `gaussian_features += grad; gaussian_denoms += grad0[:,0]` (:149-150)
 *   gwbp_backproject_view_lowres  the same after `F.interpolate(encoder_out, (H,W))` (backproject.py:108-113,236-249),
 *                            taking the encoder-resolution map itself
 *   gwbp_render_view ....... rasterize_to_pixels forward for D-channel colours (segment.py:209-220)
 *   gwbp_render_pixels ..... the `rasterization(features, render_mode="RGB+D")` call of the click prompt,
 *                            of which only ONE pixel is read (click_and_segment.py:241-262)
 *   gwbp_ratio_accumulate .. `gaussian_features += grad / (grad0[:,0:1] + 1e-12)` per view
 *                            (affordance_transfer/demo_affordance_transfer.py:768-796)
 *   gwbp_sh_colors ......... the SH -> RGB stage of `rasterization(sh_degree=3)` (backproject.py:88-100)
 *   gwbp_finalize .......... backproject.py:166-169
 *   gwbp_peer_reduce_finalize  the multi-GPU closing step (all-reduce of the accumulators) + finalise,
 *                            fused into one kernel over NVLink peer memory; gwbp_ipc_* carry the mappings
 *   gwbp_mask3d ............ segment.py:52-58
 *   gwbp_mask2d ............ segment.py:221-224
 */
#ifndef GWBP_H
#define GWBP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GWBP_TILE 16
#define GWBP_ABI_VERSION 12

/* kernel selection for gwbp_backproject_view and gwbp_render_view */
#define GWBP_KERNEL_AUTO 0
#define GWBP_KERNEL_SIMT 1 /* fp32 CUDA-core contraction (truth kernel, any D) */
#define GWBP_KERNEL_TC 2   /* tcgen05 split-bf16 contraction, fp32 TMEM accumulation */
#define GWBP_KERNEL_FPACK_READY 0x100 /* OR-ed in: `fpack` was already filled by gwbp_pack_features */

/* flags for gwbp_view_prepare */
#define GWBP_PREPARE_GSPLAT_EXACT 0 /* intersection list == gsplat-1.4.0 isect_tiles (bounding-square test) */
#define GWBP_PREPARE_TILE_CULL 1    /* additionally drop (Gaussian, tile) pairs whose alpha stays < 1/255 on the
                                       whole tile: same accumulators, ~40 % shorter list to sort and walk */
#define GWBP_PREPARE_COUNTING_BIN 2 /* tile binning by the hand-written sort-free counting path (project.cu: bin_*_kernel)
                                       instead of emit + radix sort; same flatten_ids / isect_offsets, no sorted tile keys
                                       (tile_key_bytes == 0); images of <= 12 288 tiles only.  Measured SLOWER than the
                                       radix-sort path on B200 (DESIGN.md "dead ends"), hence opt-in */

#define GWBP_PREPARE_SUPERTILE 4    /* bin into SUPERTILES of 8 x 4 tiles (128 x 64 px): one (supertile, depth-ordered) entry
                                       = (packed index, 32-bit mask of the tiles hit) per Gaussian and supertile, ~2.4x fewer
                                       entries to emit and sort, and <= 256 supertiles up to 1920 x 1088 = ONE 8-bit sort
                                       pass.  The tcgen05 back-projection kernels filter a supertile's list down to their
                                       tile on the fly (same Gaussians, same order, same results); every other consumer
                                       needs the per-tile lists.  view_info.list_kind == 1 */
#define GWBP_SUPER_W 8
#define GWBP_SUPER_H 4

typedef struct gwbp_scene {
    int64_t n;        /* Gaussians */
    const void *geo;  /* 40*n bytes written by gwbp_pack_scene: float4 (mean.xyz, opacity)[n],
                         float4 (c00,c01,c02,c11)[n], float2 (c12,c22)[n] */
} gwbp_scene;

typedef struct gwbp_camera {
    float viewmat[16]; /* row-major world->camera, as get_viewmat_from_colmap_image (utils.py:215-219) */
    float K[9];        /* row-major intrinsics */
    int32_t width, height;
    float near_plane, far_plane, radius_clip, eps2d; /* gsplat defaults 0.01, 1e10, 0.0, 0.3 */
} gwbp_camera;

/* byte offsets of every stage output inside the caller's workspace */
typedef struct gwbp_ws_layout {
    size_t total;
    size_t cnt;       /* uint64 [n+1]   (visible<<32 | tiles) per Gaussian */
    size_t scan;      /* uint64 [n+1]   exclusive prefix of cnt */
    size_t rec;       /* float4 [2*n]   unpacked projection records */
    size_t mask;      /* uint64 [n]     unpacked tile-hit masks (tile culling) */
    size_t grec;      /* float4 [2*n]   packed records: (mean2d.xy, opacity, gaussian_id bits), (conic.xyz, depth) */
    size_t erec;      /* uint4 [n]      emission records of the visible Gaussians: tile-hit mask, packed rectangle */
    size_t radii;     /* int32  [n]     packed radii */
    size_t tiles_per_gauss; /* int32 [n] packed */
    size_t dkeys0, dkeys1;  /* uint32 [n]  depth bits of the visible Gaussians (sort double buffer) */
    size_t dvals0, dvals1;  /* uint32 [n]  packed index */
    size_t cnt2;      /* uint32 [n+1]   tile counts in depth order */
    size_t base2;     /* uint32 [n+1]   exclusive prefix of cnt2 */
    size_t tkeys0, tkeys1;  /* uint32|uint16 [cap] tile id per intersection (sort double buffer; see tile_key_bytes) */
    size_t tvals0, tvals1;  /* int32  [cap] flatten_ids: packed index per intersection */
    size_t offsets;   /* int32  [tiles+1] isect_offsets (+ terminator = n_isects) */
    size_t stats;     /* int64  [16]    device counters */
    size_t bin_counts; /* uint32 [chunks][tiles] per-chunk tile histograms -> exclusive prefixes (sort-free binning) */
    size_t bin_seg;    /* uint32 [segments][tiles] */
    size_t bin_tot;    /* uint32 [tiles]  intersections per tile */
    size_t spg;       /* int32  [n]     supertile entries per packed Gaussian (GWBP_PREPARE_SUPERTILE) */
    size_t svals;     /* uint64 [cap]   second buffer of the (packed index | tile mask << 32) entries; the first is tvals0..tvals1 */
    size_t front;     /* uint64 [4 + ceil(n/256)] front-end control block: CTA ticket, intersection / visible totals,
                         one chained-scan status word per projection CTA */
    size_t sort_tmp;   /* scratch for scan / sort */
    size_t sort_tmp_bytes;
} gwbp_ws_layout;

typedef struct gwbp_view_info {
    int64_t n_vis, n_isects;
    int64_t cap_isects; /* the capacity the workspace layout was computed with */
    int32_t tile_w, tile_h;
    int32_t sorted_buf; /* which of tkeys0/tkeys1, tvals0/tvals1 holds the sorted result */
    int32_t tile_key_bytes; /* 0: sort-free binning, no tile keys materialised (flatten_ids in tvals0); 2 or 4: element
                               size of the sorted tkeys buffer of the radix-sort path (16-bit keys when tiles <= 65536);
                               supertile lists: 1, 2 or 4 */
    int32_t list_kind;      /* 0: per-tile lists (flatten_ids + isect_offsets); 1: per-supertile lists (GWBP_PREPARE_SUPERTILE):
                               `offsets` holds super_w * super_h + 1 supertile ranges over the sorted 64-bit entries */
    int32_t super_w, super_h; /* supertiles per row / column (list_kind 1) */
    int32_t reserved;
    int64_t n_entries;      /* sorted list entries: == n_isects for per-tile lists, (Gaussian, supertile) pairs otherwise */
} gwbp_view_info;

/* counters filled by gwbp_backproject_view when `stats` != NULL (device int64[4]):
 * [0] rows with non-zero weight  [1] (tile,Gaussian) entries walked  [2],[3] reserved */

int gwbp_abi_version(void);
/* kernels this library has launched so far in this process (bench.py reports the difference over its timed region) */
unsigned long long gwbp_launch_count(void);
const char *gwbp_last_error(void);

int gwbp_workspace_layout(int64_t n, int32_t width, int32_t height, int64_t cap_isects, gwbp_ws_layout *out_host);

/* means [n,3], quats [n,4] wxyz un-normalised, scales [n,3] linear, opacities [n] -> geo (40*n bytes) */
int gwbp_pack_scene(int64_t n, const float *means, const float *quats, const float *scales,
                    const float *opacities, void *geo, void *stream);

/* project + depth-sort + tile-bin one camera into `ws`.  Synchronises `stream` once (intersection count). */
int gwbp_view_prepare(const gwbp_scene *scene, const gwbp_camera *cam_host, void *ws, size_t ws_bytes,
                      int64_t cap_isects, int32_t flags, void *stream, gwbp_view_info *info_host);

/* DEBUG/PROFILING ONLY: when set to a device buffer of >= 4*4096*16 bytes, CTA 0 of the next tcgen05
 * back-projection launches records (role, event, batch, chunk, clock64) tuples into it; NULL disables. */
int gwbp_debug_set_trace(void *buf, size_t bytes);

/* PROFILING ONLY (bench.py `roofline.stages`): while enabled, CUDA events are recorded on the caller's stream at the
 * stage boundaries of gwbp_view_prepare / gwbp_pack_features* / gwbp_backproject_view (process-wide, one device).
 * gwbp_profile_read waits for the last view and returns the milliseconds of its stages (-1 = stage not run):
 *   [0] project  [1] count scan + host read-back of the totals (the one host sync)  [2] compact  [3] depth sort
 *   [4] tile binning  [5] feature re-layout  [6] fused back-projection kernel */
#define GWBP_PROFILE_STAGES 7
int gwbp_profile_enable(int on);
int gwbp_profile_read(float *ms_host, int n);

/* bytes of the packed bf16 feature buffer the GWBP_KERNEL_TC path needs (0 if D unsupported) */
size_t gwbp_fpack_bytes(int32_t width, int32_t height, int32_t d);

/* Re-layout of one feature map for the tcgen05 path (fp32, element strides -> bf16 hi/lo, tile-major).
 * gwbp_backproject_view does this itself unless GWBP_KERNEL_FPACK_READY is set; exposing it lets the
 * caller run it on a second stream, concurrently with gwbp_view_prepare of the same view (it depends
 * only on F). */
int gwbp_pack_features(int32_t width, int32_t height, const float *F, int64_t sH, int64_t sW, int64_t sD, int32_t d,
                       void *fpack, void *stream);

/* Same, fused with the upsample the reference performs first (backproject.py:110-112 bilinear,
 * :245-249 nearest): S is the ENCODER-resolution map [src_h, src_w, d] (element strides sH,sW,sD), sampled
 * with torch.nn.functional.interpolate(align_corners=False) arithmetic.  The full-resolution [H,W,d] map
 * (2.2 GB at config G) is never materialised. */
int gwbp_pack_features_lowres(int32_t width, int32_t height, const float *S, int32_t src_h, int32_t src_w, int64_t sH,
                              int64_t sW, int64_t sD, int32_t nearest, int32_t d, void *fpack, void *stream);

/* num[n,d] += sum_p w(g,p) F[p,:]; den[n] += sum_p w(g,p) for the prepared view.
 * F: fp32, element strides (sH,sW,sD) -- both [H,W,D]-contiguous and the reference's permuted
 * [D,H,W] view (backproject.py:113) are accepted as they are. */
int gwbp_backproject_view(const gwbp_scene *scene, const gwbp_camera *cam_host, const void *ws,
                          const gwbp_view_info *info_host, const float *F, int64_t sH, int64_t sW, int64_t sD,
                          int32_t d, float *num, float *den, int32_t kernel, void *fpack, int64_t *stats,
                          void *stream);

/* The same accumulation for an ENCODER-RESOLUTION map S [src_h, src_w, d] (element strides sH,sW,sD), i.e.
 *   gwbp_backproject_view(F = interpolate(S, size=(H,W), mode=bilinear|nearest))      (backproject.py:108-113, :236-249)
 * without building F or its packed copy: the weights are down-sampled on the tensor cores (the adjoint of the
 * up-sample) and contracted with the low-res map fetched by TMA (backproject_lr.cu).  `fpack` is scratch of at least
 * gwbp_fpack_bytes(width, height, d) bytes.  Geometries whose per-tile window of S exceeds 8 rows x 8 texels
 * (gwbp_lowres_adjoint_supported() == 0) silently take gwbp_pack_features_lowres + the full-resolution kernel.
 * The bf16 copy of S the TMA boxes read (2 x 59 MB at 240 x 240 x 512) depends only on S: gwbp_pack_lowres_adjoint writes
 * it into `fpack` on its own (e.g. on a second stream next to gwbp_view_prepare of the same view); pass
 * `nearest | GWBP_LOWRES_PACKED` to gwbp_backproject_view_lowres then (adjoint-supported geometries only). */
#define GWBP_LOWRES_PACKED 2
int gwbp_pack_lowres_adjoint(const float *S, int32_t src_h, int32_t src_w, int64_t sH, int64_t sW, int64_t sD, int32_t d,
                             void *fpack, void *stream);
int gwbp_lowres_adjoint_supported(int32_t width, int32_t height, int32_t src_h, int32_t src_w, int32_t d, int32_t nearest);
int gwbp_backproject_view_lowres(const gwbp_scene *scene, const gwbp_camera *cam_host, const void *ws,
                                 const gwbp_view_info *info_host, const float *S, int32_t src_h, int32_t src_w, int64_t sH,
                                 int64_t sW, int64_t sD, int32_t nearest, int32_t d, float *num, float *den, void *fpack,
                                 int64_t *stats, void *stream);

/* render[H,W,d] = sum_g w(g,p) colors[g,:] (+ (1-alpha) background), alpha[H,W] = 1-T.  Every in-image pixel
 * of `render` and `alpha` is written (no pre-zeroing needed).  kernel: GWBP_KERNEL_SIMT = fp32 CUDA cores,
 * weights regenerated per 32-channel chunk like gsplat; GWBP_KERNEL_TC = tcgen05 split-bf16 contraction, weights
 * generated once per 256 channels (needs 32 <= d, d % 4 == 0, 16-byte aligned rows); AUTO picks TC when it can
 * and d >= 64.  The TC kernel uses the workspace regions that are dead once a view is prepared (counts, scans,
 * unpacked records, hit masks) as scratch for its weight cache: `ws` is const only as far as the prepared view
 * (records, sorted lists, offsets) is concerned, and two renders of one workspace must not run concurrently. */
int gwbp_render_view(const gwbp_scene *scene, const gwbp_camera *cam_host, const void *ws,
                     const gwbp_view_info *info_host, const float *colors, int64_t color_stride, int32_t d,
                     const float *background, float *render, float *alpha, int32_t kernel, void *stream);

/* The same composite evaluated at k probe pixels only: out[i, 0:d] = sum_g w(g,p_i) colors[g,:] and, if
 * `extra` [n] != NULL, out[i, d] = sum_g w(g,p_i) extra[g] (e.g. camera depth: render_mode="RGB+D").
 * xy: device int32 [k,2] (x, y); out: [k, d + (extra != NULL)]; alpha [k] optional.  A pixel outside the
 * image yields zeros.  Replaces a full D-channel render when only clicked pixels are read
 * (click_and_segment.py:241-262: 513 channels rendered, one pixel used). */
int gwbp_render_pixels(const gwbp_scene *scene, const gwbp_camera *cam_host, const void *ws,
                       const gwbp_view_info *info_host, const float *colors, int64_t color_stride, int32_t d,
                       const float *extra, const int32_t *xy, int32_t k, float *out, float *alpha, void *stream);

/* Per-view-ratio accumulation (affordance_transfer/demo_affordance_transfer.py:768-796):
 *   acc[g,:] += (num_scale * num_v[g,:]) / (den_scale * den_v[g] + eps)   for every Gaussian the prepared view saw,
 * after which num_v[g,:] and den_v[g] are reset to 0, so the per-view scratch pair is ready for the next
 * view without a dense [n,d] pass.  den_acc [n] (optional) += den_v, which keeps the den > 0 prune mask
 * available in this mode.  num_v/den_v must have been filled by gwbp_backproject_view for THIS view
 * starting from zeros. */
int gwbp_ratio_accumulate(const gwbp_scene *scene, const gwbp_camera *cam_host, const void *ws,
                          const gwbp_view_info *info_host, float *num_v, float *den_v, float *acc, float *den_acc,
                          int32_t d, float num_scale, float den_scale, float eps, void *stream);

/* colours[g, c] = max(sum_k Y_k(normalise(mean_g - cam_pos)) * coeffs[g, k, c] + 0.5, 0), k < (degree+1)^2, degree <= 4:
 * the `rasterization(..., sh_degree=3)` colour stage (backproject.py:88-100, segment.py:197-208; gsplat-1.4.0
 * spherical_harmonics + clamp_min(. + 0.5, 0)).  means [n,3] contiguous; coeffs [n,K,3] with ELEMENT strides
 * (sN, sK, sC) so `torch.cat((features_dc, features_rest), 1)` or separate views need no copy; cam_pos_host: 3 floats
 * on the HOST (camera centre in world space); out [n,3] contiguous. */
int gwbp_sh_colors(int64_t n, int32_t degree, const float *means, const float *coeffs, int64_t sN, int64_t sK, int64_t sC,
                   const float *cam_pos_host, float *out, void *stream);

/* out[g,:] = normalise(num[g,:]/den[g]); NaN -> 0   (backproject.py:166-169); out may alias num */
int gwbp_finalize(const float *num, const float *den, float *out, int64_t n, int32_t d, void *stream);

/* ---- closing step of a view-sharded job over NVLink peer memory (replaces the NCCL all-reduce of SURVEY.md 8e + the
 * finalise of backproject.py:166-169 when every rank only needs ITS rows of the field) ----
 * gwbp_ipc_export: CUDA IPC handle (64 bytes) of the allocation that holds `ptr` + the byte offset of `ptr` inside it;
 * gwbp_ipc_open: map a peer process' allocation into this process (peer access enabled lazily), returns its base;
 * gwbp_ipc_close: unmap.  The handles travel over the host side's own channel (torch.distributed all_gather_object). */
#define GWBP_MAX_PEERS 8
#define GWBP_IPC_HANDLE_BYTES 64
int gwbp_ipc_export(const void *ptr, void *handle_out, int64_t *offset_out);
int gwbp_ipc_open(const void *handle, void **base_out);
int gwbp_ipc_close(void *base);
/* 1 if gwbp_peer_reduce_finalize covers (world, d): world <= 8, d % 4 == 0, d <= 1024 */
int gwbp_peer_reduce_supported(int32_t world, int32_t d);
/* For the rows [lo, lo + rows) this rank owns: den = den_0 + sum_{r>0}(den_r - eps); num = sum over the ranks that
 * touched the row (den_r > eps), rank order; out_feat = normalise(num/den), NaN -> 0.  num_ptrs / den_ptrs: HOST arrays
 * of `world` DEVICE pointers to every rank's full num[N,d] / den[N] (own pointers for this rank, gwbp_ipc_open mappings
 * for the peers).  out_feat [rows,d], out_num [rows,d], out_den [rows]: each optional (NULL).  The caller orders the call
 * after every rank's last view (barrier) and keeps the accumulators untouched until every rank's call has finished. */
int gwbp_peer_reduce_finalize(const void *const *num_ptrs, const void *const *den_ptrs, int32_t world, int64_t lo,
                              int64_t rows, int32_t d, float eps, float *out_feat, float *out_num, float *out_den,
                              void *stream);

/* mask[i] = max_{j<npos} s_ij > max_{j>=npos} s_ij (and s_i0 > threshold if use_threshold),
 * s = normalise(x_i) . normalise(text_j).  x [rows,d] contiguous, text [p,d]; score [rows,p] optional */
int gwbp_mask3d(const float *x, int64_t rows, int32_t d, const float *text, int32_t p, int32_t npos,
                float threshold, int32_t use_threshold, uint8_t *mask, float *score, void *stream);
/* same compare for rendered pixels (segment.py:221-224); no threshold */
int gwbp_mask2d(const float *render, int64_t pixels, int32_t d, const float *text, int32_t p, int32_t npos,
                uint8_t *mask, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GWBP_H */
