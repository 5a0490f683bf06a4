#!/usr/bin/env python
"""Config Q (BASELINE.json configs[3]): forward 512-d feature render + text cosine mask per view, garden scale.
Times (a) the exact path: D-channel render -> per-pixel normalise -> scores -> compare (segment.py:209-224),
(b) the linearity path: render the P per-Gaussian scores instead of D channels, (c) the 3-D mask."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gwbp
S = gwbp.scene
cfg = S.CONFIGS["G"]
W, H, d = cfg["width"], cfg["height"], cfg["d"]
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 6
sc = S.make_scene(cfg["n"], 0)
vm, K = S.make_cameras(cfg["views"], W, H, 0)
t = lambda a: torch.from_numpy(a).cuda()
scene = gwbp.PackedScene(t(sc.means), t(sc.quats), t(sc.scales), t(sc.opacities))
feats = torch.nn.functional.normalize(torch.randn(sc.n, d, device="cuda"), dim=1)
text = t(S.make_text_queries(3, d, 0))
out = {}
def timed(fn, n):
    fn(0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): r = fn(i + 1)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r
ms_exact, m_exact = timed(lambda i: gwbp.render_mask_2d(scene, feats, text, 1, vm[i], K, W, H, exact_render=True), nv)
scores = gwbp.gaussian_scores(feats, text)  # once per query
ms_lin, m_lin = timed(lambda i: gwbp.render_mask_2d(scene, feats, text, 1, vm[i], K, W, H, exact_render=False, scores=scores), nv)
ms_3d, _ = timed(lambda i: gwbp.get_mask3d(feats, text, 1), 5)
# forward render kernels alone, on one prepared view (prepare excluded): fp32 CUDA cores vs tcgen05
view = gwbp.View(scene, gwbp.make_camera(vm[1], K, W, H), tile_cull=True)
ms_simt, (r_simt, _) = timed(lambda i: view.render(feats, None, gwbp.KERNEL_SIMT), 3)
ms_tc, (r_tc, _) = timed(lambda i: view.render(feats, None, gwbp.KERNEL_TC), 5)
render_maxdiff = float((r_simt - r_tc).abs().max())
del r_simt, r_tc
diff = int((m_exact != m_lin).sum())
out = {"config": "Q", "views_timed": nv, "ms_per_view_exact_render": ms_exact, "frames_per_s_exact": 1e3 / ms_exact,
       "ms_per_view_score_render": ms_lin, "frames_per_s_score_render": 1e3 / ms_lin, "mask3d_ms": ms_3d,
       "mask3d_GBps": sc.n * d * 4 / ms_3d / 1e6, "pixels_differing_between_paths": diff, "pixels": W * H,
       "render_kernel_ms_simt": ms_simt, "render_kernel_ms_tcgen05": ms_tc, "render_tc_vs_simt_max_abs_diff": render_maxdiff,
       "render_output_GBps_tcgen05": W * H * d * 4 / ms_tc / 1e6}
print(json.dumps(out))
