#!/usr/bin/env python
"""Times the forward feature render kernels on one prepared config-G view (prepare excluded).
usage: render_bench.py [d] [reps]   (GWBP_RENDER_DEBUG selects experiment variants of the tcgen05 kernel)"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gwbp
S = gwbp.scene
cfg = S.CONFIGS["G"]
W, H = cfg["width"], cfg["height"]
d = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["d"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
sc = S.make_scene(cfg["n"], 0)
vm, K = S.make_cameras(cfg["views"], W, H, 0)
t = lambda a: torch.from_numpy(a).cuda()
scene = gwbp.PackedScene(t(sc.means), t(sc.quats), t(sc.scales), t(sc.opacities))
feats = torch.nn.functional.normalize(torch.randn(sc.n, d, device="cuda"), dim=1)
out = {"d": d, "debug": os.environ.get("GWBP_RENDER_DEBUG", "0")}
for cull in (True,):
    view = gwbp.View(scene, gwbp.make_camera(vm[1], K, W, H), tile_cull=cull)
    for name, k in (("tcgen05", gwbp.KERNEL_TC), ("simt", gwbp.KERNEL_SIMT)):
        if name == "simt" and out["debug"] != "0":
            continue
        view.render(feats, None, k); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): view.render(feats, None, k)
        e1.record(); torch.cuda.synchronize()
        out[f"ms_{name}"] = e0.elapsed_time(e1) / reps
    out["n_isects"] = view.n_isects
print(json.dumps(out))
# reference points for the output stream: plain fill and copy of a render-sized tensor
buf = torch.empty(H, W, d, device="cuda"); src = torch.randn(H, W, d, device="cuda")
for name, fn in (("fill", lambda: buf.zero_()), ("copy", lambda: buf.copy_(src))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({name + "_ms": e0.elapsed_time(e1) / 5, "GB": buf.numel() * 4 / 1e9}))
if os.environ.get("GWBP_RENDER_PROF"):
    # per-role cycle accounting of render_tc_kernel (debug counters behind gwbp_debug_set_trace)
    trace = torch.zeros(4 * 4096 * 2, dtype=torch.int64, device="cuda")
    lib = gwbp._lib.lib()
    lib.gwbp_debug_set_trace(trace.data_ptr(), trace.numel() * 8)
    view.render(feats, None, gwbp.KERNEL_TC); torch.cuda.synchronize()
    lib.gwbp_debug_set_trace(None, 0)
    pr = trace[:32].cpu().numpy().reshape(4, 8).astype(float)
    n_cta = 148
    names = {0: ("ALU", ["other", "popc barrier", "wait entry slot", "wait w_free", "generate+store W"]),
             1: ("LOAD", ["other", "wait entry", "issue loads", "wait x_empty", "data wait+split+store"]),
             2: ("EPI", ["other", "wait entry", "wait acc_full", "tmem ld + stores"]),
             3: ("MMA", ["other", "wait entry", "wait acc_empty", "wait w_full", "wait x_full", "issue"])}
    for r, (nm, cats) in names.items():
        tot = pr[r, :len(cats)].sum()
        print(nm, "total Mcycles/CTA %.2f |" % (tot / n_cta / 1e6), "  ".join(f"{c}: {pr[r, i] / tot * 100:.1f}%" for i, c in enumerate(cats)))
