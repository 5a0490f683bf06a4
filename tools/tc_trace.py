#!/usr/bin/env python
"""Per-role timeline of CTA 0 of bp_tc_kernel (debug): run on the GPU box."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gwbp
S = gwbp.scene
cfg = S.CONFIGS["G"]
W, H, d = cfg["width"], cfg["height"], cfg["d"]
sc = S.make_scene(cfg["n"], 0)
vm, K = S.make_cameras(cfg["views"], W, H, 0)
t = lambda a: torch.from_numpy(a).cuda()
bp = gwbp.BackProjector(t(sc.means), t(sc.quats), t(sc.scales), t(sc.opacities), d, kernel="tc")
F = S.make_feature_map_torch(0, d, H, W, "cuda", 0)
for v in range(3):
    bp.add_view(vm[v], K, W, H, F)
trace = torch.zeros(4 * 4096 * 2, dtype=torch.int64, device="cuda")
gwbp._lib.lib().gwbp_debug_set_trace(trace.data_ptr(), trace.numel() * 8)
bp.add_view(vm[3], K, W, H, F)
torch.cuda.synchronize()
gwbp._lib.lib().gwbp_debug_set_trace(None, 0)
tr = trace.cpu().numpy().reshape(4, 4096, 2)
names = {0: "ALU0", 1: "ALU7", 2: "EPI ", 3: "MMA "}
evn = {(0, 0): "batch_begin", (0, 1): "wfree_ok", (0, 2): "batch_end", (1, 0): "batch_begin", (1, 1): "wfree_ok", (1, 2): "batch_end",
       (2, 0): "chunk_begin", (2, 1): "chunk_end", (3, 0): "chunk_begin", (3, 1): "chunk_issued"}
events = []
for role in range(4):
    for k in range(4096):
        tag, clk = int(tr[role, k, 0]), int(tr[role, k, 1])
        if clk == 0:
            break
        events.append((clk, role, (tag >> 48) & 0xffff, tag & 0xffffffff, (tag >> 32) & 0xffff))
events.sort()
t0 = events[0][0]
print("n events", len(events))
lo, hi = 40, 64  # batches to print
for clk, role, ev, q, c in events:
    if lo <= q < hi:
        print(f"{clk - t0:10d}  {names[role]}  q={q:3d} c={c}  {evn[(role, ev)]}")
# summary: per-batch period from MMA chunk 0 begins
mm = [(q, clk) for clk, role, ev, q, c in events if role == 3 and ev == 0 and c == 0]
per = np.diff([c for _, c in mm])
print("MMA batch period cycles: median", np.median(per), "mean", per.mean(), "p10", np.percentile(per, 10), "p90", np.percentile(per, 90), "n", len(per))
for role in (0, 1):
    b = {q: clk for clk, r, ev, q, c in events if r == role and ev == 0}
    w = {q: clk for clk, r, ev, q, c in events if r == role and ev == 1}
    e = {q: clk for clk, r, ev, q, c in events if r == role and ev == 2}
    qs = sorted(set(b) & set(w) & set(e))
    print(names[role], "wait w_free median", np.median([w[q] - b[q] for q in qs]), "compute median", np.median([e[q] - w[q] for q in qs]),
          "batch-to-batch median", np.median(np.diff([b[q] for q in qs])))
eb = {(q, c): clk for clk, r, ev, q, c in events if r == 2 and ev == 0}
ee = {(q, c): clk for clk, r, ev, q, c in events if r == 2 and ev == 1}
ks = sorted(set(eb) & set(ee))
print("EPI chunk duration median", np.median([ee[k] - eb[k] for k in ks]))
mb = {(q, c): clk for clk, r, ev, q, c in events if r == 3 and ev == 0}
me = {(q, c): clk for clk, r, ev, q, c in events if r == 3 and ev == 1}
ks = sorted(set(mb) & set(me))
print("MMA chunk issue duration median", np.median([me[k] - mb[k] for k in ks]), "c0", np.median([me[k] - mb[k] for k in ks if k[1] == 0]), "c1", np.median([me[k] - mb[k] for k in ks if k[1] == 1]))
