#!/usr/bin/env python
"""Per-role timeline of CTA 0 of bp_lr_kernel (debug, run on the GPU box): roles ALU warp 0, converter warp 12,
epilogue warp 8, MMA warp.  Prints a window of the timeline and per-stage medians (cycles)."""
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
low = torch.nn.functional.normalize(torch.randn(d, 240, 240, device="cuda"), dim=0).permute(1, 2, 0)
for v in range(3):
    bp.add_view_lowres(vm[v], K, W, H, low)
trace = torch.zeros(4 * 4096 * 2, dtype=torch.int64, device="cuda")
gwbp._lib.lib().gwbp_debug_set_trace(trace.data_ptr(), trace.numel() * 8)
bp.add_view_lowres(vm[3], K, W, H, low)
torch.cuda.synchronize()
gwbp._lib.lib().gwbp_debug_set_trace(None, 0)
tr = trace.cpu().numpy().reshape(4, 4096, 2)
names = {0: "ALU0", 1: "CVT ", 2: "EPI ", 3: "MMA "}
evn = {(0, 0): "batch_begin", (0, 1): "wfree_ok", (0, 2): "batch_end", (1, 0): "d1_full_ok", (1, 1): "a2_empty_ok",
       (1, 2): "a2_written", (2, 0): "chunk_begin", (2, 1): "chunk_end", (3, 0): "g1_begin", (3, 1): "g1_issued",
       (3, 2): "a2_full_ok", (3, 3): "g2_chunk_issued"}
events = []
for role in range(4):
    for k in range(4096):
        tag, clk = int(tr[role, k, 0]), int(tr[role, k, 1])
        if clk == 0:
            break
        events.append((clk, role, (tag >> 48) & 0xffff, tag & 0xffffffff, (tag >> 32) & 0xffff))
events.sort()
t0 = events[0][0]
print("n events", len(events), "span cycles", events[-1][0] - t0)
lo, hi = 30, 38
for clk, role, ev, q, c in events:
    if lo <= q < hi:
        print(f"{clk - t0:10d}  {names[role]}  q={q:3d} c={c}  {evn[(role, ev)]}")
def ev(role, e, c=None):
    return {(q, cc): clk for clk, r, e_, q, cc in events if r == role and e_ == e and (c is None or cc == c)}
g1b, g1e, a2ok = ev(3, 0), ev(3, 1), ev(3, 2)
per = np.diff(sorted(g1b.values()))
print("batches", len(g1b), "MMA batch period: median", np.median(per), "mean", per.mean(), "p10", np.percentile(per, 10), "p90", np.percentile(per, 90))
ks = sorted(set(g1b) & set(g1e) & set(a2ok))
print("G1 issue duration median", np.median([g1e[k] - g1b[k] for k in ks]), " wait for A2 after G1 median", np.median([a2ok[k] - g1e[k] for k in ks]))
g2 = {}
for clk, r, e_, q, cc in events:
    if r == 3 and e_ == 3:
        g2.setdefault(q, {})[cc] = clk
print("G2 all chunks issued after a2_full_ok median", np.median([max(g2[q].values()) - a2ok[(q, 0)] for q in g2 if (q, 0) in a2ok]),
      " per chunk:", [float(np.median([g2[q][c] - (g2[q][c - 1] if c else a2ok[(q, 0)]) for q in g2 if c in g2[q] and (q, 0) in a2ok])) for c in range(3)])
ab, aw, ae = ev(0, 0), ev(0, 1), ev(0, 2)
ks = sorted(set(ab) & set(aw) & set(ae))
print("ALU0: wait w_free median", np.median([aw[k] - ab[k] for k in ks]), "compute median", np.median([ae[k] - aw[k] for k in ks]),
      "batch-to-batch median", np.median(np.diff([ab[k] for k in ks])))
eb, ee = ev(2, 0), ev(2, 1)
ks = sorted(set(eb) & set(ee))
print("EPI chunk duration median", np.median([ee[k] - eb[k] for k in ks]), "n", len(ks))
cb, cw, cd = ev(1, 0), ev(1, 1), ev(1, 2)
ks = sorted(set(cb) & set(cw) & set(cd))
print("CVT: wait a2_empty median", np.median([cw[k] - cb[k] for k in ks]), "convert median", np.median([cd[k] - cw[k] for k in ks]))
