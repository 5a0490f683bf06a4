#!/usr/bin/env python
"""Smallest adjoint (encoder-resolution) back-projection: used under compute-sanitizer when the kernel misbehaves."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gwbp
S = gwbp.scene
W, H, d = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (211, 137, 48)
enc = int(sys.argv[4]) if len(sys.argv) > 4 else 9
sc = S.make_scene(int(sys.argv[5]) if len(sys.argv) > 5 else 2000, 21)
vm, K = S.make_cameras(1, W, H, 21)
t = lambda a: torch.from_numpy(a).cuda()
low = torch.nn.functional.normalize(torch.randn(d, enc, enc, device="cuda"), dim=0).permute(1, 2, 0)
a = gwbp.BackProjector(t(sc.means), t(sc.quats), t(sc.scales), t(sc.opacities), d, kernel="tc")
b = gwbp.BackProjector(t(sc.means), t(sc.quats), t(sc.scales), t(sc.opacities), d, kernel="tc")
b.lowres_impl = "upsample"
a.add_view_lowres(vm[0], K, W, H, low, mode="bilinear")
torch.cuda.synchronize()
b.add_view_lowres(vm[0], K, W, H, low, mode="bilinear")
torch.cuda.synchronize()
seen = b.den > 1e-6
err = (a.num[seen] - b.num[seen]).norm(dim=1) / b.num[seen].norm(dim=1).clamp_min(1e-9)
print("rows", int(seen.sum()), "max rel err adjoint vs upsample", float(err.max()), "den max diff", float((a.den - b.den).abs().max()))
