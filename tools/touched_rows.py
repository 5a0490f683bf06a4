#!/usr/bin/env python
"""How many rows of the [N, D] accumulators does one rank of a view-sharded job touch?  (Sizing of a sparse closing
exchange: config G, rank 0 of 8 = views 0, 8, 16, ...)  Prints the union fraction after k views."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gwbp

S = gwbp.scene
cfg = S.CONFIGS["G"]
W, H, d = cfg["width"], cfg["height"], 16   # the touched set does not depend on D
sc = S.make_scene(cfg["n"], 0)
vm, K = S.make_cameras(cfg["views"], W, H, 0)
dev = torch.device("cuda")
t = lambda a: torch.from_numpy(a).to(dev)
bp = gwbp.BackProjector(t(sc.means), t(sc.quats), t(sc.scales), t(sc.opacities), d)
F = torch.rand(H, W, d, device=dev)
for world in (8, 4, 2):
    bp.reset()
    for i in range(20):
        bp.add_view(vm[(i * world) % cfg["views"]], K, W, H, F)
        if i in (0, 4, 9, 19):
            print(f"world {world}: after {i + 1:2d} views of rank 0: touched rows {float((bp.den > 1e-12).float().mean()):.4f} of N")
