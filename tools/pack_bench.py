#!/usr/bin/env python
"""Times the feature re-layout pass (gwbp_pack_features) for the input layouts it accepts, config-G size."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gwbp
L = gwbp._lib
W, H = 1297, 840
for d in (512, 64):
    planar = torch.randn(d, H, W, device="cuda")
    layouts = {"planar [D,H,W] view (reference)": planar.permute(1, 2, 0), "contiguous [H,W,D]": planar.permute(1, 2, 0).contiguous()}
    fp = torch.empty(gwbp.fpack_bytes(W, H, d), dtype=torch.uint8, device="cuda")
    ref = None
    for name, F in layouts.items():
        sH, sW, sD = F.stride()
        run = lambda: L.check(L.lib().gwbp_pack_features(W, H, F.data_ptr(), sH, sW, sD, d, fp.data_ptr(), None), "pack")
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        same = "" if ref is None else f" identical={bool(torch.equal(ref, fp))}"
        ref = fp.clone() if ref is None else ref
        print(f"D={d} {name}: {ms:.3f} ms  {2 * H * W * d * 4 / ms / 1e6:.0f} GB/s (r+w){same}")
