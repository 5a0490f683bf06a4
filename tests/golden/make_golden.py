"""Generates tests/golden/small_case.npz from the numpy oracle (oracle/gsplat_oracle.py).

    python tests/golden/make_golden.py

The reference cannot be imported here (gsplat-1.4.0 is un-vendored and not installable: SURVEY.md
§0.3), so these vectors come from our restatement; they pin the C oracle and the CUDA kernels to it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gwbp  # noqa: E402
from helpers import small_case  # noqa: E402
from oracle import gsplat_oracle as O  # noqa: E402

cfg = dict(n=3000, views=2, width=96, height=64, d=8, seed=1)
sc, vm, K, feats = small_case(gwbp.scene, **cfg)
W, H, d = cfg["width"], cfg["height"], cfg["d"]
num = np.zeros((sc.n, d))
den = np.zeros(sc.n)
out = {("cfg_" + k): np.int64(v) for k, v in cfg.items()}
for v in range(cfg["views"]):
    a, b = O.backproject_view(sc.means, sc.quats, sc.scales, sc.opacities, vm[v], K, W, H, feats[v])
    num += a
    den += b
proj, isect = O.view_geometry(sc.means, sc.quats, sc.scales, vm[0], K, W, H)
out.update(isect_ids_v0=isect["isect_ids"], flatten_ids_v0=isect["flatten_ids"],
           isect_offsets_v0=isect["isect_offsets"], radii_v0=proj["radii"], num=num, den=den)
f = O.finalize(num, den + 1e-12)
out["features"] = f
out["mask3d"], _ = O.mask3d(f, gwbp.scene.make_text_queries(3, d, 0), 1)
render, alpha = O.render_view(sc.means, sc.quats, sc.scales, sc.opacities, f.astype(np.float32), vm[1], K, W, H)
out["render_v1"] = render.astype(np.float32)
out["alpha_v1"] = alpha.astype(np.float32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "small_case.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
