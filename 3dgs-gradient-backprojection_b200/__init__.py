"""B200-native gradient-weighted feature back-projection for 3D Gaussian splats.

Drop-in for one path of JojiJoseph/3dgs-gradient-backprojection: the gsplat-style
`rasterization(...)` call of backproject.py / backproject_compressed.py / segment.py and the
per-Gaussian feature tensor it produces.  All arithmetic runs in lib/libgwbp.so (hand-written
sm_100a CUDA behind the C ABI of include/gwbp.h); importing the compute API without that library
raises -- there is no CPU fallback.
"""
from . import scene  # noqa: F401  (numpy-only, importable without the CUDA library)
from . import dist  # noqa: F401
from . import splats  # noqa: F401  (torch-only data-contract helpers)
from ._lib import LIB_PATH, KERNEL_AUTO, KERNEL_SIMT, KERNEL_TC  # noqa: F401
from .backproject import BackProjector, create_feature_field, DEN_EPS  # noqa: F401
from .engine import PackedScene, View, cosine_mask, finalize, make_camera, fpack_bytes  # noqa: F401
from .rasterization import rasterization, cache_clear as rasterization_cache_clear  # noqa: F401
from .sh import sh_colors  # noqa: F401
from .segment import (click_mask3d, click_prompt, gaussian_scores, get_mask3d, render_features,  # noqa: F401
                      render_mask_2d)

__all__ = ["rasterization", "BackProjector", "create_feature_field", "PackedScene", "View", "finalize",
           "cosine_mask", "get_mask3d", "click_prompt", "click_mask3d", "render_features", "render_mask_2d", "make_camera", "sh_colors", "scene", "dist", "splats"]
