"""Opt-in shim: put `<repo>/shims` on PYTHONPATH and the reference scripts' `from gsplat import
rasterization` (backproject.py:7, utils.py:5, segment.py:9) resolves to the B200 engine."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gwbp import rasterization  # noqa: E402,F401
