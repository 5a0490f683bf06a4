"""Importable alias of the package directory `3dgs-gradient-backprojection_b200/` (whose name,
mandated by the repo layout, is not a Python identifier):  `import gwbp` == that package."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("3dgs-gradient-backprojection_b200")
sys.modules[__name__] = _pkg
