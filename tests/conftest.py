import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gwbp():
    import gwbp as pkg  # alias of 3dgs-gradient-backprojection_b200/

    if not os.path.exists(pkg._lib.LIB_PATH):
        # a fresh checkout (built artefacts are git-ignored): compile the extension for sm_100a first -- nvcc
        # cross-compiles without a GPU.  The tests never run against anything but lib/libgwbp.so.
        import __graft_entry__

        __graft_entry__.build()
    return pkg


@pytest.fixture(scope="session")
def coracle():
    """The plain-C oracle, built on demand (gcc)."""
    from oracle import c_oracle

    c_oracle.build()
    return c_oracle


@pytest.fixture(scope="session")
def noracle():
    from oracle import gsplat_oracle

    return gsplat_oracle
