import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN_DIR, name))
    return load


@pytest.fixture(scope="session", autouse=True)
def _built_library(request):
    """GPU runs need libmspl_b200.so; build it in-tree if the snapshot came without it (CPU-only sessions that never touch
    the library are not forced to compile)."""
    markexpr = request.config.getoption("-m", default="") or ""
    if "gpu" in markexpr and "not gpu" not in markexpr:
        from mspl_b200 import _lib
        if not os.path.isfile(_lib.LIB_PATH):
            import __graft_entry__
            __graft_entry__.build()
