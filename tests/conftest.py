import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """Make sure the C-ABI library exists (nvcc cross-compiles without a GPU).  On the GPU box
    the in-tree .so travels with the snapshot and is up to date, so this is a no-op there."""
    from delight_b200 import build as b
    if not os.path.exists(b.LIB) or (b._stale() and os.path.exists("/usr/local/cuda/bin/nvcc")):
        try:
            b.build()
        except Exception as e:          # tests that need the library will fail loudly
            print("delight_b200 build failed: %s" % e)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load
