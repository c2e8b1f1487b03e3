import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    class G:
        def __getitem__(self, name):
            return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return G()


@pytest.fixture(scope="session")
def fe():
    """Process-wide engine on cuda:0 (GPU tests only); fails loudly if the library is missing."""
    from radarslampy_b200 import _ffi
    eng = _ffi.RadarFE(device=0)
    yield eng
    eng.close()
