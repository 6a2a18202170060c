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


@pytest.fixture(scope="session", autouse=True)
def _built():
    """CPU and GPU suites both need the in-tree artefacts (cheap when up to date)."""
    from pyminiweather_b200 import _lib
    from oracle import c_oracle
    if _lib.needs_build() and os.path.exists("/usr/local/cuda/bin/nvcc"):
        _lib.build_library()
    c_oracle.build()


def golden(name):
    return np.load(os.path.join(GOLDEN, name))
