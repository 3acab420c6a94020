import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_ind():
    return np.load(os.path.join(GOLDEN, "example_ind.npz"))


@pytest.fixture(scope="session")
def golden_ss():
    return np.load(os.path.join(GOLDEN, "example_ss.npz"))


@pytest.fixture(scope="session")
def golden_syn():
    return np.load(os.path.join(GOLDEN, "synthetic_small.npz"))


@pytest.fixture(scope="session")
def engine_lib():
    """Build (if needed) and load the CUDA engine; GPU tests must fail, not skip, without it."""
    from sushie_b200 import _capi, build

    build.build()
    return _capi.load_library()
