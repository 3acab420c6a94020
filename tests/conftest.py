import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
SCRIPTS = os.path.join(ROOT, "scripts")
if SCRIPTS not in sys.path:
    sys.path.append(SCRIPTS)  # bench_cv / bench_ss: generators of the BASELINE.json config 4 / 5 inputs


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: full-size BASELINE.json configuration (tens of seconds)")


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


def _load_golden(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.fail(f"{path} missing: run tests/golden/make_golden.py in the build container")
    return np.load(path)


@pytest.fixture(scope="session")
def golden_headline():
    return _load_golden("headline_genes.npz")


@pytest.fixture(scope="session")
def golden_config5():
    return _load_golden("config5_gene.npz")


@pytest.fixture(scope="session")
def golden_config4():
    return _load_golden("config4_summary.npz")


import contextlib


@contextlib.contextmanager
def genotype_storage(mode):
    """Device storage of individual-level genotypes for a test: "dense" = standardised fp64 matrix, "auto" = the
    product default (hard calls 0/1/2 as bit planes), "bytes" = hard calls kept one byte per entry."""
    from sushie_b200 import config

    old = (config.genotype_storage, config.bit_planes)
    config.genotype_storage = "dense" if mode == "dense" else "auto"
    config.bit_planes = mode != "bytes"
    try:
        yield mode
    finally:
        config.genotype_storage, config.bit_planes = old


STORAGE_MODES = ["dense", "auto", "bytes"]
