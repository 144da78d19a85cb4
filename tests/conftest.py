import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def pf():
    import pffft_b200
    return pffft_b200


@pytest.fixture(scope="session")
def ref():
    """the unmodified reference library built into oracle/_ref (checker only)"""
    from oracle import ref as R
    if not R.have_ref():
        pytest.skip("oracle/_ref/libpffft_ref.so not built")
    return R.ref()


@pytest.fixture(scope="session")
def R():
    from oracle import ref as R
    return R


def uniform(rng, n, dtype=np.float32):
    """uniform(-1,1) like the reference validator's frand()*2-1 (benchmarks/bench_pffft.c:318)"""
    return (rng.random(n) * 2 - 1).astype(dtype)
