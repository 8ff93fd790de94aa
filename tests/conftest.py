import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


@pytest.fixture(scope="session")
def oracle_f64():
    from oracle.loader import CpuTvl1
    o = CpuTvl1("port", np.float64)
    o.set_threads(1)
    return o


@pytest.fixture(scope="session")
def oracle_f32():
    from oracle.loader import CpuTvl1
    o = CpuTvl1("port", np.float32)
    o.set_threads(1)
    return o
