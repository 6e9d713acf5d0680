import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_c1():
    return dict(np.load(os.path.join(GOLDEN, "c1_default.npz")))


@pytest.fixture(scope="session")
def golden_c1b():
    return dict(np.load(os.path.join(GOLDEN, "c1b_h0025.npz")))


@pytest.fixture(scope="session")
def golden_kat():
    return dict(np.load(os.path.join(GOLDEN, "kat_functions.npz")))
