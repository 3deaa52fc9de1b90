import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "golden.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O._lib()
    return O


def relerr(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)
