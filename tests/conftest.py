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
def fish_fixture():
    return dict(np.load(os.path.join(GOLDEN, "fish_fixture.npz")))


@pytest.fixture(scope="session")
def fish_oracle():
    return dict(np.load(os.path.join(GOLDEN, "fish_oracle_cv2_4_13.npz")))


@pytest.fixture(scope="session")
def taps_default():
    return dict(np.load(os.path.join(GOLDEN, "taps_default.npz")))
