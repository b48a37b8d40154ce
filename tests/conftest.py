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
    config.addinivalue_line("markers", "slow: multi-second CPU oracle run")


@pytest.fixture(scope="session")
def golden_detector():
    return np.load(os.path.join(GOLDEN, "detector_xl_seed0.npz"))


@pytest.fixture(scope="session")
def golden_transformer():
    return np.load(os.path.join(GOLDEN, "transformer_seed0.npz"))


@pytest.fixture(scope="session")
def test1_tile():
    return np.load(os.path.join(GOLDEN, "test1_tile.npz"))["tile"]


@pytest.fixture(scope="session")
def detector_sd():
    from findtextcenternet_b200 import synthetic
    return synthetic.detector_state_dict(0)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
