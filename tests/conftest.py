import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def feats():
    return dict(np.load(os.path.join(GOLDEN, "features_small.npz")))


@pytest.fixture(scope="session")
def gpairs():
    return dict(np.load(os.path.join(GOLDEN, "golden_pairs.npz")))


@pytest.fixture(scope="session")
def gsynth():
    return dict(np.load(os.path.join(GOLDEN, "golden_synth.npz")))


@pytest.fixture(scope="session")
def oracle():
    from oracle import cvoracle
    cvoracle.build()
    return cvoracle
