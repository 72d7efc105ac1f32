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
def golden():
    return np.load(os.path.join(GOLDEN, "node_test_golden.npz"))


@pytest.fixture(scope="session")
def test_inputs():
    return np.load(os.path.join(GOLDEN, "test_inputs.npz"))


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle as O
    O.build()
    O.lib()
    return O


@pytest.fixture(scope="session")
def ctx():
    """A real GPU context (only requested by @pytest.mark.gpu tests)."""
    import homography_js_b200 as hg
    c = hg.Context(0)
    yield c
    c.close()
