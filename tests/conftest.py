import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_py():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "py_golden.npz"))


@pytest.fixture(scope="session")
def golden_ref():
    """Outputs of the reference's own kernels (oracle/_ref) recorded on a B200 by
    tests/golden/make_golden_gpu.py."""
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_kernels_golden.npz"))
