import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    A = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
    return A, (z["rhs"] if "rhs" in z else None), (z["sol"] if "sol" in z else None)


@pytest.fixture
def fixture_loader():
    return load_fixture
