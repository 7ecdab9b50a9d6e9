import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _single_blas_thread():
    """The oracle's matrices are tiny (39 x 39, 39 x 2000): a multi-threaded BLAS pool makes its loop 50-100x SLOWER
    (oversubscription on small gemms).  One thread for the whole test session."""
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1):
            yield
    except ImportError:
        yield


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel(a, b):
    """Norm-wise relative error ||a-b||_2 / ||b||_2 (x is sparse: never elementwise)."""
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="session")
def build_lib():
    import __graft_entry__
    __graft_entry__.build()
    return True


@pytest.fixture(scope="session")
def ir_basis(build_lib):
    from admmsolver_b200 import problems
    return problems.ir_basis()
