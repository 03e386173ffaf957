import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


@pytest.fixture(scope="session")
def port():
    from oracle import oracle as O
    return O.Port()


@pytest.fixture(scope="session")
def ref():
    """The compiled reference, when its prebuilt .so travelled here or /root/reference exists; else skip."""
    from oracle import oracle as O
    try:
        return O.Ref()
    except Exception as e:  # noqa: BLE001
        pytest.skip("compiled reference unavailable: %s" % e)


@pytest.fixture(scope="session")
def ptv():
    import proxtv_b200
    proxtv_b200.require_device()
    return proxtv_b200


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def jumps(x, tol=0.0):
    x = np.asarray(x).ravel()
    return np.nonzero(np.abs(np.diff(x)) > tol)[0]
