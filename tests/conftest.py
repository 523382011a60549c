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


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure): oracle/librb_oracle.so via ctypes."""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def spec128():
    d = np.load(os.path.join(GOLDEN, "hm12_spectrum_128.npz"))
    return d["E_eV"].copy(), d["Inu"].copy()


def relerr(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), floor)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.abs(a - b) / den
    r = np.where((a == b), 0.0, r)
    return float(np.nanmax(r)) if r.size else 0.0
