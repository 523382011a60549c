"""The two decisions of a solve_pce step without their divisions (csrc/rb_ion.cuh; reference
loop ion_solver.f90:497-528): `ne > ne_old` for `ratio_ne > 1`, and the error test decided
from reciprocal-based quotients where that is certain.  Both must coincide with what the IEEE
divisions give, for every input -- checked here on 3e6 random triples concentrated around the
thresholds (ratios within 1e-16 .. 1e-3 of 1 and of 1 +- tol), incl. zeros and huge values."""
import ctypes as C

import numpy as np
import pytest

from rabacus_b200 import _lib

pytestmark = pytest.mark.gpu


def _run(a, b, tol):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    n = a.shape[0]
    out = np.zeros(n, dtype=np.int32)
    DP = C.POINTER(C.c_double)
    _lib.check(_lib.lib().rbi_diag_ratio_err(a.ctypes.data_as(DP), b.ctypes.data_as(DP), C.c_int(n),
                                             C.c_double(tol), out.ctypes.data_as(C.POINTER(C.c_int))))
    return out


@pytest.mark.parametrize("tol", [1e-6, 1e-8, 1e-4, 1e-14, 0.3])
def test_decisions_equal_the_divisions(tol):
    rng = np.random.default_rng(3)
    n = 600000
    b = 10.0 ** rng.uniform(-12, 0, (n, 3))
    # the largest ratio sits at 1 + s (1 +- tol) (1 + eps): s, eps log-uniform in magnitude
    eps = rng.choice([-1.0, 1.0], n) * 10.0 ** rng.uniform(-17, -3, n)
    centre = rng.choice([1.0, 1.0 + tol, 1.0 - tol], n)
    top = centre * (1.0 + eps)
    others = rng.uniform(0.0, 1.0, (n, 2)) * top[:, None]
    ratios = np.concatenate([top[:, None], others], axis=1)
    perm = rng.permuted(np.tile(np.arange(3), (n, 1)), axis=1)
    ratios = np.take_along_axis(ratios, perm, axis=1)
    a = ratios * b
    # a few exact neighbours: a one ulp above / below b
    a[:1000, 0] = np.nextafter(b[:1000, 0], np.inf)
    a[1000:2000, 0] = np.nextafter(b[1000:2000, 0], 0.0)
    a[2000:3000, 0] = b[2000:3000, 0]
    out = _run(a, b, tol)
    assert np.array_equal(out & 1, (out >> 1) & 1), "certified error test != division"
    assert np.array_equal((out >> 2) & 1, (out >> 3) & 1), "a > b != a/b > 1"
    assert np.array_equal((out >> 4) & 1, (out >> 5) & 1), "a < b != a/b < 1"
    assert 0.05 < (out & 1).mean() < 0.95            # both outcomes occur


def test_decisions_special_values():
    a = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1e-310, 1.0], [1e300, 1.0, 1.0],
                  [np.nan, 1.0, 1.0], [1.0, 1.0, 1.0], [5e-324, 1.0, 1.0]])
    b = np.array([[0.0, 1.0, 1.0], [0.0, 1.0, 1.0], [1.0, 1e-310, 1.0], [1e-300, 1.0, 1.0],
                  [1.0, 1.0, 1.0], [np.inf, 1.0, 1.0], [5e-324, 1.0, 1.0]])
    out = _run(a, b, 1e-6)
    assert np.array_equal(out & 1, (out >> 1) & 1)
    assert np.array_equal((out >> 2) & 1, (out >> 3) & 1)
    assert np.array_equal((out >> 4) & 1, (out >> 5) & 1)
