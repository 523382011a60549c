"""The memo of the rate coefficients over solve_pcte's temperature bisection tree
(csrc/rb_ion.cuh: TeqTab; reference algorithm ion_solver.f90:796-850).

The memo must be invisible: a row is used only when its key IS the probe temperature
(bit for bit) and holds what the solver's own get_kchem / get_kcool return there, so every
find_Teq result has to be bit-identical with the memo on (default), at other depths and
off (RB_TEQ_DEPTH=0) -- for the speculative group solver, the one-thread-per-cell batch
solver, the lock-step thin start and the single-zone API."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import FRACTIONS, compare, run_gpu, run_oracle
from rabacus_b200 import _lib, configs
from rabacus_b200 import rabacus_fc as fc

pytestmark = pytest.mark.gpu

_DP = C.POINTER(C.c_double)


class _env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        for k, v in self.kw.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        _lib.lib().rb_cache_clear()


def _fetch(fcA, row0, n):
    out = np.empty((n, 16))
    depth = C.c_int(0)
    _lib.check(_lib.lib().rbi_teq_table_fetch(C.c_double(fcA), C.c_int(row0), C.c_int(n),
                                              out.ctypes.data_as(_DP), C.byref(depth)))
    return out, depth.value


def test_rows_hold_the_solvers_coefficients():
    """Row key = the bisection's temperature; row content = rb_get_kchem / rb_get_kcool there."""
    for fcA in (1.0, 0.0, 0.37):
        top, depth = _fetch(fcA, 0, 64)
        last, _ = _fetch(fcA, (1 << depth) - 3, 4)
        rows = np.vstack([top, last])
        T = rows[:, 15].copy()
        assert T[0] == 7.5e3 and T[-1] == 1.0e5
        # heap order: node 1 is the geometric mean, 2 / 3 the mid-points of its halves
        assert abs(T[1] / np.sqrt(7.5e3 * 1.0e5) - 1.0) < 1e-15
        assert T[2] == (7.5e3 + T[1]) * 0.5 and T[3] == (T[1] + 1.0e5) * 0.5
        assert T[6] == (T[1] + T[3]) * 0.5 and T[5] == (T[2] + T[1]) * 0.5
        kc = fc.get_kchem_v(T, fcA, fcA, fcA, 1, T.size)
        kl = fc.get_kcool_v(T, fcA, fcA, fcA, 1, T.size)
        for q in range(6):
            assert np.array_equal(rows[:, q], kc[q]), ("kchem", q)
        # rb_kcool_t order: recH2 recHe2 recHe3 cicH1 cicHe1 cicHe2 cecH1 cecHe2 bremss compton;
        # the f2py tuple (chem_cool_rates.py:332-334) ends (..., compton, bremss)
        names = _kcool_order(kl, rows)
        assert names, "kcool rows do not match rb_get_kcool"


def _kcool_order(kl, rows):
    kl = [np.asarray(a) for a in kl]
    ok = all(np.array_equal(rows[:, 6 + q], kl[q]) for q in range(8))
    brem = [a for a in kl[8:] if np.array_equal(rows[:, 14], a)]
    return ok and len(brem) == 1


def _solve(cfg, **env):
    with _env(**env):
        return run_gpu(cfg)


@pytest.mark.parametrize("make", [
    lambda: configs.c2_slab_plane(Nl=128, Nnu=64, find_Teq=True),
    lambda: configs.c3_slab_bgnd(Nl=128, Nnu=64, find_Teq=True),
    lambda: configs.c4_sphere_bgnd(Nl=160, Nnu=64, Nmu=8, find_Teq=True, nH0=1.0),
])
def test_memo_is_invisible(orc, make):
    cfg = make()
    ref = _solve(cfg, RB_TEQ_DEPTH=0)
    for env in (dict(RB_TEQ_DEPTH=None), dict(RB_TEQ_DEPTH=3), dict(RB_TEQ_DEPTH=12),
                dict(RB_TEQ_DEPTH=None, RB_SPEC_G=1), dict(RB_TEQ_DEPTH=None, RB_SPEC_G=4)):
        got = _solve(cfg, **env)
        assert got["iters"] == ref["iters"], env
        for k in FRACTIONS + ("T", "H1i_src", "H1h_src", "He2h_src"):
            assert np.array_equal(got[k], ref[k], equal_nan=True), (env, k)
    # and the default build against the oracle (T and fractions through the solver check)
    got = _solve(cfg, RB_TEQ_DEPTH=None)
    worst, bad = compare(got, run_oracle(orc, cfg))
    assert not bad, bad


def test_thin_start_lockstep_memo():
    cfg = configs.c4_sphere_bgnd(Nl=700, Nnu=64, Nmu=4, find_Teq=True, nH0=0.3)
    with _env(RB_TEQ_DEPTH=0):
        ref = run_gpu(cfg, thin=True)
    with _env(RB_TEQ_DEPTH=None):
        got = run_gpu(cfg, thin=True)
    for k in FRACTIONS + ("T",):
        assert np.array_equal(got[k], ref[k]), k


@pytest.mark.parametrize("lockstep", [False, True])
def test_single_zone_api_memo(lockstep):
    n = 4096 if not lockstep else 1500
    rng = np.random.default_rng(7)
    nH = 10.0 ** rng.uniform(-6.0, 1.0, n)
    nHe = nH * 0.079
    # HM12-like photo-rates scaled by a common attenuation per zone (heating per ionisation
    # stays physical, so that the equilibrium is bracketed by [7.5e3, 1e5] K or at the floor)
    att = 10.0 ** rng.uniform(-4.0, 0.5, n)
    ion = [8.0e-13 * att, 5.0e-13 * att, 6.0e-15 * att]
    heat = [5.0e-24 * att, 6.0e-24 * att, 2.0e-25 * att]
    f = fc.solve_pcte_v if lockstep else None

    def call(fcA):
        if lockstep:
            return f(nH, nHe, *ion, *heat, 3.0, 0.0, fcA, fcA, fcA, 1, 1e-6, n)
        from rabacus_b200.rabacus_fc import _solve_pcte
        return tuple(_solve_pcte((nH, nHe, *ion, *heat), 3.0, 0.0, (fcA, fcA, fcA), 1, 1e-6, n, 0))

    for fcA in (1.0, 0.25):
        with _env(RB_TEQ_DEPTH=0):
            ref = call(fcA)
        with _env(RB_TEQ_DEPTH=None):
            got = call(fcA)
        for a, b in zip(got, ref):
            assert np.array_equal(a, b)
        Tout = np.asarray(got[5])
        assert (Tout > 7.5e3).any() and (Tout <= 1.0e5).all()
