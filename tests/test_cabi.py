"""The C-ABI library: loads, exports every symbol the header declares, and has no
CPU fallback (compute entry points fail loudly without a CUDA device).  The
host-side helpers it exports are checked against the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from rabacus_b200 import _lib


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "rabacus_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rb_[A-Za-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == names        # the Python loader's list is the header's list


def test_internal_symbols_are_not_in_the_public_header():
    """Measurement / diagnostic entry points live in csrc/rb_internal.h (rbi_*), exported by
    the library but not declared by the drop-in boundary."""
    L = _lib.lib()
    txt = open(os.path.join(ROOT, "rabacus_b200", "csrc", "rb_internal.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = sorted(set(re.findall(r"\b(rbi_[A-Za-z0-9_]+)\s*\(", txt)))
    assert names == sorted(_lib.INTERNAL)
    assert not [n for n in names if not hasattr(L, n)]
    pub = open(os.path.join(ROOT, "include", "rabacus_b200.h")).read()
    assert "rbi_" not in pub and "rb_diag" not in pub and "rb_microbench" not in pub


def test_every_entry_cites_the_reference_interface():
    txt = open(os.path.join(ROOT, "include", "rabacus_b200.h")).read()
    for ref in ("sphere_bgnd.py:216-244", "sphere_stromgren.py:215-241", "slab_bgnd.py:198-220",
                "slab_plane.py:178-200", "sphere_base.py:182-193", "ion_solver.py:150-159",
                "chem_cool_rates.py:104-106", "legendre_polynomial.f90:815-864"):
        assert ref in txt, ref


def test_version_and_strerror():
    L = _lib.lib()
    assert b"sm_100a" in L.rb_version()
    assert L.rb_strerror(0) == b"ok"
    assert b"monotonic" in L.rb_strerror(1)
    assert b"not supported" in L.rb_strerror(9)


@pytest.mark.skipif(_lib.device_count() > 0, reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    from rabacus_b200 import RabacusB200Error, rabacus_fc as fc
    with pytest.raises(RabacusB200Error) as e:
        fc.get_kchem_v(np.array([1e4, 2e4]), 1.0, 1.0, 1.0, 1, 2)
    assert e.value.code == 10 and "no CPU fallback" in str(e.value)
    Nl, Nnu = 8, 4
    with pytest.raises(RabacusB200Error) as e:
        fc.sphere_bgnd.sphere_bgnd_solve(
            np.linspace(0, 1e22, Nl + 1), np.full(Nl, 1e-3), np.full(Nl, 1e-4), np.full(Nl, 1e4),
            4, np.array([13.6, 20., 30., 60.]), np.ones(4) * 1e-22, 1, 1.0, 1, 1, 0, 0, 1., 1., 1.,
            0.0, 0.0, 1e-4, Nl, Nnu)
    assert e.value.code == 10
    with pytest.raises(RabacusB200Error):
        fc.solve_pce_s(1e-3, 1e-4, 1e-13, 1e-13, 1e-13, 1e-20, 1e-20, 1e-20, 1e-12, 1e-12, 1e-14,
                       1e-8)


def test_argument_checks_happen_before_the_device_is_touched():
    from rabacus_b200 import RabacusB200Error, rabacus_fc as fc
    Nl = 4
    args = [np.linspace(0, 1e22, Nl + 1), np.full(Nl, 1e-3), np.full(Nl, 1e-4), np.full(Nl, 1e4),
            4, np.array([13.6, 20.]), np.ones(2) * 1e-22, 3, 1.0, 1, 1, 0, 0, 1., 1., 1., 0.0, 0.0,
            1e-4, Nl, 2]
    args[9] = 2                                   # photo fits other than Verner 96 do not exist
    with pytest.raises(RabacusB200Error) as e:
        fc.sphere_bgnd.sphere_bgnd_solve(*args)
    assert e.value.code == 9
    args[9] = 1
    args[7] = 5                                   # rec_meth outside {fixed, outward, radial, isotropic}
    with pytest.raises(RabacusB200Error) as e:
        fc.sphere_bgnd.sphere_bgnd_solve(*args)
    assert e.value.code == 8
    args[7] = 1
    args[3] = np.full(Nl, 1e4, dtype=np.float32)  # f2py intent(inout) needs float64
    with pytest.raises(TypeError):
        fc.sphere_bgnd.sphere_bgnd_solve(*args)


def test_host_helpers_match_oracle(orc):
    from rabacus_b200 import rabacus_fc as fc
    E = 13.6 * 10 ** np.linspace(0, np.log10(400.0), 300)
    for X in range(3):
        assert np.array_equal(fc.return_sigma(X, E), orc.v96_sigma(X, E))
        assert fc.return_E_th(X) == orc.v96_Eth(X)
    for n in (1, 2, 7, 32, 256):
        x, w = fc.p_quadrature_rule(n)
        xo, wo = orc.p_quadrature_rule(n)
        assert np.array_equal(x, xo) and np.array_equal(w, wo)   # same Golub-Welsch / implicit QL


def test_struct_layouts_match_header():
    # rb_info: 2 ints, 8 doubles, 1 long long ; rb_plan_desc: 6 ints, 2 doubles, 3 ints
    # (+ padding), 3 doubles, 1 int (+ padding)
    assert C.sizeof(_lib.rb_info) == 8 + 8 * 8 + 8
    assert C.sizeof(_lib.rb_plan_desc) == 24 + 16 + 16 + 24 + 8


def test_photon_energies_match_the_host_producer():
    """rb_photon_energies = rad_src.photon_energies (rad_src/source.py:178-259): numpy's
    linspace / power / argmin statement by statement; pinned to the reference's 128-point
    fixture through tests/test_oracle_pins.py's check of rad_src."""
    from rabacus_b200 import rad_src
    L = _lib.lib()
    for Nnu in (8, 64, 128, 512):
        for q_max in (10.0, 400.0):
            ref, idx = rad_src.photon_energies(1.0, q_max, Nnu)
            E = np.empty(Nnu)
            assert L.rb_photon_energies(C.c_double(1.0), C.c_double(q_max), C.c_int(Nnu),
                                        E.ctypes.data_as(C.POINTER(C.c_double))) == 0
            assert np.max(np.abs(E / ref - 1.0)) <= 4.5e-16           # libm pow vs numpy's: <= 2 ulp
            for name, Eth in (("H1", 13.60), ("He1", 24.59), ("He2", 54.42)):
                assert E[idx[name]] == ref[idx[name]]                   # thresholds snapped exactly
                assert abs(E[idx[name]] - Eth) < 1e-13
    E = np.empty(16)
    for bad in ((2.0, 400.0, 16), (1.0, 3.0, 16), (1.0, 400.0, 1)):     # the reference raises here
        assert L.rb_photon_energies(C.c_double(bad[0]), C.c_double(bad[1]), C.c_int(bad[2]),
                                    E.ctypes.data_as(C.POINTER(C.c_double))) == 8
