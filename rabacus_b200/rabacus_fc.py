"""Drop-in stand-in for the reference's f2py extension module ``rabacus_fc``.

The reference's wrapper classes do ``import rabacus_fc`` and call e.g.
``rabacus_fc.sphere_bgnd.sphere_bgnd_solve(...)`` (``rabacus/f2py/sphere_bgnd.py:223``).
This module exposes sub-namespaces with the same names, the same positional
signatures and the same tuple returns, backed by the CUDA C ABI
(``include/rabacus_b200.h``).  Like the f2py ``intent(inout)`` arguments, the
``T`` array (and for the spheres ``Edges, nH, nHe``) must be contiguous
``float64`` ndarrays; ``T`` is updated in place.

Errors are Python exceptions (:class:`rabacus_b200._lib.RabacusB200Error`)
instead of the reference's process-killing Fortran ``stop``.
"""
import ctypes as C
import types

import numpy as np

from . import _lib
from ._lib import check, lib, rb_info

_DP = C.POINTER(C.c_double)


def _inout(a, name):
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous
            and a.flags.writeable):
        raise TypeError("%s must be a writeable contiguous float64 ndarray "
                        "(f2py intent(inout))" % name)
    return a.ctypes.data_as(_DP)


def _in(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_DP)


def _bc(a, n):
    a = np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,)))
    return a, a.ctypes.data_as(_DP)


def _outs(n, count):
    arrs = [np.zeros(n) for _ in range(count)]
    return arrs, [a.ctypes.data_as(_DP) for a in arrs]


# info of the most recent geometric solve (iterations, timings): the reference
# documents an ``itr`` attribute (sphere_bgnd.py:118) but never returns it.
last_info = {}


def _record(info):
    last_info.clear()
    last_info.update(info.as_dict())


def _sphere_solve(fn, Edges, nH, nHe, T, Nmu, E_eV, shape, i_rec_meth, fixed_fcA,
                  i_photo_fit, i_rate_fit, i_find_Teq, i_thin, em_H1_fac, em_He1_fac,
                  em_He2_fac, z, Hz, tol, Nl, Nnu):
    E, pE = _in(E_eV)
    S, pS = _in(shape)
    if E.size != Nnu or S.size != Nnu or np.size(nH) != Nl or np.size(Edges) != Nl + 1:
        raise ValueError("array sizes do not match Nl / Nnu")
    outs, pouts = _outs(Nl, 17)
    info = rb_info()
    rc = fn(_inout(Edges, "Edges"), _inout(nH, "nH"), _inout(nHe, "nHe"), _inout(T, "T"),
            C.c_int(int(Nmu)), pE, pS, C.c_int(int(i_rec_meth)), C.c_double(float(fixed_fcA)),
            C.c_int(int(i_photo_fit)), C.c_int(int(i_rate_fit)), C.c_int(int(i_find_Teq)),
            C.c_int(int(i_thin)), C.c_double(float(em_H1_fac)), C.c_double(float(em_He1_fac)),
            C.c_double(float(em_He2_fac)), C.c_double(float(z)), C.c_double(float(Hz)),
            C.c_double(float(tol)), C.c_int(int(Nl)), C.c_int(int(Nnu)), *pouts,
            C.byref(info))
    check(rc)
    _record(info)
    return tuple(outs)


def sphere_bgnd_solve(Edges, nH, nHe, T, Nmu, E_eV, shape, i_rec_meth, fixed_fcA,
                      i_photo_fit, i_rate_fit, i_find_Teq, i_thin, em_H1_fac, em_He1_fac,
                      em_He2_fac, z, Hz, tol, Nl, Nnu):
    """``rabacus_fc.sphere_bgnd.sphere_bgnd_solve`` (sphere_bgnd.py:218-244;
    sphere_bgnd.f90:107-155) -> 17-tuple (xH1..xHe3, 6 ion/heat src+rec pairs)."""
    return _sphere_solve(lib().rb_sphere_bgnd_solve, Edges, nH, nHe, T, Nmu, E_eV, shape,
                         i_rec_meth, fixed_fcA, i_photo_fit, i_rate_fit, i_find_Teq, i_thin,
                         em_H1_fac, em_He1_fac, em_He2_fac, z, Hz, tol, Nl, Nnu)


def sphere_stromgren_solve(Edges, nH, nHe, T, Nmu, E_eV, shape, i_rec_meth, fixed_fcA,
                           i_photo_fit, i_rate_fit, i_find_Teq, i_thin, em_H1_fac,
                           em_He1_fac, em_He2_fac, z, Hz, tol, Nl, Nnu):
    """``rabacus_fc.sphere_stromgren.sphere_stromgren_solve``
    (sphere_stromgren.py:215-241; sphere_stromgren.f90:111-158)."""
    return _sphere_solve(lib().rb_sphere_stromgren_solve, Edges, nH, nHe, T, Nmu, E_eV,
                         shape, i_rec_meth, fixed_fcA, i_photo_fit, i_rate_fit, i_find_Teq,
                         i_thin, em_H1_fac, em_He1_fac, em_He2_fac, z, Hz, tol, Nl, Nnu)


def slab_bgnd_solve(Edges, nH, nHe, T, E_eV, shape, i_rec_meth, fixed_fcA, i_photo_fit,
                    i_rate_fit, i_find_Teq, i_thin, z, Hz, tol, Nl, Nnu):
    """``rabacus_fc.slab_bgnd.slab_bgnd_solve`` (slab_bgnd.py:198-220;
    slab_bgnd.f90:97-129).  Only ``T`` is intent(inout)."""
    Ed, pEd = _in(Edges)
    H, pH = _in(nH)
    He, pHe = _in(nHe)
    E, pE = _in(E_eV)
    S, pS = _in(shape)
    outs, pouts = _outs(Nl, 17)
    info = rb_info()
    rc = lib().rb_slab_bgnd_solve(
        pEd, pH, pHe, _inout(T, "T"), pE, pS, C.c_int(int(i_rec_meth)),
        C.c_double(float(fixed_fcA)), C.c_int(int(i_photo_fit)), C.c_int(int(i_rate_fit)),
        C.c_int(int(i_find_Teq)), C.c_int(int(i_thin)), C.c_double(float(z)),
        C.c_double(float(Hz)), C.c_double(float(tol)), C.c_int(int(Nl)), C.c_int(int(Nnu)),
        *pouts, C.byref(info))
    check(rc)
    _record(info)
    return tuple(outs)


def slab_plane_solve(Edges, nH, nHe, T, E_eV, shape, i_2side, i_rec_meth, fixed_fcA,
                     i_photo_fit, i_rate_fit, i_find_Teq, i_thin, z, Hz, tol, Nl, Nnu):
    """``rabacus_fc.slab_plane.slab_plane_solve`` (slab_plane.py:178-200,419-441;
    slab_plane.f90:105-137)."""
    Ed, pEd = _in(Edges)
    H, pH = _in(nH)
    He, pHe = _in(nHe)
    E, pE = _in(E_eV)
    S, pS = _in(shape)
    outs, pouts = _outs(Nl, 17)
    info = rb_info()
    rc = lib().rb_slab_plane_solve(
        pEd, pH, pHe, _inout(T, "T"), pE, pS, C.c_int(int(i_2side)), C.c_int(int(i_rec_meth)),
        C.c_double(float(fixed_fcA)), C.c_int(int(i_photo_fit)), C.c_int(int(i_rate_fit)),
        C.c_int(int(i_find_Teq)), C.c_int(int(i_thin)), C.c_double(float(z)),
        C.c_double(float(Hz)), C.c_double(float(tol)), C.c_int(int(Nl)), C.c_int(int(Nnu)),
        *pouts, C.byref(info))
    check(rc)
    _record(info)
    return tuple(outs)


def set_column_vs_impact(xH1, xH2, xHe1, xHe2, xHe3, Edges, nH, nHe, T, Nl):
    """``rabacus_fc.sphere_base.set_column_vs_impact`` (sphere_base.py:182-193;
    sphere_base.f90:1115-1205) -> (NH, NHe, NH1, NHe1, NHe2)."""
    ins = [_in(a) for a in (xH1, xH2, xHe1, xHe2, xHe3, Edges, nH, nHe, T)]
    outs, pouts = _outs(Nl, 5)
    check(lib().rb_set_column_vs_impact(*[p for _, p in ins], C.c_int(int(Nl)), *pouts))
    return tuple(outs)


# ----------------------------------------------------------------- single zone
def _get_kchem(T, fcA_H2, fcA_He2, fcA_He3, irate, nn):
    ins = [_bc(a, nn) for a in (T, fcA_H2, fcA_He2, fcA_He3)]
    outs, pouts = _outs(nn, 6)
    check(lib().rb_get_kchem(*[p for _, p in ins], C.c_int(int(irate)), C.c_int(int(nn)),
                             *pouts))
    return outs


def _get_kcool(T, fcA_H2, fcA_He2, fcA_He3, irate, nn):
    ins = [_bc(a, nn) for a in (T, fcA_H2, fcA_He2, fcA_He3)]
    outs, pouts = _outs(nn, 10)
    check(lib().rb_get_kcool(*[p for _, p in ins], C.c_int(int(irate)), C.c_int(int(nn)),
                             *pouts))
    return outs


def _scalarise(outs):
    return tuple(float(o[0]) for o in outs)


def get_kchem_s(T, fcA_H2, fcA_He2, fcA_He3, irate, nn=1):
    """chem_cool_rates.f90:45-88"""
    return _scalarise(_get_kchem(T, fcA_H2, fcA_He2, fcA_He3, irate, 1))


def get_kchem_v(T, fcA_H2, fcA_He2, fcA_He3, irate, nn):
    """chem_cool_rates.f90:91-140"""
    return tuple(_get_kchem(T, fcA_H2, fcA_He2, fcA_He3, irate, nn))


def get_kcool_s(T, fcA_H2, fcA_He2, fcA_He3, irate, nn=1):
    """chem_cool_rates.f90:147-204"""
    return _scalarise(_get_kcool(T, fcA_H2, fcA_He2, fcA_He3, irate, 1))


def get_kcool_v(T, fcA_H2, fcA_He2, fcA_He3, irate, nn):
    """chem_cool_rates.f90:207-263"""
    return tuple(_get_kcool(T, fcA_H2, fcA_He2, fcA_He3, irate, nn))


def _solve_ce(reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, nn):
    ins = [_bc(a, nn) for a in (reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2)]
    outs, pouts = _outs(nn, 5)
    check(lib().rb_solve_ce(*[p for _, p in ins], C.c_int(int(nn)), *pouts))
    return outs


def solve_ce_s(reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, nn=1):
    """ion_solver.f90:339-374"""
    return _scalarise(_solve_ce(reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, 1))


def solve_ce_v(reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, nn):
    """ion_solver.f90:377-412"""
    return tuple(_solve_ce(reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, nn))


def _solve_pce(args, tol, nn, lockstep):
    ins = [_bc(a, nn) for a in args]
    outs, pouts = _outs(nn, 5)
    check(lib().rb_solve_pce(*[p for _, p in ins], C.c_double(float(tol)), C.c_int(int(nn)),
                             C.c_int(int(lockstep)), *pouts))
    return outs


def solve_pce_s(nH, nHe, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, H1i, He1i, He2i, tol,
                nn=1):
    """ion_solver.f90:440-562"""
    return _scalarise(_solve_pce((nH, nHe, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, H1i,
                                  He1i, He2i), tol, 1, 0))


def solve_pce_v(nH, nHe, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, H1i, He1i, He2i, tol, nn):
    """ion_solver.f90:565-696 (lock-step: all zones iterate until all converge)"""
    return tuple(_solve_pce((nH, nHe, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, H1i, He1i,
                             He2i), tol, nn, 1))


def _solve_pcte(args, z, Hz, fcs, irate, tol, nn, lockstep):
    ins = [_bc(a, nn) for a in args]
    f = [_bc(a, nn) for a in fcs]
    outs, pouts = _outs(nn, 6)
    check(lib().rb_solve_pcte(*[p for _, p in ins], C.c_double(float(z)), C.c_double(float(Hz)),
                              *[p for _, p in f], C.c_int(int(irate)), C.c_double(float(tol)),
                              C.c_int(int(nn)), C.c_int(int(lockstep)), *pouts))
    return outs


def solve_pcte_s(nH, nHe, H1i, He1i, He2i, H1h, He1h, He2h, z, Hz, fcA_H2, fcA_He2, fcA_He3,
                 irate, tol, nn=1):
    """ion_solver.f90:729-880"""
    return _scalarise(_solve_pcte((nH, nHe, H1i, He1i, He2i, H1h, He1h, He2h), z, Hz,
                                  (fcA_H2, fcA_He2, fcA_He3), irate, tol, 1, 0))


def solve_pcte_v(nH, nHe, H1i, He1i, He2i, H1h, He1h, He2h, z, Hz, fcA_H2, fcA_He2, fcA_He3,
                 irate, tol, nn):
    """ion_solver.f90:884-1056 (lock-step)"""
    return tuple(_solve_pcte((nH, nHe, H1i, He1i, He2i, H1h, He1h, He2h), z, Hz,
                             (fcA_H2, fcA_He2, fcA_He3), irate, tol, nn, 1))


# host-side helpers (no GPU needed)
def return_sigma(species, E_eV):
    """photo_xsections.f90:100-180 (species 0=H1, 1=He1, 2=He2)"""
    E, pE = _in(np.atleast_1d(E_eV))
    out = np.empty_like(E)
    check(lib().rb_photo_sigma(C.c_int(int(species)), pE, C.c_int(E.size),
                               out.ctypes.data_as(_DP)))
    return out


def return_E_th(species):
    """photo_xsections.f90:29-97"""
    return lib().rb_photo_Eth(int(species))


def p_quadrature_rule(n):
    """legendre_polynomial.f90:815-864"""
    x = np.empty(n)
    w = np.empty(n)
    check(lib().rb_gauss_legendre(C.c_int(int(n)), x.ctypes.data_as(_DP),
                                  w.ctypes.data_as(_DP)))
    return x, w


# sub-namespaces named like the Fortran modules f2py exposes
sphere_bgnd = types.SimpleNamespace(sphere_bgnd_solve=sphere_bgnd_solve)
sphere_stromgren = types.SimpleNamespace(sphere_stromgren_solve=sphere_stromgren_solve)
slab_bgnd = types.SimpleNamespace(slab_bgnd_solve=slab_bgnd_solve)
slab_plane = types.SimpleNamespace(slab_plane_solve=slab_plane_solve)
sphere_base = types.SimpleNamespace(set_column_vs_impact=set_column_vs_impact)
ion_solver = types.SimpleNamespace(
    solve_ce_s=solve_ce_s, solve_ce_v=solve_ce_v, solve_pce_s=solve_pce_s,
    solve_pce_v=solve_pce_v, solve_pcte_s=solve_pcte_s, solve_pcte_v=solve_pcte_v)
chem_cool_rates = types.SimpleNamespace(
    get_kchem_s=get_kchem_s, get_kchem_v=get_kchem_v, get_kcool_s=get_kcool_s,
    get_kcool_v=get_kcool_v)
photo_xsections = types.SimpleNamespace(
    return_sigma_H1=lambda E, ifit=1, nn=None: return_sigma(0, E),
    return_sigma_He1=lambda E, ifit=1, nn=None: return_sigma(1, E),
    return_sigma_He2=lambda E, ifit=1, nn=None: return_sigma(2, E),
    return_E_H1_th=lambda ifit=1: return_E_th(0),
    return_E_He1_th=lambda ifit=1: return_E_th(1),
    return_E_He2_th=lambda ifit=1: return_E_th(2))
legendre_polynomial = types.SimpleNamespace(p_quadrature_rule=p_quadrature_rule)
