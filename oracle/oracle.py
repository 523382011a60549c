"""ctypes binding of the CPU oracle (``oracle/librb_oracle.so``).

TEST INFRASTRUCTURE ONLY.  Importable from ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline legs; the product package ``rabacus_b200`` never
imports this module.  See ``oracle/rb_oracle.h`` for what each entry restates
(reference file:line) and for the parity-pinning statement.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librb_oracle.so")

ORDER_GS, ORDER_JACOBI = 0, 1
FLAVOUR_HOISTED, FLAVOUR_FAITHFUL = 0, 1

ERRORS = {
    1: "Edges not monotonic",
    2: "MAX_ITER in outer sweep loop",
    3: "solve_pce: iter > MAX_ITER",
    4: "solve_pcte: iter > MAX_ITER",
    5: "Teq not bracketed",
    6: "NaN in solve_ce",
    7: "ray geometry undefined (hollow sphere / segment out of range)",
    8: "bad argument / unsupported option",
}


class OracleError(RuntimeError):
    pass


def build(force=False):
    """Compile the oracle with its Makefile (gcc, -ffp-contract=off, OpenMP)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", _HERE, "librb_oracle.so"])
    return _LIB_PATH


_O3_PATH = os.path.join(_HERE, "librb_oracle_o3.so")


def _host_tag():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def build_timing():
    """The -O3 -mtune=native build that bench.py times (oracle/Makefile).  -mtune=native is
    a property of the host that compiles: the library is rebuilt when the CPU model differs
    from the one recorded beside it (the .so travels from the build container to the GPU
    box).  Returns (path, note)."""
    tag_file = _O3_PATH + ".host"
    tag = _host_tag()
    have = None
    if os.path.exists(tag_file):
        with open(tag_file) as f:
            have = f.read().strip()
    note = "gcc -O3 -mtune=native -fopenmp -ffp-contract=off, built on this host (%s)" % tag
    if not os.path.exists(_O3_PATH) or have != tag:
        try:
            subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "librb_oracle_o3.so"])
            with open(tag_file, "w") as f:
                f.write(tag)
        except Exception as e:  # no compiler on this host: keep the library that travelled
            if not os.path.exists(_O3_PATH):
                raise
            note = ("gcc -O3 -mtune=native -fopenmp, built on another host (%s): rebuild here "
                    "failed (%r)" % (have, e))
    return _O3_PATH, note


_lib = None
_libs = {}
_timing = False


def _load(path):
    L = C.CDLL(path)
    L.orc_e1xa.restype = C.c_double
    L.orc_e1xa.argtypes = [C.c_double]
    L.orc_e2xa.restype = C.c_double
    L.orc_e2xa.argtypes = [C.c_double]
    L.orc_v96_Eth.restype = C.c_double
    L.orc_v96_Eth.argtypes = [C.c_int]
    L.orc_integrate_trap.restype = C.c_double
    L.orc_ds_segment.restype = C.c_double
    L.orc_last_sweep_seconds.restype = C.c_double
    L.orc_set_num_threads.argtypes = [C.c_int]
    L.orc_set_num_threads.restype = None
    L.orc_set_coefficient_hook.argtypes = [C.c_void_p]
    L.orc_set_coefficient_hook.restype = None
    return L


def lib():
    """The parity build, or the timing build inside ``with timing_build():``."""
    if _timing:
        if "o3" not in _libs:
            path, note = build_timing()
            _libs["o3"] = _load(path)
            _libs["o3_note"] = note
        return _libs["o3"]
    if "o2" not in _libs:
        build()
        _libs["o2"] = _load(_LIB_PATH)
    return _libs["o2"]


class timing_build(object):
    """``with oracle.timing_build() as note:`` -- route every call of this module through
    librb_oracle_o3.so (bench.py's CPU legs; tests compare the two builds bit for bit)."""

    def __enter__(self):
        global _timing
        self._prev = _timing
        _timing = True
        lib()
        return _libs["o3_note"]

    def __exit__(self, *exc):
        global _timing
        _timing = self._prev
        return False


class _Opts(C.Structure):
    _fields_ = [("order", C.c_int), ("flavour", C.c_int), ("max_sweeps", C.c_int)]


def _d(a):
    """contiguous float64 copy/view + pointer"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _p(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _check(rc):
    if rc != 0:
        raise OracleError(ERRORS.get(rc, "error %d" % rc))


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    """OpenMP team size of the active build (overrides OMP_NUM_THREADS; torchrun sets 1)."""
    lib().orc_set_num_threads(int(n))
    return lib().orc_num_threads()


class coefficient_hook(object):
    """``with oracle.coefficient_hook(fnptr):`` -- the oracle's solvers take the rate /
    cooling coefficients of a temperature from the C function ``fnptr`` (signature
    ``void(double T, double fA, double fB, double fC, double *out16)``) instead of its own
    fits.  The GPU parity tests install the host build of the CUDA library's evaluation
    (same bits as the device), which isolates the solvers from the <= 1-ulp difference
    between two pow() implementations.  Applies to the active build only."""

    def __init__(self, fnptr):
        self.fn = C.cast(fnptr, C.c_void_p)

    def __enter__(self):
        lib().orc_set_coefficient_hook(self.fn)
        return self

    def __exit__(self, *exc):
        lib().orc_set_coefficient_hook(C.c_void_p(0))
        return False


def last_sweep_seconds():
    """Wall time of the sweep loop of the last sphere_bgnd_solve (thin start excluded)."""
    return float(lib().orc_last_sweep_seconds())


# ---------------------------------------------------------------- atomic data
def v96_Eth(species):
    return lib().orc_v96_Eth(int(species))


def v96_sigma(species, E_eV):
    E, pE = _d(np.atleast_1d(E_eV))
    out = np.empty_like(E)
    lib().orc_v96_sigma(C.c_int(species), pE, C.c_int(E.size), _p(out))
    return out


def hg97_all(T):
    T = np.atleast_1d(np.asarray(T, dtype=np.float64))
    out = np.empty((T.size, 23))
    row = np.empty(23)
    for i, t in enumerate(T):
        lib().orc_hg97_all(C.c_double(t), _p(row))
        out[i] = row
    return out


def _bc(n, *arrs):
    return [np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,)))
            for a in arrs]


def get_kchem(T, fcA_H2, fcA_He2, fcA_He3):
    T = np.atleast_1d(np.asarray(T, dtype=np.float64))
    n = T.size
    T, a, b, c = _bc(n, T, fcA_H2, fcA_He2, fcA_He3)
    outs = [np.empty(n) for _ in range(6)]
    lib().orc_get_kchem(_p(T), _p(a), _p(b), _p(c), C.c_int(n), *[_p(o) for o in outs])
    return tuple(outs)


def get_kcool(T, fcA_H2, fcA_He2, fcA_He3):
    T = np.atleast_1d(np.asarray(T, dtype=np.float64))
    n = T.size
    T, a, b, c = _bc(n, T, fcA_H2, fcA_He2, fcA_He3)
    outs = [np.empty(n) for _ in range(10)]
    lib().orc_get_kcool(_p(T), _p(a), _p(b), _p(c), C.c_int(n), *[_p(o) for o in outs])
    return tuple(outs)


def e1xa(x):
    return np.array([lib().orc_e1xa(float(v)) for v in np.atleast_1d(x)])


def e2xa(x):
    return np.array([lib().orc_e2xa(float(v)) for v in np.atleast_1d(x)])


def p_quadrature_rule(n):
    t = np.empty(n)
    w = np.empty(n)
    lib().orc_p_quadrature_rule(C.c_int(n), _p(t), _p(w))
    return t, w


def integrate_trap(x, y):
    x, px = _d(x)
    y, py = _d(y)
    return lib().orc_integrate_trap(px, py, C.c_int(x.size))


# ---------------------------------------------------------------- single zone
def solve_ce(reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2):
    n = np.atleast_1d(reH2).size
    ins = _bc(n, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2)
    outs = [np.empty(n) for _ in range(5)]
    _check(lib().orc_solve_ce(*[_p(a) for a in ins], C.c_int(n), *[_p(o) for o in outs]))
    return tuple(outs)


def solve_pce(nH, nHe, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, H1i, He1i, He2i,
              tol, vector=False):
    n = np.atleast_1d(nH).size
    ins = _bc(n, nH, nHe, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, H1i, He1i, He2i)
    outs = [np.empty(n) for _ in range(5)]
    fn = lib().orc_solve_pce_v if vector else lib().orc_solve_pce_s
    _check(fn(*[_p(a) for a in ins], C.c_double(tol), C.c_int(n), *[_p(o) for o in outs]))
    return tuple(outs)


def solve_pcte(nH, nHe, H1i, He1i, He2i, H1h, He1h, He2h, z, Hz, fcA_H2, fcA_He2,
               fcA_He3, tol, vector=False):
    n = np.atleast_1d(nH).size
    a = _bc(n, nH, nHe, H1i, He1i, He2i, H1h, He1h, He2h)
    f = _bc(n, fcA_H2, fcA_He2, fcA_He3)
    outs = [np.empty(n) for _ in range(6)]
    fn = lib().orc_solve_pcte_v if vector else lib().orc_solve_pcte_s
    _check(fn(*[_p(x) for x in a], C.c_double(z), C.c_double(Hz), *[_p(x) for x in f],
              C.c_double(tol), C.c_int(n), *[_p(o) for o in outs]))
    return tuple(outs)  # xH1,xH2,xHe1,xHe2,xHe3,T


# ---------------------------------------------------------------- geometry
def m_for_ray(Edges, il, mu):
    E, pE = _d(Edges)
    return lib().orc_m_for_ray(pE, C.c_int(il), C.c_double(mu), C.c_int(E.size - 1))


def ds_segment(Edges, il, mu, j):
    E, pE = _d(Edges)
    err = C.c_int(0)
    ds = lib().orc_ds_segment(pE, C.c_int(il), C.c_double(mu), C.c_int(j),
                              C.c_int(E.size - 1), C.byref(err))
    _check(err.value)
    return ds


def sphere_ray_columns(Edges, nH, nHe, xH1, xHe1, xHe2, il, mu):
    E, pE = _d(Edges)
    arrs = [_d(a) for a in (nH, nHe, xH1, xHe1, xHe2)]
    out = [C.c_double(0), C.c_double(0), C.c_double(0)]
    _check(lib().orc_sphere_ray_columns(pE, *[p for _, p in arrs], C.c_int(il),
                                        C.c_double(mu), C.c_int(E.size - 1),
                                        *[C.byref(o) for o in out]))
    return tuple(o.value for o in out)


_OUT_NAMES = ("xH1", "xH2", "xHe1", "xHe2", "xHe3",
              "H1i_src", "He1i_src", "He2i_src", "H1i_rec", "He1i_rec", "He2i_rec",
              "H1h_src", "He1h_src", "He2h_src", "H1h_rec", "He1h_rec", "He2h_rec")


class Result(dict):
    """17 output arrays + T + iters + counters, attribute access."""
    __getattr__ = dict.__getitem__


def _finish(res, outs, T, iters, counters):
    for name, arr in zip(_OUT_NAMES, outs):
        res[name] = arr
    res["T"] = T
    res["iters"] = iters.value
    res["counters"] = dict(zip(("segments", "sqrt", "exp", "cells"),
                               [int(c) for c in counters]))
    return res


def _opts(order, flavour, max_sweeps):
    return _Opts(int(order), int(flavour), int(max_sweeps))


def sphere_bgnd_solve(Edges, nH, nHe, T, Nmu, E_eV, shape, i_rec_meth=1,
                      fixed_fcA=1.0, i_find_Teq=0, i_thin=0, z=0.0, Hz=0.0,
                      tol=1e-4, order=ORDER_GS, flavour=FLAVOUR_HOISTED,
                      max_sweeps=0, em_fac=(1.0, 1.0, 1.0)):
    """Restates sphere_bgnd.f90:107-314.  Inputs are copied (the f2py call
    mutates T in place; here the updated T is returned as ``res.T``)."""
    Edges = np.array(Edges, dtype=np.float64)
    nH = np.array(nH, dtype=np.float64)
    nHe = np.array(nHe, dtype=np.float64)
    T = np.array(T, dtype=np.float64)
    E, pE = _d(E_eV)
    S, pS = _d(shape)
    Nl, Nnu = nH.size, E.size
    outs = [np.zeros(Nl) for _ in range(17)]
    iters = C.c_int(0)
    counters = (C.c_longlong * 4)(0, 0, 0, 0)
    o = _opts(order, flavour, max_sweeps)
    rc = lib().orc_sphere_bgnd_solve(
        _p(Edges), _p(nH), _p(nHe), _p(T), C.c_int(Nmu), pE, pS, C.c_int(i_rec_meth),
        C.c_double(fixed_fcA), C.c_int(1), C.c_int(1), C.c_int(i_find_Teq),
        C.c_int(i_thin), C.c_double(em_fac[0]), C.c_double(em_fac[1]),
        C.c_double(em_fac[2]), C.c_double(z), C.c_double(Hz), C.c_double(tol),
        C.c_int(Nl), C.c_int(Nnu), *[_p(a) for a in outs], C.byref(o),
        C.byref(iters), counters)
    _check(rc)
    return _finish(Result(), outs, T, iters, counters)


def sphere_stromgren_solve(Edges, nH, nHe, T, Nmu, E_eV, shape, i_rec_meth=1,
                           fixed_fcA=1.0, i_find_Teq=0, i_thin=0, z=0.0, Hz=0.0,
                           tol=1e-4, flavour=FLAVOUR_HOISTED, max_sweeps=0,
                           em_fac=(1.0, 1.0, 1.0)):
    """Restates sphere_stromgren.f90:111-320."""
    Edges = np.array(Edges, dtype=np.float64)
    nH = np.array(nH, dtype=np.float64)
    nHe = np.array(nHe, dtype=np.float64)
    T = np.array(T, dtype=np.float64)
    E, pE = _d(E_eV)
    S, pS = _d(shape)
    Nl, Nnu = nH.size, E.size
    outs = [np.zeros(Nl) for _ in range(17)]
    iters = C.c_int(0)
    counters = (C.c_longlong * 4)(0, 0, 0, 0)
    o = _opts(ORDER_JACOBI, flavour, max_sweeps)
    rc = lib().orc_sphere_stromgren_solve(
        _p(Edges), _p(nH), _p(nHe), _p(T), C.c_int(Nmu), pE, pS, C.c_int(i_rec_meth),
        C.c_double(fixed_fcA), C.c_int(1), C.c_int(1), C.c_int(i_find_Teq),
        C.c_int(i_thin), C.c_double(em_fac[0]), C.c_double(em_fac[1]),
        C.c_double(em_fac[2]), C.c_double(z), C.c_double(Hz), C.c_double(tol),
        C.c_int(Nl), C.c_int(Nnu), *[_p(a) for a in outs], C.byref(o),
        C.byref(iters), counters)
    _check(rc)
    return _finish(Result(), outs, T, iters, counters)


def slab_bgnd_solve(Edges, nH, nHe, T, E_eV, shape, i_rec_meth=1, fixed_fcA=1.0,
                    i_find_Teq=0, i_thin=0, z=0.0, Hz=0.0, tol=1e-4,
                    flavour=FLAVOUR_HOISTED, max_sweeps=0):
    """Restates slab_bgnd.f90:97-288."""
    Edges, pEd = _d(Edges)
    nH, pnH = _d(nH)
    nHe, pnHe = _d(nHe)
    T = np.array(T, dtype=np.float64)
    E, pE = _d(E_eV)
    S, pS = _d(shape)
    Nl, Nnu = nH.size, E.size
    outs = [np.zeros(Nl) for _ in range(17)]
    iters = C.c_int(0)
    counters = (C.c_longlong * 4)(0, 0, 0, 0)
    o = _opts(ORDER_JACOBI, flavour, max_sweeps)
    rc = lib().orc_slab_bgnd_solve(
        pEd, pnH, pnHe, _p(T), pE, pS, C.c_int(i_rec_meth), C.c_double(fixed_fcA),
        C.c_int(1), C.c_int(1), C.c_int(i_find_Teq), C.c_int(i_thin), C.c_double(z),
        C.c_double(Hz), C.c_double(tol), C.c_int(Nl), C.c_int(Nnu),
        *[_p(a) for a in outs], C.byref(o), C.byref(iters), counters)
    _check(rc)
    return _finish(Result(), outs, T, iters, counters)


def slab_plane_solve(Edges, nH, nHe, T, E_eV, shape, i_2side=0, i_rec_meth=1,
                     fixed_fcA=1.0, i_find_Teq=0, i_thin=0, z=0.0, Hz=0.0,
                     tol=1e-4, flavour=FLAVOUR_HOISTED, max_sweeps=0):
    """Restates slab_plane.f90:105-283."""
    Edges, pEd = _d(Edges)
    nH, pnH = _d(nH)
    nHe, pnHe = _d(nHe)
    T = np.array(T, dtype=np.float64)
    E, pE = _d(E_eV)
    S, pS = _d(shape)
    Nl, Nnu = nH.size, E.size
    outs = [np.zeros(Nl) for _ in range(17)]
    iters = C.c_int(0)
    counters = (C.c_longlong * 4)(0, 0, 0, 0)
    o = _opts(ORDER_JACOBI, flavour, max_sweeps)
    rc = lib().orc_slab_plane_solve(
        pEd, pnH, pnHe, _p(T), pE, pS, C.c_int(i_2side), C.c_int(i_rec_meth),
        C.c_double(fixed_fcA), C.c_int(1), C.c_int(1), C.c_int(i_find_Teq),
        C.c_int(i_thin), C.c_double(z), C.c_double(Hz), C.c_double(tol),
        C.c_int(Nl), C.c_int(Nnu), *[_p(a) for a in outs], C.byref(o),
        C.byref(iters), counters)
    _check(rc)
    return _finish(Result(), outs, T, iters, counters)


def set_column_vs_impact(xH1, xH2, xHe1, xHe2, xHe3, Edges, nH, nHe, T):
    """Restates sphere_base.f90:1115-1205.  Returns (NH, NHe, NH1, NHe1, NHe2)."""
    ins = [np.array(a, dtype=np.float64) for a in
           (xH1, xH2, xHe1, xHe2, xHe3, Edges, nH, nHe, T)]
    Nl = ins[0].size
    outs = [np.zeros(Nl) for _ in range(5)]
    _check(lib().orc_set_column_vs_impact(*[_p(a) for a in ins], C.c_int(Nl),
                                          *[_p(o) for o in outs]))
    return tuple(outs)
