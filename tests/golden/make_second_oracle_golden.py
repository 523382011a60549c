#!/usr/bin/env python
"""Generate ``second_oracle_solves.npz``: whole solves of small problems of every geometry by
the independent numpy derivation ``oracle/second_oracle.py`` (float64, written from the
Fortran, not from the C oracle).

    python tests/golden/make_second_oracle_golden.py

The fixture lets the GPU tests compare the CUDA path with a derivation that shares no code
with ``oracle/rb_oracle_*.c`` (tests/test_zz_gpu_second_oracle_golden.py), and the CPU tests
check it against the C oracle (tests/test_second_oracle.py).  Inputs are stored with the
outputs, so nothing but numpy is needed to use it.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import second_oracle as so  # noqa: E402
from rabacus_b200 import _lib, configs  # noqa: E402

FR = ("xH1", "xH2", "xHe1", "xHe2", "xHe3")
RT = ("H1i_src", "He1i_src", "He2i_src", "H1h_src", "He1h_src", "He2h_src")
RC = ("H1i_rec", "He1i_rec", "He2i_rec", "H1h_rec", "He1h_rec", "He2h_rec")
REC_ID = {None: 1, "outward": 2, "radial": 3, "isotropic": 4, "ray": 2}

# name: (geometry, find_Teq, rec_meth, em_fac)
CASES = {
    "sphere_bgnd": ("sphere_bgnd", False, None, (1.0, 1.0, 1.0)),
    "sphere_bgnd_Teq": ("sphere_bgnd", True, None, (1.0, 1.0, 1.0)),
    "sphere_bgnd_isotropic": ("sphere_bgnd", False, "isotropic", (1.0, 0.9, 1.1)),
    "sphere_bgnd_radial": ("sphere_bgnd", False, "radial", (1.0, 1.0, 1.0)),
    "sphere_stromgren": ("sphere_stromgren", False, None, (1.0, 1.0, 1.0)),
    "slab_plane_1_Teq": ("slab_plane_1", True, None, (1.0, 1.0, 1.0)),
    "slab_plane_2": ("slab_plane_2", False, None, (1.0, 1.0, 1.0)),
    "slab_bgnd_Teq": ("slab_bgnd", True, None, (1.0, 1.0, 1.0)),
    "slab_bgnd_ray": ("slab_bgnd", False, "ray", (1.0, 1.0, 1.0)),
}


def gauss_legendre(n):
    """the product's host rule (rb_gauss_legendre: equal to the C oracle's p_quadrature_rule,
    tests/test_cabi.py) -- a CPU helper, no device needed"""
    import ctypes as C
    x, w = np.zeros(n), np.zeros(n)
    DP = C.POINTER(C.c_double)
    _lib.check(_lib.lib().rb_gauss_legendre(C.c_int(n), x.ctypes.data_as(DP), w.ctypes.data_as(DP)))
    return x, w


def build_case(name):
    geom, find_Teq, meth, em_fac = CASES[name]
    if geom == "sphere_bgnd":
        cfg = configs.c4_sphere_bgnd(16, 16, 2, find_Teq, nH0=0.3)
    elif geom == "sphere_stromgren":
        cfg = configs.c1_iliev_stromgren(16)
    elif geom == "slab_bgnd":
        cfg = configs.c3_slab_bgnd(16, 16, find_Teq)
    else:
        cfg = configs.c2_slab_plane(16, 16, find_Teq, two_sided=(geom == "slab_plane_2"))
    fcA = 1.0 if meth else cfg["fixed_fcA"]
    F = np.float64
    if geom == "sphere_bgnd":
        rv = lambda v: np.asarray(v)[::-1].copy()  # noqa: E731
        mu, w = gauss_legendre(cfg["Nmu"])
        out = so.sphere_bgnd_solve(rv(cfg["Edges"]), rv(cfg["nH"]), rv(cfg["nHe"]), rv(cfg["T"]),
                                   mu, w, cfg["E_eV"], cfg["shape"], fcA, find_Teq, cfg["z"], 0.0,
                                   1e-4, F, rec_meth=meth, em_fac=em_fac)
        x, T, src = [rv(v) for v in out[0]], rv(out[1]), [rv(v) for v in out[2]]
        rec = [rv(v) for v in out[4]] if meth else None
    else:
        out = so.solve_column_geometry(geom, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"],
                                       cfg["E_eV"], cfg["shape"], fcA, find_Teq, cfg["z"], 0.0,
                                       1e-4, F, rec_meth=meth)
        x, T, src = list(out[0]), out[1], list(out[2])
        rec = list(out[4]) if meth else None
    Nl = len(cfg["nH"])
    d = {"geometry": geom, "find_Teq": int(find_Teq), "i_rec_meth": REC_ID[meth],
         "em_fac": np.array(em_fac), "fixed_fcA": fcA, "z": cfg["z"], "Nmu": cfg["Nmu"],
         "Edges": cfg["Edges"], "nH": cfg["nH"], "nHe": cfg["nHe"], "T0": cfg["T"],
         "E_eV": cfg["E_eV"], "shape": cfg["shape"], "iters": out[3],
         "T": np.asarray(T, dtype=np.float64)}
    for q, k in enumerate(FR):
        d[k] = np.asarray(x[q], dtype=np.float64)
    for q, k in enumerate(RT):
        d[k] = np.asarray(src[q], dtype=np.float64)
    for q, k in enumerate(RC):
        d[k] = np.asarray(rec[q], dtype=np.float64) if rec is not None else np.zeros(Nl)
    return d


def main():
    flat = {}
    for name in CASES:
        d = build_case(name)
        print("%-24s %3d sweeps" % (name, d["iters"]))
        for k, v in d.items():
            flat["%s/%s" % (name, k)] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "second_oracle_solves.npz"), **flat)


if __name__ == "__main__":
    main()
