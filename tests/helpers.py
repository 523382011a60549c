"""Shared helpers: run one named configuration through the oracle and through
the CUDA C ABI with identical inputs."""
import numpy as np

from rabacus_b200 import _lib, configs
from rabacus_b200.plan import Plan

OUT17 = ("xH1", "xH2", "xHe1", "xHe2", "xHe3", "H1i_src", "He1i_src", "He2i_src",
         "H1i_rec", "He1i_rec", "He2i_rec", "H1h_src", "He1h_src", "He2h_src",
         "H1h_rec", "He1h_rec", "He2h_rec")
CHECKED = ("xH1", "xH2", "xHe1", "xHe2", "xHe3", "T", "H1i_src", "He1i_src", "He2i_src",
           "H1h_src", "He1h_src", "He2h_src")
CHECKED_REC = CHECKED + ("H1i_rec", "He1i_rec", "He2i_rec", "H1h_rec", "He1h_rec")


def run_oracle(orc, cfg, tol=1e-4, max_sweeps=0, order=None, thin=False, flavour=0):
    g = cfg["geometry"]
    kw = dict(fixed_fcA=cfg["fixed_fcA"], i_find_Teq=1 if cfg["find_Teq"] else 0,
              i_thin=1 if thin else 0, z=cfg["z"], Hz=cfg.get("Hz", 0.0), tol=tol,
              max_sweeps=max_sweeps, flavour=flavour, i_rec_meth=cfg.get("i_rec_meth", 1))
    a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"])
    em = cfg.get("em_fac", (1.0, 1.0, 1.0))
    if g == "sphere_bgnd":
        res = orc.sphere_bgnd_solve(*a, cfg["Nmu"], cfg["E_eV"], cfg["shape"], em_fac=em,
                                    order=orc.ORDER_JACOBI if order is None else order, **kw)
    elif g == "sphere_stromgren":
        res = orc.sphere_stromgren_solve(*a, max(1, cfg["Nmu"]), cfg["E_eV"], cfg["shape"],
                                         em_fac=em, **kw)
    elif g == "slab_bgnd":
        res = orc.slab_bgnd_solve(*a, cfg["E_eV"], cfg["shape"], **kw)
    else:
        res = orc.slab_plane_solve(*a, cfg["E_eV"], cfg["shape"],
                                   i_2side=1 if g == "slab_plane_2" else 0, **kw)
    res["_orc"] = orc
    return res


def run_gpu(cfg, tol=1e-4, max_sweeps=0, thin=False, libm_pow=False, far_field=None, order=None):
    """Through the plan API of the C ABI (same code path as the one-shot entry
    points, plus max_sweeps for sweep-level parity)."""
    Nl, Nnu = cfg["nH"].size, cfg["E_eV"].size
    plan = Plan(configs.GEOM_IDS[cfg["geometry"]], 1, Nl, Nnu, cfg["Nmu"],
                find_Teq=cfg["find_Teq"], fixed_fcA=cfg["fixed_fcA"], tol=tol,
                max_sweeps=max_sweeps, i_rec_meth=cfg.get("i_rec_meth", 1),
                em_fac=cfg.get("em_fac", (1.0, 1.0, 1.0)))
    plan.set_problem(0, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"],
                     cfg["shape"], cfg["z"], cfg.get("Hz", 0.0))
    plan.upload()
    if libm_pow:
        plan.set_libm_pow(True)
    if far_field is not None:
        plan.set_far_field(far_field)
    if order is not None:
        plan.set_sweep_order(order)
    if thin:
        plan.thin()
        plan.sync()
        info = plan.last_info()
    else:
        info = plan.solve()
    res = plan.get_problem(0)
    res["info"] = info
    res["_run"] = dict(cfg=cfg, tol=tol, thin=thin, libm_pow=libm_pow)
    plan.close()
    return res


FRACTIONS = ("xH1", "xH2", "xHe1", "xHe2", "xHe3")
TOL_FRAC = 1.0e-2  # sphere_bgnd.f90:46: the per-cell solvers run at tol*TOL_FRAC
# Rates below this are compared as zeros: exp(-tau) of the reference's libm is denormal /
# zero there (tau > ~640 behind a wall of neutral gas) and so are its products with the
# per-frequency weights; the CUDA path gives exact zeros for rays that are dark at every
# frequency (k_nu_rates).  Nothing this small reaches any other number.
RATE_FLOOR = 1.0e-280


def frac_rtol(tol, rtol=1e-10):
    """Bound for quantities that are LINEAR in minority ionisation fractions when no
    solver-level check is possible (recombination-photon rates; GPU-vs-GPU A/B runs):
    1e-10, or 1 % of the reference solver's own stopping tolerance tol*TOL_FRAC where
    that is larger.  See :func:`solver_check` for how the fractions themselves are pinned."""
    return max(rtol, 1.0e-2 * tol * TOL_FRAC)


def solver_check(orc, res_gpu):
    """The parity statement for the ionisation fractions (and T with find_Teq).

    The reference's per-cell algorithm is ill-conditioned by construction:
    ``analytic_H1``'s discriminant cancels in neutral gas, the n_e bisection stops on
    ``max(x/x_old)`` rather than on a residual, and where photo-ionisation is negligible
    its first branch ``ratio_ne > 1`` is decided by the last bit of the inputs.  Two
    builds of the reference itself therefore agree on minority fractions only as far as
    their rates and libm agree in the last bit -- an end-to-end bound on the fractions
    measures that noise, not the solver.  What CAN be demanded exactly: given the
    CUDA path's own photo-rates (which are separately held to 1e-10, measured <= 1e-13)
    and the device's rate-coefficient bits (host build of the same code, installed as
    the oracle's coefficient hook), the ORACLE's scalar solvers must return the CUDA
    fractions and temperatures BIT FOR BIT, in every cell."""
    from rabacus_b200 import _lib
    run = res_gpu["_run"]
    cfg, tol, thin = run["cfg"], run["tol"], run["thin"]
    if run["libm_pow"]:
        return {}
    Nl = cfg["nH"].size
    two = cfg["geometry"] in ("slab_plane_2", "slab_bgnd")
    n = Nl // 2 if two else Nl
    sl = slice(0, n)
    fcA = cfg["fixed_fcA"] if cfg.get("i_rec_meth", 1) == 1 else 1.0
    tot = {}
    for k in ("H1i", "He1i", "He2i", "H1h", "He1h", "He2h"):
        tot[k] = (res_gpu[k + "_src"] + res_gpu[k + "_rec"])[sl]
    nH, nHe = np.asarray(cfg["nH"], dtype=np.float64)[sl], np.asarray(cfg["nHe"])[sl]
    tol_in = tol * TOL_FRAC
    bad = {}
    with orc.coefficient_hook(_lib.lib().rbi_host_coefficients):
        with np.errstate(all="ignore"):
            if cfg["find_Teq"]:
                out = orc.solve_pcte(nH, nHe, tot["H1i"], tot["He1i"], tot["He2i"], tot["H1h"],
                                     tot["He1h"], tot["He2h"], cfg["z"], cfg.get("Hz", 0.0),
                                     fcA, fcA, fcA, tol_in, vector=thin)
                names = FRACTIONS + ("T",)
            else:
                T = np.asarray(cfg["T"], dtype=np.float64)[sl]
                k = orc.get_kchem(T, fcA, fcA, fcA)
                out = orc.solve_pce(nH, nHe, *k, tot["H1i"], tot["He1i"], tot["He2i"], tol_in,
                                    vector=thin)
                names = FRACTIONS
    for name, ref in zip(names, out):
        got = np.asarray(res_gpu[name])[sl]
        if not np.array_equal(got, ref, equal_nan=True):
            with np.errstate(all="ignore"):
                r = np.where(got == ref, 0.0, np.abs(got - ref) / np.abs(ref))
            bad["solver:" + name] = (int(np.sum(got != ref)), float(np.nanmax(r)))
        if two and not thin:   # the mirrored half is a copy
            full = np.asarray(res_gpu[name])
            if not np.array_equal(full[Nl - n:][::-1], full[:n], equal_nan=True):
                bad["mirror:" + name] = 1
    return bad


def compare(res_gpu, res_orc, names=CHECKED, rtol=1e-10, tol=1e-4):
    """(worst, bad).  Photo-rates and T: relative error <= rtol against the oracle run
    (values below RATE_FLOOR compare as zeros).  Ionisation fractions: when ``res_orc`` is
    an oracle run, the solver-level bit-for-bit statement of :func:`solver_check`; their
    end-to-end deviation is only recorded in ``worst``.  Recombination-photon rates (linear
    in minority fractions of the previous iterate) and GPU-vs-GPU comparisons:
    :func:`frac_rtol`."""
    worst = {}
    bad = {}
    orc = res_orc.get("_orc") if hasattr(res_orc, "get") else None
    solver_level = orc is not None and "_run" in res_gpu
    for n in names:
        a = np.asarray(res_gpu[n])
        b = np.asarray(res_orc[n])
        is_rate = not (n in FRACTIONS or n == "T")
        floor = RATE_FLOOR if is_rate else 0.0
        with np.errstate(divide="ignore", invalid="ignore"):
            r = np.abs(a - b) / np.maximum(np.abs(b), floor)
        r = np.where((a == b) | ((np.abs(a) < floor) & (np.abs(b) < floor)), 0.0, r)
        worst[n] = float(np.nanmax(r))
        if n in FRACTIONS and solver_level:
            continue
        bound = frac_rtol(tol, rtol) if (n in FRACTIONS or n.endswith("_rec")) else rtol
        if not worst[n] <= bound:
            bad[n] = worst[n]
    if solver_level:
        bad.update(solver_check(orc, res_gpu))
    return worst, bad


def load_second_oracle_golden():
    """tests/golden/second_oracle_solves.npz (made by make_second_oracle_golden.py from the
    independent numpy derivation) -> {case: dict of inputs and outputs}"""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                             "second_oracle_solves.npz"))
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        v = z[key]
        cases.setdefault(name, {})[field] = v.item() if v.ndim == 0 else v
    return cases


def golden_cfg(g):
    """the config dict run_gpu / run_oracle take, from a golden case"""
    return dict(geometry=str(g["geometry"]), Edges=g["Edges"].copy(), nH=g["nH"].copy(),
                nHe=g["nHe"].copy(), T=g["T0"].copy(), E_eV=g["E_eV"].copy(),
                shape=g["shape"].copy(), fixed_fcA=float(g["fixed_fcA"]),
                find_Teq=bool(g["find_Teq"]), z=float(g["z"]), Nmu=int(g["Nmu"]),
                i_rec_meth=int(g["i_rec_meth"]), em_fac=tuple(float(v) for v in g["em_fac"]))


def golden_diff(res, g, names):
    """worst relative difference over ``names`` (exact zeros must match as zeros)"""
    worst = 0.0
    for k in names:
        a, b = np.asarray(res[k], dtype=np.float64), np.asarray(g[k], dtype=np.float64)
        nz = b != 0.0
        assert np.all(a[~nz] == 0.0), k
        if nz.any():
            worst = max(worst, float(np.max(np.abs(a[nz] / b[nz] - 1.0))))
    return worst
