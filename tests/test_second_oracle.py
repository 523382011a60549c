"""The C oracle pinned by an independent second derivation (oracle/second_oracle.py: numpy,
written from the Fortran, not from rb_oracle_*.c).  In float64 the two codings execute the
reference's statements in the same order, so fractions, temperatures and the bisection
paths must coincide; in longdouble (64-bit mantissa) the second derivation is the "truth"
both float64 codings are measured against.  Bounds: rates <= 1e-13, float64 bisection
results bit-identical (a 1-ulp difference between numpy's and C's pow() in a rate
coefficient is allowed for: <= 1e-14 there)."""
import numpy as np
import pytest

from oracle import second_oracle as so
from rabacus_b200 import configs

FR = ("xH1", "xH2", "xHe1", "xHe2", "xHe3")
RT = ("H1i_src", "He1i_src", "He2i_src", "H1h_src", "He1h_src", "He2h_src")


def _zones(n, seed):
    rng = np.random.default_rng(seed)
    T = 10 ** rng.uniform(3.9, 5.0, n)
    nH = 10 ** rng.uniform(-5.0, 1.0, n)
    H1i = 10 ** rng.uniform(-17.0, -11.0, n)
    return T, nH, nH * 0.08, H1i, H1i * 0.6, H1i * 0.02


def test_rate_coefficients_against_the_second_derivation(orc):
    T = 10 ** np.linspace(3.0, 7.0, 60)
    for fc in ((1.0, 1.0, 1.0), (0.0, 0.0, 0.0), (0.3, 0.77, 0.5)):
        kc, kl = orc.get_kchem(T, *fc), orc.get_kcool(T, *fc)
        for i, t in enumerate(T):
            truth = so.get_kchem(t, fc, np.longdouble) + so.get_kcool(t, fc, np.longdouble)
            f64 = so.get_kchem(t, fc, np.float64) + so.get_kcool(t, fc, np.float64)
            for q, col in enumerate(list(kc) + list(kl)):
                # vs 64-bit mantissa; exp(-lam/2) carries |lam/2| ulp of its argument (lam <= 570)
                assert abs(col[i] / float(truth[q]) - 1.0) <= 5e-14, (t, q)
                # vs the float64 twin: numpy's log10/pow differ from glibc's by <= 1 ulp, which the
                # case A/B mixing 10**(f log10 a + ..) amplifies by |ln a| ~ 30
                assert abs(col[i] / float(f64[q]) - 1.0) <= 2e-14, (t, q)


def test_solve_pce_same_fractions_and_same_bisection_path(orc):
    T, nH, nHe, H1i, He1i, He2i = _zones(60, 1)
    k = orc.get_kchem(T, 1.0, 1.0, 1.0)
    x = orc.solve_pce(nH, nHe, *k, H1i, He1i, He2i, 1e-6)
    for i in range(T.size):
        ki = [a[i] for a in k]
        p64, pld = [], []
        x64 = so.solve_pce(nH[i], nHe[i], ki, H1i[i], He1i[i], He2i[i], 1e-6, np.float64, p64)
        xld = so.solve_pce(nH[i], nHe[i], ki, H1i[i], He1i[i], He2i[i], 1e-6, np.longdouble, pld)
        assert all(x[q][i] == x64[q] for q in range(5)), i        # bit-identical
        assert p64 == pld and len(p64) >= 1, i                    # same branches, same length
        # float64 against 64-bit-mantissa arithmetic on the SAME path: the conditioning of the
        # reference algorithm itself (analytic_H1's discriminant cancels, ion_solver.f90:172-183)
        assert max(abs(float(xld[q]) / x[q][i] - 1.0) for q in range(5)) <= 1e-10


def test_solve_pcte_same_temperature_and_fractions(orc):
    T, nH, nHe, H1i, He1i, He2i = _zones(24, 2)
    heat = (H1i * 5e-12, He1i * 8e-12, He2i * 2e-11)
    o = orc.solve_pcte(nH, nHe, H1i, He1i, He2i, *heat, 3.0, 1e-18, 1.0, 0.5, 0.0, 1e-6)
    for i in range(T.size):
        p = (H1i[i], He1i[i], He2i[i], heat[0][i], heat[1][i], heat[2][i])
        p64, pld = [], []
        x64, T64 = so.solve_pcte(nH[i], nHe[i], p, 3.0, 1e-18, (1.0, 0.5, 0.0), 1e-6,
                                 np.float64, p64)
        xld, Tld = so.solve_pcte(nH[i], nHe[i], p, 3.0, 1e-18, (1.0, 0.5, 0.0), 1e-6,
                                 np.longdouble, pld)
        assert o[5][i] == T64 and p64 == pld, i                  # same T bisection path
        assert abs(float(Tld) / o[5][i] - 1.0) <= 1e-15
        for q in range(5):
            assert abs(o[q][i] / x64[q] - 1.0) <= 1e-14, (i, q)
            assert abs(o[q][i] / float(xld[q]) - 1.0) <= 1e-10, (i, q)


def test_helium_free_zone_first_branch_is_decided_by_the_last_bit(orc):
    """Why fractions cannot be held to 1e-10 end to end (helpers.solver_check).  Without helium
    analytic_H1 (ion_solver.f90:172-183) is the exact fixed point of solve_pce, so the first
    ratio_ne (:503) is 1 within a few ulp and the branch :515-521 is decided by rounding.
    The scalar solver stops after that one step, but the lock-step `_v` twin (the optically
    thin start, :565-696) keeps every zone bisecting until the slowest has converged: a
    helium-free zone then returns from whichever bracket the noise chose.  Shown in the
    oracle alone, no GPU involved: nudging H1i by ONE ulp moves xH1 of helium-free zones by
    up to ~1e-7 (a good part of them by more than 1e-10), zones with helium by < 1e-14."""
    rng = np.random.default_rng(4)
    n, tol = 200, 1e-6
    T = np.full(n, 1.0e4)
    nH = 10 ** rng.uniform(-4.0, 0.0, n)
    H1i = 10 ** rng.uniform(-14.0, -11.0, n)
    nHe = nH * 0.08
    nHe[::2] = 0.0
    k = orc.get_kchem(T, 1.0, 1.0, 1.0)
    # the first ratio_ne of a helium-free zone: one implicit step from the analytic start
    for i in range(0, n, 2):
        reH2, ciH1 = k[0][i], k[3][i]
        RR = (ciH1 + reH2) * nH[i]
        QQ = -(H1i[i] + reH2 * nH[i] + RR)
        PP = reH2 * nH[i]
        x1 = PP / (-0.5 * (QQ - np.sqrt(max(QQ * QQ - 4.0 * RR * PP, 0.0))))
        ne0 = (1.0 - x1) * nH[i]
        ne1 = (ciH1 * ne0 + H1i[i]) / (reH2 * ne0 + ciH1 * ne0 + H1i[i]) * nH[i]
        assert abs(ne1 / ne0 - 1.0) <= 1e-11, i
    args = (nH, nHe) + tuple(k)
    base = orc.solve_pce(*args, H1i, H1i * 0.6, H1i * 0.02, tol, vector=True)
    moved = np.zeros(n)
    for h in (np.nextafter(H1i, np.inf), np.nextafter(H1i, 0.0)):
        x = orc.solve_pce(*args, h, H1i * 0.6, H1i * 0.02, tol, vector=True)
        moved = np.maximum(moved, np.abs(x[0] / base[0] - 1.0))
    assert moved[1::2].max() <= 1e-14                   # zones with helium: well conditioned
    assert (moved[::2] > 1e-10).sum() >= 10              # helium-free: decided by the last bit
    assert 1e-9 < moved[::2].max() <= 100 * tol


def _sphere_cfg(kind, Nl, Nnu, Nmu, find_Teq):
    cfg = configs.c4_sphere_bgnd(Nl, Nnu, Nmu, find_Teq, nH0=0.3)
    if kind == "log_clumpy":      # log-spaced radii, lognormal clumps (2 dex), a void shell
        rng = np.random.default_rng(5)
        cfg["Edges"] = np.concatenate([[0.0], cfg["Edges"][-1] * 10 ** np.linspace(-3, 0, Nl)])
        cfg["nH"] = cfg["nH"] * 10 ** rng.normal(0.0, 2.0, Nl)
        cfg["nH"][Nl // 3] = 1e-12 * cfg["nH"].max()
        cfg["nHe"] = cfg["nH"] * 0.08
    return cfg


@pytest.mark.parametrize("kind", ["uniform", "log_clumpy"])
@pytest.mark.parametrize("find_Teq", [False, True])
def test_one_jacobi_sweep_of_sphere_bgnd(orc, kind, find_Teq):
    Nl, Nnu, Nmu = 28, 16, 5
    cfg = _sphere_cfg(kind, Nl, Nnu, Nmu, find_Teq)
    a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], Nmu, cfg["E_eV"], cfg["shape"])
    kw = dict(i_find_Teq=int(find_Teq), z=3.0, order=orc.ORDER_JACOBI)
    thin = orc.sphere_bgnd_solve(*a, i_thin=1, **kw)
    one = orc.sphere_bgnd_solve(*a, max_sweeps=1, **kw)
    mu, w = orc.p_quadrature_rule(Nmu)
    rv = lambda v: np.asarray(v)[::-1].copy()  # noqa: E731  (the sweep runs outer shell first)
    for F, rt_tol, x_tol in ((np.float64, 2e-15, 1e-13), (np.longdouble, 1e-13, 1e-10)):
        out = so.sphere_bgnd_jacobi_sweep(
            rv(cfg["Edges"]), rv(cfg["nH"]), rv(cfg["nHe"]), rv(thin["T"]),
            [rv(thin[k]) for k in FR], [rv(thin[k]) for k in RT], mu, w, cfg["E_eV"],
            cfg["shape"], 1.0, find_Teq, 3.0, 0.0, 1e-4, F)
        for il, (xs, Tn, acc) in out.items():
            io = Nl - 1 - il
            for q, k in enumerate(RT):
                assert abs(float(acc[q]) / one[k][io] - 1.0) <= rt_tol, (F, il, k)
            for q, k in enumerate(FR):
                assert abs(float(xs[q]) / one[k][io] - 1.0) <= x_tol, (F, il, k)
            assert abs(float(Tn) / one["T"][io] - 1.0) <= 1e-15, (F, il)


@pytest.mark.parametrize("find_Teq", [False, True])
def test_one_sweep_of_slab_bgnd(orc, find_Teq):
    Nl, Nnu = 26, 16
    cfg = configs.c3_slab_bgnd(Nl, Nnu, find_Teq)
    rng = np.random.default_rng(9)
    cfg["nH"] = cfg["nH"] * 10 ** rng.normal(0.0, 1.0, Nl)      # asymmetric, clumpy
    cfg["nHe"] = cfg["nH"] * 0.08
    a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"])
    kw = dict(i_find_Teq=int(find_Teq), z=3.0)
    thin = orc.slab_bgnd_solve(*a, i_thin=1, **kw)
    one = orc.slab_bgnd_solve(*a, max_sweeps=1, **kw)
    for F, rt_tol, x_tol in ((np.float64, 2e-15, 1e-13), (np.longdouble, 1e-13, 1e-10)):
        out = so.slab_bgnd_sweep(cfg["Edges"], cfg["nH"], cfg["nHe"], thin["T"],
                                 [thin[k] for k in FR], cfg["E_eV"], cfg["shape"], 1.0,
                                 find_Teq, 3.0, 0.0, 1e-4, F)
        assert sorted(out) == list(range(Nl // 2))
        for i, (xs, Tn, acc) in out.items():
            for q, k in enumerate(RT):
                assert abs(float(acc[q]) / one[k][i] - 1.0) <= rt_tol, (F, i, k)
                assert one[k][Nl - 1 - i] == one[k][i]             # the mirrored copy
            for q, k in enumerate(FR):
                assert abs(float(xs[q]) / one[k][i] - 1.0) <= x_tol, (F, i, k)
            assert abs(float(Tn) / one["T"][i] - 1.0) <= 1e-15, (F, i)


@pytest.mark.parametrize("geom,find_Teq", [("slab_plane_1", False), ("slab_plane_1", True),
                                           ("slab_plane_2", True), ("sphere_stromgren", False),
                                           ("sphere_stromgren_hm12", True)])
def test_one_sweep_of_the_column_geometries(orc, geom, find_Teq):
    """SlabPln, Slab2Pln and SphereStromgren (monochromatic Iliev source and a polychromatic
    one): thin start by the C oracle, then one sweep by both derivations."""
    Nl, Nnu = 24, 16
    rng = np.random.default_rng(13)
    if geom.startswith("slab"):
        cfg = configs.c2_slab_plane(Nl, Nnu, find_Teq, two_sided=(geom == "slab_plane_2"))
        cfg["nH"] = cfg["nH"] * 10 ** rng.normal(0.0, 1.0, Nl)       # clumpy, asymmetric
        cfg["nHe"] = cfg["nH"] * 0.08
        a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"])
        kw = dict(i_find_Teq=int(find_Teq), z=3.0, i_2side=int(geom == "slab_plane_2"))
        solve = orc.slab_plane_solve
        g = geom
    else:
        cfg = configs.c1_iliev_stromgren(Nl)
        if geom.endswith("hm12"):                                    # polychromatic point source
            from rabacus_b200 import rad_src
            src = rad_src.PointSource(1.0, 400.0, "hm12", Nnu=Nnu, z=3.0)
            src.normalize_Ln(5.0e48)
            cfg["E_eV"], cfg["shape"] = src.E.copy(), src.shape.copy()
            cfg["nHe"] = cfg["nH"] * 0.08
        cfg["nH"] = cfg["nH"] * 10 ** rng.normal(0.0, 0.5, Nl)
        a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], 1, cfg["E_eV"], cfg["shape"])
        kw = dict(i_find_Teq=int(find_Teq), z=3.0, fixed_fcA=cfg["fixed_fcA"])
        solve = orc.sphere_stromgren_solve
        g = "sphere_stromgren"
    thin = solve(*a, i_thin=1, **kw)
    one = solve(*a, max_sweeps=1, **kw)
    fcA = cfg["fixed_fcA"]
    ncell = Nl // 2 if g == "slab_plane_2" else Nl
    # float64: same statements, same bisection path; longdouble: the "truth" -- a minority
    # fraction carries the n_e bisection's own stopping noise (tol_inner = 1e-6 on successive
    # iterates), measured up to 1.6e-10 here
    for F, rt_tol, x_tol in ((np.float64, 4e-15, 1e-13), (np.longdouble, 1e-13, 1e-9)):
        out = so.column_sweep(g, cfg["Edges"], cfg["nH"], cfg["nHe"], thin["T"],
                              [thin[k] for k in FR], cfg["E_eV"], cfg["shape"], fcA, find_Teq,
                              3.0, 0.0, 1e-4, F)
        assert sorted(out) == list(range(ncell))
        for i, (xs, Tn, acc) in out.items():
            for q, k in enumerate(RT):
                ref = one[k][i]
                assert (ref == 0.0 and float(acc[q]) == 0.0) or \
                    abs(float(acc[q]) / ref - 1.0) <= rt_tol, (F, i, k, float(acc[q]), ref)
            for q, k in enumerate(FR):
                assert abs(float(xs[q]) / one[k][i] - 1.0) <= x_tol, (F, i, k)
            assert abs(float(Tn) / one["T"][i] - 1.0) <= 1e-15, (F, i)


@pytest.mark.parametrize("find_Teq", [False, True])
def test_one_gauss_seidel_sweep_of_sphere_bgnd(orc, find_Teq):
    """the reference's own (single-thread) order, sphere_bgnd.f90:753-913: the oracle's
    order=gs against the second derivation run the same way"""
    Nl, Nnu, Nmu = 24, 16, 4
    cfg = _sphere_cfg("uniform", Nl, Nnu, Nmu, find_Teq)
    a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], Nmu, cfg["E_eV"], cfg["shape"])
    kw = dict(i_find_Teq=int(find_Teq), z=3.0, order=orc.ORDER_GS)
    thin = orc.sphere_bgnd_solve(*a, i_thin=1, **kw)
    one = orc.sphere_bgnd_solve(*a, max_sweeps=1, **kw)
    jac = orc.sphere_bgnd_solve(*a, max_sweeps=1, i_find_Teq=int(find_Teq), z=3.0,
                                order=orc.ORDER_JACOBI)
    assert np.max(np.abs(jac["xH1"] / one["xH1"] - 1.0)) > 1e-9      # the orders do differ
    mu, w = orc.p_quadrature_rule(Nmu)
    rv = lambda v: np.asarray(v)[::-1].copy()  # noqa: E731
    out = so.sphere_bgnd_jacobi_sweep(
        rv(cfg["Edges"]), rv(cfg["nH"]), rv(cfg["nHe"]), rv(thin["T"]),
        [rv(thin[k]) for k in FR], [rv(thin[k]) for k in RT], mu, w, cfg["E_eV"],
        cfg["shape"], 1.0, find_Teq, 3.0, 0.0, 1e-4, np.float64, gauss_seidel=True)
    for il, (xs, Tn, acc) in out.items():
        io = Nl - 1 - il
        for q, k in enumerate(RT):
            assert abs(float(acc[q]) / one[k][io] - 1.0) <= 4e-15, (il, k)
        for q, k in enumerate(FR):
            assert abs(float(xs[q]) / one[k][io] - 1.0) <= 1e-13, (il, k)
        assert abs(float(Tn) / one["T"][io] - 1.0) <= 1e-15, il


@pytest.mark.parametrize("meth", ["outward", "radial", "isotropic"])
def test_sphere_recombination_photon_closures(orc, meth):
    """set_recomb_photons_outward / _radial / _isotropic (sphere_base.f90:79-1083): the closure
    the first sweep of a rec_meth solve evaluates on the thin state, against the second
    derivation (float64: same statements; longdouble: the truth)."""
    Nl, Nnu, Nmu = 20, 16, 4
    cfg = _sphere_cfg("log_clumpy", Nl, Nnu, Nmu, False)
    a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], Nmu, cfg["E_eV"], cfg["shape"])
    em_fac = (1.0, 0.8, 1.25) if meth == "isotropic" else (1.0, 1.0, 1.0)
    kw = dict(z=3.0, order=orc.ORDER_JACOBI, i_rec_meth={"outward": 2, "radial": 3, "isotropic": 4}[meth],
              em_fac=em_fac)
    thin = orc.sphere_bgnd_solve(*a, i_thin=1, **kw)
    one = orc.sphere_bgnd_solve(*a, max_sweeps=1, **kw)
    assert cfg["Edges"][1] > cfg["Edges"][0]
    rv = lambda v: np.asarray(v)[::-1].copy()  # noqa: E731
    names = ("H1i_rec", "He1i_rec", "He2i_rec", "H1h_rec", "He1h_rec", "He2h_rec")
    mu, w = orc.p_quadrature_rule(Nmu)
    for F, tol in ((np.float64, 1e-13), (np.longdouble, 1e-12)):
        if meth == "radial":                    # the routine's own orientation: increasing edges
            got = so.recomb_radial(cfg["Edges"], cfg["nH"], cfg["nHe"], thin["T"],
                                   [thin[k] for k in FR], F)
            got = [rv(g) for g in got]
        else:
            args = (rv(cfg["Edges"]), rv(cfg["nH"]), rv(cfg["nHe"]), rv(thin["T"]),
                    [rv(thin[k]) for k in FR])
            got = (so.recomb_outward(*args, F) if meth == "outward"
                   else so.recomb_isotropic(*args, mu, w, em_fac, F))
        for g, k in zip(got, names):
            ref = rv(one[k])
            g = np.asarray(g, dtype=np.float64)
            assert np.all((ref == 0.0) == (g == 0.0)), k
            nz = ref != 0.0
            if nz.any():
                err = np.max(np.abs(g[nz] / ref[nz] - 1.0))
                assert err <= tol, (F, k, err)
    assert np.any(one["H1i_rec"] > 0.0) and np.all(one["He2h_rec"] == 0.0)


@pytest.mark.parametrize("geom", ["slab_plane_1", "slab_plane_2", "slab_bgnd"])
def test_slab_recombination_photon_closure(orc, geom):
    """set_recomb_photons (slab_base.f90:163-560), rec_meth = ray, on the thin state."""
    Nl, Nnu = 24, 16
    rng = np.random.default_rng(17)
    if geom == "slab_bgnd":
        cfg = configs.c3_slab_bgnd(Nl, Nnu, False)
        solve, kw = orc.slab_bgnd_solve, dict(z=3.0, i_rec_meth=2)
    else:
        cfg = configs.c2_slab_plane(Nl, Nnu, False, two_sided=(geom == "slab_plane_2"))
        solve, kw = orc.slab_plane_solve, dict(z=3.0, i_rec_meth=2, i_2side=int(geom == "slab_plane_2"))
    cfg["nH"] = cfg["nH"] * 10 ** rng.normal(0.0, 1.0, Nl)
    cfg["nHe"] = cfg["nH"] * 0.08
    a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"])
    thin = solve(*a, i_thin=1, **kw)
    one = solve(*a, max_sweeps=1, **kw)
    ncell = Nl if geom == "slab_plane_1" else Nl // 2
    names = ("H1i_rec", "He1i_rec", "He2i_rec", "H1h_rec", "He1h_rec", "He2h_rec")
    # longdouble: the routine recovers each layer's optical depth as a DIFFERENCE of the running
    # sums tau_lo(k+1) - tau_lo(k) (:400-427), which in float64 carries eps * tau_lo / dtau
    for F, tol in ((np.float64, 1e-13), (np.longdouble, 1e-10)):
        got = so.recomb_slab(cfg["Edges"], cfg["nH"], cfg["nHe"], thin["T"], [thin[k] for k in FR], F)
        for g, k in zip(got, names):
            ref = np.asarray(one[k])[:ncell]
            g = np.asarray(g, dtype=np.float64)[:ncell]
            if ncell < Nl:                      # slab_bgnd.f90:842-865: rates reflected too
                assert np.array_equal(np.asarray(one[k])[::-1][:ncell], ref), k
            assert np.all((ref == 0.0) == (g == 0.0)), k
            nz = ref != 0.0
            if nz.any():
                err = np.max(np.abs(g[nz] / ref[nz] - 1.0))
                assert err <= tol, (F, k, err)
    assert np.any(one["H1i_rec"] > 0.0) and np.all(one["He2h_rec"] == 0.0)


@pytest.mark.parametrize("kind", ["uniform", "log_clumpy"])
def test_column_vs_impact(orc, kind):
    """set_column_vs_impact (sphere_base.f90:1115-1205) on a thin state, increasing and
    decreasing edges (the routine reverses and restores)."""
    Nl, Nnu, Nmu = 30, 16, 4
    cfg = _sphere_cfg(kind, Nl, Nnu, Nmu, False)
    a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], Nmu, cfg["E_eV"], cfg["shape"])
    thin = orc.sphere_bgnd_solve(*a, i_thin=1, z=3.0, order=orc.ORDER_JACOBI)
    x = [thin[k] for k in FR]
    ref = orc.set_column_vs_impact(*x, cfg["Edges"], cfg["nH"], cfg["nHe"], thin["T"])
    rv = lambda v: np.asarray(v)[::-1].copy()  # noqa: E731
    ref_dec = orc.set_column_vs_impact(*[rv(v) for v in x], rv(cfg["Edges"]), rv(cfg["nH"]),
                                       rv(cfg["nHe"]), rv(thin["T"]))
    for r, d in zip(ref, ref_dec):
        assert np.array_equal(r, rv(d))
    for F, tol in ((np.float64, 0.0), (np.longdouble, 1e-13)):
        got = so.column_vs_impact(rv(cfg["Edges"]), rv(cfg["nH"]), rv(cfg["nHe"]), [rv(v) for v in x], F)
        for g, r in zip(got, ref):
            g = rv(np.asarray(g, dtype=np.float64))
            assert np.all(r > 0.0)
            assert np.max(np.abs(g / r - 1.0)) <= tol, (F, np.max(np.abs(g / r - 1.0)))


@pytest.mark.parametrize("find_Teq", [False, True])
def test_thin_start_with_the_lock_step_solvers(orc, find_Teq):
    """sphere_bgnd_optically_thin (sphere_bgnd.f90:379-499) with solve_pce_v / solve_pcte_v
    (ion_solver.f90:565-696, 884-1056): every zone iterates until ALL have converged, so a
    zone's result depends on its neighbours' iteration counts."""
    Nl, Nnu, Nmu = 18, 16, 4
    cfg = _sphere_cfg("log_clumpy", Nl, Nnu, Nmu, find_Teq)
    a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], Nmu, cfg["E_eV"], cfg["shape"])
    thin = orc.sphere_bgnd_solve(*a, i_thin=1, i_find_Teq=int(find_Teq), z=3.0,
                                 order=orc.ORDER_JACOBI)
    for F, x_tol in ((np.float64, 1e-13), (np.longdouble, 1e-9)):
        x, T, r6 = so.sphere_bgnd_thin(cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"],
                                       1.0, find_Teq, 3.0, 0.0, 1e-4, F)
        for q, kname in enumerate(RT):
            assert np.max(np.abs(float(r6[q]) / thin[kname] - 1.0)) <= 4e-15, (F, kname)
        for i in range(Nl):
            for q, kname in enumerate(FR):
                assert abs(float(x[i][q]) / thin[kname][i] - 1.0) <= x_tol, (F, i, kname)
            assert abs(float(T[i]) / thin["T"][i] - 1.0) <= 1e-15, (F, i)
    # the lock-step result is NOT the scalar solver's: at least one zone differs in its last bits
    if not find_Teq:
        k = [so.get_kchem(cfg["T"][i], (1.0,) * 3, np.float64) for i in range(Nl)]
        diff = 0
        for i in range(Nl):
            xs = so.solve_pce(cfg["nH"][i], cfg["nHe"][i], k[i], r6[0], r6[1], r6[2], 1e-6, np.float64)
            diff += int(any(float(xs[q]) != thin[kname][i] for q, kname in enumerate(FR)))
        assert diff > 0


@pytest.mark.parametrize("order,find_Teq", [("jacobi", False), ("jacobi", True), ("gs", False)])
def test_whole_sphere_bgnd_solve(orc, order, find_Teq):
    """sphere_bgnd_solve (sphere_bgnd.f90:107-314): thin start, the 1/2 (old + sum) recurrence of
    the rates across sweeps (Q1), the n_e convergence test -- the same sweep count and the same
    converged state from both derivations, in the Jacobi order and in the reference's own
    sequential order."""
    Nl, Nnu, Nmu = 12, 12, 2
    cfg = _sphere_cfg("uniform", Nl, Nnu, Nmu, find_Teq)
    a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], Nmu, cfg["E_eV"], cfg["shape"])
    ref = orc.sphere_bgnd_solve(*a, i_find_Teq=int(find_Teq), z=3.0,
                                order=orc.ORDER_JACOBI if order == "jacobi" else orc.ORDER_GS)
    rv = lambda v: np.asarray(v)[::-1].copy()  # noqa: E731
    mu, w = orc.p_quadrature_rule(Nmu)
    x, T, src, sweeps = so.sphere_bgnd_solve(rv(cfg["Edges"]), rv(cfg["nH"]), rv(cfg["nHe"]),
                                             rv(cfg["T"]), mu, w, cfg["E_eV"], cfg["shape"], 1.0,
                                             find_Teq, 3.0, 0.0, 1e-4, np.float64,
                                             gauss_seidel=(order == "gs"))
    assert sweeps == ref["iters"] and sweeps > 5
    for q, k in enumerate(RT):
        assert np.max(np.abs(rv(src[q]).astype(np.float64) / ref[k] - 1.0)) <= 1e-12, k
    for q, k in enumerate(FR):
        assert np.max(np.abs(rv(x[q]).astype(np.float64) / ref[k] - 1.0)) <= 1e-11, k
    assert np.max(np.abs(rv(T).astype(np.float64) / ref["T"] - 1.0)) <= 1e-13


@pytest.mark.parametrize("geom,find_Teq", [("slab_plane_1", False), ("slab_plane_1", True),
                                           ("slab_plane_2", False), ("slab_bgnd", False),
                                           ("slab_bgnd", True), ("sphere_stromgren", False)])
def test_whole_solves_of_the_column_geometries(orc, geom, find_Teq):
    """Drivers of SlabPln, Slab2Pln, Slab2Bgnd and SphereStromgren from the thin start (lock-step
    solvers over all layers) to convergence: the same sweep count and state from both
    derivations."""
    Nl, Nnu = 16, 12
    kw = dict(i_find_Teq=int(find_Teq), z=3.0)
    if geom == "slab_bgnd":
        cfg = configs.c3_slab_bgnd(Nl, Nnu, find_Teq)
        a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"])
        ref = orc.slab_bgnd_solve(*a, **kw)
    elif geom.startswith("slab"):
        cfg = configs.c2_slab_plane(Nl, Nnu, find_Teq, two_sided=(geom == "slab_plane_2"))
        a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"])
        ref = orc.slab_plane_solve(*a, i_2side=int(geom == "slab_plane_2"), **kw)
    else:
        cfg = configs.c1_iliev_stromgren(Nl)
        a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], 1, cfg["E_eV"], cfg["shape"])
        ref = orc.sphere_stromgren_solve(*a, fixed_fcA=cfg["fixed_fcA"], **kw)
    x, T, src, sweeps = so.solve_column_geometry(geom, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"],
                                                 cfg["E_eV"], cfg["shape"], cfg["fixed_fcA"],
                                                 find_Teq, 3.0, 0.0, 1e-4, np.float64)
    assert sweeps == ref["iters"] and sweeps > 2, (sweeps, ref["iters"])
    for q, k in enumerate(RT):
        r = np.asarray(ref[k])
        g = np.asarray(src[q], dtype=np.float64)
        nz = r != 0.0
        assert np.all(g[~nz] == 0.0)
        if nz.any():
            assert np.max(np.abs(g[nz] / r[nz] - 1.0)) <= 1e-12, k
    assert np.all(np.asarray(ref["H1i_src"]) > 0.0)
    for q, k in enumerate(FR):
        assert np.max(np.abs(np.asarray(x[q], dtype=np.float64) / ref[k] - 1.0)) <= 1e-11, k
    assert np.max(np.abs(np.asarray(T, dtype=np.float64) / ref["T"] - 1.0)) <= 1e-13


RC = ("H1i_rec", "He1i_rec", "He2i_rec", "H1h_rec", "He1h_rec", "He2h_rec")


@pytest.mark.parametrize("geom,meth", [("sphere_bgnd", "outward"), ("sphere_bgnd", "radial"),
                                       ("sphere_bgnd", "isotropic"), ("slab_plane_1", "ray"),
                                       ("slab_plane_2", "ray"), ("slab_bgnd", "ray"),
                                       ("sphere_stromgren", "outward")])
def test_whole_solves_with_recombination_photons(orc, geom, meth):
    """rec_meth != fixed from the thin start to convergence: the closure at the top of every
    sweep on the pre-sweep state, case-A coefficients, total = source + recombination rates."""
    irec = {"outward": 2, "radial": 3, "isotropic": 4, "ray": 2}[meth]
    kw = dict(z=3.0, i_rec_meth=irec)
    if geom == "sphere_bgnd":
        Nl, Nnu, Nmu = 12, 12, 2
        cfg = _sphere_cfg("uniform", Nl, Nnu, Nmu, False)
        a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], Nmu, cfg["E_eV"], cfg["shape"])
        ref = orc.sphere_bgnd_solve(*a, order=orc.ORDER_JACOBI, em_fac=(1.0, 0.9, 1.1), **kw)
        rv = lambda v: np.asarray(v)[::-1].copy()  # noqa: E731
        mu, w = orc.p_quadrature_rule(Nmu)
        x, T, src, sweeps, rec = so.sphere_bgnd_solve(
            rv(cfg["Edges"]), rv(cfg["nH"]), rv(cfg["nHe"]), rv(cfg["T"]), mu, w, cfg["E_eV"],
            cfg["shape"], 1.0, False, 3.0, 0.0, 1e-4, np.float64, rec_meth=meth,
            em_fac=(1.0, 0.9, 1.1))
        x, src, rec, T = [rv(v) for v in x], [rv(v) for v in src], [rv(v) for v in rec], rv(T)
    else:
        Nl, Nnu = 16, 12
        if geom == "slab_bgnd":
            cfg = configs.c3_slab_bgnd(Nl, Nnu, False)
            a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"])
            ref = orc.slab_bgnd_solve(*a, **kw)
        elif geom.startswith("slab"):
            cfg = configs.c2_slab_plane(Nl, Nnu, False, two_sided=(geom == "slab_plane_2"))
            a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"])
            ref = orc.slab_plane_solve(*a, i_2side=int(geom == "slab_plane_2"), **kw)
        else:
            cfg = configs.c1_iliev_stromgren(Nl)
            cfg["nHe"] = cfg["nH"] * 0.08
            a = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], 1, cfg["E_eV"], cfg["shape"])
            ref = orc.sphere_stromgren_solve(*a, fixed_fcA=1.0, **kw)
        x, T, src, sweeps, rec = so.solve_column_geometry(
            geom, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"], 1.0,
            False, 3.0, 0.0, 1e-4, np.float64, rec_meth=meth)
    assert sweeps == ref["iters"] and sweeps > 2, (sweeps, ref["iters"])
    for names, arr, tol in ((RT, src, 1e-11), (RC, rec, 1e-10), (FR, x, 1e-10)):
        for q, k in enumerate(names):
            r = np.asarray(ref[k])
            g = np.asarray(arr[q], dtype=np.float64)
            nz = r != 0.0
            assert np.all(g[~nz] == 0.0), k
            if nz.any():
                assert np.max(np.abs(g[nz] / r[nz] - 1.0)) <= tol, (k, np.max(np.abs(g[nz] / r[nz] - 1.0)))
    assert np.any(np.asarray(ref["H1i_rec"]) > 0.0)


def test_committed_golden_solves_match_the_c_oracle(orc):
    """tests/golden/second_oracle_solves.npz (whole solves by the numpy derivation, float64):
    the C oracle on the stored inputs gives the same sweep counts and states -- the fixture the
    GPU tests compare the CUDA path with (test_zz_gpu_second_oracle_golden.py)."""
    from helpers import golden_cfg, golden_diff, load_second_oracle_golden, run_oracle
    cases = load_second_oracle_golden()
    assert len(cases) == 9
    for name, g in cases.items():
        ref = run_oracle(orc, golden_cfg(g))
        assert ref["iters"] == int(g["iters"]), name
        assert golden_diff(ref, g, RT) <= 1e-11, name
        assert golden_diff(ref, g, RC) <= 1e-10, name
        assert golden_diff(ref, g, FR) <= 1e-10, name
        assert golden_diff(ref, g, ("T",)) <= 1e-13, name
