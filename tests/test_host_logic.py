"""Host-side logic that needs no GPU: input producers, configuration builders,
solver-class argument handling."""
import numpy as np
import pytest

from rabacus_b200 import configs, rad_src


def test_photon_grid_is_segmented_at_thresholds():
    E, idx = rad_src.photon_energies(1.0, 400.0, 128)
    assert E[0] == 13.6 and abs(E[-1] / (13.6 * 400) - 1) < 1e-12
    assert E[idx["H1"]] == 13.6 and E[idx["He1"]] == 24.59 and abs(E[idx["He2"]] - 54.42) < 1e-12
    assert np.all(np.diff(E) > 0)
    with pytest.raises(ValueError):
        rad_src.photon_energies(1.5, 400.0, 64)      # must cover the H I threshold
    with pytest.raises(ValueError):
        rad_src.photon_energies(1.0, 3.0, 64)        # must cover the He II threshold


def test_hm12_table_interpolation_is_exact_at_table_redshifts():
    tab = rad_src.HM12Table()
    E = rad_src.C_CM_S * rad_src.H_EV_S / (10 ** tab.log_lam_cm[100:110])  # table wavelengths
    for iz in (0, 17, 40):
        got = tab.spectrum_E(tab.z[iz] + 1e-12, E)
        assert np.allclose(np.log10(got), tab.logInu[100:110, iz], atol=1e-8)


def test_sources_and_normalisation():
    pt = rad_src.PointSource(1.0, 1.0, "monochromatic")
    pt.normalize_Ln(5.0e48)
    assert pt.Nnu == 1 and pt.monochromatic
    Ln = pt.shape[0] / pt.E[0] * rad_src.ERG_2_EV
    assert abs(Ln / 5.0e48 - 1) < 1e-14
    pl = rad_src.PlaneSource(1.0, 400.0, "hm12", Nnu=64, z=3.0)
    bg = rad_src.BackgroundSource(1.0, 400.0, "hm12", Nnu=64, z=3.0)
    assert np.allclose(pl.shape, 4 * np.pi * bg.shape)           # plane.py:312-313
    assert np.allclose(pl.thin_rates(), bg.thin_rates(), rtol=1e-14)
    pl.normalize_H1i(1e-12)
    assert abs(pl.thin_rates()[0] / 1e-12 - 1) < 1e-13
    th = rad_src.BackgroundSource(1.0, 400.0, "thermal", Nnu=32, T_eff=1e5)
    pw = rad_src.BackgroundSource(1.0, 400.0, "powerlaw", Nnu=32, alpha=-1.5)
    assert th.shape.argmax() < 16 and abs(pw.shape[-1] / 400 ** -1.5 - 1) < 1e-12
    with pytest.raises(ValueError):
        rad_src.BackgroundSource(1.0, 2.0, "monochromatic")


@pytest.mark.parametrize("name,Nl", [("c1", 128), ("c2", 256), ("c3", 1024), ("c4", 4096)])
def test_named_configurations(name, Nl):
    cfg = {"c1": configs.c1_iliev_stromgren, "c2": configs.c2_slab_plane,
           "c3": configs.c3_slab_bgnd, "c4": configs.c4_sphere_bgnd}[name]()
    assert cfg["nH"].size == Nl and cfg["Edges"].size == Nl + 1
    assert cfg["E_eV"].size == cfg["shape"].size
    assert np.all(cfg["shape"] > 0) and np.all(cfg["nH"] > 0)
    if name == "c4":
        assert cfg["E_eV"].size == 512 and cfg["Nmu"] == 256
        assert abs(cfg["nH"][0] - 10.0) < 1e-12 and cfg["nH"][-1] < 1e-2     # r^-2 envelope


def test_parameter_sweep_builder():
    probs = configs.c5_sweep(n_nH=2, n_R=2, n_z=3, Nl=16, Nnu=32, Nmu=4)
    assert len(probs) == 12
    assert len({p["z"] for p in probs}) == 3 and len({p["Edges"][-1] for p in probs}) == 2


def test_solver_classes_validate_arguments_before_calling_the_library():
    import rabacus_b200 as ra
    src = ra.BackgroundSource(1.0, 400.0, "hm12", Nnu=32, z=3.0)
    Nl = 8
    Ed, T, nH = np.linspace(0, 1e22, Nl + 1), np.full(Nl, 1e4), np.full(Nl, 1e-3)
    with pytest.raises(ValueError, match="source type"):
        ra.SphereStromgren(Ed, T, nH, nH * 0.1, src)
    with pytest.raises(ValueError, match="rec_meth"):
        ra.SphereBgnd(Ed, T, nH, nH * 0.1, src, rec_meth="thresh")
    with pytest.raises(ValueError, match="redshift"):
        ra.SphereBgnd(Ed, T, nH, nH * 0.1, src, find_Teq=True)
    with pytest.raises(AssertionError):
        ra.SphereBgnd(Ed[:-1], T, nH, nH * 0.1, src)
    with pytest.raises(ValueError, match="even"):
        ra.Slab2Bgnd(Ed[:-1], T[:-1], nH[:-1], nH[:-1] * 0.1, src)


# ----------------------------------------------------------------- rb_fastpow.cuh (host build)
def _diag_pow(x, y, on_device=0):
    import ctypes as C
    from rabacus_b200 import _lib
    out = np.empty_like(x)
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    rc = _lib.lib().rbi_diag_pow(P(x), P(y), C.c_int(x.size), P(out), C.c_int(on_device))
    assert rc == 0
    return out


def _diag_hg97(T, f1, f2, f3, on_device=0, libm=0):
    import ctypes as C
    from rabacus_b200 import _lib
    out = np.empty((16, T.size))
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    rc = _lib.lib().rbi_diag_hg97(P(T), P(f1), P(f2), P(f3), C.c_int(T.size), P(out),
                                  C.c_int(on_device), C.c_int(libm))
    assert rc == 0
    return out


def fastpow_samples(n, seed=1):
    """Bases and exponents of the Hui & Gnedin fits: lam = 2 T_X / T, 1 + (lam/t3)^t4,
    T, values near 1; exponents in [-4, 4]."""
    rng = np.random.default_rng(seed)
    x = np.concatenate([10 ** rng.uniform(-6, 6.5, n // 2), 1 + 10 ** rng.uniform(-8, 7, n // 4),
                        rng.uniform(0.5, 2.0, n - n // 2 - n // 4)])
    return x, rng.uniform(-4, 4, n)


def test_fastpow_is_within_one_ulp_of_libm_and_half_an_ulp_of_the_truth():
    """x**y by exp_dd(y log_dd(x)) (csrc/rb_fastpow.cuh): <= 1 ulp from glibc's pow (what the
    reference's gfortran `**` calls) on 2e5 samples, and <= 0.53 ulp from the exact value
    (mpmath, 200 bits) on 3000 of them."""
    import mpmath as mp
    x, y = fastpow_samples(200000)
    out = _diag_pow(x, y)
    ref = np.power(x, y)
    ulp = np.spacing(np.abs(ref))
    assert np.max(np.abs(out - ref) / ulp) <= 1.0
    mp.mp.prec = 200
    worst = 0.0
    for i in range(0, x.size, x.size // 3000):
        t = mp.power(mp.mpf(float(x[i])), mp.mpf(float(y[i])))
        worst = max(worst, float(abs((mp.mpf(float(out[i])) - t) / mp.mpf(float(ulp[i])))))
    assert worst <= 0.53, worst


def test_fastpow_special_arguments_fall_back_to_libm():
    x = np.array([0.0, 5e-324, 1e-310, np.inf, 2.0, 2.0, 1e300, 1e-300, np.nan, 1.0])
    y = np.array([2.0, 0.5, 3.0, -1.0, 2000.0, -2000.0, 3.0, 3.0, 1.0, 1e300])
    out = _diag_pow(x, y)
    with np.errstate(all="ignore"):
        ref = np.power(x, y)
    assert np.array_equal(out, ref, equal_nan=True)


def test_hg97_fast_evaluation_matches_the_libm_statement_form(orc):
    """get_kchem_s / get_kcool_s with the shared sub-expressions and the table-driven powers
    against the oracle's statement-by-statement libm form (chem_cool_rates.f90:45-204):
    <= 2e-14 (the case A/B mixing 10**(f log10 a + ..) amplifies one ulp of its argument
    by |log a| ~ 30), over the solver's temperature range and far outside it (T < 1 K and
    T > 1e12 K take the libm path and agree exactly)."""
    rng = np.random.default_rng(2)
    n = 20000
    T = 10 ** rng.uniform(0.0, 12.0, n)
    T[: n // 2] = 10 ** rng.uniform(np.log10(7.5e3), 5.0, n // 2)
    T[-4:] = [0.5, 1e13, 1e-3, 3e14]
    f = [rng.choice([0.0, 1.0, 0.3, 0.77], n) for _ in range(3)]
    out = _diag_hg97(T, *f)
    ref = np.array(list(orc.get_kchem(T, *f)) + list(orc.get_kcool(T, *f)))
    with np.errstate(all="ignore"):
        rel = np.where(out == ref, 0.0, np.abs(out - ref) / np.abs(ref))
    assert np.nanmax(rel) <= 2e-14
    assert np.array_equal(out[:, -4:], ref[:, -4:], equal_nan=True)


def test_hg97_libm_statement_form_is_bit_identical_to_the_oracle(orc):
    """libm=1 selects the statement-by-statement libm form of the fits; built by the host
    compiler it calls the same glibc pow/exp/log10 as the oracle's C and must agree with
    it bit for bit -- the two codings of chem_cool_rates.f90:45-204 pin each other."""
    rng = np.random.default_rng(3)
    n = 5000
    T = 10 ** rng.uniform(2.0, 9.0, n)
    f = [rng.choice([0.0, 1.0, 0.5], n) for _ in range(3)]
    out = _diag_hg97(T, *f, libm=1)
    ref = np.array(list(orc.get_kchem(T, *f)) + list(orc.get_kcool(T, *f)))
    assert np.array_equal(out, ref)


# ----------------------------------------------------------------- slab post-processing
def test_mean_molecular_weight_limits():
    """jeans.py:45-64 with the wrapper's default Yp = 0.248: neutral gas 4/(4 - 3 Yp),
    fully ionised 4/(8 - 5 Yp)."""
    from rabacus_b200 import post
    Yp = 0.248
    assert post.mean_molecular_weight(0.0, 0.0, 0.0) == 4.0 / (4.0 - 3.0 * Yp)
    assert abs(post.mean_molecular_weight(1.0, 0.0, 1.0) - 4.0 / (8.0 - 5.0 * Yp)) < 1e-15
    mu = post.mean_molecular_weight(np.array([0.0, 0.5, 1.0]), np.zeros(3), np.zeros(3))
    assert np.all(np.diff(mu) < 0)


def test_slab_averages_follow_the_reference_definitions():
    """slab_base.py:120-283, 369-431: <q>_w = trap(z_c, w q) / trap(z_c, w) for the six
    weightings and the central-layer sample, on a synthetic slab."""
    from rabacus_b200 import post

    class Obj(object):
        pass
    o = Obj()
    Nl = 16
    rng = np.random.default_rng(4)
    o.Nl = Nl
    Edges = np.linspace(0.0, 3.0, Nl + 1) ** 1.3
    o.dz = Edges[1:] - Edges[:-1]
    o.z_c = Edges[:-1] + 0.5 * o.dz
    for name in set(post.AVERAGED_FIELDS + post.CENTRAL_FIELDS):
        setattr(o, name, rng.uniform(0.1, 1.0, Nl))
    o.T = np.full(Nl, 1.0e4)                      # a constant field averages to itself
    avg = post.slab_averages(o)
    for w in ("v_w", "nH_w", "nHe_w", "nH1_w", "nHe1_w", "nHe2_w"):
        assert abs(getattr(avg, w).T / 1.0e4 - 1.0) < 1e-14
    wa = o.nH * o.xH1
    ref = rad_src.trap(o.z_c, wa * o.H1i_src) / rad_src.trap(o.z_c, wa)
    assert avg.nH1_w.H1i_src == ref
    lo, hi = o.xHe2.min(), o.xHe2.max()
    assert lo <= avg.v_w.xHe2 <= hi
    assert avg.cen.ne == o.ne[Nl // 2] and avg.cen.He2h_rec == o.He2h_rec[Nl // 2]


def test_slab_wrapper_post_processing_sets_reference_attributes(monkeypatch):
    """The slab wrappers' __post__ attributes (slab_base.py:286-437) from a synthetic solver
    result; the GPU-backed ChemistryRates is stubbed so that this runs without a device."""
    from rabacus_b200 import solvers

    class KA(object):
        def __init__(self, T, *a, **k):
            for n in ("reH2", "reHe2", "reHe3", "ciH1", "ciHe1", "ciHe2"):
                setattr(self, n, np.full(np.size(T), 1.0e-13))
    monkeypatch.setattr(solvers, "ChemistryRates", KA)
    Nl = 12
    rng = np.random.default_rng(6)
    s = object.__new__(solvers.Slab2Bgnd)
    s.Nl = Nl
    s.Edges = np.linspace(0.0, 1.0e21, Nl + 1)
    s.nH = np.full(Nl, 1.0e-2); s.nHe = np.full(Nl, 1.0e-3); s.T = np.full(Nl, 1.0e4)
    s.atomic_fit_name = "hg97"
    out = [rng.uniform(0.05, 0.9, Nl) for _ in range(17)]
    monkeypatch.setattr(solvers.rabacus_fc, "last_info", {"iters": 7}, raising=False)
    s._attach(out)
    s._post()
    assert s.itr == 7
    assert np.array_equal(s.dz, s.Edges[1:] - s.Edges[:-1]) and s.z_c.size == Nl
    assert s.mu.shape == (Nl,) and np.all(s.mu > 0.5) and np.all(s.mu < 1.3)
    assert np.isclose(s.logNH1_thru, np.log10(np.sum(s.dz * s.nH * s.xH1)))
    for w in ("v_w", "nH_w", "nHe_w", "nH1_w", "nHe1_w", "nHe2_w"):
        assert np.isfinite(getattr(s.avg, w).H1h_rec)
    assert s.avg.cen.xH1 == s.xH1[Nl // 2]
