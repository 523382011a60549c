"""GPU parity tests: the CUDA hot path (through the C ABI) against the CPU oracle
on identical inputs.  FP64 tolerance stated by BASELINE.json's north_star:
relative error <= 1e-10.

Tolerances (helpers.compare): photo-ionisation / heating rates and T: 1e-10
relative (measured <= 1e-12; T bit-exact); sweep counts: equal; ionisation
fractions: the oracle's scalar solvers, fed with the CUDA path's photo-rates and the
device's rate-coefficient bits, must reproduce the CUDA fractions BIT FOR BIT in every
cell (helpers.solver_check explains why an end-to-end bound on minority fractions would
measure last-bit noise of the reference algorithm instead).

Protocol:
  (i)   sweep level : same input state -> 6 rates, x, T agree <= 1e-10
  (ii)  solve level : same order and tol -> same sweep count and <= 1e-10
  (iii) SphereBgnd vs the reference's Gauss-Seidel order: both run to a tight
        tolerance so they sit on the common fixed point.
"""
import numpy as np
import pytest

from helpers import CHECKED, CHECKED_REC, compare, frac_rtol, run_gpu, run_oracle
from rabacus_b200 import configs

pytestmark = pytest.mark.gpu

RTOL = 1.0e-10


def _small_sphere(Nl=96, Nnu=128, Nmu=16, find_Teq=False, nH0=1.0, spec=None):
    return configs.c4_sphere_bgnd(Nl, Nnu, Nmu, find_Teq, nH0=nH0, spectrum=spec)


# ----------------------------------------------------------------- single zone
def test_kchem_kcool_match_oracle(orc):
    from rabacus_b200 import rabacus_fc as fc
    T = np.logspace(np.log10(7.5e3), 5.0, 257)
    for fcA in (1.0, 0.0, 0.37):
        got = fc.get_kchem_v(T, fcA, fcA, fcA, 1, T.size)
        ref = orc.get_kchem(T, fcA, fcA, fcA)
        for g, r in zip(got, ref):
            assert np.max(np.abs(g / r - 1)) < 1e-13
        got = fc.get_kcool_v(T, fcA, fcA, fcA, 1, T.size)
        ref = orc.get_kcool(T, fcA, fcA, fcA)
        for g, r in zip(got, ref):
            assert np.max(np.abs(g / r - 1)) < 1e-13


def _zone_inputs(n=257):
    T = np.logspace(4.0, 5.0, n)
    nH = np.logspace(-5.0, 1.0, n)
    nHe = nH * 0.25 * 0.24 / 0.76
    H1i, He1i, He2i = 8.0e-13, 5.0e-13, 6.0e-15          # ~HM12 z=3 photo-rates
    H1h, He1h, He2h = 5.0e-24, 6.0e-24, 2.0e-25
    return T, nH, nHe, (H1i, He1i, He2i), (H1h, He1h, He2h)


@pytest.mark.parametrize("lockstep", [False, True])
def test_solve_pce_matches_oracle(orc, lockstep):
    from rabacus_b200 import rabacus_fc as fc
    T, nH, nHe, ion, _ = _zone_inputs()
    k = orc.get_kchem(T, 1.0, 1.0, 1.0)
    n = T.size
    f = fc.solve_pce_v if lockstep else None
    if lockstep:
        got = f(nH, nHe, *k, *ion, 1e-8, n)
    else:
        got = [np.empty(n) for _ in range(5)]
        for i in range(n):
            r = fc.solve_pce_s(nH[i], nHe[i], *[a[i] for a in k], *ion, 1e-8)
            for q in range(5):
                got[q][i] = r[q]
    ref = orc.solve_pce(nH, nHe, *k, *ion, 1e-8, vector=lockstep)
    for g, r in zip(got, ref):
        assert np.max(np.abs(g / r - 1)) < RTOL


@pytest.mark.parametrize("lockstep", [False, True])
def test_solve_pcte_matches_oracle(orc, lockstep):
    from rabacus_b200 import rabacus_fc as fc
    _, nH, nHe, ion, heat = _zone_inputs(129)
    n = nH.size
    if lockstep:
        got = fc.solve_pcte_v(nH, nHe, *ion, *heat, 3.0, 0.0, 1.0, 1.0, 1.0, 1, 1e-7, n)
    else:
        from rabacus_b200.rabacus_fc import _solve_pcte
        got = _solve_pcte((nH, nHe) + ion + heat, 3.0, 0.0, (1.0, 1.0, 1.0), 1, 1e-7, n, 0)
    ref = orc.solve_pcte(nH, nHe, *ion, *heat, 3.0, 0.0, 1.0, 1.0, 1.0, 1e-7, vector=lockstep)
    for g, r in zip(got, ref):
        assert np.max(np.abs(g / r - 1)) < RTOL


@pytest.mark.parametrize("n", [513, 600, 2100, 8192])
def test_lockstep_solvers_many_zones_per_thread(orc, n):
    """The `_v` twins with more zones than one block has threads (4 and 16 zones
    per thread): every zone must walk the oracle's lock-step path."""
    from rabacus_b200 import rabacus_fc as fc
    nH = np.logspace(-4.0, -1.0, n)
    nHe = nH * 0.08
    ion, heat = (8e-13, 5e-13, 6e-15), (5e-24, 6e-24, 2e-25)
    k = orc.get_kchem(np.full(n, 1.0e4), 1.0, 1.0, 1.0)
    got = fc.solve_pce_v(nH, nHe, *k, *ion, 1e-8, n)
    ref = orc.solve_pce(nH, nHe, *k, *ion, 1e-8, vector=True)
    for g, r in zip(got, ref):
        assert np.max(np.abs(g / r - 1)) < RTOL
    got = fc.solve_pcte_v(nH, nHe, *ion, *heat, 3.0, 0.0, 1.0, 1.0, 1.0, 1, 1e-7, n)
    ref = orc.solve_pcte(nH, nHe, *ion, *heat, 3.0, 0.0, 1.0, 1.0, 1.0, 1e-7, vector=True)
    assert np.array_equal(got[5], ref[5])          # T: same bisection path
    for g, r in zip(got[:5], ref[:5]):
        assert np.max(np.abs(g / r - 1)) < frac_rtol(1e-7 / 1e-2)


@pytest.mark.parametrize("G", [1, 4, 8, 16, 32])
def test_speculative_T_bisection_equals_serial(orc, spec128, G, monkeypatch):
    """k_cell_solve_spec (G lanes per cell probing the next log2 G bisection levels
    at once) must return the serial solve_pcte_s result: T bit-identical to the
    oracle for every group size, fractions within the solver tolerance policy."""
    monkeypatch.setenv("RB_SPEC_G", str(G))
    cfg = _small_sphere(96, 128, 8, True, spec=spec128)
    for kw in (dict(max_sweeps=1), dict(max_sweeps=4)):
        g = run_gpu(cfg, **kw)
        o = run_oracle(orc, cfg, **kw)
        assert np.array_equal(g["T"], o["T"])
        worst, bad = compare(g, o)
        assert not bad, (G, kw, bad)
    cfg = configs.c2_slab_plane(64, 128, True, True)   # floor cells + mirrored halves
    g = run_gpu(cfg)
    o = run_oracle(orc, cfg)
    assert g["iters"] == o["iters"]
    assert np.array_equal(g["T"], o["T"])
    worst, bad = compare(g, o)
    assert not bad, (G, bad)


def test_speculative_ne_bisection_is_bit_identical_to_serial(orc, spec128, monkeypatch):
    """k_cell_solve_pce_spec (fixed T: G lanes per cell probing the next log2 G steps of
    the n_e bisection at once) against the one-thread-per-cell kernel: every output
    bit-identical for every group size, on a sphere, a one-sided slab and a mirrored one."""
    cases = [_small_sphere(96, 128, 8, False, spec=spec128),
             configs.c2_slab_plane(64, 128, False, False),
             configs.c3_slab_bgnd(64, 128, False)]
    for cfg in cases:
        monkeypatch.setenv("RB_SPEC_G", "1")
        ref = run_gpu(cfg)
        o = run_oracle(orc, cfg)
        assert ref["iters"] == o["iters"]
        worst, bad = compare(ref, o)
        assert not bad, bad
        for G in (4, 8, 16, 32):
            monkeypatch.setenv("RB_SPEC_G", str(G))
            g = run_gpu(cfg)
            assert g["iters"] == ref["iters"], G
            for n in CHECKED:
                assert np.array_equal(g[n], ref[n]), (cfg["geometry"], G, n)


@pytest.mark.parametrize("Nl", [700, 2100, 4500])
@pytest.mark.parametrize("find_Teq", [False, True])
def test_thin_state_across_a_cluster(orc, spec128, Nl, find_Teq):
    """Optically thin start with the lock-step solvers spread over a thread-block
    cluster (2, 8 CTAs; 2 cells per thread at 4500): any() must span all cells."""
    cfg = _small_sphere(Nl, 128, 4, find_Teq, spec=spec128)
    g = run_gpu(cfg, thin=True)
    o = run_oracle(orc, cfg, thin=True)
    assert np.array_equal(g["T"], o["T"])
    worst, bad = compare(g, o)
    assert not bad, bad


def test_solver_bit_identical_given_same_coefficients(orc):
    """With identical rate coefficients the CUDA bisection walks exactly the
    oracle's path: fractions are bit-identical (no libm call inside solve_pce)."""
    from rabacus_b200.rabacus_fc import _solve_pce
    T, nH, nHe, ion, _ = _zone_inputs(513)
    nH = nH * 10.0  # include neutral, collisionally dominated zones
    k = orc.get_kchem(T, 1.0, 1.0, 1.0)
    got = _solve_pce((nH, nHe) + tuple(k) + ion, 1e-6, T.size, 0)
    ref = orc.solve_pce(nH, nHe, *k, *ion, 1e-6)
    for g, r in zip(got, ref):
        assert np.array_equal(g, r)


# ----------------------------------------------------------------- SphereBgnd
@pytest.mark.parametrize("Nmu", [8, 16, 5, 33])
@pytest.mark.parametrize("find_Teq", [False, True])
def test_sphere_bgnd_thin_and_one_sweep(orc, spec128, Nmu, find_Teq):
    cfg = _small_sphere(96, 128, Nmu, find_Teq, spec=spec128)
    for kw in (dict(thin=True), dict(max_sweeps=1), dict(max_sweeps=3)):
        g = run_gpu(cfg, **kw)
        o = run_oracle(orc, cfg, **kw)
        worst, bad = compare(g, o)
        assert not bad, (kw, bad)


def test_sphere_bgnd_decreasing_edges_same_as_increasing(orc, spec128):
    cfg = _small_sphere(64, 128, 8, spec=spec128)
    rev = dict(cfg)
    for k in ("Edges", "nH", "nHe", "T"):
        rev[k] = cfg[k][::-1].copy()
    a = run_gpu(cfg, max_sweeps=2)
    b = run_gpu(rev, max_sweeps=2)
    for n in CHECKED:
        assert np.array_equal(a[n], b[n][::-1])


@pytest.mark.parametrize("find_Teq", [False, True])
def test_sphere_bgnd_full_solve_same_iterations(orc, spec128, find_Teq):
    cfg = _small_sphere(128, 128, 16, find_Teq, nH0=3.0, spec=spec128)
    g = run_gpu(cfg, tol=1e-4)
    o = run_oracle(orc, cfg, tol=1e-4)
    assert g["iters"] == o["iters"]
    worst, bad = compare(g, o)
    assert not bad, bad


def test_sphere_bgnd_vs_reference_gauss_seidel_fixed_point(orc, spec128):
    """(iii): Jacobi on the GPU and the reference's sequential Gauss-Seidel order
    converge to the same fixed point; compared at a tight tolerance."""
    cfg = _small_sphere(48, 128, 8, False, nH0=0.3, spec=spec128)
    g = run_gpu(cfg, tol=1e-12)
    o = run_oracle(orc, cfg, tol=1e-12, order=orc.ORDER_GS)
    worst, bad = compare(g, o, rtol=1e-9, tol=1e-12)
    assert not bad, bad
    assert max(worst[n] for n in ("xH1", "xH2", "xHe1", "xHe2", "xHe3")) <= 1e-8
    # and the looser O(tol) agreement at the default tolerance
    g = run_gpu(cfg, tol=1e-4)
    o = run_oracle(orc, cfg, tol=1e-4, order=orc.ORDER_GS)
    worst, bad = compare(g, o, rtol=5e-3)
    assert not bad, bad
    assert max(worst[n] for n in ("xH1", "xH2", "xHe1", "xHe2", "xHe3")) <= 5e-3


def test_sphere_bgnd_hm12_512_frequencies(orc):
    cfg = configs.c4_sphere_bgnd(128, 512, 32, False, nH0=1.0)
    g = run_gpu(cfg, max_sweeps=2)
    o = run_oracle(orc, cfg, max_sweeps=2)
    worst, bad = compare(g, o)
    assert not bad, bad


def test_c4_full_size_sweeps_match_oracle(orc):
    """BASELINE.json configs[3] at full size (4096 shells x 512 frequencies x 256 rays):
    optically thin start + two Jacobi sweeps against the oracle (a few seconds of
    OpenMP on the host), then size-independent properties of the converged solve:
    every fraction in [0, 1] and closed per element, neutral fraction rising inwards
    through the self-shielding transition, source rates bounded by the thin rates."""
    cfg = configs.c4_sphere_bgnd(4096, 512, 256, False)
    g = run_gpu(cfg, max_sweeps=2)
    o = run_oracle(orc, cfg, max_sweeps=2)
    worst, bad = compare(g, o)
    assert not bad, bad
    thin = run_gpu(cfg, thin=True)
    full = run_gpu(cfg)
    assert 10 < full["iters"] < 500
    for a, b_, c in (("xH1", "xH2", None), ("xHe1", "xHe2", "xHe3")):
        tot = full[a] + full[b_] + (full[c] if c else 0.0)
        assert np.max(np.abs(tot - 1.0)) < 1e-12
        assert np.all(full[a] >= 0.0) and np.all(full[a] <= 1.0)
    # Edges increase outwards in this config: the centre is shielded, the surface is not
    assert full["xH1"][0] > 0.9 and full["xH1"][-1] < 1e-2
    for n in ("H1i_src", "He1i_src", "He2i_src", "H1h_src", "He1h_src", "He2h_src"):
        # the fixed point of src <- (src + sum_mu w rate)/2 is sum_mu w rate
        # = 2 x the angle mean, so the bound is twice the optically thin rate
        assert np.all(full[n] <= 2.0 * thin[n] * (1 + 1e-12)), n
        assert np.all(full[n] >= 0.0)


def test_one_shot_entry_point_matches_plan(spec128):
    from rabacus_b200 import rabacus_fc as fc
    cfg = _small_sphere(64, 128, 8, True, spec=spec128)
    T = cfg["T"].copy()
    Ed, nH, nHe = cfg["Edges"].copy(), cfg["nH"].copy(), cfg["nHe"].copy()
    out = fc.sphere_bgnd.sphere_bgnd_solve(Ed, nH, nHe, T, 8, cfg["E_eV"], cfg["shape"], 1, 1.0,
                                           1, 1, 1, 0, 1.0, 1.0, 1.0, 3.0, 0.0, 1e-4, 64, 128)
    g = run_gpu(cfg, tol=1e-4)
    assert np.array_equal(out[0], g["xH1"])
    assert np.array_equal(T, g["T"])          # intent(inout) T carries Teq
    assert np.array_equal(Ed, cfg["Edges"])   # restored orientation
    assert fc.last_info["iters"] == g["iters"]
    for q in (8, 9, 10, 14, 15, 16):
        assert not out[q].any()               # *_rec are zero for rec_meth=fixed


# ----------------------------------------------------------------- other geometries
def test_c1_iliev_stromgren_sphere(orc):
    cfg = configs.c1_iliev_stromgren(128)
    g = run_gpu(cfg, tol=1e-4)
    o = run_oracle(orc, cfg, tol=1e-4)
    assert g["iters"] == o["iters"]
    worst, bad = compare(g, o)
    assert not bad, bad
    # Stromgren radius for this test problem is ~5.4 kpc: neutral outside
    assert g["xH1"][0] < 1e-2 and g["xH1"][-1] > 0.5


@pytest.mark.parametrize("find_Teq", [False, True])
@pytest.mark.parametrize("two_sided", [False, True])
def test_c2_slab_plane(orc, find_Teq, two_sided):
    cfg = configs.c2_slab_plane(256, 128, find_Teq, two_sided)
    g = run_gpu(cfg, tol=1e-4)
    o = run_oracle(orc, cfg, tol=1e-4)
    assert g["iters"] == o["iters"]
    worst, bad = compare(g, o)
    assert not bad, bad


@pytest.mark.parametrize("find_Teq", [False, True])
def test_c3_slab_bgnd(orc, find_Teq):
    cfg = configs.c3_slab_bgnd(256, 128, find_Teq)
    g = run_gpu(cfg, tol=1e-4)
    o = run_oracle(orc, cfg, tol=1e-4)
    assert g["iters"] == o["iters"]
    worst, bad = compare(g, o)
    assert not bad, bad
    assert np.array_equal(g["xH1"], g["xH1"][::-1])  # mirrored


def test_c3_slab_bgnd_full_size_one_sweep(orc):
    cfg = configs.c3_slab_bgnd(1024, 512, False)
    g = run_gpu(cfg, max_sweeps=2)
    o = run_oracle(orc, cfg, max_sweeps=2)
    worst, bad = compare(g, o)
    assert not bad, bad


def test_slab_asymmetric_density_and_odd_layers(orc, spec128):
    """nH need not be symmetric; the mirror is unconditional and an odd middle
    layer is never updated."""
    cfg = configs.c3_slab_bgnd(65, 128, False)
    cfg["E_eV"], cfg["shape"] = spec128
    cfg["nH"] = cfg["nH"] * np.linspace(0.5, 3.0, 65)
    cfg["nHe"] = cfg["nH"] * 0.08
    g = run_gpu(cfg, max_sweeps=3)
    o = run_oracle(orc, cfg, max_sweeps=3)
    worst, bad = compare(g, o)
    assert not bad, bad


def test_set_column_vs_impact(orc, spec128):
    from rabacus_b200 import rabacus_fc as fc
    cfg = _small_sphere(200, 128, 8, spec=spec128)
    g = run_gpu(cfg, max_sweeps=2)
    args = (g["xH1"], g["xH2"], g["xHe1"], g["xHe2"], g["xHe3"], cfg["Edges"], cfg["nH"],
            cfg["nHe"], cfg["T"])
    got = fc.sphere_base.set_column_vs_impact(*args, 200)
    ref = orc.set_column_vs_impact(*args)
    for a, b in zip(got, ref):
        assert np.max(np.abs(a / b - 1)) < RTOL


# ----------------------------------------------------------------- errors, batches
def test_errors_are_exceptions(spec128):
    from rabacus_b200 import RabacusB200Error, rabacus_fc as fc
    cfg = _small_sphere(16, 128, 4, spec=spec128)
    bad = cfg["Edges"].copy()
    bad[5] = bad[7]
    bad[6] = bad[7] * 2
    args = lambda Ed, rec: (Ed, cfg["nH"].copy(), cfg["nHe"].copy(), cfg["T"].copy(), 4,
                            cfg["E_eV"], cfg["shape"], rec, 1.0, 1, 1, 0, 0, 1.0, 1.0, 1.0,
                            0.0, 0.0, 1e-4, 16, 128)
    with pytest.raises(RabacusB200Error) as e:
        fc.sphere_bgnd.sphere_bgnd_solve(*args(bad, 1))
    assert e.value.code == 1
    with pytest.raises(RabacusB200Error) as e:
        fc.sphere_bgnd.sphere_bgnd_solve(*args(cfg["Edges"].copy(), 5))
    assert e.value.code == 8
    # the hollow sphere also trips the ray geometry of the isotropic photon transfer
    with pytest.raises(RabacusB200Error) as e:
        fc.sphere_bgnd.sphere_bgnd_solve(*args(np.linspace(1.0, 2.0, 17) * 3e21, 4))
    assert e.value.code == 7
    # hollow sphere: innermost edge > impact parameter of some ray (geometry.f90:61-71)
    hollow = np.linspace(1.0, 2.0, 17) * 3e21
    with pytest.raises(RabacusB200Error) as e:
        fc.sphere_bgnd.sphere_bgnd_solve(*args(hollow, 1))
    assert e.value.code == 7


def test_batched_problems_match_individual_solves(spec128):
    """Independent problems in one plan (the C5 parameter-sweep path): every
    problem stops at its own sweep count and equals its stand-alone solve."""
    from rabacus_b200.plan import Plan
    cfgs = [_small_sphere(64, 128, 8, False, nH0=n0, spec=spec128) for n0 in (0.05, 0.5, 5.0)]
    cfgs[1]["Edges"] = cfgs[1]["Edges"] * 0.5
    plan = Plan(0, len(cfgs), 64, 128, 8, tol=1e-4)
    for b, c in enumerate(cfgs):
        plan.set_problem(b, c["Edges"], c["nH"], c["nHe"], c["T"], c["E_eV"], c["shape"])
    plan.upload()
    plan.solve()
    its = []
    for b, c in enumerate(cfgs):
        r = plan.get_problem(b)
        s = run_gpu(c, tol=1e-4)
        its.append(r["iters"])
        assert r["iters"] == s["iters"]
        for n in CHECKED:
            assert np.array_equal(r[n], s[n]), n
    assert len(set(its)) > 1  # they really did stop at different sweeps
    plan.close()


@pytest.mark.parametrize("find_Teq", [False, True])
def test_batch_tail_launch_shapes_do_not_change_results(spec128, find_Teq):
    """48 problems whose sweep counts differ widely: while they drop out, the launch shapes
    follow the number still iterating (live_problems in rb_api.cu: lanes per cell, CTAs and
    chunks per problem, cells per unit; graphs re-captured).  Every problem must still equal
    its stand-alone solve bit for bit, at its own sweep count."""
    from rabacus_b200.plan import Plan
    nH0s = np.logspace(-1.5, 1.5, 48)
    cfgs = [_small_sphere(64, 128, 8, find_Teq, nH0=float(n0), spec=spec128) for n0 in nH0s]
    plan = Plan(0, len(cfgs), 64, 128, 8, find_Teq=find_Teq, tol=1e-4)
    for b, c in enumerate(cfgs):
        plan.set_problem(b, c["Edges"], c["nH"], c["nHe"], c["T"], c["E_eV"], c["shape"], c["z"])
    plan.upload()
    plan.solve()
    its = []
    for b in range(0, len(cfgs), 5):
        r = plan.get_problem(b)
        s = run_gpu(cfgs[b], tol=1e-4)
        its.append(r["iters"])
        assert r["iters"] == s["iters"], b
        for n in CHECKED:
            assert np.array_equal(r[n], s[n]), (b, n)
    assert max(its) >= 1.5 * min(its)   # a real tail
    plan.close()


def test_cell_sharded_sweeps_equal_single_sweep(spec128):
    """The multi-GPU decomposition (cells dealt cyclically) emulated on
    one GPU: sweeping the shards one after another equals one full sweep."""
    from rabacus_b200.plan import Plan
    cfg = _small_sphere(96, 128, 8, True, spec=spec128)
    outs = []
    for world in (1, 4):
        plan = Plan(0, 1, 96, 128, 8, find_Teq=True, tol=1e-4)
        plan.set_problem(0, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"],
                         cfg["shape"], 3.0, 0.0)
        plan.upload()
        plan.thin()
        for _ in range(3):
            # Jacobi: every shard must read the PRE-sweep state, so stage the
            # shard results and apply them together, as the all-gather does.
            if world == 1:
                plan.sweep_local(0, 1)
            else:
                import torch
                fields = [plan.field_tensor(i) for i in range(12)]
                plan.sync()
                before = [f.clone() for f in fields]
                staged = [f.clone() for f in fields]
                for r in range(world):
                    for f, b in zip(fields, before):
                        f.copy_(b)
                    torch.cuda.synchronize()
                    plan.sweep_local(r, world)
                    plan.sync()
                    for f, s in zip(fields, staged):
                        s[r::world] = f[r::world]
                for f, s in zip(fields, staged):
                    f.copy_(s)
                torch.cuda.synchronize()
            plan.finish_sweep()
        outs.append(plan.get_problem(0))
        plan.close()
    for n in CHECKED:
        assert np.array_equal(outs[0][n], outs[1][n]), n


# ----------------------------------------------------------------- recombination photons
@pytest.mark.parametrize("meth", [2, 3, 4])
@pytest.mark.parametrize("find_Teq", [False, True])
def test_sphere_bgnd_recombination_photons(orc, spec128, meth, find_Teq):
    """rec_meth = outward / radial / isotropic (sphere_base.f90:79-1083) inside the
    SphereBgnd sweep: case-A rates, total = source + recombination photo-rates."""
    cfg = _small_sphere(64, 128, 8, find_Teq, nH0=1.0, spec=spec128)
    cfg["i_rec_meth"] = meth
    cfg["em_fac"] = (1.0, 0.7, 1.3) if meth == 4 else (1.0, 1.0, 1.0)
    for kw in (dict(max_sweeps=1), dict(max_sweeps=3), dict()):
        g = run_gpu(cfg, **kw)
        o = run_oracle(orc, cfg, **kw)
        assert g["iters"] == o["iters"], kw
        worst, bad = compare(g, o, names=CHECKED_REC)
        assert not bad, (meth, kw, bad)
        assert float(np.max(o["H1i_rec"])) > 0.0


@pytest.mark.parametrize("meth", [2, 3, 4])
def test_stromgren_recombination_photons(orc, meth):
    """The same three methods on the innermost-first orientation (SphereStromgren),
    including input given with decreasing Edges."""
    cfg = configs.c1_iliev_stromgren(64)
    cfg["i_rec_meth"] = meth
    cfg["Nmu"] = 8
    g = run_gpu(cfg)
    o = run_oracle(orc, cfg)
    assert g["iters"] == o["iters"]
    worst, bad = compare(g, o, names=CHECKED_REC)
    assert not bad, (meth, bad)
    rev = dict(cfg)
    for k in ("Edges", "nH", "nHe", "T"):
        rev[k] = cfg[k][::-1].copy()
    g2 = run_gpu(rev)
    for n in CHECKED_REC:
        assert np.array_equal(g2[n][::-1], g[n]), n


@pytest.mark.parametrize("geom", ["slab_plane_1", "slab_plane_2", "slab_bgnd"])
@pytest.mark.parametrize("find_Teq", [False, True])
def test_slab_recombination_photons(orc, geom, find_Teq):
    """rec_meth = "ray" for the slabs (slab_base.f90:163-560): E1 kernels on the
    layer edges, trapezoid rule below and above every layer."""
    if geom == "slab_bgnd":
        cfg = configs.c3_slab_bgnd(96, 128, find_Teq)
    else:
        cfg = configs.c2_slab_plane(96, 128, find_Teq, geom == "slab_plane_2")
    cfg["i_rec_meth"] = 2
    for kw in (dict(max_sweeps=2), dict()):
        g = run_gpu(cfg, **kw)
        o = run_oracle(orc, cfg, **kw)
        assert g["iters"] == o["iters"], kw
        worst, bad = compare(g, o, names=CHECKED_REC)
        assert not bad, (geom, kw, bad)


def test_public_classes_with_recombination_photons(orc):
    """The wrapper classes (same constructor keywords as ra.SphereBgnd / ra.Slab2Bgnd)
    with rec_meth strings: results equal the oracle's, *_rec and the totals are set."""
    import rabacus_b200 as ra
    from rabacus_b200.rad_src import KPC_CM
    src = ra.BackgroundSource(1.0, 400.0, "hm12", Nnu=64, z=3.0)
    Nl = 48
    Edges = np.linspace(0.0, 12.0 * KPC_CM, Nl + 1)
    r = 0.5 * (Edges[1:] + Edges[:-1])
    nH = 0.5 * np.where(r < KPC_CM, 1.0, (r / KPC_CM) ** -2.0)
    nHe = nH * 0.08
    T = np.full(Nl, 1.0e4)
    s = ra.SphereBgnd(Edges, T, nH, nHe, src, rec_meth="isotropic", Nmu=8, em_He1_fac=0.5)
    o = orc.sphere_bgnd_solve(Edges, nH, nHe, T, 8, src.E, src.shape, i_rec_meth=4,
                              em_fac=(1.0, 0.5, 1.0), order=orc.ORDER_JACOBI)
    assert s.itr == o["iters"]
    got = {n: getattr(s, n) for n in CHECKED_REC if n != "T"}
    got["T"] = s.T
    worst, bad = compare(got, o, names=CHECKED_REC)
    assert not bad, bad
    assert np.array_equal(s.H1i, s.H1i_src + s.H1i_rec) and np.all(s.H1i_rec > 0.0)
    sl = ra.Slab2Bgnd(Edges, T, np.full(Nl, 5e-3), np.full(Nl, 4e-4), src, rec_meth="ray")
    o = orc.slab_bgnd_solve(Edges, np.full(Nl, 5e-3), np.full(Nl, 4e-4), T, src.E, src.shape,
                            i_rec_meth=2)
    got = {n: getattr(sl, n) for n in CHECKED_REC if n != "T"}
    got["T"] = sl.T
    worst, bad = compare(got, o, names=CHECKED_REC)
    assert not bad, bad
    with pytest.raises(ValueError):
        ra.Slab2Bgnd(Edges, T, nH, nHe, src, rec_meth="isotropic")


def test_graph_replay_equals_direct_launches(orc, spec128, monkeypatch):
    """rb_plan_solve replays one captured CUDA graph per in-flight sweep slot after the
    first sweep; the results must be those of direct kernel launches bit for bit, also
    when a cached plan is re-armed with other flags (graphs are re-captured)."""
    from rabacus_b200 import rabacus_fc as fc
    cases = [_small_sphere(96, 128, 8, False, spec=spec128),
             _small_sphere(96, 128, 8, True, spec=spec128),
             configs.c3_slab_bgnd(64, 128, False), configs.c1_iliev_stromgren(64)]
    for cfg in cases:
        monkeypatch.setenv("RB_NO_GRAPHS", "1")
        ref = run_gpu(cfg)
        monkeypatch.delenv("RB_NO_GRAPHS")
        g = run_gpu(cfg)
        assert g["iters"] == ref["iters"]
        for n in CHECKED:
            assert np.array_equal(g[n], ref[n]), (cfg["geometry"], n)
    # one-shot entry point: the cached plan is reused with a different tolerance and find_Teq
    cfg = _small_sphere(64, 128, 8, False, spec=spec128)
    outs = []
    for tol, teq in ((1e-4, 0), (1e-6, 0), (1e-4, 1), (1e-4, 0)):
        T = cfg["T"].copy()
        o = fc.sphere_bgnd.sphere_bgnd_solve(
            cfg["Edges"].copy(), cfg["nH"].copy(), cfg["nHe"].copy(), T, 8, cfg["E_eV"],
            cfg["shape"], 1, 1.0, 1, 1, teq, 0, 1.0, 1.0, 1.0, 3.0, 0.0, tol, 64, 128)
        outs.append((o, dict(fc.last_info)["iters"]))
    assert outs[1][1] > outs[0][1]                      # the tighter tolerance took effect
    assert outs[3][1] == outs[0][1]
    for a, b_ in zip(outs[0][0], outs[3][0]):
        assert np.array_equal(a, b_)


def _one_shot_iso(cfg, em):
    from rabacus_b200 import rabacus_fc as fc
    T = cfg["T"].copy()
    Nl, Nnu = cfg["nH"].size, cfg["E_eV"].size
    o = fc.sphere_bgnd.sphere_bgnd_solve(
        cfg["Edges"].copy(), cfg["nH"].copy(), cfg["nHe"].copy(), T, cfg["Nmu"], cfg["E_eV"],
        cfg["shape"], 4, 1.0, 1, 1, 0, 0, em[0], em[1], em[2], cfg["z"], 0.0, 1e-4, Nl, Nnu)
    return dict(zip(("xH1", "xH2", "xHe1", "xHe2", "xHe3", "H1i_src", "He1i_src", "He2i_src",
                     "H1i_rec", "He1i_rec", "He2i_rec", "H1h_src", "He1h_src", "He2h_src",
                     "H1h_rec", "He1h_rec", "He2h_rec"), o)), dict(fc.last_info)["iters"]


def test_one_shot_isotropic_rescans_em_fac_on_a_cached_plan(orc, spec128):
    """Scanning em_*_fac on a fixed grid through the one-shot ABI re-arms one cached plan:
    the captured sweep graphs bake the emission factors in (kernel arguments by value) and
    must be re-captured, or sweeps >= 2 run with the previous call's factors.  Every call
    must equal a fresh plan's result bit for bit, and the oracle's within tolerance."""
    from rabacus_b200 import _lib
    cfg = _small_sphere(64, 128, 8, False, nH0=1.0, spec=spec128)
    cfg["i_rec_meth"] = 4
    _lib.lib().rb_cache_clear()
    for em in ((1.0, 1.0, 1.0), (0.3, 2.0, 0.5), (1.0, 1.0, 1.0), (0.0, 0.0, 0.0)):
        got, iters = _one_shot_iso(cfg, em)          # cached plan from the 2nd call on
        cfg["em_fac"] = em
        fresh = run_gpu(cfg)                          # a new plan every time
        assert iters == fresh["iters"], em
        for n in CHECKED_REC:
            if n != "T":
                assert np.array_equal(got[n], fresh[n]), (em, n)
        o = run_oracle(orc, cfg)
        assert iters == o["iters"], em
        worst, bad = compare(fresh, o, names=CHECKED_REC)
        assert not bad, (em, bad)


def test_explicit_zero_emission_factors_switch_the_emission_off(orc, spec128):
    """sphere_base.f90:955 multiplies the emissivities by em_*_fac: zeros mean no
    recombination emission (all *_rec rates exactly zero, case-A rates), not "unset"."""
    cfg = _small_sphere(48, 128, 8, False, nH0=1.0, spec=spec128)
    cfg["i_rec_meth"] = 4
    got, _ = _one_shot_iso(cfg, (0.0, 0.0, 0.0))
    for n in ("H1i_rec", "He1i_rec", "He2i_rec", "H1h_rec", "He1h_rec", "He2h_rec"):
        assert np.all(got[n] == 0.0), n
    fixedA = dict(cfg, i_rec_meth=1, fixed_fcA=1.0)
    ref = run_gpu(fixedA)
    for n in ("xH1", "xHe1", "xHe2", "H1i_src", "He1i_src", "He2i_src"):
        assert np.array_equal(got[n], ref[n]), n


@pytest.mark.parametrize("Nl,Nnu,Nmu", [(1, 16, 4), (2, 16, 1), (3, 1, 2), (33, 2, 3),
                                        (7500, 16, 4), (8192, 8, 2)])
def test_sphere_bgnd_edge_sizes(orc, Nl, Nnu, Nmu):
    """Smallest and largest grids: one and two shells, one ray, one and two frequencies
    (the monochromatic branch), and grids whose shell records do not fit in shared memory
    (the global-memory column kernel, two cells per thread in the thin-state cluster)."""
    from rabacus_b200.rad_src import BackgroundSource
    if Nnu == 1:
        src = BackgroundSource(1.5, 1.5, "monochromatic")
        src.normalize_n(1.0e-5)
        spec = (src.E.copy(), src.shape.copy())
    else:
        spec = configs.hm12_background(Nnu, 3.0)
    cfg = configs.c4_sphere_bgnd(Nl, Nnu, Nmu, False, nH0=0.3, spectrum=spec)
    for kw in (dict(thin=True), dict(max_sweeps=2)):
        g = run_gpu(cfg, **kw)
        o = run_oracle(orc, cfg, **kw)
        worst, bad = compare(g, o)
        assert not bad, (kw, bad)


def test_hot_and_cold_cells(orc, spec128):
    """A wide spread of fixed temperatures in one problem (7.9e3 .. 1.6e5 K)."""
    cfg = _small_sphere(64, 128, 8, False, spec=spec128)
    cfg["T"] = np.logspace(3.9, 5.2, 64)
    g = run_gpu(cfg, max_sweeps=3)
    o = run_oracle(orc, cfg, max_sweeps=3)
    worst, bad = compare(g, o)
    assert not bad, bad


def test_absent_species(orc, spec128):
    """Cells without helium (zero_absent_species, ion_solver.f90:548-560).  In such a cell
    the reference's starting guess (analytic_H1, exact for pure hydrogen) already is the
    fixed point of the n_e bisection, so ratio_ne = 1 +- 1 ulp and the FIRST branch of
    solve_pce (:515-521) is decided by rounding noise (shown directly in
    tests/test_second_oracle.py::test_helium_free_zone_first_branch_is_decided_by_the_last_bit);
    the iteration then returns from the far bracket and stops at its own tolerance
    (1e-2 tol = 1e-6 on successive fractions).  Any two builds of the reference agree on
    xH1 there only to that tolerance, and from the second sweep on the columns -- hence the
    rates of every cell behind -- inherit it: rates are held to 100 x the solver tolerance
    here.  The fractions themselves are pinned exactly: the oracle's solver, given the CUDA
    rates, returns the CUDA fractions bit for bit (helpers.solver_check); He fractions are
    exactly zero on both sides."""
    cfg = _small_sphere(64, 128, 8, False, spec=spec128)
    cfg["nHe"] = cfg["nHe"].copy()
    cfg["nHe"][::3] = 0.0
    for kw in (dict(thin=True), dict(max_sweeps=1), dict(max_sweeps=3)):
        g = run_gpu(cfg, **kw)
        o = run_oracle(orc, cfg, **kw)
        worst, bad = compare(g, o, rtol=1e-10 if kw.get("thin") else 1e-4)
        assert not bad, (kw, bad)
        assert max(worst[n] for n in ("xH1", "xH2")) <= 1e-4    # |dx/x| ~ solver tolerance
        for n in ("xHe1", "xHe2", "xHe3"):
            assert np.all(g[n][::3] == 0.0) and np.all(o[n][::3] == 0.0)


# ----------------------------------------------------------------- table-driven powers (rb_fastpow.cuh)
def test_fastpow_device_equals_host_build_bit_for_bit():
    """pow_fast is IEEE add / multiply / fma and table reads only: the sm_100a build and the
    host build of the same source give identical bits (the host build is checked against
    mpmath in tests/test_host_logic.py)."""
    from test_host_logic import _diag_pow, fastpow_samples
    x, y = fastpow_samples(200000, seed=5)
    assert np.array_equal(_diag_pow(x, y, 1), _diag_pow(x, y, 0))


def test_device_coefficients_equal_the_host_build_bit_for_bit():
    """The 16 rate / cooling coefficients of a temperature are IEEE add / multiply / divide /
    sqrt / fma and table reads only (rb_fastpow.cuh: x**y, exp, log10 from shared tables), so
    the sm_100a build and the host build return the same bits.  This is what allows the
    host build to stand in for the device as the oracle's coefficient hook
    (helpers.solver_check)."""
    from test_host_logic import _diag_hg97
    rng = np.random.default_rng(11)
    n = 200000
    T = 10 ** rng.uniform(3.0, 9.0, n)
    T[: n // 2] = 10 ** rng.uniform(np.log10(7.5e3), 5.0, n // 2)
    f = [rng.choice([0.0, 1.0, 0.3, 0.77], n) for _ in range(3)]
    assert np.array_equal(_diag_hg97(T, *f, on_device=1), _diag_hg97(T, *f, on_device=0))


def test_hg97_device_coefficients_match_oracle(orc):
    """All 16 rate / cooling coefficients of one temperature probe, CUDA fast evaluation
    vs the oracle's libm statement form, over the solver's range and outside it."""
    from test_host_logic import _diag_hg97
    rng = np.random.default_rng(7)
    n = 100000
    T = 10 ** rng.uniform(0.0, 12.0, n)
    T[: n // 2] = 10 ** rng.uniform(np.log10(7.5e3), 5.0, n // 2)
    f = [rng.choice([0.0, 1.0, 0.3, 0.77], n) for _ in range(3)]
    out = _diag_hg97(T, *f, on_device=1)
    ref = np.array(list(orc.get_kchem(T, *f)) + list(orc.get_kcool(T, *f)))
    with np.errstate(all="ignore"):
        rel = np.where(out == ref, 0.0, np.abs(out - ref) / np.abs(ref))
    assert np.nanmax(rel) <= 2e-14


def test_solve_with_table_driven_powers_equals_libm_form(spec128):
    """A/B over whole find_Teq solves: the Hui & Gnedin fits through rb_fastpow.cuh vs through
    libdevice's pow() (a per-plan switch, csrc/rb_internal.h): same sweep count, temperatures
    bit-identical, rates <= 1e-12, fractions within the fraction policy."""
    cfgs = [_small_sphere(96, 128, 8, True, nH0=1.0, spec=spec128),
            configs.c2_slab_plane(128, 128, True), configs.c3_slab_bgnd(128, 128, True)]
    for cfg in cfgs:
        fast = run_gpu(cfg)
        slow = run_gpu(cfg, libm_pow=True)
        assert fast["iters"] == slow["iters"]
        assert np.array_equal(fast["T"], slow["T"])
        worst, bad = compare(fast, slow)
        assert not bad, bad
