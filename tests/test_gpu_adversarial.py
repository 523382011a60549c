"""GPU parity on inputs chosen to stress the column walk and the cell solvers (VERDICT r01
"input diversity"): log-spaced and randomly perturbed Edges, clumpy density (seeded lognormal,
sigma = 2 dex), shells without gas, a 1e12 density contrast, an ionisation front resolved by
one cell; every geometry, SphereBgnd with Nmu in {2, 33, 256}.  The summed-by-parts /
far-field evaluation of the columns re-associates the reference's running sums: this is where
its cancellation behaviour would show.  Tolerances: helpers.compare (rates / T 1e-10,
fractions the stated policy); the measured worst cases are written to
$RB_PARITY_REPORT (committed as profiles/r02_parity_adversarial.txt)."""
import os

import numpy as np
import pytest

from helpers import CHECKED, compare, run_gpu, run_oracle
from rabacus_b200 import configs
from rabacus_b200._lib import RabacusB200Error

pytestmark = pytest.mark.gpu

KINDS = ("log_edges", "jitter_edges", "clumpy", "contrast12", "sharp_front", "clumpy_log")


def adversarial(cfg, kind, seed=11):
    cfg = dict(cfg)
    Nl = cfg["nH"].size
    rng = np.random.default_rng(seed)
    E = cfg["Edges"].copy()
    L = E[-1] - E[0]
    nH = cfg["nH"].copy()
    he_ratio = cfg["nHe"][0] / cfg["nH"][0]
    if kind in ("log_edges", "clumpy_log"):
        # four decades in radius / depth; spheres keep the centre at 0
        inner = E[0] + L * 10 ** np.linspace(-4.0, 0.0, Nl)
        E = np.concatenate([[E[0]], inner])
    if kind == "jitter_edges":
        w = rng.uniform(0.05, 1.0, Nl)          # shell widths vary by a factor 20
        E = E[0] + L * np.concatenate([[0.0], np.cumsum(w) / w.sum()])
    if kind in ("clumpy", "clumpy_log"):
        nH = nH * 10 ** rng.normal(0.0, 2.0, Nl)
    if kind == "contrast12":
        nH = 10 ** rng.uniform(-8.0, 4.0, Nl)
    if kind == "sharp_front":
        nH = np.where(np.arange(Nl) < Nl // 2, 1e-4, 3.0e1)   # thin gas against a wall
        if cfg["geometry"].startswith("sphere_bgnd"):
            nH = nH[::-1].copy()                               # wall in the centre
    cfg["Edges"], cfg["nH"] = E, nH
    cfg["nHe"] = nH * he_ratio
    return cfg


def base_config(geom, Nl=72, Nnu=48, Nmu=8, find_Teq=False):
    if geom == "sphere_bgnd":
        return configs.c4_sphere_bgnd(Nl, Nnu, Nmu, find_Teq, nH0=0.5)
    if geom == "sphere_stromgren":
        c = configs.c1_iliev_stromgren(Nl)
        c["nHe"] = c["nH"] * 0.08
        return c
    if geom == "slab_bgnd":
        return configs.c3_slab_bgnd(Nl, Nnu, find_Teq)
    return configs.c2_slab_plane(Nl, Nnu, find_Teq, geom == "slab_plane_2")


def _report(tag, worst):
    path = os.environ.get("RB_PARITY_REPORT")
    if path:
        with open(path, "a") as f:
            f.write("%-58s %s\n" % (tag, " ".join("%s=%.1e" % (k, worst[k]) for k in CHECKED)))


def _check(orc, cfg, tag, **kw):
    g = run_gpu(cfg, **kw)
    o = run_oracle(orc, cfg, **kw)
    if not kw.get("thin"):
        assert g["iters"] == o["iters"], (tag, g["iters"], o["iters"])
    worst, bad = compare(g, o)
    _report(tag, worst)
    assert not bad, (tag, bad)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("geom", ["sphere_bgnd", "sphere_stromgren", "slab_plane_1",
                                  "slab_plane_2", "slab_bgnd"])
def test_adversarial_inputs_every_geometry(orc, geom, kind):
    for teq in (False, True):
        cfg = adversarial(base_config(geom, find_Teq=teq), kind)
        tag = "%s %s %s" % (geom, kind, "Teq" if teq else "fixedT")
        _check(orc, cfg, tag + " 2 sweeps", max_sweeps=2)
        _check(orc, cfg, tag + " solve", max_sweeps=40)    # to convergence or 40 sweeps


@pytest.mark.parametrize("Nmu", [2, 33, 256])
@pytest.mark.parametrize("kind", KINDS)
def test_adversarial_sphere_bgnd_ray_counts(orc, Nmu, kind):
    Nl = 40 if Nmu == 256 else 72
    cfg = adversarial(base_config("sphere_bgnd", Nl=Nl, Nmu=Nmu), kind, seed=Nmu)
    _check(orc, cfg, "sphere_bgnd Nmu=%d %s 3 sweeps" % (Nmu, kind), max_sweeps=3)


def test_adversarial_large_sphere_one_sweep(orc):
    """The production code path of C4 (shared-memory records, one CTA per SM, LPT list) on a
    clumpy log-spaced sphere: 1024 shells x 64 frequencies x 32 rays, thin start + 2 sweeps."""
    cfg = adversarial(configs.c4_sphere_bgnd(1024, 64, 32, False, nH0=1.0), "clumpy_log")
    _check(orc, cfg, "sphere_bgnd 1024x64x32 clumpy_log", max_sweeps=2)


def test_shells_without_gas(orc):
    """nH = nHe = 0 in some shells: the reference's solvers return zero fractions there
    (ion_solver.f90:548-559), contribute no column, and its convergence test never passes
    (0/0): sweep-level parity, and both sides stop at the outer iteration limit."""
    for geom in ("sphere_bgnd", "slab_bgnd"):
        cfg = base_config(geom, Nl=48, Nnu=32, Nmu=6)
        cfg["nH"][5:9] = 0.0
        cfg["nH"][30] = 0.0
        cfg["nHe"] = cfg["nH"] * 0.08
        g = run_gpu(cfg, max_sweeps=3)
        o = run_oracle(orc, cfg, max_sweeps=3)
        worst, bad = compare(g, o)
        _report("%s empty shells 3 sweeps" % geom, worst)
        assert not bad, bad
        for n in ("xH1", "xH2", "xHe1", "xHe2", "xHe3"):
            assert np.all(g[n][5:9] == 0.0)
            if geom == "sphere_bgnd":
                assert g[n][30] == 0.0
            else:   # the second half of a two-sided slab is the mirror image of the first
                assert g[n][30] == g[n][17] and np.all(g[n][48 - 9:48 - 5] == 0.0)
    cfg = base_config("sphere_bgnd", Nl=24, Nnu=16, Nmu=4)
    cfg["nH"][3] = 0.0
    cfg["nHe"] = cfg["nH"] * 0.08
    with pytest.raises(RabacusB200Error) as e:
        run_gpu(cfg)
    assert e.value.code == 2                       # RB_ERR_MAX_ITER_OUTER
    with pytest.raises(orc.OracleError):
        run_oracle(orc, cfg)


@pytest.mark.parametrize("kind", ("uniform",) + KINDS)
@pytest.mark.parametrize("shape", [(72, 48, 8), (300, 32, 33), (40, 16, 2), (1024, 32, 16)])
def test_far_field_columns_equal_the_full_walk(orc, kind, shape):
    """The far-field moment table (rb_kernels.cuh: k_far_moments, far_eval) against the walk
    over every shell edge (the round-1 column kernel, per-plan switch in rb_internal.h): the
    same sums, associated differently.  Ray optical depths feed exp(-tau): the rates after
    sweeps 1 and 3 must agree to 1e-12, and each arm separately with the oracle."""
    Nl, Nnu, Nmu = shape
    cfg = base_config("sphere_bgnd", Nl=Nl, Nnu=Nnu, Nmu=Nmu)
    if kind != "uniform":
        cfg = adversarial(cfg, kind, seed=Nl)
    for sweeps in (1, 3):
        a = run_gpu(cfg, max_sweeps=sweeps, far_field=True)
        b = run_gpu(cfg, max_sweeps=sweeps, far_field=False)
        worst = 0.0
        for n in CHECKED:
            if n.endswith("_src"):
                with np.errstate(all="ignore"):
                    r = np.where(a[n] == b[n], 0.0, np.abs(a[n] - b[n]) / np.abs(b[n]))
                worst = max(worst, float(np.nanmax(r)))
        _report("far vs full walk %s %dx%dx%d %d sweeps" % (kind, Nl, Nnu, Nmu, sweeps),
                dict({k: 0.0 for k in CHECKED}, H1i_src=worst))
        assert worst <= 1e-12, (kind, shape, sweeps, worst)
    if Nl <= 300:
        o = run_oracle(orc, cfg, max_sweeps=3)
        for arm in (a, b):
            w, bad = compare(arm, o)
            assert not bad, bad
