"""SphereBgnd in the reference's own sweep order (rabacus/f2py/sphere_bgnd.f90:753-913 run
with OMP_NUM_THREADS=1: cells outermost first, each seeing the cells before it already
updated).  The CUDA path offers it as a non-default switch (rb_plan_set_sweep_order); here it
is held to the oracle's ``order=gs`` -- the restatement of that loop -- at the DEFAULT
tolerance: same sweep count, photo-rates and T <= 1e-10, fractions through the solver check
(VERDICT r1 "What's missing" #4: Jacobi agrees with this order only to O(tol))."""
import numpy as np
import pytest

from helpers import CHECKED, CHECKED_REC, compare, run_gpu, run_oracle
from rabacus_b200 import configs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("Nl,Nnu,Nmu,find_Teq,nH0", [
    (48, 64, 8, False, 0.3),
    (96, 128, 5, False, 1.0),      # odd Nmu: the mu ~ 0 node
    (160, 64, 16, True, 1.0),
    (512, 32, 4, False, 3.0),      # the size named by the verdict (Nl <= 512)
])
def test_gauss_seidel_solve_matches_reference_order(orc, Nl, Nnu, Nmu, find_Teq, nH0):
    cfg = configs.c4_sphere_bgnd(Nl, Nnu, Nmu, find_Teq, nH0=nH0)
    g = run_gpu(cfg, tol=1e-4, order="gs")
    o = run_oracle(orc, cfg, tol=1e-4, order=orc.ORDER_GS)
    assert g["iters"] == o["iters"], (g["iters"], o["iters"])
    worst, bad = compare(g, o)
    assert not bad, (bad, worst)
    # and it is NOT the Jacobi answer: the two orders differ by O(tol) before convergence
    j = run_gpu(cfg, tol=1e-4)
    d = np.max(np.abs(j["xH1"] / g["xH1"] - 1.0))
    assert d > 1e-9, "orders coincide: the switch did nothing"
    assert d < 5e-2


def test_gauss_seidel_single_sweeps(orc):
    """sweep-level: thin start + 1, 2, 3 sweeps, decreasing and increasing input edges"""
    cfg = configs.c4_sphere_bgnd(80, 64, 6, False, nH0=2.0)
    for n in (1, 2, 3):
        g = run_gpu(cfg, max_sweeps=n, order="gs")
        o = run_oracle(orc, cfg, max_sweeps=n, order=orc.ORDER_GS)
        worst, bad = compare(g, o)
        assert not bad, (n, bad)
    rev = dict(cfg)
    rev["Edges"] = cfg["Edges"][::-1].copy()
    for k in ("nH", "nHe", "T"):
        rev[k] = cfg[k][::-1].copy()
    g2 = run_gpu(rev, max_sweeps=3, order="gs")
    for k in CHECKED:
        assert np.array_equal(np.asarray(g2[k])[::-1], np.asarray(g[k])), k


def test_gauss_seidel_with_recombination_photons(orc):
    cfg = configs.c4_sphere_bgnd(64, 64, 8, False, nH0=1.0)
    cfg["i_rec_meth"] = 2   # outward-only: the photon-transfer closure runs once per sweep
    g = run_gpu(cfg, order="gs")
    o = run_oracle(orc, cfg, order=orc.ORDER_GS)
    assert g["iters"] == o["iters"]
    worst, bad = compare(g, o, names=CHECKED_REC)
    assert not bad, bad


def test_gauss_seidel_only_for_sphere_bgnd_and_not_sharded():
    from rabacus_b200._lib import RabacusB200Error
    from rabacus_b200.plan import Plan
    slab = Plan(4, 1, 32, 16, 0)
    with pytest.raises(RabacusB200Error) as e:
        slab.set_sweep_order("gs")
    assert e.value.code == 9
    slab.close()
    sph = Plan(0, 1, 32, 16, 4)
    sph.set_sweep_order("gs")
    cfg = configs.c4_sphere_bgnd(32, 16, 4, False)
    sph.set_problem(0, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"], 3.0)
    sph.upload()
    sph.thin()
    with pytest.raises(RabacusB200Error) as e:
        sph.sweep_local(1, 2)          # a cell shard
    assert e.value.code == 9
    sph.close()


def test_public_class_sweep_order_switch(orc):
    from rabacus_b200 import rad_src, solvers
    cfg = configs.c4_sphere_bgnd(64, 64, 8, False, nH0=1.0)
    src = rad_src.BackgroundSource(1.0, 400.0, "hm12", Nnu=64, z=3.0)
    s = solvers.SphereBgnd(cfg["Edges"], cfg["T"], cfg["nH"], cfg["nHe"], src, z=3.0, Nmu=8,
                           sweep_order="gs")
    o = run_oracle(orc, cfg, order=orc.ORDER_GS)
    assert np.max(np.abs(s.H1i / o["H1i_src"] - 1.0)) <= 1e-10
    j = solvers.SphereBgnd(cfg["Edges"], cfg["T"], cfg["nH"], cfg["nHe"], src, z=3.0, Nmu=8)
    assert np.max(np.abs(j.H1i / s.H1i - 1.0)) > 1e-9
