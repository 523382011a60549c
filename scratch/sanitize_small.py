"""Small-grid pass over every kernel family of the hot path: the four geometries, fixed T
and find_Teq, every recombination-photon closure, a batch with the device-side input
producers, chunked column walks, the Gauss-Seidel order, the single-zone API and
set_column_vs_impact.  Written for compute-sanitizer; that tool is closed on the GPU pool
this was developed on, so it is run against the CHECKED build instead
(RB_LIBRARY=rabacus_b200/librabacus_b200_checked.so: device assertions on every computed
index, guard words behind every plan buffer -- Plan.close() raises if one was overwritten),
twice, demanding bit-identical results from both passes (run-to-run determinism)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from rabacus_b200 import _lib, configs, rabacus_fc as fc
from rabacus_b200.plan import Plan
import helpers

_seen = {}


def ok(name, res, quiet):
    bad = [k for k in helpers.CHECKED if not np.all(np.isfinite(np.asarray(res[k])))]
    sig = tuple(np.asarray(res[k]).tobytes() for k in helpers.CHECKED)
    same = _seen.setdefault(name, sig) == sig
    if not quiet:
        print("%-58s sweeps %3d  %s" % (name, res.get("iters", -1), "ok" if not bad else "NON-FINITE %s" % bad))
        sys.stdout.flush()
    assert not bad, name
    assert same, "NON-DETERMINISTIC between passes: %s" % name


def one_pass(quiet):
    cases = [
        ("C1 SphereStromgren mono 64", configs.c1_iliev_stromgren(64), {}),
        ("C2 SlabPln 64x32", configs.c2_slab_plane(64, 32, False), {}),
        ("C2 SlabPln 64x32 find_Teq", configs.c2_slab_plane(64, 32, True), {}),
        ("C2 Slab2Pln 64x32", configs.c2_slab_plane(64, 32, False, two_sided=True), {}),
        ("C3 Slab2Bgnd 64x32", configs.c3_slab_bgnd(64, 32, False), {}),
        ("C3 Slab2Bgnd 65x32 find_Teq (odd)", configs.c3_slab_bgnd(65, 32, True), {}),
        ("C4 SphereBgnd 96x32x8", configs.c4_sphere_bgnd(96, 32, 8, False, nH0=1.0), {}),
        ("C4 SphereBgnd 96x32x8 find_Teq", configs.c4_sphere_bgnd(96, 32, 8, True, nH0=1.0), {}),
        ("C4 SphereBgnd 300x32x9 (far rows + ragged walks)", configs.c4_sphere_bgnd(300, 32, 9, False, nH0=3.0), {}),
        ("C4 SphereBgnd 96x32x7 walked (no far field)", configs.c4_sphere_bgnd(96, 32, 7, False, nH0=1.0),
         dict(far_field=False)),
        ("C4 SphereBgnd 70x160x6 (Nnu > 128 kernel)", configs.c4_sphere_bgnd(70, 160, 6, False, nH0=1.0), {}),
        ("C4 SphereBgnd 70x520x6 (Nnu > 512: two passes)", configs.c4_sphere_bgnd(70, 520, 6, False, nH0=1.0), {}),
        ("C4 SphereBgnd 64x32x8 Gauss-Seidel", configs.c4_sphere_bgnd(64, 32, 8, False, nH0=1.0),
         dict(order="gs")),
    ]
    for name, cfg, kw in cases:
        ok(name, helpers.run_gpu(cfg, **kw), quiet)
    for meth, nm in ((2, "outward"), (3, "radial"), (4, "isotropic")):
        cfg = configs.c4_sphere_bgnd(64, 32, 6, False, nH0=1.0)
        cfg["i_rec_meth"] = meth
        ok("SphereBgnd 64x32x6 rec_meth=%s" % nm, helpers.run_gpu(cfg), quiet)
    cfg = configs.c3_slab_bgnd(64, 32, False)
    cfg["i_rec_meth"] = 2
    ok("Slab2Bgnd 64x32 rec_meth=ray", helpers.run_gpu(cfg), quiet)

    # chunked column walks (cell shards of a small grid) + combine
    os.environ["RB_COLS_CHUNKS"] = "4"
    ok("C4 SphereBgnd 96x32x8 walked, 4 chunks",
       helpers.run_gpu(configs.c4_sphere_bgnd(96, 32, 8, False, nH0=1.0), far_field=False), quiet)
    del os.environ["RB_COLS_CHUNKS"]

    # batch with the device-side input producers, fixed T and find_Teq (one thread per cell)
    for teq in (False, True):
        idx = np.arange(0, 8192, 8192 // 12)
        plan = Plan(0, idx.size, 64, 32, 4, find_Teq=teq)
        configs.c5_stage_device(plan, idx, find_Teq=teq)
        plan.upload()
        os.environ["RB_SPEC_G"] = "1"
        plan.solve()
        del os.environ["RB_SPEC_G"]
        res = plan.get_all()
        res["iters"] = int(res["iters"].max())
        plan.close()          # checked build: raises if a guard word was overwritten
        ok("C5 family batch 12 x 64x32x4 find_Teq=%s" % teq, res, quiet)

    # single-zone API and post-processing
    n = 300
    nH = np.logspace(-5, 0, n); nHe = nH * 0.08
    att = np.logspace(-3, 0, n)
    for lock in (0, 1):
        out = fc._solve_pcte((nH, nHe, 8e-13 * att, 5e-13 * att, 6e-15 * att, 5e-24 * att, 6e-24 * att,
                              2e-25 * att), 3.0, 0.0, (1.0, 1.0, 1.0), 1, 1e-6, n, lock)
        assert all(np.all(np.isfinite(o)) for o in out)
        k = fc.get_kchem_v(np.full(n, 2e4), 1.0, 1.0, 1.0, 1, n)
        out = fc._solve_pce((nH, nHe) + tuple(k) + (8e-13 * att, 5e-13 * att, 6e-15 * att), 1e-6, n, lock)
        assert all(np.all(np.isfinite(o)) for o in out)
    cfg = configs.c4_sphere_bgnd(64, 32, 8, False, nH0=1.0)
    r = helpers.run_gpu(cfg)
    cols = fc.sphere_base.set_column_vs_impact(r["xH1"], r["xH2"], r["xHe1"], r["xHe2"], r["xHe3"],
                                               cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], 64)
    assert all(np.all(np.isfinite(c)) for c in cols)
    if not quiet:
        print("%-58s ok" % "single-zone solve_pce / solve_pcte (scalar + lock-step)")
        print("%-58s ok" % "set_column_vs_impact")


probe = Plan(0, 1, 8, 4, 2)
n_guard, _ = probe.check_guards()
probe.close()
print("library: %s  (%s)" % (_lib.LIB_PATH, "CHECKED build: device assertions + buffer guards"
                            if n_guard >= 0 else "release build: no assertions, no guards"))
one_pass(False)
one_pass(True)
print("second pass: every case bit-identical to the first")
print("ALL CASES DONE")
