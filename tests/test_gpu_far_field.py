"""k_far_moments_split against k_far_moments (rb_kernels.cuh): the moment table of the far
field with the NT recurrences of a (segment, absorber) dealt over five threads, against one
thread carrying all of them.  Same operations in the same order for every moment: every output
of a solve must be bit-identical."""
import os

import numpy as np
import pytest

from helpers import run_gpu
from rabacus_b200 import configs

pytestmark = pytest.mark.gpu

FIELDS = ("xH1", "xH2", "xHe1", "xHe2", "xHe3", "T", "H1i_src", "He1i_src", "He2i_src",
          "H1h_src", "He1h_src", "He2h_src")


def _with_split(on, fn):
    old = os.environ.get("RB_FAR_SPLIT")
    os.environ["RB_FAR_SPLIT"] = "1" if on else "0"
    try:
        return fn()
    finally:
        if old is None:
            del os.environ["RB_FAR_SPLIT"]
        else:
            os.environ["RB_FAR_SPLIT"] = old


@pytest.mark.parametrize("shape", [(40, 16, 4), (300, 32, 33), (512, 128, 32), (1000, 16, 8),
                                   (4096, 32, 16), (6000, 16, 4)])
@pytest.mark.parametrize("clumpy", [False, True])
def test_split_moment_scan_is_bit_identical(shape, clumpy):
    Nl, Nnu, Nmu = shape
    cfg = configs.c4_sphere_bgnd(Nl, Nnu, Nmu, False)
    if clumpy:
        rng = np.random.default_rng(Nl)
        cfg["nH"] = cfg["nH"] * 10 ** rng.normal(0.0, 1.5, Nl)
        cfg["nHe"] = cfg["nH"] * 0.08
        e = np.sort(cfg["Edges"][-1] * rng.random(Nl - 1))
        cfg["Edges"] = np.concatenate(([0.0], e, [cfg["Edges"][-1]]))
    a = _with_split(True, lambda: run_gpu(cfg, max_sweeps=3))
    b = _with_split(False, lambda: run_gpu(cfg, max_sweeps=3))
    for n in FIELDS:
        assert np.array_equal(a[n], b[n]), (shape, clumpy, n)
    assert np.all(a["H1i_src"] > 0.0)


def test_split_moment_scan_batch_is_bit_identical():
    """a batch (one CTA per problem, no cluster) through the family producer"""
    from rabacus_b200.plan import Plan

    def run():
        n = 24
        plan = Plan(configs.GEOM_IDS["sphere_bgnd"], n, 512, 128, 32, tol=1e-4, max_sweeps=4)
        u = np.arange(n) / n
        plan.set_sphere_family(10.0 ** (-3.0 + 2.5 * u), (5.0 + 20.0 * u[::-1]) * configs.KPC_CM,
                               2.0 + 2.0 * u)
        plan.upload()
        plan.solve()
        out = plan.get_all()
        plan.close()
        return out
    a = _with_split(True, run)
    b = _with_split(False, run)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def _with_env(name, value, fn):
    old = os.environ.get(name)
    os.environ[name] = str(value)
    try:
        return fn()
    finally:
        if old is None:
            del os.environ[name]
        else:
            os.environ[name] = old


@pytest.mark.parametrize("variant", [1, 6, 9])
def test_column_launch_shapes_are_bit_identical(variant):
    """k_sphere_columns_smem<THREADS, U, 1>: the launch shape is chosen from the number of
    ray-pair items per SM (full launches 512 threads, cell shards 384 / 256; rb_api.cu
    launch_sphere_columns) -- same operations in the same order, so every output of a solve
    must be the same bits whatever the shape."""
    cfg = configs.c4_sphere_bgnd(4096, 16, 16, False)
    a = _with_env("RB_COLS_VARIANT", 7, lambda: run_gpu(cfg, max_sweeps=2))
    b = _with_env("RB_COLS_VARIANT", variant, lambda: run_gpu(cfg, max_sweeps=2))
    for n in FIELDS:
        assert np.array_equal(a[n], b[n]), (variant, n)
