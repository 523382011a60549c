"""bench.py's host-side contract that needs no GPU: the CPU (reference) arm uses every host
thread whatever OMP_NUM_THREADS the launcher exported, both arms emit the same ``config``
dict, the timing build of the oracle is bit-identical to the parity build, and the batched
C5 staging equals the per-problem builder."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT
from rabacus_b200 import configs


def test_reference_arm_ignores_the_launchers_thread_count_and_shares_the_config():
    env = dict(os.environ, OMP_NUM_THREADS="1")   # what torchrun exports
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "2", "--warmup", "1", "--Nl", "96", "--Nnu", "32", "--Nmu", "8"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                      # exactly one JSON line on stdout
    line = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC
    assert line["cpu_baseline"]["cores"] == bench.host_threads()
    assert "OpenMP threads: %d" % bench.host_threads() in out.stderr
    assert line["e2e"]["value"] == line["value"] and line["n_gpus"] == 2
    assert "-O3 -mtune=native" in line["cpu_baseline"]["build"]

    class A(object):
        Nl, Nnu, Nmu, find_Teq = 96, 32, 8, 0
    assert line["config"] == bench.bench_config(A)      # the GPU arm prints this same dict


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_timing_build_of_the_oracle_is_bit_identical_to_the_parity_build(orc, spec128):
    """librb_oracle_o3.so (-O3 -mtune=native, what bench.py times) against librb_oracle.so
    (-O2, the parity checker): same sources, no FMA contraction in either."""
    cfg = configs.c4_sphere_bgnd(64, 128, 8, True, nH0=1.0, spectrum=spec128)
    args = (cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], 8, cfg["E_eV"], cfg["shape"])
    kw = dict(i_find_Teq=1, z=3.0, order=orc.ORDER_JACOBI)
    a = orc.sphere_bgnd_solve(*args, **kw)
    with orc.timing_build() as note:
        assert "-O3" in note
        n = orc.set_num_threads(3)
        assert n == 3 and orc.num_threads() == 3
        b = orc.sphere_bgnd_solve(*args, **kw)
        assert orc.last_sweep_seconds() > 0.0
    assert a["iters"] == b["iters"]
    for k in ("xH1", "xH2", "xHe1", "xHe2", "xHe3", "T", "H1i_src", "He1i_src", "He2i_src",
              "H1h_src", "He1h_src", "He2h_src"):
        assert np.array_equal(a[k], b[k]), k


def test_batched_c5_staging_equals_the_per_problem_builder():
    probs = configs.c5_sweep(n_nH=4, n_R=16, n_z=16, Nl=64, Nnu=32, Nmu=4)
    idx = np.array([0, 1, 15, 16, 255, 256, 700, 1023])
    a = configs.c5_batch_arrays(idx, Nl=64, Nnu=32, n_nH=4)
    for i, k in enumerate(idx):
        c = probs[k]
        q = a["spec_index"][i]
        assert np.array_equal(a["Edges"][i], c["Edges"]) and np.array_equal(a["nH"][i], c["nH"])
        assert np.array_equal(a["nHe"][i], c["nHe"]) and np.array_equal(a["T"][i], c["T"])
        assert np.array_equal(a["E"][q], c["E_eV"]) and np.array_equal(a["shape"][q], c["shape"])
        assert a["z"][i] == c["z"]
