"""Cell sharding over NVLink peer memory on real GPUs: runs the torchrun checks of
tests/multigpu/ when the box has at least two GPUs (skipped on the single-GPU test box; the
N>1 host logic is covered on CPU by tests/test_sharding_gloo.py, and bench.py --gpus N
carries a driver-visible ``parity_vs_n1`` bit).  Records of the same checks on 2 GPUs:
profiles/r02_check_cell_sharding_n2.log, r02_soak_n2.log, r02_dead_peer_n2.log."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, port, args=(), timeout=600, env=None):
    from rabacus_b200 import _lib
    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "multigpu", script)] + list(args)
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=e)


def test_cell_sharding_over_peer_memory_two_ranks():
    """11 cases (4 geometries, find_Teq, every photon-transfer closure): peer path and NCCL
    path bit-identical to the one-GPU solve, equal sweep counts."""
    out = _torchrun("check_cell_sharding.py", 29577)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "cell-sharding check OK" in out.stdout


def test_peer_exchange_soak_two_ranks():
    """~1000 sweeps through the flag barrier + remote-store exchange on one connected plan:
    every solve bit-identical to one GPU."""
    out = _torchrun("soak_peer.py", 29578, ["--solves", "24"])
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "soak OK" in out.stdout


def test_dead_peer_times_out_instead_of_hanging():
    """A rank that stops sweeping surfaces as RB_ERR_PEER_TIMEOUT (12) on its peer."""
    out = _torchrun("dead_peer.py", 29579, timeout=120, env={"RB_PEER_TIMEOUT_MS": "400"})
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "dead-peer check OK" in out.stdout
