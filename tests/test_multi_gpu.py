"""Cell sharding over NVLink peer memory on real GPUs: runs
tests/multigpu/check_cell_sharding.py under torchrun when the box has at least two
GPUs (skipped on the single-GPU test box; the N>1 host logic is covered on CPU by
tests/test_sharding_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_cell_sharding_over_peer_memory_two_ranks():
    from rabacus_b200 import _lib
    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29577",
           os.path.join(root, "tests", "multigpu", "check_cell_sharding.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "cell-sharding check OK" in out.stdout
