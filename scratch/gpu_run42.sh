#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02z3_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02z3_pytest_gpu.log
timeout 300 python scratch/rank_share2.py 2>&1 | grep "rank 0" | tee gpurun_out/r02z3_rank_share.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z3_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02z3_smoke.log
