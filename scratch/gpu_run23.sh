#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02s_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02s_pytest_gpu.log
python scratch/rank_share2.py | grep -E "stride (1|-2|-4|-8) "
timeout 300 python scratch/c5_stages.py 8192
