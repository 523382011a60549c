#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02q_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02q_pytest_gpu.log
for l in 6 10 6 10; do RB_NU_LTAB=$l timeout 300 python scratch/c5_stages.py 2048; done
RB_NU_LTAB=10 timeout 300 python scratch/c5_stages.py 8192
RB_NU_LTAB=10 timeout 300 python scratch/c5_stages.py 8192 --find_Teq
