#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02y_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02y_pytest_gpu.log
timeout 600 python scratch/rec_timing.py 2>&1 | tail -8
