#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export RB_LIBRARY=$PWD/rabacus_b200/librabacus_b200_checked.so
timeout 900 python scratch/sanitize_small.py > gpurun_out/r02_checked_small.log 2>&1; echo "checked small rc=$?"; tail -6 gpurun_out/r02_checked_small.log
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_checked_pytest.log 2>&1; echo "checked pytest rc=$?"; tail -6 gpurun_out/r02_checked_pytest.log
unset RB_LIBRARY
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02h_pytest_gpu.log 2>&1; echo "release pytest rc=$?"; tail -4 gpurun_out/r02h_pytest_gpu.log
