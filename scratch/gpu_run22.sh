#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02r_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02r_pytest_gpu.log
(cd rabacus_b200 && python build.py --checked > /dev/null 2>&1)
RB_LIBRARY=$PWD/rabacus_b200/librabacus_b200_checked.so timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02r_checked_pytest.log 2>&1; echo "checked pytest rc=$?"; tail -3 gpurun_out/r02r_checked_pytest.log
RB_LIBRARY=$PWD/rabacus_b200/librabacus_b200_checked.so timeout 900 python scratch/sanitize_small.py > gpurun_out/r02r_checked_small.log 2>&1; echo "checked small rc=$?"; tail -3 gpurun_out/r02r_checked_small.log
python -c "import __graft_entry__ as g; g.smoke()"
