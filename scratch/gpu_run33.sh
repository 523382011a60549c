#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02z_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/r02z_smoke.log
