#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_adversarial.txt
RB_PARITY_REPORT=$PWD/gpurun_out/r02_parity_adversarial.txt timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/r02_smoke.log
