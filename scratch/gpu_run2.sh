#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_adversarial.txt
RB_PARITY_REPORT=$PWD/gpurun_out/r02_parity_adversarial.txt timeout 1500 python -m pytest tests/test_gpu_adversarial.py -q > gpurun_out/r02_adv.log 2>&1
echo "adv rc=$?"; tail -25 gpurun_out/r02_adv.log
