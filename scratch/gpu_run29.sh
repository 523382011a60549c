#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02x_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02x_pytest_gpu.log
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('ms/step %.4f nu %.4f cols %.4f'%(d['ms_per_step'], r['kernels']['k_nu_rates']['ms'], r['kernels']['k_sphere_columns']['ms']))"; done
timeout 300 python scratch/c5_stages.py 8192
