#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for l in 8 10 8 10; do
RB_NU_LTAB=$l timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('LTAB $l ms/step %.4f nu %.4f cols %.4f'%(d['ms_per_step'], r['kernels']['k_nu_rates']['ms'], r['kernels']['k_sphere_columns']['ms']))"
done
RB_NU_LTAB=10 timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02p_pytest_ltab10.log 2>&1; echo "pytest ltab10 rc=$?"; tail -3 gpurun_out/r02p_pytest_ltab10.log
RB_NU_LTAB=10 timeout 600 ncu --set full --clock-control none -k regex:'k_nu_rates' -s 2 -c 1 -o gpurun_out/r02p_nu10 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-e2e > /dev/null 2>&1; echo "ncu rc=$?"
