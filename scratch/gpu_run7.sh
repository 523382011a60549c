#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_teq_memo.py tests/test_gpu_inputs.py -m gpu -q -s > gpurun_out/r02e_pytest_new.log 2>&1
echo "pytest new rc=$?"; tail -25 gpurun_out/r02e_pytest_new.log
for g in 4 8 16 32; do
  RB_SPEC_G=$g timeout 300 python bench.py --steps 10 --warmup 3 --find_Teq 1 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('G $g c4 find_Teq ms/step %.4f cell %.4f'%(d['ms_per_step'], r['ms_cell_solve']))"
done
