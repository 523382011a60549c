#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_teq_memo.py -m gpu -q -x > gpurun_out/r02d_pytest_memo.log 2>&1
echo "pytest memo rc=$?"; tail -15 gpurun_out/r02d_pytest_memo.log
for d in 0 12 16 18 20 22; do
  RB_TEQ_DEPTH=$d timeout 300 python bench.py --steps 10 --warmup 3 --find_Teq 1 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('depth $d c4 find_Teq ms/step %.4f cell %.4f'%(d['ms_per_step'], r['ms_cell_solve']))"
done
for d in 0 18 22; do
  RB_TEQ_DEPTH=$d timeout 300 python scratch/c5_stages.py 2048 --find_Teq
done
timeout 300 python scratch/c5_stages.py 2048
timeout 300 python scratch/c5_stages.py 8192
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02d_c5_launches.csv python scratch/c5_stages.py 1024 --sweeps 4 > gpurun_out/r02d_c5_ncu.log 2>&1
echo "ncu rc=$?"
