#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_far_field.py -x -q -m gpu > gpurun_out/r02z_far_split_test.log 2>&1; echo "far test rc=$?"; tail -5 gpurun_out/r02z_far_split_test.log
for s in 1 0; do
RB_FAR_SPLIT=$s timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02z_bench_split$s.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02z_bench_split$s.json')); print('split=$s ms_per_step', d['ms_per_step'], d['roofline']['kernels'], d['roofline']['ms_cell_solve'], d['roofline']['ms_finish'])"
done
RB_FAR_SPLIT=1 timeout 300 python scratch/rank_share2.py 2>&1 | grep "stride -8\|stride 1 " | head -4
RB_FAR_SPLIT=0 timeout 300 python scratch/rank_share2.py 2>&1 | grep "stride -8\|stride 1 " | head -4
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02z2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02z2_pytest_gpu.log
