#!/bin/bash
# (1) group size of the fixed-T speculative cell solver at C4, (2) checked build over the far-field tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for g in 0 8 16 32; do
RB_SPEC_G=$g timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/tmp_g$g.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/tmp_g$g.json')); print('RB_SPEC_G=$g ms_per_step %.4f cell %.4f finish %.4f' % (d['ms_per_step'], d['roofline']['ms_cell_solve'], d['roofline']['ms_finish']))"
done
export RB_LIBRARY=$PWD/rabacus_b200/librabacus_b200_checked.so
timeout 900 python -m pytest tests/test_gpu_far_field.py tests/test_gpu_adversarial.py -m gpu -q -k "far or split" > gpurun_out/r02z_checked_far.log 2>&1; echo "checked far rc=$?"; tail -4 gpurun_out/r02z_checked_far.log
timeout 600 python scratch/sanitize_small.py > gpurun_out/r02z_checked_small.log 2>&1; echo "checked small rc=$?"; tail -4 gpurun_out/r02z_checked_small.log
