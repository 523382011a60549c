#!/bin/bash
for v in 6 7 8 9; do
  RB_COL_VARIANT=$v python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('col variant $v', 'step %.3f'%d['ms_per_step'], d['roofline']['kernels'])"
done
for v in 0 1 2 3 4; do
  RB_COL_VARIANT=6 RB_NU_VARIANT=$v python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('nu variant $v', 'step %.3f'%d['ms_per_step'], d['roofline']['kernels'])"
done
