#!/bin/bash
for v in 0 1; do
  RB_COLS_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant $v', d['roofline']['kernels'], d['ms_per_step'])"
done
