#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag 2>/dev/null
timeout 300 python -m pytest tests/test_gpu_inputs.py -m gpu -q 2>&1 | tail -2
for i in 1 2 3; do timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
for k in ('c5','c5_find_Teq'): print(k,{a:round(d[k].get(a),4) for a in ('wall_s','staging_s','solve_s','readback_s')})"; done
