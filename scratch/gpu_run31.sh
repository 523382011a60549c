#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for v in 0 1 2 3 4; do RB_NU_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('variant $v ms/step %.4f nu %.4f'%(d['ms_per_step'], r['kernels']['k_nu_rates']['ms']))"; done
