#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02u_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02u_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --find_Teq 1 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('c4 find_Teq ms/step %.4f cell %.4f'%(d['ms_per_step'], r['ms_cell_solve']))"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('c4 fixed T ms/step %.4f cell %.4f'%(d['ms_per_step'], r['ms_cell_solve']))"
timeout 300 python scratch/c5_stages.py 8192 --find_Teq
timeout 300 python scratch/c5_stages.py 8192
python - <<'PY'
import sys, time, ctypes as C
sys.path.insert(0,'.')
import bench
z=bench.zones_leg()
for k in ('1e+06','1e+07'): print(k, 'pce kernel ms %.3f pcte kernel ms %.3f'%(z[k]['solve_pce']['kernel_ms'], z[k]['solve_pcte']['kernel_ms']))
PY
