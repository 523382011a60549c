#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02f_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r02f_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('c4 ms/step %.4f'%d['ms_per_step'], r['kernels'], r['ms_cell_solve'], r['ms_finish'])"
timeout 300 python scratch/c5_stages.py 2048
timeout 300 python scratch/c5_stages.py 8192
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r02f_c5_launches.csv python scratch/c5_stages.py 1024 --sweeps 4 > gpurun_out/r02f_c5_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02f_c5_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i+1; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[start:]:
    if len(r)<=vi: continue
    n=r[ki].split('(')[0]; v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1e3
    agg.setdefault(n,[]).append(v)
for n,v in agg.items(): print('%-60s n=%3d  %s'%(n[:60],len(v),' '.join('%.0f'%x for x in v[:6])))
PY
