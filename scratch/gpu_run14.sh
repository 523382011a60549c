#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02j_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02j_pytest_gpu.log
python scratch/rank_share2.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02j_share_launches.csv python scratch/rank_share2.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02j_share_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i+1; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
seq=[]
for r in rows[start:]:
    if len(r)<=vi: continue
    n=r[ki].split('(')[0]; v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1e3
    seq.append((n,v))
for n,v in seq[20:26]+seq[-6:]: print('%-60s %8.1f us'%(n[:60],v))
PY
