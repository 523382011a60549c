#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python scratch/rank_share2.py
python scratch/rank_share2.py --find_Teq
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02i_share8_launches.csv python scratch/rank_share2.py > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02i_share8_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i+1; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
seq=[]
for r in rows[start:]:
    if len(r)<=vi: continue
    n=r[ki].split('(')[0]; v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1e3
    seq.append((n,v))
# print the last 12 launches (stride 8)
for n,v in seq[-12:]: print('%-60s %8.1f us'%(n[:60],v))
PY
