import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from helpers import *
from oracle import oracle as orc
from rabacus_b200 import configs
d=np.load('tests/golden/hm12_spectrum_128.npz'); spec=(d['E_eV'],d['Inu'])
cfg=configs.c4_sphere_bgnd(128,128,16,False,nH0=3.0,spectrum=spec)
g=run_gpu(cfg,tol=1e-4); o=run_oracle(orc,cfg,tol=1e-4)
print(g['iters'],o['iters'])
for n in CHECKED:
    a=g[n]; b=o[n]
    r=np.abs(a-b)/np.abs(b); i=np.nanargmax(r)
    print(n, i, r[i], a[i], b[i])
for k in (1,2,5,10,20):
    g=run_gpu(cfg,max_sweeps=k); o=run_oracle(orc,cfg,max_sweeps=k)
    w,_=compare(g,o); print(k, max(w.values()), max(w,key=w.get))
