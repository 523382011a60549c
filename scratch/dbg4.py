import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from helpers import *
from oracle import oracle as orc
from rabacus_b200 import configs
d=np.load('tests/golden/hm12_spectrum_128.npz'); spec=(d['E_eV'],d['Inu'])
def report(tag,cfg,**kw):
    g=run_gpu(cfg,**kw); o=run_oracle(orc,cfg,**kw)
    s=[]
    for n in CHECKED:
        a=g[n]; b=o[n]
        with np.errstate(all='ignore'):
            r=np.abs(a-b)/np.abs(b)
        r=np.where(a==b,0,r); i=np.nanargmax(r)
        s.append('%s:%.1e@%d(%.1e)'%(n,r[i],i,b[i]))
    print(tag,g['iters'],o['iters'],' '.join(s))
for nH0 in (0.1,1.0,3.0,10.0):
    for Teq in (False,True):
        cfg=configs.c4_sphere_bgnd(128,128,16,Teq,nH0=nH0,spectrum=spec)
        report('sph nH0=%g Teq=%d'%(nH0,Teq),cfg,tol=1e-4)
cfg=configs.c4_sphere_bgnd(512,128,32,True,nH0=10.0,spectrum=spec)
report('sph512 Teq',cfg,tol=1e-4)
cfg=configs.c4_sphere_bgnd(512,128,32,False,nH0=10.0,spectrum=spec)
report('sph512',cfg,tol=1e-4)
for Teq in (False,True):
    report('c2 Teq=%d'%Teq, configs.c2_slab_plane(256,128,Teq,False), tol=1e-4)
    report('c3 Teq=%d'%Teq, configs.c3_slab_bgnd(256,128,Teq), tol=1e-4)
