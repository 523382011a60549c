import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from oracle import oracle as orc
from rabacus_b200 import rabacus_fc as fc
ion=(8e-13,5e-13,6e-15); heat=(5e-24,6e-24,2e-25)
for n in (100, 512, 513, 600, 2100):
    nH=np.logspace(-4,-1,n); nHe=nH*0.08
    k=orc.get_kchem(np.full(n,1e4),1.,1.,1.)
    g=fc.solve_pce_v(nH,nHe,*k,*ion,1e-8,n); o=orc.solve_pce(nH,nHe,*k,*ion,1e-8,vector=True)
    print(n,'pce_v', max(np.abs(a/b-1).max() for a,b in zip(g,o)))
    g=fc.solve_pcte_v(nH,nHe,*ion,*heat,3.0,0.0,1.,1.,1.,1,1e-7,n); o=orc.solve_pcte(nH,nHe,*ion,*heat,3.0,0.0,1.,1.,1.,1e-7,vector=True)
    print(n,'pcte_v', max(np.abs(a/b-1).max() for a,b in zip(g,o)), g[5][:3], o[5][:3])
