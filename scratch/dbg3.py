import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from helpers import *
from oracle import oracle as orc
from rabacus_b200 import configs, rabacus_fc as fc
d=np.load('tests/golden/hm12_spectrum_128.npz'); spec=(d['E_eV'],d['Inu'])
cfg=configs.c4_sphere_bgnd(128,128,16,False,nH0=3.0,spectrum=spec)
o=run_oracle(orc,cfg,tol=1e-4)
k=orc.get_kchem(cfg['T'],1.,1.,1.)
kg=fc.get_kchem_v(cfg['T'],1.,1.,1.,1,128)
print('kchem diff',[float(np.abs(a/b-1).max()) for a,b in zip(kg,k)])
base=orc.solve_pce(cfg['nH'],cfg['nHe'],*k,o['H1i_src'],o['He1i_src'],o['He2i_src'],1e-6)
from rabacus_b200.rabacus_fc import _solve_pce
for kk,lab in ((k,'oracle k'),(kg,'gpu k')):
    got=_solve_pce((cfg['nH'],cfg['nHe'])+tuple(kk)+(o['H1i_src'],o['He1i_src'],o['He2i_src']),1e-6,128,0)
    for q,nm in enumerate(('xH1','xH2','xHe1','xHe2','xHe3')):
        rel=np.abs(got[q]/base[q]-1); i=np.argmax(rel)
        print(lab,nm,rel[i],i,got[q][i],base[q][i])
