import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from helpers import *
from oracle import oracle as orc
from rabacus_b200 import configs
cfg=configs.c3_slab_bgnd(1024,512,True)
prev=None
for k in range(1,19):
    g=run_gpu(cfg,max_sweeps=k); o=run_oracle(orc,cfg,max_sweeps=k)
    dT=np.abs(g['T']/o['T']-1); i=int(np.argmax(dT))
    ne_g=g['xH2']*cfg['nH']+(g['xHe2']+2*g['xHe3'])*cfg['nHe']; ne_o=o['xH2']*cfg['nH']+(o['xHe2']+2*o['xHe3'])*cfg['nHe']
    dne=np.abs(ne_g/ne_o-1)
    line='k=%2d iters g/o %d/%d  maxdT %.2e@%d (T=%.6f vs %.6f) max dne %.2e'%(k,g['iters'],o['iters'],dT[i],i,g['T'][i],o['T'][i],dne.max())
    if prev is not None:
        ch_g=np.abs(ne_g/prev[0]-1).max(); ch_o=np.abs(ne_o/prev[1]-1).max()
        line+='  conv change g %.6e o %.6e'%(ch_g,ch_o)
    prev=(ne_g,ne_o)
    print(line)
