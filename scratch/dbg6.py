import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from helpers import *
from oracle import oracle as orc
from rabacus_b200 import configs
for Nl in (256,512,520,1024,2048):
    for Teq in (False,True):
        cfg=configs.c3_slab_bgnd(Nl,128,Teq)
        g=run_gpu(cfg,thin=True); o=run_oracle(orc,cfg,thin=True)
        w,_=compare(g,o)
        g1=run_gpu(cfg,max_sweeps=1); o1=run_oracle(orc,cfg,max_sweeps=1)
        w1,_=compare(g1,o1)
        print(Nl,Teq,'thin worst %.2e (%s)'%(max(w.values()),max(w,key=w.get)),' 1 sweep worst %.2e (%s)'%(max(w1.values()),max(w1,key=w1.get)), 'Tthin', g['T'][:2], o['T'][:2])
