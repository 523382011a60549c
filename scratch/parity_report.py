import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from oracle import oracle as orc
from rabacus_b200 import configs
from helpers import run_gpu, run_oracle, compare, CHECKED
cases = [('sphere 96x128x16 1 sweep', configs.c4_sphere_bgnd(96,128,16,False,nH0=1.0), dict(max_sweeps=1)),
         ('sphere 96x128x16 solve', configs.c4_sphere_bgnd(96,128,16,False,nH0=1.0), dict()),
         ('sphere 512x128x32 Teq solve', configs.c4_sphere_bgnd(512,128,32,True,nH0=3.0), dict()),
         ('sphere 4096x512x256 2 sweeps', configs.c4_sphere_bgnd(4096,512,256,False), dict(max_sweeps=2)),
         ('slab2bgnd 1024x512 solve', configs.c3_slab_bgnd(1024,512,False), dict()),
         ('slabpln 256x128 Teq solve', configs.c2_slab_plane(256,128,True), dict()),
         ('stromgren 128 mono solve', configs.c1_iliev_stromgren(128), dict())]
for name,cfg,kw in cases:
    g=run_gpu(cfg,**kw); o=run_oracle(orc,cfg,**kw)
    worst,bad=compare(g,o)
    rates=max(worst[n] for n in CHECKED if n.endswith('_src'))
    fr=max(worst[n] for n in ('xH1','xH2','xHe1','xHe2','xHe3'))
    print('%-32s sweeps %3d/%3d  rates %.1e  T %.1e  fractions %.1e'%(name,g['iters'],o['iters'],rates,worst['T'],fr))
