import sys, os, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from oracle import oracle as orc
from rabacus_b200 import configs
from helpers import run_gpu, run_oracle, compare
spec = tuple(np.load('tests/golden/hm12_spectrum_128.npz')[k] for k in ('E_eV','Inu')) if False else configs.hm12_background(128,3.0)
def case(name, mod, **kw):
    cfg = configs.c4_sphere_bgnd(64,128,8,False,nH0=1.0,spectrum=spec)
    mod(cfg)
    for G in ('1','32'):
        os.environ['RB_SPEC_G']=G
        g=run_gpu(cfg,**kw); o=run_oracle(orc,cfg,**kw)
        w,bad=compare(g,o)
        i=int(np.argmax(np.abs(g['xH1']/o['xH1']-1)))
        print(name,kw,'G',G,'worst',{k:'%.1e'%v for k,v in w.items() if v>1e-10}, 'cell',i, cfg['nHe'][i], cfg['T'][i])
def m1(c): c['nHe']=c['nHe'].copy(); c['nHe'][::3]=0.0
def m2(c): c['T']=np.logspace(3.9,5.2,64)
def m3(c): c['T']=np.logspace(3.9,4.9,64)
case('nHe0',m1,thin=True); case('nHe0',m1,max_sweeps=1)
case('Tspread',m2,thin=True); case('Tspread',m2,max_sweeps=1); case('Tspread<1e5',m3,max_sweeps=1)
