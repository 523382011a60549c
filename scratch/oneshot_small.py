import sys, time, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from rabacus_b200 import configs, rabacus_fc as fc
from oracle import oracle as orc
from helpers import run_oracle
def call(cfg):
    g=cfg['geometry']; Nl=cfg['nH'].size; Nnu=cfg['E_eV'].size
    T=cfg['T'].copy()
    if g=='sphere_stromgren':
        return fc.sphere_stromgren.sphere_stromgren_solve(cfg['Edges'].copy(),cfg['nH'].copy(),cfg['nHe'].copy(),T,1,cfg['E_eV'],cfg['shape'],1,cfg['fixed_fcA'],1,1,int(cfg['find_Teq']),0,1.,1.,1.,cfg['z'],0.0,1e-4,Nl,Nnu)
    if g=='slab_bgnd':
        return fc.slab_bgnd.slab_bgnd_solve(cfg['Edges'],cfg['nH'],cfg['nHe'],T,cfg['E_eV'],cfg['shape'],1,cfg['fixed_fcA'],1,1,int(cfg['find_Teq']),0,cfg['z'],0.0,1e-4,Nl,Nnu)
    return fc.slab_plane.slab_plane_solve(cfg['Edges'],cfg['nH'],cfg['nHe'],T,cfg['E_eV'],cfg['shape'],1 if g=='slab_plane_2' else 0,1,cfg['fixed_fcA'],1,1,int(cfg['find_Teq']),0,cfg['z'],0.0,1e-4,Nl,Nnu)
for name,cfg in (('C1 stromgren 128 mono', configs.c1_iliev_stromgren(128)),
                 ('C2 slabpln 256x128 fixedT', configs.c2_slab_plane(256,128,False)),
                 ('C2 slabpln 256x128 Teq', configs.c2_slab_plane(256,128,True)),
                 ('C3 slab2bgnd 1024x512', configs.c3_slab_bgnd(1024,512,False)),
                 ('C3 slab2bgnd 1024x512 Teq', configs.c3_slab_bgnd(1024,512,True))):
    call(cfg); ts=[]
    for _ in range(5):
        t=time.perf_counter(); call(cfg); ts.append(time.perf_counter()-t)
    info=dict(fc.last_info)
    t=time.perf_counter(); o=run_oracle(orc,cfg); tc=time.perf_counter()-t
    print('%-28s one-shot wall %.2f ms (device %.2f ms, %d sweeps, %d launches) | CPU oracle %d thr %.2f ms'%(name, min(ts)*1e3, info['ms_device'], info['iters'], info['n_launches'], orc.num_threads(), tc*1e3))
