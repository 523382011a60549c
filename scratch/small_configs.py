import sys, time, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from helpers import run_gpu, run_oracle
from oracle import oracle as orc
from rabacus_b200 import configs
from rabacus_b200.plan import Plan
def timeit(f, n=3):
    f(); ts=[]
    for _ in range(n):
        t=time.perf_counter(); r=f(); ts.append(time.perf_counter()-t)
    return min(ts), r
for name,cfg in (('C1 stromgren 128 mono', configs.c1_iliev_stromgren(128)),
                 ('C2 slabpln 256x128 fixedT', configs.c2_slab_plane(256,128,False)),
                 ('C2 slabpln 256x128 Teq', configs.c2_slab_plane(256,128,True)),
                 ('C3 slab2bgnd 1024x512', configs.c3_slab_bgnd(1024,512,False)),
                 ('C3 slab2bgnd 1024x512 Teq', configs.c3_slab_bgnd(1024,512,True)),
                 ('C4 sphere 4096x512x256 Teq', configs.c4_sphere_bgnd(4096,512,256,True))):
    tg, g = timeit(lambda: run_gpu(cfg, tol=1e-4))
    if 'C4' in name:
        print('%-30s gpu %.2f ms (%d sweeps, %d launches) stages %s'%(name, tg*1e3, g['iters'], g['info']['n_launches'], {k:round(v,2) for k,v in g['info'].items() if k.startswith('ms_')}))
        continue
    tc, o = timeit(lambda: run_oracle(orc, cfg, tol=1e-4), 1)
    print('%-30s gpu %.2f ms (%d sweeps, %d launches)  cpu-oracle %.2f ms (%d sweeps, %d thr)  ms_dev %.2f'%(name, tg*1e3, g['iters'], g['info']['n_launches'], tc*1e3, o['iters'], orc.num_threads(), g['info']['ms_device']))
# C5 batch
for B in (64, 512):
    probs = configs.c5_sweep(n_nH=8, n_R=8, n_z=B//64, Nl=512, Nnu=128, Nmu=32)
    plan = Plan(0, len(probs), 512, 128, 32, tol=1e-4)
    for b,c in enumerate(probs):
        plan.set_problem(b, c['Edges'], c['nH'], c['nHe'], c['T'], c['E_eV'], c['shape'], c['z'])
    t=time.perf_counter(); plan.upload(); info=plan.solve(); dt=time.perf_counter()-t
    print('C5 batch B=%d: %.1f ms, max sweeps %d, evals %.3e, evals/s %.3e, stages %s'%(len(probs), dt*1e3, info['iters'], info['evals'], info['evals']/dt, {k:round(v,2) for k,v in info.items() if k.startswith('ms_')}))
    plan.close()
