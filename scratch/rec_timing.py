import sys, time, numpy as np
sys.path.insert(0,'.')
from rabacus_b200 import configs
from rabacus_b200.plan import Plan
def run(cfg, meth, n=3):
    Nl, Nnu = cfg['nH'].size, cfg['E_eV'].size
    plan = Plan(configs.GEOM_IDS[cfg['geometry']], 1, Nl, Nnu, cfg['Nmu'], find_Teq=cfg['find_Teq'], tol=1e-4, max_sweeps=-10**9, i_rec_meth=meth)
    plan.set_problem(0, cfg['Edges'], cfg['nH'], cfg['nHe'], cfg['T'], cfg['E_eV'], cfg['shape'], cfg['z'])
    plan.upload(); plan.thin(); plan.bench_steps(1, 0)
    i0=plan.last_info(); ms=plan.bench_steps(n, 0); i1=plan.last_info()
    plan.close()
    return ms/n, {k:(i1[k]-i0[k])/n for k in ('ms_rec','ms_columns','ms_nu','ms_cell')}
c4 = configs.c4_sphere_bgnd(4096,512,256,False)
for meth in (1,2,3,4):
    print('C4 rec_meth', meth, run(c4, meth))
c3 = configs.c3_slab_bgnd(1024,512,False)
for meth in (1,2):
    print('C3 rec_meth', meth, run(c3, meth, 5))
c1 = configs.c1_iliev_stromgren(128); c1['Nmu']=32
for meth in (1,2,3,4):
    print('C1 rec_meth', meth, run(c1, meth, 5))
