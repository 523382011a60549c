import sys, time, os, numpy as np
sys.path.insert(0,'.')
from rabacus_b200 import configs
from rabacus_b200.plan import Plan
which = sys.argv[1]
if which == 'c5':
    probs = configs.c5_sweep(n_nH=8, n_R=8, n_z=8, Nl=512, Nnu=128, Nmu=32)
    plan = Plan(0, len(probs), 512, 128, 32, tol=1e-4)
    for b,c in enumerate(probs):
        plan.set_problem(b, c['Edges'], c['nH'], c['nHe'], c['T'], c['E_eV'], c['shape'], c['z'])
    for rep in range(2):
        t=time.perf_counter(); plan.upload(); info=plan.solve(); dt=time.perf_counter()-t
    print('C5 B=%d NU=%s W=%s: %.1f ms, evals/s %.3e, cols %.1f nu %.1f cell %.1f'%(len(probs), os.environ.get('RB_NU_VARIANT'), os.environ.get('RB_COLS_WAVES'), dt*1e3, info['evals']/dt, info['ms_columns'], info['ms_nu'], info['ms_cell']))
elif which == 'c4':
    c = configs.c4_sphere_bgnd(4096,512,256,False)
    plan = Plan(0, 1, 4096, 512, 256, tol=1e-4, max_sweeps=-10**9)
    plan.set_problem(0, c['Edges'], c['nH'], c['nHe'], c['T'], c['E_eV'], c['shape'], c['z'])
    plan.upload(); plan.thin(); plan.bench_steps(3, 0)
    i0=plan.last_info(); ms=plan.bench_steps(10, 0); i1=plan.last_info()
    print('C4 NU=%s: step %.3f ms cols %.3f nu %.3f cell %.3f'%(os.environ.get('RB_NU_VARIANT'), ms/10, (i1['ms_columns']-i0['ms_columns'])/10, (i1['ms_nu']-i0['ms_nu'])/10, (i1['ms_cell']-i0['ms_cell'])/10))
else:
    # host phases of an uncached solve
    for name,cfg in (('C2 Teq', configs.c2_slab_plane(256,128,True)), ('C4', configs.c4_sphere_bgnd(4096,512,256,False))):
        for rep in range(3):
            t0=time.perf_counter()
            plan = Plan(configs.GEOM_IDS[cfg['geometry']], 1, cfg['nH'].size, cfg['E_eV'].size, cfg['Nmu'], find_Teq=cfg['find_Teq'], tol=1e-4, max_sweeps=2)
            t1=time.perf_counter()
            plan.set_problem(0, cfg['Edges'], cfg['nH'], cfg['nHe'], cfg['T'], cfg['E_eV'], cfg['shape'], cfg['z'])
            t2=time.perf_counter(); plan.upload(); t3=time.perf_counter(); info=plan.solve(); t4=time.perf_counter()
            r=plan.get_problem(0); t5=time.perf_counter(); plan.close(); t6=time.perf_counter()
            print(name, 'create %.2f set %.2f upload %.2f solve %.2f (dev %.2f) get %.2f close %.2f ms'%tuple(1e3*x for x in (t1-t0,t2-t1,t3-t2,t4-t3,info['ms_device']*1e-3,t5-t4,t6-t5)))
