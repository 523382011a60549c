import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabacus_b200 import configs
from rabacus_b200.plan import Plan
cfg = configs.c4_sphere_bgnd(4096, 512, 256, False)
plan = Plan(0, 1, 4096, 512, 256, find_Teq=False, tol=1e-4, max_sweeps=-(10**9))
plan.set_problem(0, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"], cfg["z"], 0.0)
plan.upload(); plan.thin(); plan.sync()
for _ in range(4):
    plan.sweep_local(7, -8); plan.finish_sweep()
