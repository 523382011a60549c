"""find_Teq timings: C4 SphereBgnd solve (one problem) with stage times."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rabacus_b200 import configs
from rabacus_b200.plan import Plan
cfg = configs.c4_sphere_bgnd(4096, 512, 256, True)
plan = Plan(configs.GEOM_IDS[cfg["geometry"]], 1, 4096, 512, 256, find_Teq=True, tol=1e-4)
plan.set_problem(0, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"], cfg["z"], 0.0)
for rep in range(3):
    plan.upload()
    t0 = time.perf_counter(); info = plan.solve(); dt = time.perf_counter() - t0
    print("C4 find_Teq solve: %.1f ms wall, %d sweeps, stages ms: cols %.1f nu %.1f cell %.1f finish %.1f"
          % (dt * 1e3, info["iters"], info["ms_columns"], info["ms_nu"], info["ms_cell"], info["ms_finish"]), flush=True)
plan.close()
