"""Stage times of the small find_Teq configurations (C2, C3): thin start, one timed sweep."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabacus_b200 import configs
from rabacus_b200.plan import Plan
import torch
for name, cfg in (("C2 slabpln 256x128 Teq", configs.c2_slab_plane(256, 128, True)),
                  ("C3 slab2bgnd 1024x512 Teq", configs.c3_slab_bgnd(1024, 512, True)),
                  ("C4-like sphere 512x128x32 Teq", configs.c4_sphere_bgnd(512, 128, 32, True))):
    Nl, Nnu = cfg["nH"].size, cfg["E_eV"].size
    plan = Plan(configs.GEOM_IDS[cfg["geometry"]], 1, Nl, Nnu, cfg["Nmu"], find_Teq=True, tol=1e-4)
    plan.set_problem(0, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"], cfg["z"], 0.0)
    for rep in range(2):
        plan.upload()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); plan.thin(); plan.sync(); t_thin = time.perf_counter() - t0
        a = plan.last_info()
        t0 = time.perf_counter()
        for _ in range(10):
            plan.sweep_local(0, 1); plan.finish_sweep()
        t_sw = (time.perf_counter() - t0) / 10
        b = plan.last_info()
    d = {k: (b[k] - a[k]) / 10 for k in ("ms_columns", "ms_nu", "ms_cell", "ms_finish")}
    print("%-30s thin %.3f ms | sweep wall %.3f ms: cols %.3f nu %.3f cell %.3f finish %.3f"
          % (name, t_thin * 1e3, t_sw * 1e3, d["ms_columns"], d["ms_nu"], d["ms_cell"], d["ms_finish"]), flush=True)
    plan.close()
