"""One rank's share of a cell-sharded C4 sweep, emulated on one GPU (round 2: far-field
columns): rb_plan_sweep_local(rank, stride) for stride = world size; per-stage device ms."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabacus_b200 import configs
from rabacus_b200.plan import Plan
teq = "--find_Teq" in sys.argv
cfg = configs.c4_sphere_bgnd(4096, 512, 256, teq)
plan = Plan(configs.GEOM_IDS[cfg["geometry"]], 1, 4096, 512, 256, find_Teq=teq, tol=1e-4, max_sweeps=-(10**9))
plan.set_problem(0, cfg["Edges"], cfg["nH"], cfg["nHe"], cfg["T"], cfg["E_eV"], cfg["shape"], cfg["z"], 0.0)
plan.upload(); plan.thin(); plan.sync()
def run(rank, stride, n=12):
    for _ in range(3):
        plan.sweep_local(rank, stride); plan.finish_sweep()
    a = plan.last_info()
    for _ in range(n):
        plan.sweep_local(rank, stride); plan.finish_sweep()
    b = plan.last_info()
    return {k: (b[k] - a[k]) / n for k in ("ms_columns", "ms_nu", "ms_cell", "ms_finish")}
deals = [(1, 0)] + [(sgn * w, r) for w in (2, 4, 8) for sgn in (1, -1) for r in (0, w - 1)]
for stride, rank in deals:
    if True:
        r = run(rank, stride)
        tot = sum(r.values())
        print("stride %d rank %d  cols %.4f nu %.4f cell %.4f finish %.4f  total %.4f  (ideal %.4f)"
              % (stride, rank, r["ms_columns"], r["ms_nu"], r["ms_cell"], r["ms_finish"], tot, 0.0), flush=True)
