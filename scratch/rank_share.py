"""One rank's share of a cell-sharded C4 sweep, emulated on one GPU:
rb_plan_sweep_local(rank, stride) for stride = world size; per-stage device ms."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabacus_b200 import configs
from rabacus_b200.plan import Plan
cfg = configs.c4_sphere_bgnd(4096, 512, 256, False)
plan = Plan(configs.GEOM_IDS[cfg["geometry"]], 1, 4096, 512, 256, find_Teq=False, tol=1e-4, max_sweeps=-(10**9))
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
for stride in (8, 4):
    for chunks in ("", "1", "2", "4", "8"):
        if chunks: os.environ["RB_COLS_CHUNKS"] = chunks
        else: os.environ.pop("RB_COLS_CHUNKS", None)
        for rank in (0, stride - 1):
            r = run(rank, stride)
            print("stride %d rank %d chunks %-7s cols %.4f nu %.4f cell %.4f finish %.4f"
                  % (stride, rank, chunks or "default", r["ms_columns"], r["ms_nu"], r["ms_cell"], r["ms_finish"]), flush=True)
