"""Steady-state stage times of the C5 batch (a slice of it): B problems dealt from the full
8192-problem sweep, a full solve; prints the summed device time per stage and per sweep."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabacus_b200 import configs
from rabacus_b200.plan import Plan, deal_problems

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
find_Teq = "--find_Teq" in sys.argv
max_sweeps = int(sys.argv[sys.argv.index("--sweeps") + 1]) if "--sweeps" in sys.argv else 0
idx = deal_problems(8192, 0, 8192 // B)
plan = Plan(0, len(idx), 512, 128, 32, find_Teq=find_Teq, tol=1e-4, max_sweeps=max_sweeps)
configs.c5_stage(plan, np.asarray(idx), find_Teq=find_Teq)
for rep in range(2):
    plan.upload()
    t0 = time.perf_counter()
    info = plan.solve()
    dt = time.perf_counter() - t0
print(json.dumps({"B": len(idx), "find_Teq": find_Teq, "wall_s": round(dt, 4), "iters": info["iters"],
                  "ms": {k: round(info[k], 2) for k in ("ms_device", "ms_columns", "ms_nu", "ms_cell", "ms_finish")},
                  "evals_per_s": info["evals"] / dt, "teq_depth": os.environ.get("RB_TEQ_DEPTH", "default")}))
