"""A/B: Hui & Gnedin fits through rb_fastpow vs libm pow; C5-like find_Teq batch and C4."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rabacus_b200 import configs, _lib
from rabacus_b200.plan import Plan, FIELDS
L = _lib.lib()
nprob = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
probs = configs.c5_sweep(find_Teq=True)[:: 8192 // nprob]
res = {}
modes = [int(c) for c in os.environ.get("RB_AB_MODES", "1010")]
for libm in modes:
    plan = Plan(0, len(probs), 512, 128, 32, find_Teq=True, tol=1e-4)
    plan.set_libm_pow(libm)
    for b, c in enumerate(probs):
        plan.set_problem(b, c["Edges"], c["nH"], c["nHe"], c["T"], c["E_eV"], c["shape"], c["z"])
    plan.upload()
    t0 = time.perf_counter(); info = plan.solve(); dt = time.perf_counter() - t0
    out = [plan.get_problem(b) for b in range(0, len(probs), 97)]
    plan.close()
    print("libm=%d: %d problems, wall %.3f s, max sweeps %d, per-sweep ms cols %.1f nu %.1f cell %.1f"
          % (libm, len(probs), dt, info["iters"], info["ms_columns"], info["ms_nu"], info["ms_cell"]), flush=True)
    res[libm] = out
if len(set(modes)) < 2:
    sys.exit(0)
worst = {k: 0.0 for k in FIELDS}
for a, b in zip(res[0], res[1]):
    for k in FIELDS:
        with np.errstate(all="ignore"):
            r = np.where(a[k] == b[k], 0.0, np.abs(a[k] - b[k]) / np.abs(b[k]))
        worst[k] = max(worst[k], float(np.nanmax(r)))
print("fast vs libm, max rel diff:", {k: "%.1e" % v for k, v in worst.items()})
print("T identical:", all(np.array_equal(a["T"], b["T"]) for a, b in zip(res[0], res[1])))
