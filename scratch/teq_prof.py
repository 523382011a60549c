import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabacus_b200 import configs, _lib
from rabacus_b200.plan import Plan
libm = int(sys.argv[1]); nprob = int(sys.argv[2])
probs = configs.c5_sweep(find_Teq=True)[:: 8192 // nprob]
plan = Plan(0, len(probs), 512, 128, 32, find_Teq=True, tol=1e-4, max_sweeps=3)
for b, c in enumerate(probs):
    plan.set_problem(b, c["Edges"], c["nH"], c["nHe"], c["T"], c["E_eV"], c["shape"], c["z"])
plan.upload()
plan.set_libm_pow(libm)
info = plan.solve()
print(info)
