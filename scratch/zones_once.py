"""rb_solve_pce / rb_solve_pcte at 1e6 zones once (the target of the ncu capture of the
single-zone batch kernels, profiles/r02_zones_ncu.txt)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rabacus_b200 import rabacus_fc as fc
n = 10 ** 6
u = (np.arange(n, dtype=np.float64) + 0.5) / n
nH = 10.0 ** (-4.0 + 6.0 * u); nHe = nH * 0.08
T = 10.0 ** (3.9 + 1.1 * ((u * 7919.0) % 1.0))
H1i = 10.0 ** (-16.0 + 4.0 * ((u * 104729.0) % 1.0)); He1i, He2i = 0.6 * H1i, 0.02 * H1i
k = fc.get_kchem_v(T, 1.0, 1.0, 1.0, 1, n)
for rep in range(2):
    x = fc._solve_pce((nH, nHe) + tuple(k) + (H1i, He1i, He2i), 1e-6, n, 0)
    y = fc._solve_pcte((nH, nHe, H1i, He1i, He2i, H1i * 5e-12, He1i * 8e-12, He2i * 2e-11), 3.0, 0.0,
                       (1.0, 1.0, 1.0), 1, 1e-6, n, 0)
print("ok", float(np.min(y[5])), float(np.max(y[5])))
