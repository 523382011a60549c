import ctypes as C, numpy as np, mpmath as mp, sys
sys.path.insert(0, '/root/repo')
from rabacus_b200 import _lib
L = _lib.lib()
def P(a): return a.ctypes.data_as(C.POINTER(C.c_double))
L.rbi_diag_pow.argtypes = [C.POINTER(C.c_double)]*2 + [C.c_int, C.POINTER(C.c_double), C.c_int]
rng = np.random.default_rng(1)
n = 400000
# bases / exponents as in the fits: lam in [1e-6, 2e6], (1+u), T, 10 ; y in [-4, 4]
x = np.concatenate([10**rng.uniform(-6, 6.5, n//2), 1 + 10**rng.uniform(-8, 7, n//4), rng.uniform(0.5, 2.0, n//4)])
y = rng.uniform(-4, 4, n)
out = np.empty(n)
assert L.rbi_diag_pow(P(x), P(y), n, P(out), 0) == 0
ref = np.power(x, y)
ulp = np.spacing(np.abs(ref))
d = (out - ref)/ulp
print("vs glibc pow: max |diff| ulp", np.abs(d).max(), "fraction differing", np.mean(d != 0))
# true ulp error on a subsample
mp.mp.prec = 200
idx = rng.choice(n, 20000, replace=False)
worst = 0; worst_g = 0
for i in idx:
    t = mp.power(mp.mpf(x[i]), mp.mpf(y[i]))
    e = abs((mp.mpf(out[i]) - t) / mp.mpf(ulp[i]))
    g = abs((mp.mpf(ref[i]) - t) / mp.mpf(ulp[i]))
    worst = max(worst, e); worst_g = max(worst_g, g)
print("true error: fast %.4f ulp, glibc %.4f ulp" % (worst, worst_g))
