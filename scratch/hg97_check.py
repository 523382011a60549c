import ctypes as C, numpy as np, sys
sys.path.insert(0, '/root/repo')
from rabacus_b200 import _lib
from oracle import oracle as orc
L = _lib.lib()
def P(a): return a.ctypes.data_as(C.POINTER(C.c_double))
rng = np.random.default_rng(2)
n = 200000
T = 10**rng.uniform(0.0, 12.0, n)
T[:n//2] = 10**rng.uniform(np.log10(7.5e3), 5.0, n//2)
T[-5:] = [0.5, 1e13, 1.0, 1e12, 7500.0]
f = [rng.choice([0.0, 1.0, 0.3, 0.77], n) for _ in range(3)]
out = np.empty((16, n))
assert L.rbi_diag_hg97(P(T), P(f[0]), P(f[1]), P(f[2]), n, P(out), 0, 0) == 0
ref = np.array(list(orc.get_kchem(T, *f)) + list(orc.get_kcool(T, *f)))
names = "reH2 reHe2 reHe3 ciH1 ciHe1 ciHe2 recH2 recHe2 recHe3 cicH1 cicHe1 cicHe2 cecH1 cecHe2 bremss compton".split()
with np.errstate(all="ignore"):
    rel = np.abs(out - ref) / np.abs(ref)
rel = np.where(out == ref, 0.0, rel)
for q, nm in enumerate(names):
    print("%-8s max rel %.3e (solver range %.3e)  identical %.3f" % (nm, np.nanmax(rel[q]), np.nanmax(rel[q][:n//2]), np.mean(out[q] == ref[q])))
