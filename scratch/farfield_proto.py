"""Prototype of the far-field series for the ray-column sums (numpy, CPU):
sum_{k<=K} sqrt(E_k^2 - p^2) D_k  =  E_K sum_n b_n u^n Mt_n[K],  u = (p/E_K)^2,
Mt_n[K] = sum_{k<=K} D_k (E_K/E_k)^(2n-1)   (recurrence in K)."""
import sys
import numpy as np
import mpmath as mp
sys.path.insert(0, '/root/repo')
from rabacus_b200 import configs

def cheb_coeffs(umax, deg):
    """monomial coefficients (in u) of the Chebyshev interpolant of sqrt(1-u) on [0,umax]"""
    mp.mp.prec = 300
    n = deg + 1
    nodes = [mp.cos(mp.pi * (2 * j + 1) / (2 * n)) for j in range(n)]
    us = [(x + 1) * umax / 2 for x in nodes]
    f = [mp.sqrt(1 - u) for u in us]
    # solve Vandermonde in high precision
    A = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            A[i, j] = us[i] ** j
    b = mp.lu_solve(A, mp.matrix(f))
    return [b[i] for i in range(n)]

def check(ratio, deg, Nl=4096, nmu=16):
    umax = 1.0 / ratio ** 2
    b = cheb_coeffs(umax, deg)
    bd = np.array([float(x) for x in b])
    # max error of the polynomial itself
    uu = np.linspace(0, umax, 2001)
    mp.mp.prec = 200
    err = max(abs(float(mp.polyval(b[::-1], mp.mpf(float(u))) - mp.sqrt(1 - mp.mpf(float(u))))) for u in uu[::20])
    cfg = configs.c4_sphere_bgnd(Nl, 16, 2 * nmu)
    E = cfg["Edges"][::-1].copy()          # decreasing
    rng = np.random.default_rng(0)
    c = cfg["nH"][::-1] * rng.uniform(1e-6, 1.0, Nl)   # absorber density per shell (outer first)
    D = np.empty(Nl + 1); D[0] = c[0]; D[1:Nl] = c[1:] - c[:-1]; D[Nl] = -c[-1]
    # moments by recurrence (double)
    nn = np.arange(deg + 1)
    Mt = np.zeros((Nl + 1, deg + 1))
    Mt[0] = D[0]
    for K in range(1, Nl + 1):
        fac = (E[K] / E[K - 1]) if E[K - 1] > 0 else 0.0
        Mt[K] = Mt[K - 1] * fac ** (2 * nn - 1) + D[K]
    x, w = np.polynomial.legendre.leggauss(2 * nmu)
    worst = 0.0
    Eld = E.astype(np.longdouble); Dld = D.astype(np.longdouble)
    for il in range(0, Nl, 97):
        r = 0.5 * (E[il] + E[il + 1])
        for mu in x[:nmu]:
            p = r * np.sqrt(1 - mu * mu)
            m = np.argmax(E <= p) - 1
            K = np.searchsorted(-E, -ratio * p, side="right") - 1   # last k with E_k >= ratio p
            K = min(K, m)
            if K < 0:
                continue
            u = (p / E[K]) ** 2
            h = 0.0
            for n in range(deg, -1, -1):
                h = h * u + bd[n] * Mt[K, n]
            far = E[K] * h
            exact = np.sum(np.sqrt(Eld[:K + 1] ** 2 - np.longdouble(p) ** 2) * Dld[:K + 1])
            # scale: the whole-ray sum
            tot = np.sum(np.sqrt(Eld[:m + 1] ** 2 - np.longdouble(p) ** 2) * Dld[:m + 1])
            worst = max(worst, abs(float(far - exact)) / abs(float(tot)))
    print("ratio %.2f deg %d: poly err %.1e, worst far-field error / whole column %.2e" % (ratio, deg, err, worst))

for ratio, deg in ((2.0, 14), (2.0, 16), (1.5, 20), (1.5, 22), (1.25, 28), (1.25, 32)):
    check(ratio, deg)
