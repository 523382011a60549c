"""Independent second derivation of the hot path -- TEST INFRASTRUCTURE ONLY.

Written from the reference's Fortran (file:line cited per function), NOT from
oracle/rb_oracle_*.c, in numpy with a selectable working precision ``F``
(``np.float64``: same IEEE operations in the reference's own statement order, so the
bisection paths can be compared step by step; ``np.longdouble``: 64-bit mantissa, the
"truth" the float64 codings are measured against).  It keeps the reference's loop
structure (per-segment ds_segment calls, six separate frequency integrals), so it is slow
and only used at small sizes by tests/test_second_oracle.py to pin the C oracle:
one sweep of every geometry -- SphereBgnd (Jacobi order and the reference's sequential
Gauss-Seidel order), Slab2Bgnd, SlabPln, Slab2Pln, SphereStromgren (monochromatic and
polychromatic point source) --, the single-zone solvers solve_pce_s, solve_pcte_s, the
lock-step solve_pce_v / solve_pcte_v behind the optically thin start, the recombination-photon
closures (outward, radial, isotropic; slab), set_column_vs_impact, and the drivers of all
five solvers from the thin start to convergence (equal sweep counts).
"""
import numpy as np

# rabacus/f2py/physical_constants.f90:7-21
PI, KB, H_ERG, EV2ERG, ERG2EV = (3.141592653589793, 1.3806488e-16, 6.62606957e-27,
                                 1.60217653e-12, 6.241509479607719e11)
T_H1, T_He1, T_He2 = 1.57807e5, 2.85335e5, 6.31515e5      # hui_gnedin_97.f90:6-8
TFLOOR, TMAX, MAX_ITER = 7.5e3, 1.0e5, 5000                # ion_solver.f90:13-15
# rabacus/atomic/verner/photox/photo.dat:1-3  (Eth E0 sigma0 ya P yw y0 y1)
V96 = ((13.60, 0.4298, 5.475e4, 32.88, 2.963, 0.0, 0.0, 0.0),
       (24.59, 13.61, 949.2, 1.469, 3.188, 2.039, 0.4434, 2.136),
       (54.42, 1.720, 1.369e4, 32.88, 2.963, 0.0, 0.0, 0.0))


def sigma_v96(X, E, F):
    """verner_96.f90:133-151"""
    Eth, E0, s0, ya, P, yw, y0, y1 = (F(v) for v in V96[X])
    E = np.asarray(E, dtype=F)
    x = E / E0 - y0
    y = np.sqrt(x * x + y1 * y1)
    t1 = (x - F(1)) * (x - F(1)) + yw * yw
    t2 = y ** (F(0.5) * P - F(5.5))
    t3 = (F(1) + np.sqrt(y / ya)) ** (-P)
    return np.where(E < Eth, F(0), (s0 * F(1.0e-18)) * (t1 * t2 * t3))


def _ferland(lam, t1, t2, t3, t4, t5, F):      # hui_gnedin_97.f90:21-45 and twins
    return F(t1) * lam ** F(t2) / (F(1) + (lam / F(t3)) ** F(t4)) ** F(t5)


def _lotz(T, lam, t1, t2, t3, t4, t5, F):      # hui_gnedin_97.f90:106-150
    return (F(t1) * T ** F(-1.5) * np.exp(-lam * F(0.5)) * lam ** F(t2)
            / (F(1) + (lam / F(t3)) ** F(t4)) ** F(t5))


def _mix(a, b, f, F):                          # chem_cool_rates.f90:67
    return F(10) ** (np.log10(a) * f + np.log10(b) * (F(1) - f))


def _di(T, F):                                 # hui_gnedin_97.f90:67-74
    lam = F(2) * F(T_He2) / T
    return (F(1.90e-3) * T ** F(-1.5) * np.exp(F(-0.75) * lam * F(0.5))
            * (F(1) + F(0.3) * np.exp(F(-0.15) * lam * F(0.5))))


def get_kchem(T, fc, F):
    """chem_cool_rates.f90:45-88 -> (reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2)"""
    T, fc = F(T), [F(v) for v in fc]
    l1, l2, l3 = F(2) * F(T_H1) / T, F(2) * F(T_He1) / T, F(2) * F(T_He2) / T
    reH2 = _mix(_ferland(l1, 1.269e-13, 1.503, 0.522, 0.470, 1.923, F),
                _ferland(l1, 2.753e-14, 1.500, 2.740, 0.407, 2.242, F), fc[0], F)
    reHe2 = _mix(F(3.0e-14) * l2 ** F(0.654), F(1.26e-14) * l2 ** F(0.750), fc[1], F)
    reHe2 = reHe2 + _di(T, F)
    reHe3 = _mix(_ferland(l3, F(2) * F(1.269e-13), 1.503, 0.522, 0.470, 1.923, F),
                 _ferland(l3, F(2) * F(2.753e-14), 1.500, 2.740, 0.407, 2.242, F), fc[2], F)
    return (reH2, reHe2, reHe3, _lotz(T, l1, 21.11, -1.089, 0.354, 0.874, 1.101, F),
            _lotz(T, l2, 32.38, -1.146, 0.416, 0.987, 1.056, F),
            _lotz(T, l3, 19.95, -1.089, 0.553, 0.735, 1.275, F))


def get_kcool(T, fc, F):
    """chem_cool_rates.f90:147-204 with hui_gnedin_97.f90:162-322 ->
    (recH2, recHe2, recHe3, cicH1, cicHe1, cicHe2, cecH1, cecHe2, bremss, compton)"""
    T, fc = F(T), [F(v) for v in fc]
    l1, l2, l3 = F(2) * F(T_H1) / T, F(2) * F(T_He1) / T, F(2) * F(T_He2) / T

    def rec(lam, t1, t2, t3, t4, t5):
        return t1 * T * lam ** F(t2) / (F(1) + (lam / F(t3)) ** F(t4)) ** F(t5)
    recH2 = _mix(rec(l1, F(1.778e-29), 1.965, 0.541, 0.502, 2.697),
                 rec(l1, F(3.435e-30), 1.970, 2.250, 0.376, 3.720), fc[0], F)
    recHe2 = _mix(F(KB) * T * (F(3.0e-14) * l2 ** F(0.654)),
                  F(KB) * T * (F(1.26e-14) * l2 ** F(0.750)), fc[1], F)
    recHe2 = recHe2 + F(0.75) * F(KB) * F(T_He2) * _di(T, F)
    recHe3 = _mix(rec(l3, F(8) * F(1.778e-29), 1.965, 0.541, 0.502, 2.697),
                  rec(l3, F(8) * F(3.435e-30), 1.970, 2.250, 0.376, 3.720), fc[2], F)
    cic = (F(KB) * F(T_H1) * _lotz(T, l1, 21.11, -1.089, 0.354, 0.874, 1.101, F),
           F(KB) * F(T_He1) * _lotz(T, l2, 32.38, -1.146, 0.416, 0.987, 1.056, F),
           F(KB) * F(T_He2) * _lotz(T, l3, 19.95, -1.089, 0.553, 0.735, 1.275, F))
    den = F(1) + np.sqrt(T * F(1.0e-5))
    cecH1 = F(7.5e-19) * np.exp(F(-0.75) * l1 * F(0.5)) / den
    cecHe2 = F(5.54e-17) * (F(1) / T) ** F(0.397) * np.exp(F(-0.75) * l3 * F(0.5)) / den
    gff = F(1.1) + F(0.34) * np.exp(-(F(5.5) - np.log10(T)) ** F(2.0) / F(3))
    return (recH2, recHe2, recHe3) + cic + (cecH1, cecHe2,
                                            F(1.43e-27) * np.sqrt(T) * gff, F(5.65e-36))


def _fpc(x):                                   # ion_solver.f90:68-80
    t = x[0] + x[1]
    h = [x[0] / t, x[1] / t]
    t = x[2] + x[3] + x[4]
    return h + [x[2] / t, x[3] / t, x[4] / t]


def _implicit(ne, k, H1i, He1i, He2i):         # ion_solver.f90:238-276
    R2, CG1 = k[0] * ne, k[3] * ne + H1i
    DH = R2 + CG1
    x = [R2 / DH, CG1 / DH]
    R2, R3 = k[1] * ne, k[2] * ne
    CG1, CG2 = k[4] * ne + He1i, k[5] * ne + He2i
    DHe = (R2 * R3) + (R3 * CG1) + (CG1 * CG2)
    return _fpc(x + [R2 * R3 / DHe, R3 * CG1 / DHe, CG1 * CG2 / DHe])


def solve_pce(nH, nHe, k, H1i, He1i, He2i, tol, F, path=None):
    """ion_solver.f90:440-562; ``path`` collects the branch taken at :515-521."""
    nH, nHe, k = F(nH), F(nHe), [F(v) for v in k]
    H1i, He1i, He2i, tol = F(H1i), F(He1i), F(He2i), F(tol)
    DD = k[0] + k[3]                                        # solve_ce_s :339-374
    x = [k[0] / DD, k[3] / DD]
    DD = k[1] * k[2] + k[2] * k[4] + k[4] * k[5]
    x = _fpc(x + [k[1] * k[2] / DD, k[2] * k[4] / DD, k[4] * k[5] / DD])
    ne_left = x[1] * nH + (x[3] + F(2) * x[4]) * nHe
    ne_right = nH + F(2) * nHe
    y = (x[3] + F(2) * x[4]) * nHe / nH
    RR = (k[3] + k[0]) * nH                                 # analytic_H1_s :161-185
    QQ = -(H1i + k[0] * nH + RR * (F(1) + y))
    PP = k[0] * nH * (F(1) + y)
    dd = QQ * QQ - F(4) * RR * PP
    dd = dd if dd >= 0 else F(0)
    x[0] = PP / (-F(0.5) * (QQ - np.sqrt(dd)))
    x[1] = F(1) - x[0]
    ne = x[1] * nH + (x[3] + F(2) * x[4]) * nHe
    ne_old, o = ne, (x[0], x[2], x[4])
    err, it = F(np.inf), 0
    while err > tol:
        x = _implicit(ne, k, H1i, He1i, He2i)
        ne = x[1] * nH + (x[3] + F(2) * x[4]) * nHe
        ratio_ne = ne / ne_old
        err = abs(max(x[0] / o[0], x[2] / o[1], x[4] / o[2]) - F(1))
        up = bool(ratio_ne > F(1))
        if path is not None:
            path.append(up)
        if up:
            ne, ne_left = (ne_old + ne_right) * F(0.5), ne_old
        else:
            ne, ne_right = (ne_left + ne_old) * F(0.5), ne_old
        ne_old, o = ne, (x[0], x[2], x[4])
        it += 1
        assert it <= MAX_ITER
    if nH <= 0:
        x[0] = x[1] = F(0)
    if nHe <= 0:
        x[2] = x[3] = x[4] = F(0)
    return x


def _dudT(nH, nHe, T, p, z, Hz, fc, tol, F):
    """get_dudT_s ion_solver.f90:1089-1134, get_heat/get_cool chem_cool_rates.f90:268-332"""
    x = solve_pce(nH, nHe, get_kchem(T, fc, F), p[0], p[1], p[2], tol, F)
    c = get_kcool(T, fc, F)
    nH, nHe, z, Hz, T = F(nH), F(nHe), F(z), F(Hz), F(T)
    heat = x[0] * nH * F(p[3]) + x[2] * nHe * F(p[4]) + x[3] * nHe * F(p[5])
    ne = x[1] * nH + (x[3] + F(2) * x[4]) * nHe
    cool = (c[0] * ne * x[1] * nH + c[1] * ne * x[3] * nHe + c[2] * ne * x[4] * nHe
            + c[3] * ne * x[0] * nH + c[4] * ne * x[2] * nHe + c[5] * ne * x[3] * nHe
            + c[6] * ne * x[0] * nH + c[7] * ne * x[3] * nHe)
    zp = F(1) + z
    cool = cool + c[9] * ((zp * zp) * (zp * zp)) * (T - F(2.725) * zp) * ne
    cool = cool + c[8] * ne * (nH * x[1] + nHe * x[3] + F(4) * nHe * x[4])
    cool = cool + F(3) * Hz * F(KB) * T * (nH + nHe + ne)
    return heat - cool


def solve_pcte(nH, nHe, p, z, Hz, fc, tol, F, path=None):
    """ion_solver.f90:729-880; p = (H1i, He1i, He2i, H1h, He1h, He2h). -> (x[5], T)"""
    Tl, Tr = F(TFLOOR), F(TMAX)
    if _dudT(nH, nHe, Tl, p, z, Hz, fc, tol, F) < 0:
        return solve_pce(nH, nHe, get_kchem(Tl, fc, F), p[0], p[1], p[2], tol, F), Tl
    assert not _dudT(nH, nHe, Tr, p, z, Hz, fc, tol, F) > 0, "Teq not bracketed"
    T = F(10) ** ((np.log10(Tl) + np.log10(Tr)) * F(0.5))
    err, it = F(np.inf), 0
    while err > F(tol):
        neg = bool(_dudT(nH, nHe, T, p, z, Hz, fc, tol, F) < 0)
        if path is not None:
            path.append(neg)
        Told = T
        if neg:
            Tr, T = T, (Tl + T) * F(0.5)
        else:
            Tl, T = T, (T + Tr) * F(0.5)
        err = abs(F(1) - T / Told)
        it += 1
        assert it <= MAX_ITER
    return solve_pce(nH, nHe, get_kchem(T, fc, F), p[0], p[1], p[2], tol, F), T


def _ssum(a):
    """Fortran sum() without reassociation: strictly left to right (np.sum is pairwise)."""
    return np.cumsum(a)[-1] if a.size else a.dtype.type(0)


def _trap(x, y):                               # utils.f90:11-21
    return x.dtype.type(0.5) * _ssum((y[1:] + y[:-1]) * (x[1:] - x[:-1]))


def _six_rates(E, shape, sig, att, F):
    """source_background.f90:287-354 (+ :186-241 for att = 1): the six integrals, each
    evaluated on its own as the reference does."""
    Eth = [F(v[0]) for v in V96]
    if E.size == 1:                             # :111-112, :302-304
        yy = [F(4) * F(PI) * shape / E * F(ERG2EV) * s * att for s in sig]
        return [y[0] for y in yy] + [(y * (E - t) * F(EV2ERG))[0] for y, t in zip(yy, Eth)]
    base = F(4) * F(PI) * shape / (E * F(H_ERG))
    ion = [_trap(E, base * s * att) for s in sig]
    heat = [_trap(E, base * s * att * (E - t) * F(EV2ERG)) for s, t in zip(sig, Eth)]
    return ion + heat


def _ds_segment(E, il, mu, j, F):              # geometry.f90:90-168 (search repeated per call)
    p = (E[il] + E[il + 1]) * F(0.5) * np.sqrt(F(1) - mu * mu)
    m = next(i for i in range(1, E.size) if E[i] <= p) - 1
    s = lambda i: np.sqrt(E[i] * E[i] - p * p)  # noqa: E731
    if j == m:
        return F(2) * s(j)
    i1 = j if j < m else m - abs(m - j)
    return abs(s(i1 + 1) - s(i1))


def _cell_solve(nH, nHe, T, r6, fcA, find_Teq, z, Hz, tol_in, F):
    if find_Teq:
        return solve_pcte(nH, nHe, r6, z, Hz, (fcA,) * 3, tol_in, F)
    return solve_pce(nH, nHe, get_kchem(T, (fcA,) * 3, F), r6[0], r6[1], r6[2], tol_in, F), F(T)


def _plus(r6, rec, i, F):
    """total rates of cell i: source + recombination photons (sphere_bgnd.f90:858-884)"""
    return list(r6) if rec is None else [a + F(rec[q][i]) for q, a in enumerate(r6)]


def sphere_bgnd_jacobi_sweep(Edges, nH, nHe, T, x, src, mu, wmu, E_eV, shape, fcA, find_Teq,
                             z, Hz, tol, F, cells=None, gauss_seidel=False, rec=None):
    """One sweep of sphere_bgnd.f90:578-918 from state ``x`` [5][Nl] and source rates ``src``
    [6][Nl] (Edges decreasing).  Returns (x_new, T_new, src_new) for ``cells`` (default all).
    Jacobi order by default; ``gauss_seidel``: the loop exactly as the reference runs it with
    one thread (:753-913) -- every cell's new fractions and temperature are stored back before
    the next (inner) cell is traced."""
    E = np.asarray(Edges, dtype=F)
    Ev, sh = np.asarray(E_eV, dtype=F), np.asarray(shape, dtype=F)
    nH_, nHe_ = np.asarray(nH, dtype=F), np.asarray(nHe, dtype=F)
    x = np.array(x, dtype=F)
    T = np.array(T, dtype=F)
    sig = [sigma_v96(X, Ev, F) for X in range(3)]
    sth = [sigma_v96(X, np.array([V96[X][0]]), F)[0] for X in range(3)]   # :213-215
    ra = [sig[X] / sth[X] for X in range(3)]
    Nl = nH_.size
    out = {}
    for il in (range(Nl) if cells is None else cells):
        acc = [F(v) for v in np.asarray(src)[:, il]]
        for imu in range(len(mu)):
            m_ = F(mu[imu])
            p = (E[il] + E[il + 1]) * F(0.5) * np.sqrt(F(1) - m_ * m_)
            m = next(i for i in range(1, Nl + 1) if E[i] <= p) - 1     # geometry.f90:61-71
            j0 = il if m_ <= 0 else 2 * m - il                          # :769-773
            N = [F(0), F(0), F(0)]
            for j in range(j0, 2 * m + 1):
                jnl = m - abs(m - j)
                ds = _ds_segment(E, il, m_, j, F)
                if j == j0:
                    ds = ds * F(0.5)
                N[0] = N[0] + ds * nH_[jnl] * x[0][jnl]
                N[1] = N[1] + ds * nHe_[jnl] * x[2][jnl]
                N[2] = N[2] + ds * nHe_[jnl] * x[3][jnl]
            tau = sum((N[X] * sth[X]) * ra[X] for X in range(3))        # :794-800, s_b.f90:267
            r6 = _six_rates(Ev, sh, sig, np.exp(-tau), F)
            acc = [a + r * F(wmu[imu]) for a, r in zip(acc, r6)]        # :837-848
        acc = [a * F(0.5) for a in acc]                                 # :850-856 (Q1)
        xs, Tn = _cell_solve(nH_[il], nHe_[il], T[il], _plus(acc, rec, il, F), fcA, find_Teq, z,
                             Hz, F(tol) * F(1.0e-2), F)
        out[il] = (xs, Tn, acc)
        if gauss_seidel:                                                # :886-907 in place
            for q in range(5):
                x[q][il] = xs[q]
            T[il] = Tn
    return out


def _e1xa(t, F):                               # zhang_jin_f.f90:15-91
    if t == 0:
        return F(1.0e300)
    if t <= 1:
        return (-np.log(t) + ((((F(1.07857e-3) * t - F(9.76004e-3)) * t + F(5.519968e-2)) * t
                               - F(0.24991055)) * t + F(0.99999193)) * t - F(0.57721566))
    es1 = (((t + F(8.5733287401)) * t + F(18.059016973)) * t + F(8.6347608925)) * t \
        + F(0.2677737343)
    es2 = (((t + F(9.5733223454)) * t + F(25.6329561486)) * t + F(21.0996530827)) * t \
        + F(3.9584969228)
    return np.exp(-t) / t * es1 / es2


def _e2xa(t, F):                               # zhang_jin_f.f90:94-173 (E1 inlined there)
    return np.exp(-t) - t * _e1xa(t, F)


def slab_bgnd_sweep(Edges, nH, nHe, T, x, E_eV, shape, fcA, find_Teq, z, Hz, tol, F,
                    cells=None, rec=None):
    """One sweep of slab_bgnd.f90:532-883 (first half; the mirror is a copy) from state x."""
    E = np.asarray(Edges, dtype=F)
    Ev, sh = np.asarray(E_eV, dtype=F), np.asarray(shape, dtype=F)
    nH_, nHe_, x = np.asarray(nH, dtype=F), np.asarray(nHe, dtype=F), np.asarray(x, dtype=F)
    sig = [sigma_v96(X, Ev, F) for X in range(3)]
    sth = [sigma_v96(X, np.array([V96[X][0]]), F)[0] for X in range(3)]
    ra = [sig[X] / sth[X] for X in range(3)]
    dl = E[1:] - E[:-1]                                                  # :612
    dt = [dl * nH_ * x[0] * sth[0], dl * nHe_ * x[2] * sth[1], dl * nHe_ * x[3] * sth[2]]
    Nl = nH_.size
    out = {}
    for i in (range(Nl // 2) if cells is None else cells):
        r6 = [F(0)] * 6
        for side in (0, 1):                                             # slab_base.f90:63-115
            th = [(_ssum(d[:i]) if side == 0 else _ssum(d[i + 1:])) + F(0.5) * d[i] for d in dt]
            tau = th[0] * ra[0] + th[1] * ra[1] + th[2] * ra[2]        # :701-713, s_b.f90:380
            att = (np.ones_like(tau) if _ssum(tau) == 0                  # s_b.f90:380-383
                   else np.array([_e2xa(t, F) for t in tau], dtype=F))
            r6 = [a + r * F(0.5) for a, r in zip(r6, _six_rates(Ev, sh, sig, att, F))]
        xs, Tn = _cell_solve(nH_[i], nHe_[i], T[i], _plus(r6, rec, i, F), fcA, find_Teq, z, Hz,
                             F(tol) * F(1.0e-2), F)
        out[i] = (xs, Tn, r6)
    return out


def _src_rates(kind, E, shape, sig, att, F, r=None):
    """The six integrals of a plane ('plane': source_plane.f90:178-342) or point ('point':
    source_point.f90:226-395) source seen through attenuation ``att``, each evaluated on its
    own as the reference does; ``r``: shell-centre radius (point source dilution)."""
    Eth = [F(v[0]) for v in V96]
    four, pi = F(4.0), F(PI)
    if E.size == 1:                             # monochromatic branches (nn == 1)
        Fn = shape[0] / E[0] * F(ERG2EV)        # s_plane.f90:101 / s_point.f90:104
        if kind == "point":
            Fn = Fn / (four * pi * r * r)       # s_point.f90:149-158
        ion = [Fn * s[0] * att[0] for s in sig]
        heat = [Fn * s[0] * (E[0] - t) * F(EV2ERG) * att[0] for s, t in zip(sig, Eth)]
        return ion + heat
    if kind == "point":
        base = shape / (E * F(H_ERG) * four * pi * r * r)   # s_point.f90:346
    else:
        base = shape / (E * F(H_ERG))                       # s_plane.f90:296
    ion = [_trap(E, base * s * att) for s in sig]
    heat = [_trap(E, base * s * att * (E - t) * F(EV2ERG)) for s, t in zip(sig, Eth)]
    return ion + heat


def column_sweep(geom, Edges, nH, nHe, T, x, E_eV, shape, fcA, find_Teq, z, Hz, tol, F,
                 cells=None, rec=None):
    """One (Jacobi) sweep of the three column geometries from state ``x`` [5][Nl]:
    'slab_plane_1' (slab_plane.f90:551-830, source at the z = 0 face), 'slab_plane_2'
    (:832-1164: half of the flux from either side; only the first half is solved, the rest
    is the mirrored copy) and 'sphere_stromgren' (sphere_stromgren.f90:595-878, increasing
    Edges, source at the centre).  Optical depths to a layer: the layers below (slab_base.f90:
    63-115 / sphere_base.f90:1232-1367, plain left-to-right sums) plus half of its own."""
    E = np.asarray(Edges, dtype=F)
    Ev, sh = np.asarray(E_eV, dtype=F), np.asarray(shape, dtype=F)
    nH_, nHe_, x = np.asarray(nH, dtype=F), np.asarray(nHe, dtype=F), np.asarray(x, dtype=F)
    sig = [sigma_v96(X, Ev, F) for X in range(3)]
    sth = [sigma_v96(X, np.array([V96[X][0]]), F)[0] for X in range(3)]
    ra = [sig[X] / sth[X] for X in range(3)]
    Nl = nH_.size
    point = geom == "sphere_stromgren"
    dl = np.abs(E[:-1] - E[1:]) if point else E[1:] - E[:-1]     # sphere_base.f90:1535 / :403
    r_c = (E[:-1] + E[1:]) * F(0.5)
    dt = [dl * nH_ * x[0] * sth[0], dl * nHe_ * x[2] * sth[1], dl * nHe_ * x[3] * sth[2]]
    two = geom == "slab_plane_2"
    out = {}
    for i in (range(Nl // 2 if two else Nl) if cells is None else cells):
        def att_from(side):
            th = [(_ssum(d[:i]) if side == 0 else _ssum(d[i + 1:])) + F(0.5) * d[i] for d in dt]
            return np.exp(-(th[0] * ra[0] + th[1] * ra[1] + th[2] * ra[2]))
        kind = "point" if point else "plane"
        if two:                                                   # slab_plane.f90:1011-1048
            lo = _src_rates(kind, Ev, sh, sig, att_from(0), F)
            hi = _src_rates(kind, Ev, sh, sig, att_from(1), F)
            r6 = [a * F(0.5) + b * F(0.5) for a, b in zip(lo, hi)]
        else:
            r6 = _src_rates(kind, Ev, sh, sig, att_from(0), F, r=r_c[i] if point else None)
        xs, Tn = _cell_solve(nH_[i], nHe_[i], T[i], _plus(r6, rec, i, F), fcA, find_Teq, z, Hz,
                             F(tol) * F(1.0e-2), F)
        out[i] = (xs, Tn, r6)
    return out


def _rec_setup(Edges, nH, nHe, T, x, F):
    """The part the three sphere closures share (sphere_base.f90:170-246, :483-547, :890-957):
    threshold cross sections, recombinations to the ground level (case A minus case B) and the
    emission coefficients [photons / (s cm^3 sr)]."""
    E = np.asarray(Edges, dtype=F)
    nH_, nHe_, T_ = (np.asarray(a, dtype=F) for a in (nH, nHe, T))
    x = np.asarray(x, dtype=F)
    Nl = nH_.size
    Eth = [F(v[0]) for v in V96]
    sth = [sigma_v96(X, np.array([V96[X][0]]), F)[0] for X in range(3)]
    s_H1 = [sth[0], sigma_v96(0, np.array([Eth[1]]), F)[0], sigma_v96(0, np.array([Eth[2]]), F)[0]]
    s_He1 = [None, sth[1], sigma_v96(1, np.array([Eth[2]]), F)[0]]
    kA = [get_kchem(T_[i], (F(1),) * 3, F) for i in range(Nl)]
    kB = [get_kchem(T_[i], (F(0),) * 3, F) for i in range(Nl)]
    re1 = np.array([[kA[i][q] - kB[i][q] for i in range(Nl)] for q in range(3)], dtype=F)
    ne = x[1] * nH_ + (x[3] + F(2) * x[4]) * nHe_
    fp = F(4.0) * F(PI)
    em = [(re1[0] * nH_ * x[1] * ne) / fp, (re1[1] * nHe_ * x[3] * ne) / fp,
          (re1[2] * nHe_ * x[4] * ne) / fp]
    return E, nH_, nHe_, x, Nl, Eth, sth, s_H1, s_He1, em


def _rec_rates(Fl, Eth, sth, s_H1, s_He1, scale, F):
    """Fluxes (or 4 pi J) at the three thresholds -> the six rates (sphere_base.f90:319-342,
    :712-733, :1046-1067), quirks included: He1h carries no excess-energy factor, He2h = 0."""
    H1i = (Fl[0] * s_H1[0] + Fl[1] * s_H1[1] + Fl[2] * s_H1[2]) * scale
    He1i = (Fl[1] * s_He1[1] + Fl[2] * s_He1[2]) * scale
    He2i = (Fl[2] * sth[2]) * scale
    H1h = (Fl[1] * (Eth[1] - Eth[0]) * s_H1[1] + Fl[2] * (Eth[2] - Eth[0]) * s_H1[2]) * scale * F(EV2ERG)
    He1h = (Fl[1] * s_He1[1] + Fl[2] * s_He1[2]) * scale * F(EV2ERG)
    return H1i, He1i, He2i, H1h, He1h, np.zeros(Fl.shape[1], dtype=F)


def _rec_shells(E, nH_, nHe_, x, sth, em, F):
    """dtau at the thresholds per shell, mid radii, flux leaving each shell (:166-246)."""
    half, fp = F(0.5), F(4.0) * F(PI)
    dr = np.abs(E[:-1] - E[1:])
    r_c = (E[:-1] + E[1:]) * half
    dt = [dr * nH_ * x[0] * sth[0], dr * nHe_ * x[2] * sth[1], dr * nHe_ * x[3] * sth[2]]
    area = fp * r_c * r_c
    vol = np.abs(F(4.0) / F(3.0) * F(PI) * (E[:-1] ** 3 - E[1:] ** 3))
    F0 = [e * fp * vol / area for e in em]
    return r_c, dt, F0


def _rec_add(Fl, isl, irl, tau, r_c, F0, sth, s_H1, s_He1, w):
    geo = (r_c[irl] / r_c[isl]) * (r_c[irl] / r_c[isl])
    t0 = tau[0] * (s_H1[0] / sth[0])
    Fl[0, isl] = Fl[0, isl] + w * F0[0][irl] * geo * np.exp(-t0)
    t1 = tau[0] * (s_H1[1] / sth[0]) + tau[1] * (s_He1[1] / sth[1])
    Fl[1, isl] = Fl[1, isl] + w * F0[1][irl] * geo * np.exp(-t1)
    t2 = tau[0] * (s_H1[2] / sth[0]) + tau[1] * (s_He1[2] / sth[1]) + tau[2] * (sth[2] / sth[2])
    Fl[2, isl] = Fl[2, isl] + w * F0[2][irl] * geo * np.exp(-t2)


def recomb_outward(Edges, nH, nHe, T, x, F):
    """set_recomb_photons_outward (sphere_base.f90:79-346): recombination photons of every
    shell stream radially outward.  ``Edges`` decreasing (outermost shell first), ``x``
    [5][Nl].  Returns (H1i, He1i, He2i, H1h, He1h, He2h) per solve layer."""
    E, nH_, nHe_, x, Nl, Eth, sth, s_H1, s_He1, em = _rec_setup(Edges, nH, nHe, T, x, F)
    r_c, dt, F0 = _rec_shells(E, nH_, nHe_, x, sth, em, F)
    half = F(0.5)
    Fl = np.zeros((3, Nl), dtype=F)
    for isl in range(Nl):
        tau = [F(0), F(0), F(0)]
        for irl in range(isl, Nl):                                  # :262-312
            for X in range(3):
                if irl == isl or irl == isl + 1:
                    tau[X] = tau[X] + dt[X][irl] * half
                else:
                    tau[X] = tau[X] + dt[X][irl - 1] * half
                    tau[X] = tau[X] + dt[X][irl] * half
            _rec_add(Fl, isl, irl, tau, r_c, F0, sth, s_H1, s_He1, F(1))
    return _rec_rates(Fl, Eth, sth, s_H1, s_He1, F(1), F)


def recomb_radial(Edges, nH, nHe, T, x, F):
    """set_recomb_photons_radial (sphere_base.f90:392-750): half of every shell's photons go
    radially inward (through the centre and out the other side), half outward.  ``Edges``
    INCREASING here (the routine's own orientation, 1-based layers in the Fortran).  The far
    side of the inward walk (layers beyond the solve layer) adds no optical depth -- the
    routine's own statement (:575-590) -- and is reproduced."""
    E, nH_, nHe_, x, Nl, Eth, sth, s_H1, s_He1, em = _rec_setup(Edges, nH, nHe, T, x, F)
    r_c, dt, F0 = _rec_shells(E, nH_, nHe_, x, sth, em, F)
    half = F(0.5)
    Fl = np.zeros((3, Nl), dtype=F)
    for isl in range(1, Nl + 1):                                    # 1-based as in the Fortran
        tau = [F(0), F(0), F(0)]
        for il in range(isl, -Nl - 1, -1):                          # :567-632
            if il == 0:
                continue
            irl = abs(il)
            for X in range(3):
                if irl == isl or irl == isl - 1:
                    tau[X] = tau[X] + dt[X][irl - 1] * half
                elif irl < isl - 1:
                    tau[X] = tau[X] + dt[X][irl] * half
                    tau[X] = tau[X] + dt[X][irl - 1] * half
            _rec_add(Fl, isl - 1, irl - 1, tau, r_c, F0, sth, s_H1, s_He1, half)
        tau = [F(0), F(0), F(0)]
        for irl in range(isl, Nl + 1):                              # :640-703
            for X in range(3):
                if irl == isl or irl == isl + 1:
                    tau[X] = tau[X] + dt[X][irl - 1] * half
                else:
                    tau[X] = tau[X] + dt[X][irl - 1] * half
                    tau[X] = tau[X] + dt[X][irl - 2] * half
            _rec_add(Fl, isl - 1, irl - 1, tau, r_c, F0, sth, s_H1, s_He1, half)
    return _rec_rates(Fl, Eth, sth, s_H1, s_He1, F(1), F)


def recomb_isotropic(Edges, nH, nHe, T, x, mu, wmu, em_fac, F):
    """set_recomb_photons_isotropic (sphere_base.f90:800-1083): formal solution along the
    ``Nmu`` rays of every solve layer at the three thresholds, then the angle average.
    ``Edges`` decreasing; ``mu``, ``wmu`` the Gauss-Legendre rule (:963)."""
    E, nH_, nHe_, x, Nl, Eth, sth, s_H1, s_He1, em = _rec_setup(Edges, nH, nHe, T, x, F)
    half = F(0.5)
    em = [e * F(f) for e, f in zip(em, em_fac)]                     # :955-957
    dk = [nH_ * x[0] * s_H1[0],
          nH_ * x[0] * s_H1[1] + nHe_ * x[2] * s_He1[1],
          nH_ * x[0] * s_H1[2] + nHe_ * x[2] * s_He1[2] + nHe_ * x[3] * sth[2]]   # :906-918
    J = np.zeros((3, Nl), dtype=F)
    for il in range(Nl):
        I = np.zeros((3, len(mu)), dtype=F)
        for imu in range(len(mu)):
            m_ = F(mu[imu])
            p = (E[il] + E[il + 1]) * half * np.sqrt(F(1) - m_ * m_)
            m = next(i for i in range(1, Nl + 1) if E[i] <= p) - 1
            j0 = il if m_ <= 0 else 2 * m - il
            tau = [F(0), F(0), F(0)]
            for j in range(j0, 2 * m + 1):                          # :1003-1023
                jnl = m - abs(m - j)
                ds = _ds_segment(E, il, m_, j, F)
                if j == j0:
                    ds = ds * half
                for q in range(3):
                    tau[q] = tau[q] + ds * dk[q][jnl]
                    I[q, imu] = I[q, imu] + em[q][jnl] * ds * np.exp(-tau[q])
        for q in range(3):
            J[q, il] = _ssum(I[q] * np.asarray(wmu, dtype=F))       # :1029-1031
    J = J * half
    return _rec_rates(J, Eth, sth, s_H1, s_He1, F(4.0) * F(PI), F)


def recomb_slab(Edges, nH, nHe, T, x, F):
    """set_recomb_photons of the slabs (slab_base.f90:163-560): plane-parallel diffuse field,
    J at the centre of every layer from the emission on the edges seen through E1 of the
    optical depth, trapezoid in z below and above, plus the layer's own contribution."""
    E, nH_, nHe_, x, Nl, Eth, sth, s_H1, s_He1, em = _rec_setup(Edges, nH, nHe, T, x, F)
    half = F(0.5)
    dl = E[1:] - E[:-1]
    dt = [dl * nH_ * x[0] * sth[0], dl * nHe_ * x[2] * sth[1], dl * nHe_ * x[3] * sth[2]]
    ra = [(F(1),), (s_H1[1] / s_H1[0], F(1)), (s_H1[2] / s_H1[0], s_He1[2] / s_He1[1], F(1))]
    J = np.zeros((3, Nl), dtype=F)
    for q in range(3):
        lo = np.zeros(Nl + 1, dtype=F)                               # :333-351
        for k in range(1, Nl + 1):
            acc = lo[k - 1]
            for X in range(q + 1):
                acc = acc + dt[X][k - 1] * ra[q][X]
            lo[k] = acc
        e = np.zeros(Nl + 1, dtype=F)                                # :367-384
        e[0], e[Nl] = em[q][0] * half, em[q][Nl - 1] * half
        e[1:Nl] = (em[q][:-1] + em[q][1:]) * half
        for k in range(Nl):                                          # :393-512
            tau = (lo[k + 1] - lo[k]) * half
            J_lo = dl[k] * em[q][k] * _e1xa(tau, F) * half
            J_hi = J_lo
            g = np.zeros(Nl + 1, dtype=F)
            if k != 0:
                tau = (lo[k + 1] - lo[k]) * half
                g[k] = e[k] * _e1xa(tau, F)
                for i in range(k - 1, -1, -1):
                    tau = tau + (lo[i + 1] - lo[i])
                    g[i] = e[i] * _e1xa(tau, F)
                J_lo = J_lo + _trap(E[:k + 1], g[:k + 1])
            if k != Nl - 1:
                tau = (lo[k + 1] - lo[k]) * half
                g[k + 1] = e[k + 1] * _e1xa(tau, F)
                for i in range(k + 1, Nl):
                    tau = tau + (lo[i + 1] - lo[i])
                    g[i + 1] = e[i + 1] * _e1xa(tau, F)
                J_hi = J_hi + _trap(E[k + 1:], g[k + 1:])
            J[q, k] = (J_lo + J_hi) * F(4.0) * F(PI) * half          # :523-532
    ev = F(EV2ERG)
    H1i = J[2] * s_H1[2] + J[1] * s_H1[1] + J[0] * s_H1[0]           # :538-556, as written there:
    He1i = J[2] * s_He1[2] + J[1] * s_He1[1]                         # He1h uses E_He2 - E_H1
    He2i = J[2] * sth[2]
    H1h = J[2] * (Eth[2] - Eth[0]) * s_H1[2] * ev + J[1] * (Eth[1] - Eth[0]) * s_H1[1] * ev
    He1h = J[2] * (Eth[2] - Eth[0]) * s_He1[2] * ev
    return H1i, He1i, He2i, H1h, He1h, np.zeros(Nl, dtype=F)


def column_vs_impact(Edges, nH, nHe, x, F):
    """set_column_vs_impact (sphere_base.f90:1115-1205): total and absorber columns along the
    full chord (mu = 0, j = 0 .. 2m) through the centre radius of every shell.  ``Edges``
    decreasing.  Returns (NH, NHe, NH1, NHe1, NHe2)."""
    E = np.asarray(Edges, dtype=F)
    nH_, nHe_, x = np.asarray(nH, dtype=F), np.asarray(nHe, dtype=F), np.asarray(x, dtype=F)
    Nl = nH_.size
    out = np.zeros((5, Nl), dtype=F)
    for il in range(Nl):
        p = (E[il] + E[il + 1]) * F(0.5) * np.sqrt(F(1) - F(0) * F(0))
        m = next(i for i in range(1, Nl + 1) if E[i] <= p) - 1
        for j in range(0, 2 * m + 1):
            jnl = m - abs(m - j)
            ds = _ds_segment(E, il, F(0), j, F)
            out[0, il] = out[0, il] + ds * nH_[jnl]
            out[1, il] = out[1, il] + ds * nHe_[jnl]
            out[2, il] = out[2, il] + ds * nH_[jnl] * x[0][jnl]
            out[3, il] = out[3, il] + ds * nHe_[jnl] * x[2][jnl]
            out[4, il] = out[4, il] + ds * nHe_[jnl] * x[3][jnl]
    return tuple(out)


def _ce(k, F):                                  # solve_ce ion_solver.f90:339-374 / :377-412
    DD = k[0] + k[3]
    x = [k[0] / DD, k[3] / DD]
    DD = k[1] * k[2] + k[2] * k[4] + k[4] * k[5]
    return _fpc(x + [k[1] * k[2] / DD, k[2] * k[4] / DD, k[4] * k[5] / DD])


def solve_pce_v(nH, nHe, k, H1i, He1i, He2i, tol, F):
    """solve_pce_v (ion_solver.f90:565-696): the lock-step form -- every zone keeps bisecting
    until ALL zones meet the tolerance, a zone with ratio_ne == 1 keeps its newly computed n_e
    (neither `where` fires).  ``k``: per zone the six coefficients.  Returns per zone x[5]."""
    n = len(nH)
    nH, nHe = [F(v) for v in nH], [F(v) for v in nHe]
    k = [[F(v) for v in kk] for kk in k]
    G = [[F(H1i[i]), F(He1i[i]), F(He2i[i])] for i in range(n)]
    tol = F(tol)
    x = [_ce(k[i], F) for i in range(n)]
    left = [x[i][1] * nH[i] + (x[i][3] + F(2) * x[i][4]) * nHe[i] for i in range(n)]
    right = [nH[i] + F(2) * nHe[i] for i in range(n)]
    ne = [None] * n
    with np.errstate(all="ignore"):
        for i in range(n):                                      # analytic_H1_v :186-210
            y = (x[i][3] + F(2) * x[i][4]) * nHe[i] / nH[i]
            RR = (k[i][3] + k[i][0]) * nH[i]
            QQ = -(G[i][0] + k[i][0] * nH[i] + RR * (F(1) + y))
            PP = k[i][0] * nH[i] * (F(1) + y)
            dd = QQ * QQ - F(4) * RR * PP
            dd = F(0) if dd < 0 else dd
            x[i][0] = PP / (-F(0.5) * (QQ - np.sqrt(dd)))
            x[i][1] = F(1) - x[i][0]
            ne[i] = x[i][1] * nH[i] + (x[i][3] + F(2) * x[i][4]) * nHe[i]
        ne_old = list(ne)
        old = [(x[i][0], x[i][2], x[i][4]) for i in range(n)]
        err = [F(np.inf)] * n
        it = 0
        while any(e > tol for e in err):
            for i in range(n):
                x[i] = _implicit(ne[i], k[i], *G[i])
                ne[i] = x[i][1] * nH[i] + (x[i][3] + F(2) * x[i][4]) * nHe[i]
                ratio = ne[i] / ne_old[i]
                err[i] = abs(max(x[i][0] / old[i][0], x[i][2] / old[i][1], x[i][4] / old[i][2]) - F(1))
                if ratio > F(1):
                    ne[i], left[i] = (ne_old[i] + right[i]) * F(0.5), ne_old[i]
                elif ratio < F(1):
                    ne[i], right[i] = (left[i] + ne_old[i]) * F(0.5), ne_old[i]
                ne_old[i], old[i] = ne[i], (x[i][0], x[i][2], x[i][4])
            it += 1
            assert it <= MAX_ITER
    for i in range(n):
        if nH[i] <= 0:
            x[i][0] = x[i][1] = F(0)
        if nHe[i] <= 0:
            x[i][2] = x[i][3] = x[i][4] = F(0)
    return x


def _dudT_v(nH, nHe, T, p, z, Hz, fc, tol, F):
    """get_dudT_v (ion_solver.f90:1136-1185): lock-step solve_pce_v inside."""
    n = len(nH)
    x = solve_pce_v(nH, nHe, [get_kchem(T[i], fc, F) for i in range(n)],
                    [p[i][0] for i in range(n)], [p[i][1] for i in range(n)],
                    [p[i][2] for i in range(n)], tol, F)
    out = []
    for i in range(n):
        c = get_kcool(T[i], fc, F)
        nh, nhe, zz, hz, t = F(nH[i]), F(nHe[i]), F(z), F(Hz), F(T[i])
        xi = x[i]
        heat = xi[0] * nh * F(p[i][3]) + xi[2] * nhe * F(p[i][4]) + xi[3] * nhe * F(p[i][5])
        ne = xi[1] * nh + (xi[3] + F(2) * xi[4]) * nhe
        cool = (c[0] * ne * xi[1] * nh + c[1] * ne * xi[3] * nhe + c[2] * ne * xi[4] * nhe
                + c[3] * ne * xi[0] * nh + c[4] * ne * xi[2] * nhe + c[5] * ne * xi[3] * nhe
                + c[6] * ne * xi[0] * nh + c[7] * ne * xi[3] * nhe)
        zp = F(1) + zz
        cool = cool + c[9] * ((zp * zp) * (zp * zp)) * (t - F(2.725) * zp) * ne
        cool = cool + c[8] * ne * (nh * xi[1] + nhe * xi[3] + F(4) * nhe * xi[4])
        cool = cool + F(3) * hz * F(KB) * t * (nh + nhe + ne)
        out.append(heat - cool)
    return out


def solve_pcte_v(nH, nHe, p, z, Hz, fc, tol, F):
    """solve_pcte_v (ion_solver.f90:884-1056): lock-step temperature bisection; zones whose
    balance is negative at the floor stay there.  ``p``: per zone the six rates.
    Returns (per zone x[5], per zone T)."""
    n = len(nH)
    Tl, Tr = [F(TFLOOR)] * n, [F(TMAX)] * n
    dmin = _dudT_v(nH, nHe, Tl, p, z, Hz, fc, tol, F)
    floor = [d < 0 for d in dmin]
    dmax = _dudT_v(nH, nHe, Tr, p, z, Hz, fc, tol, F)
    assert not any(d > 0 and not f for d, f in zip(dmax, floor)), "Teq not bracketed"
    T = [Tl[i] if floor[i] else F(10) ** ((np.log10(Tl[i]) + np.log10(Tr[i])) * F(0.5))
         for i in range(n)]
    err, it = [F(np.inf)] * n, 0
    while any(e > F(tol) for e in err):
        d = _dudT_v(nH, nHe, T, p, z, Hz, fc, tol, F)
        Told = list(T)
        for i in range(n):
            if floor[i]:
                pass
            elif d[i] < 0:
                Tr[i], T[i] = T[i], (Tl[i] + T[i]) * F(0.5)
            elif d[i] > 0:
                Tl[i], T[i] = T[i], (T[i] + Tr[i]) * F(0.5)
            err[i] = abs(F(1) - T[i] / Told[i])
        it += 1
        assert it <= MAX_ITER
    x = solve_pce_v(nH, nHe, [get_kchem(T[i], fc, F) for i in range(n)], [q[0] for q in p],
                    [q[1] for q in p], [q[2] for q in p], tol, F)
    return x, T


def sphere_bgnd_thin(nH, nHe, T, E_eV, shape, fcA, find_Teq, z, Hz, tol, F):
    """sphere_bgnd_optically_thin (sphere_bgnd.f90:379-499): tau = 0 rates in every shell, then
    the lock-step solvers at tol * TOL_FRAC.  Returns (per shell x[5], T, the six rates)."""
    Ev, sh = np.asarray(E_eV, dtype=F), np.asarray(shape, dtype=F)
    sig = [sigma_v96(X, Ev, F) for X in range(3)]
    r6 = _six_rates(Ev, sh, sig, np.ones_like(Ev), F)
    n = len(nH)
    tin = F(tol) * F(1.0e-2)
    if find_Teq:
        x, Tn = solve_pcte_v(nH, nHe, [r6] * n, z, Hz, (fcA,) * 3, tin, F)
    else:
        x = solve_pce_v(nH, nHe, [get_kchem(T[i], (fcA,) * 3, F) for i in range(n)],
                        [r6[0]] * n, [r6[1]] * n, [r6[2]] * n, tin, F)
        Tn = [F(t) for t in T]
    return x, Tn, r6


def sphere_rec(meth, Edges, nH, nHe, T, x, mu, wmu, em_fac, F):
    """The closure ``meth`` ('outward', 'radial', 'isotropic') on a state given with
    DECREASING edges: [6][Nl] rates (radial: the routine's own orientation is increasing)."""
    if meth == "radial":
        rv = lambda a: np.asarray(a)[::-1]  # noqa: E731
        out = recomb_radial(rv(Edges), rv(nH), rv(nHe), rv(T), [rv(v) for v in x], F)
        return [rv(o) for o in out]
    if meth == "outward":
        return list(recomb_outward(Edges, nH, nHe, T, x, F))
    return list(recomb_isotropic(Edges, nH, nHe, T, x, mu, wmu, em_fac, F))


def sphere_bgnd_solve(Edges, nH, nHe, T, mu, wmu, E_eV, shape, fcA, find_Teq, z, Hz, tol, F,
                      gauss_seidel=False, max_iter=500, rec_meth=None, em_fac=(1.0, 1.0, 1.0)):
    """sphere_bgnd_solve (sphere_bgnd.f90:107-314): thin start, then sweeps until
    all(|ne_new / ne_old - 1| < tol) (:284-296; the limit is tested before the increment).
    ``Edges`` decreasing.  ``rec_meth``: the recombination-photon closure evaluated at the top
    of every sweep on the pre-sweep state (:676-745), its rates added to the source rates of
    every cell.  Returns (x [5][Nl], T, src [6][Nl], sweeps[, rec [6][Nl]])."""
    Nl = len(nH)
    xs, Tn, r6 = sphere_bgnd_thin(nH, nHe, T, E_eV, shape, fcA, find_Teq, z, Hz, tol, F)
    rec = None
    x = np.array([[xs[i][q] for i in range(Nl)] for q in range(5)], dtype=F)
    T = np.array(Tn, dtype=F)
    src = np.array([[r6[q]] * Nl for q in range(6)], dtype=F)
    nH_, nHe_ = np.asarray(nH, dtype=F), np.asarray(nHe, dtype=F)
    ne_old = x[1] * nH_ + (x[3] + F(2) * x[4]) * nHe_
    it = 0
    while True:
        if rec_meth:
            rec = sphere_rec(rec_meth, Edges, nH, nHe, T, x, mu, wmu, em_fac, F)
        out = sphere_bgnd_jacobi_sweep(Edges, nH, nHe, T, x, src, mu, wmu, E_eV, shape, fcA,
                                       find_Teq, z, Hz, tol, F, gauss_seidel=gauss_seidel,
                                       rec=rec)
        for il, (xc, Tc, acc) in out.items():
            for q in range(5):
                x[q][il] = xc[q]
            T[il] = Tc
            for q in range(6):
                src[q][il] = acc[q]
        assert it <= max_iter
        it += 1
        ne = x[1] * nH_ + (x[3] + F(2) * x[4]) * nHe_
        done = bool(np.all(np.abs(ne / ne_old - F(1)) < F(tol)))
        ne_old = ne
        if done:
            return (x, T, src, it) if rec is None else (x, T, src, it, np.array(rec, dtype=F))


def solve_column_geometry(geom, Edges, nH, nHe, T, E_eV, shape, fcA, find_Teq, z, Hz, tol, F,
                          rec_meth=None):
    """The drivers of the other four geometries from the thin start to convergence:
    'slab_plane_1' / 'slab_plane_2' (slab_plane.f90:105-345), 'sphere_stromgren'
    (sphere_stromgren.f90:111-383, increasing Edges) and 'slab_bgnd' (slab_bgnd.f90:97-289: no
    error at the sweep limit, it stops after 200).  Thin start: tau = 0 rates (per shell for the
    point source) and the lock-step solvers over all layers; every sweep is Jacobi; the
    two-sided slabs solve the first half and mirror fractions, temperature and rates.
    Returns (x [5][Nl], T, src [6][Nl], sweeps)."""
    E = np.asarray(Edges, dtype=F)
    Ev, sh = np.asarray(E_eV, dtype=F), np.asarray(shape, dtype=F)
    sig = [sigma_v96(X, Ev, F) for X in range(3)]
    Nl = len(nH)
    one = np.ones_like(Ev)
    r_c = (E[:-1] + E[1:]) * F(0.5)
    if geom == "slab_bgnd":
        p = [_six_rates(Ev, sh, sig, one, F)] * Nl
    elif geom == "sphere_stromgren":
        p = [_src_rates("point", Ev, sh, sig, one, F, r=r_c[i]) for i in range(Nl)]
    else:
        p = [_src_rates("plane", Ev, sh, sig, one, F)] * Nl
    tin = F(tol) * F(1.0e-2)
    if find_Teq:
        xs, Tn = solve_pcte_v(nH, nHe, p, z, Hz, (fcA,) * 3, tin, F)
    else:
        xs = solve_pce_v(nH, nHe, [get_kchem(T[i], (fcA,) * 3, F) for i in range(Nl)],
                         [q[0] for q in p], [q[1] for q in p], [q[2] for q in p], tin, F)
        Tn = T
    x = np.array([[xs[i][q] for i in range(Nl)] for q in range(5)], dtype=F)
    T = np.array(Tn, dtype=F)
    src = np.array([[p[i][q] for i in range(Nl)] for q in range(6)], dtype=F)
    nH_, nHe_ = np.asarray(nH, dtype=F), np.asarray(nHe, dtype=F)
    ne_old = x[1] * nH_ + (x[3] + F(2) * x[4]) * nHe_
    mirror = geom in ("slab_bgnd", "slab_plane_2")
    it = 0
    rec = None
    while True:
        if rec_meth == "ray":                   # slabs: slab_base.f90:163-560, pre-sweep state
            rec = [np.array(r, dtype=F) for r in recomb_slab(Edges, nH, nHe, T, x, F)]
        elif rec_meth:                          # Stromgren sphere: increasing edges
            rv = lambda a: np.asarray(a)[::-1]  # noqa: E731
            rec = [rv(r) for r in sphere_rec(rec_meth, rv(Edges), rv(nH), rv(nHe), rv(T),
                                             [rv(v) for v in x], None, None, (1.0,) * 3, F)]
        if geom == "slab_bgnd":
            out = slab_bgnd_sweep(Edges, nH, nHe, T, x, E_eV, shape, fcA, find_Teq, z, Hz, tol, F,
                                  rec=rec)
        else:
            out = column_sweep(geom, Edges, nH, nHe, T, x, E_eV, shape, fcA, find_Teq, z, Hz,
                               tol, F, rec=rec)
        if rec is not None and mirror:          # slab_bgnd.f90:842-865: rates reflected too
            for r in rec:
                for il in out:
                    r[Nl - 1 - il] = r[il]
        for il, (xc, Tc, acc) in out.items():
            for j in ((il, Nl - 1 - il) if mirror else (il,)):
                for q in range(5):
                    x[q][j] = xc[q]
                T[j] = Tc
                for q in range(6):
                    src[q][j] = acc[q]
        if geom != "slab_bgnd":
            assert it <= 500                    # tested before the increment
        it += 1
        ne = x[1] * nH_ + (x[3] + F(2) * x[4]) * nHe_
        done = bool(np.all(np.abs(ne / ne_old - F(1)) < F(tol)))
        ne_old = ne
        if done or (geom == "slab_bgnd" and it >= 200):
            return (x, T, src, it) if rec is None else (x, T, src, it, np.array(rec, dtype=F))
