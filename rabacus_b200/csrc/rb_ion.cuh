// rb_ion.cuh -- per-cell ionisation / temperature equilibrium (host+device).
//
// Mirrors the *scalar* solvers of the reference, which are what every sweep
// calls once per cell:
//   solve_ce_s      rabacus/f2py/ion_solver.f90:339-374
//   analytic_H1_s   :161-185        implicit_ionfracs_s :238-276   fpc_s :68-80
//   solve_pce_s     :440-562  (bisection on n_e)
//   get_dudT_s      :1089-1134
//   solve_pcte_s    :729-880  (bisection on T in [7.5e3, 1e5] K)
// The bisection sequences (brackets, mid-points, stopping rules) are kept
// identical so that the same inputs walk the same branch path.
#pragma once
#include "rb_atomic.cuh"

namespace rb {

enum : int {
  ION_OK = 0,
  ION_ERR_MAX_ITER_PCE = RB_ERR_MAX_ITER_PCE,
  ION_ERR_MAX_ITER_PCTE = RB_ERR_MAX_ITER_PCTE,
  ION_ERR_T_NOT_BRACKETED = RB_ERR_T_NOT_BRACKETED,
  ION_ERR_NAN = RB_ERR_NAN
};

struct PhotoRates {
  double H1i, He1i, He2i, H1h, He1h, He2h;
};
// case-A fractions of the three recombining ions (all equal in the geometric
// drivers; independent in the single-zone API)
struct CaseA {
  double H2, He2, He3;
  int libm = 0;  // != 0: Hui & Gnedin fits through libm pow() (diagnostic A/B, rb_internal.h)
};

RB_HD void fpc(rb_frac_t &x) {
  double total = x.xH1 + x.xH2;
  x.xH1 = x.xH1 / total;
  x.xH2 = x.xH2 / total;
  total = x.xHe1 + x.xHe2 + x.xHe3;
  x.xHe1 = x.xHe1 / total;
  x.xHe2 = x.xHe2 / total;
  x.xHe3 = x.xHe3 / total;
}

RB_HD int solve_ce(const rb_kchem_t &k, rb_frac_t &x) {
  double DD = k.reH2 + k.ciH1;
  x.xH1 = k.reH2 / DD;
  x.xH2 = k.ciH1 / DD;
  DD = k.reHe2 * k.reHe3 + k.reHe3 * k.ciHe1 + k.ciHe1 * k.ciHe2;
  x.xHe1 = k.reHe2 * k.reHe3 / DD;
  x.xHe2 = k.reHe3 * k.ciHe1 / DD;
  x.xHe3 = k.ciHe1 * k.ciHe2 / DD;
  fpc(x);
  if (isnan(x.xH1) || isnan(x.xH2) || isnan(x.xHe1) || isnan(x.xHe2) ||
      isnan(x.xHe3))
    return ION_ERR_NAN;
  return ION_OK;
}

RB_HD double analytic_H1(double nH, double y, double reH2, double ciH1,
                         double H1i) {
  const double RR = (ciH1 + reH2) * nH;
  const double QQ = -(H1i + reH2 * nH + RR * (1.0 + y));
  const double PP = reH2 * nH * (1.0 + y);
  double dd = QQ * QQ - 4.0 * RR * PP;
  if (dd < 0.0) dd = 0.0;
  const double q = -0.5 * (QQ - sqrt(dd));
  return PP / q;
}

RB_HD void implicit_ionfracs(double ne, const rb_kchem_t &k, double H1i,
                             double He1i, double He2i, rb_frac_t &x) {
  double R2 = k.reH2 * ne;
  double CG1 = k.ciH1 * ne + H1i;
  const double DH = R2 + CG1;
  x.xH1 = R2 / DH;
  x.xH2 = CG1 / DH;
  R2 = k.reHe2 * ne;
  const double R3 = k.reHe3 * ne;
  CG1 = k.ciHe1 * ne + He1i;
  const double CG2 = k.ciHe2 * ne + He2i;
  const double DHe = (R2 * R3) + (R3 * CG1) + (CG1 * CG2);
  x.xHe1 = R2 * R3 / DHe;
  x.xHe2 = R3 * CG1 / DHe;
  x.xHe3 = CG1 * CG2 / DHe;
  fpc(x);
}

RB_HD double ne_of(double nH, double nHe, const rb_frac_t &x) {
  return x.xH2 * nH + (x.xHe2 + 2.0 * x.xHe3) * nHe;
}

// State of one solve_pce bisection, so that the scalar loop below and the
// lock-step (vector) kernels share the step.
struct PceState {
  double ne, ne_old, ne_left, ne_right, xH1_old, xHe1_old, xHe3_old, err;
};

RB_HD int pce_init(double nH, double nHe, const rb_kchem_t &k, double H1i,
                   rb_frac_t &x, PceState &s) {
  const int rc = solve_ce(k, x);
  if (rc) return rc;
  s.ne_left = ne_of(nH, nHe, x);   // ne_min  :468
  s.ne_right = nH + 2.0 * nHe;     // ne_max  :469
  const double y = (x.xHe2 + 2.0 * x.xHe3) * nHe / nH;
  x.xH1 = analytic_H1(nH, y, k.reH2, k.ciH1, H1i);
  x.xH2 = 1.0 - x.xH1;
  s.ne = ne_of(nH, nHe, x);
  s.ne_old = s.ne;
  s.xH1_old = x.xH1;
  s.xHe1_old = x.xHe1;
  s.xHe3_old = x.xHe3;
  s.err = 1.7976931348623157e308;  // huge(one)
  return ION_OK;
}

// one pass of the loop body :497-528.  `vector_rule` selects the `_v` twin's
// treatment of ratio_ne == 1 (:640-652): ne is then left at the fresh value.
RB_HD void pce_step(double nH, double nHe, const rb_kchem_t &k, double H1i,
                    double He1i, double He2i, rb_frac_t &x, PceState &s,
                    bool vector_rule) {
  implicit_ionfracs(s.ne, k, H1i, He1i, He2i, x);
  s.ne = ne_of(nH, nHe, x);
  const double ratio_ne = s.ne / s.ne_old;
  const double ratio_x = fmax(fmax(x.xH1 / s.xH1_old, x.xHe1 / s.xHe1_old),
                              x.xHe3 / s.xHe3_old);
  s.err = fabs(ratio_x - 1.0);
  if (ratio_ne > 1.0) {
    s.ne = (s.ne_old + s.ne_right) * 0.5;
    s.ne_left = s.ne_old;
  } else if (!vector_rule || ratio_ne < 1.0) {
    s.ne = (s.ne_left + s.ne_old) * 0.5;
    s.ne_right = s.ne_old;
  }
  s.ne_old = s.ne;
  s.xH1_old = x.xH1;
  s.xHe1_old = x.xHe1;
  s.xHe3_old = x.xHe3;
}

RB_HD void zero_absent_species(double nH, double nHe, rb_frac_t &x) {
  if (nH <= 0.0) {
    x.xH1 = 0.0;
    x.xH2 = 0.0;
  }
  if (nHe <= 0.0) {
    x.xHe1 = 0.0;
    x.xHe2 = 0.0;
    x.xHe3 = 0.0;
  }
}

RB_HDN int solve_pce(double nH, double nHe, const rb_kchem_t &k, double H1i,
                    double He1i, double He2i, double tol, rb_frac_t &x) {
  PceState s;
  int rc = pce_init(nH, nHe, k, H1i, x, s);
  if (rc) return rc;
  int iter = 0;
  while (s.err > tol) {
    pce_step(nH, nHe, k, H1i, He1i, He2i, x, s, false);
    iter = iter + 1;
    if (iter > RB_ION_MAX_ITER) return ION_ERR_MAX_ITER_PCE;
  }
  zero_absent_species(nH, nHe, x);
  return ION_OK;
}

RB_HDN int get_dudT(double nH, double nHe, double T, const PhotoRates &p,
                   double z, double Hz, const CaseA &fc, double tol, double &dudT) {
  rb_kchem_t kc;
  rb_kcool_t kl;
  get_kchem_kcool(T, fc.H2, fc.He2, fc.He3, kc, kl, fc.libm);
  rb_frac_t x;
  const int rc = solve_pce(nH, nHe, kc, p.H1i, p.He1i, p.He2i, tol, x);
  if (rc) return rc;
  const double heat =
      get_heat(nH, nHe, p.H1h, p.He1h, p.He2h, x.xH1, x.xHe1, x.xHe2);
  const double cool = get_cool(nH, nHe, T, kl, x, z, Hz);
  dudT = heat - cool;
  return ION_OK;
}

// get_dudT that also hands back the fractions of its solve_pce: when the T
// bisection stops at a temperature that has been probed, these ARE the final
// fractions (solve_pcte's last statement re-solves at that same T).
RB_HDN int probe_T(double nH, double nHe, double T, const PhotoRates &p, double z,
                   double Hz, const CaseA &fc, double tol, double &dudT, rb_frac_t &x) {
  rb_kchem_t kc;
  rb_kcool_t kl;
  get_kchem_kcool(T, fc.H2, fc.He2, fc.He3, kc, kl, fc.libm);
  const int rc = solve_pce(nH, nHe, kc, p.H1i, p.He1i, p.He2i, tol, x);
  if (rc) return rc;
  const double heat =
      get_heat(nH, nHe, p.H1h, p.He1h, p.He2h, x.xH1, x.xHe1, x.xHe2);
  const double cool = get_cool(nH, nHe, T, kl, x, z, Hz);
  dudT = heat - cool;
  return ION_OK;
}

// State of solve_pcte's bisection (:796-850) and its two branches.
struct TBisect {
  double Tl, Tr, T, err;
};
RB_HD void tbisect_step(TBisect &s, bool neg) {
  const double Told = s.T;
  if (neg) {
    s.Tr = s.T;
    s.T = (s.Tl + s.T) * 0.5;
  } else {
    s.Tl = s.T;
    s.T = (s.T + s.Tr) * 0.5;
  }
  s.err = fabs(1.0 - s.T / Told);
}

RB_HDN int solve_pcte(double nH, double nHe, const PhotoRates &p, double z,
                     double Hz, const CaseA &fc, double tol, rb_frac_t &x,
                     double &Tout) {
  double Tleft = RB_TFLOOR, Tright = RB_TMAX;
  double dudT, dudTmin, dudTmax;
  int rc = get_dudT(nH, nHe, Tleft, p, z, Hz, fc, tol, dudTmin);
  if (rc) return rc;
  if (dudTmin < 0.0) {  // :782-794: equilibrium below the floor
    const rb_kchem_t kc = get_kchem(Tleft, fc.H2, fc.He2, fc.He3, fc.libm);
    rc = solve_pce(nH, nHe, kc, p.H1i, p.He1i, p.He2i, tol, x);
    Tout = Tleft;
    return rc;
  }
  rc = get_dudT(nH, nHe, Tright, p, z, Hz, fc, tol, dudTmax);
  if (rc) return rc;
  if (dudTmin < 0.0 || dudTmax > 0.0) return ION_ERR_T_NOT_BRACKETED;

  double T = (log10(Tleft) + log10(Tright)) * 0.5;
  T = pow(10.0, T);
  double err = 1.7976931348623157e308;
  int iter = 0;
  while (err > tol) {
    rc = get_dudT(nH, nHe, T, p, z, Hz, fc, tol, dudT);
    if (rc) return rc;
    const double Told = T;
    if (dudT < 0.0) {
      Tright = T;
      T = (Tleft + T) * 0.5;
    } else {
      Tleft = T;
      T = (T + Tright) * 0.5;
    }
    err = fabs(1.0 - T / Told);
    iter = iter + 1;
    if (iter > RB_ION_MAX_ITER) return ION_ERR_MAX_ITER_PCTE;
  }
  const rb_kchem_t kc = get_kchem(T, fc.H2, fc.He2, fc.He3, fc.libm);
  rc = solve_pce(nH, nHe, kc, p.H1i, p.He1i, p.He2i, tol, x);
  if (rc) return rc;
  zero_absent_species(nH, nHe, x);
  Tout = T;
  return ION_OK;
}

}  // namespace rb
