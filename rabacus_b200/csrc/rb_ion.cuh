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

// ---------------------------------------------------------------------------
// The two DECISIONS of a solve_pce step without their divisions.
//
// A step of the n_e bisection (ion_solver.f90:497-528) uses four quotients only to decide:
//   ratio_ne = ne / ne_old  ->  ratio_ne > 1 (< 1)         which bracket moves
//   err = |max(xH1/xH1_old, xHe1/xHe1_old, xHe3/xHe3_old) - 1|  ->  err > tol   loop test
// The iterates themselves (mid-points of the bracket) never see the quotients.  So the
// reference's path is kept bit for bit as long as the two DECISIONS are the reference's:
//  * RN(a/b) > 1  <=>  a > b  and  RN(a/b) < 1  <=>  a < b  for IEEE doubles (a > b gives
//    a/b >= 1 + ulp(b)/b > 1 + 2^-53, which rounds to >= 1 + 2^-52; a < b gives a/b <=
//    1 - 2^-53, representable; zeros / NaN: both sides false): the comparison IS the decision.
//  * err > tol is decided from quotients formed with a refined reciprocal (relative error
//    < 1e-11) whenever the outcome is certain -- |max - 1| further than 2e-10 from tol --
//    and from the IEEE divisions otherwise (also for any non-finite intermediate).
// An IEEE double division is ~25 instructions on this GPU, and with the rate coefficients of
// a probe memoised the divisions were what was left of a temperature probe: 14 -> 10 per step.
// ---------------------------------------------------------------------------
RB_HD bool ratio_err_exceeds(double a1, double b1, double a2, double b2, double a3, double b3,
                             double tol) {
#ifdef __CUDA_ARCH__
  double r1, r2, r3;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r1) : "d"(b1));
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r2) : "d"(b2));
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r3) : "d"(b3));
  r1 = fma(fma(-b1, r1, 1.0), r1, r1);   // one Newton step: 2^-20 -> 2^-40
  r2 = fma(fma(-b2, r2, 1.0), r2, r2);
  r3 = fma(fma(-b3, r3, 1.0), r3, r3);
  const double m = fmax(fmax(a1 * r1, a2 * r2), a3 * r3);
  const double d = fabs(m - 1.0);
  const double slack = 2.0e-10 * (1.0 + d);
  if (d > tol + slack) return true;     // (false for NaN / inf: those take the exact path)
  if (d < tol - slack) return false;
#endif
  return fabs(fmax(fmax(a1 / b1, a2 / b2), a3 / b3) - 1.0) > tol;   // :507-512
}

// State of one solve_pce bisection, so that the scalar loop below and the
// lock-step (vector) kernels share the step.  `err` carries only what the loop tests read
// from it: huge while the step's error exceeds the tolerance, 0 once it does not.
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
                    bool vector_rule, double tol) {
  implicit_ionfracs(s.ne, k, H1i, He1i, He2i, x);
  s.ne = ne_of(nH, nHe, x);
  const bool up = s.ne > s.ne_old, down = s.ne < s.ne_old;   // ratio_ne > 1, < 1 (:503)
  s.err = ratio_err_exceeds(x.xH1, s.xH1_old, x.xHe1, s.xHe1_old, x.xHe3, s.xHe3_old, tol)
              ? 1.7976931348623157e308 : 0.0;
  if (up) {
    s.ne = (s.ne_old + s.ne_right) * 0.5;
    s.ne_left = s.ne_old;
  } else if (!vector_rule || down) {
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
    pce_step(nH, nHe, k, H1i, He1i, He2i, x, s, false, tol);
    iter = iter + 1;
    if (iter > RB_ION_MAX_ITER) return ION_ERR_MAX_ITER_PCE;
  }
  zero_absent_species(nH, nHe, x);
  return ION_OK;
}


// ---------------------------------------------------------------------------
// Memo of the rate coefficients over the temperatures solve_pcte can probe.
//
// solve_pcte_s (ion_solver.f90:796-850) always probes TFLOOR and TMAX, starts its bisection
// at 10^((log10 TFLOOR + log10 TMAX)/2) and halves: whatever the cell, the d-th probe
// temperature is one of 2^(d-1) fixed doubles -- the nodes of one binary tree (heap index:
// root = 1, child of n = 2n after a negative dudT, 2n + 1 otherwise).  For the first
// `depth` levels the 16 coefficients of get_kchem_s + get_kcool_s (they depend on T and the
// plan's case-A fractions only) are evaluated ONCE per (fcA, libm flag) by k_teq_table with
// the very device function the solvers call, so a hit returns the bits the solver would have
// computed; the row's key is the temperature itself, compared bit for bit, so a row can never
// be used for another temperature (a miss evaluates the fits as before).
// rows: [2^depth + 1][16] = {6 kchem, 8 kcool (recH2 .. cecHe2), bremss, T};
// row 0 = TFLOOR, rows 1 .. 2^depth - 1 = tree nodes, row 2^depth = TMAX.
// ---------------------------------------------------------------------------
constexpr int kTeqStride = 16;
struct TeqTab {
  const double *rows;   // nullptr: no table (host code, single-zone API with per-zone fcA)
  int depth, pad_;
};
RB_HD int teq_rows(const TeqTab &tb) { return tb.rows ? (1 << tb.depth) : 0; }   // = row of TMAX
// heap index of the next probe (-1 once the path leaves the table)
RB_HD int teq_child(const TeqTab &tb, int node, bool neg) {
  if (node <= 0 || node >= (1 << 29)) return -1;
  const int c = 2 * node + (neg ? 0 : 1);
  return c < teq_rows(tb) ? c : -1;
}
RB_HD bool teq_hit(const TeqTab &tb, int row, double T, const double *&r) {
  if (row < 0 || tb.rows == nullptr || row > teq_rows(tb)) return false;
  RB_ASSERT(row <= (1 << tb.depth));
  r = tb.rows + (size_t)row * kTeqStride;
#ifdef __CUDA_ARCH__
  return __ldg(r + 15) == T;
#else
  return r[15] == T;
#endif
}
RB_HD void teq_load_chem(const double *r, rb_kchem_t &kc) {
#ifdef __CUDA_ARCH__
  const double2 a = __ldg(reinterpret_cast<const double2 *>(r));
  const double2 b = __ldg(reinterpret_cast<const double2 *>(r) + 1);
  const double2 c = __ldg(reinterpret_cast<const double2 *>(r) + 2);
  kc.reH2 = a.x; kc.reHe2 = a.y; kc.reHe3 = b.x; kc.ciH1 = b.y; kc.ciHe1 = c.x; kc.ciHe2 = c.y;
#else
  kc.reH2 = r[0]; kc.reHe2 = r[1]; kc.reHe3 = r[2]; kc.ciH1 = r[3]; kc.ciHe1 = r[4]; kc.ciHe2 = r[5];
#endif
}
RB_HD void teq_load_cool(const double *r, rb_kcool_t &kl) {
#ifdef __CUDA_ARCH__
  const double2 a = __ldg(reinterpret_cast<const double2 *>(r) + 3);
  const double2 b = __ldg(reinterpret_cast<const double2 *>(r) + 4);
  const double2 c = __ldg(reinterpret_cast<const double2 *>(r) + 5);
  const double2 d = __ldg(reinterpret_cast<const double2 *>(r) + 6);
  const double e = __ldg(r + 14);
  kl.recH2 = a.x; kl.recHe2 = a.y; kl.recHe3 = b.x; kl.cicH1 = b.y; kl.cicHe1 = c.x;
  kl.cicHe2 = c.y; kl.cecH1 = d.x; kl.cecHe2 = d.y; kl.bremss = e;
#else
  kl.recH2 = r[6]; kl.recHe2 = r[7]; kl.recHe3 = r[8]; kl.cicH1 = r[9]; kl.cicHe1 = r[10];
  kl.cicHe2 = r[11]; kl.cecH1 = r[12]; kl.cecHe2 = r[13]; kl.bremss = r[14];
#endif
  kl.compton = 5.65e-36;   // chem_cool_rates.f90:200 (a constant)
}
RB_HD void teq_store(double *r, double T, const rb_kchem_t &kc, const rb_kcool_t &kl) {
  r[0] = kc.reH2; r[1] = kc.reHe2; r[2] = kc.reHe3; r[3] = kc.ciH1; r[4] = kc.ciHe1; r[5] = kc.ciHe2;
  r[6] = kl.recH2; r[7] = kl.recHe2; r[8] = kl.recHe3; r[9] = kl.cicH1; r[10] = kl.cicHe1;
  r[11] = kl.cicHe2; r[12] = kl.cecH1; r[13] = kl.cecHe2; r[14] = kl.bremss; r[15] = T;
}
// get_kchem_s + get_kcool_s / get_kchem_s / get_kcool_s at T, from row `row` when its key is T
RB_HD void kchem_kcool_memo(const TeqTab &tb, int row, double T, const CaseA &fc,
                            rb_kchem_t &kc, rb_kcool_t &kl) {
  const double *r;
  if (teq_hit(tb, row, T, r)) { teq_load_chem(r, kc); teq_load_cool(r, kl); return; }
  get_kchem_kcool(T, fc.H2, fc.He2, fc.He3, kc, kl, fc.libm);
}
RB_HD rb_kchem_t kchem_memo(const TeqTab &tb, int row, double T, const CaseA &fc) {
  const double *r;
  rb_kchem_t kc;
  if (teq_hit(tb, row, T, r)) { teq_load_chem(r, kc); return kc; }
  return get_kchem(T, fc.H2, fc.He2, fc.He3, fc.libm);
}
RB_HD rb_kcool_t kcool_memo(const TeqTab &tb, int row, double T, const CaseA &fc) {
  const double *r;
  rb_kcool_t kl;
  if (teq_hit(tb, row, T, r)) { teq_load_cool(r, kl); return kl; }
  return get_kcool(T, fc.H2, fc.He2, fc.He3, fc.libm);
}
// the bisection's first temperature (ion_solver.f90:814-815); one function, so that the
// solvers and the table builder hold the same bits
RB_HDN double teq_root_T() {
  double T = (log10(RB_TFLOOR) + log10(RB_TMAX)) * 0.5;
  T = pow(10.0, T);
  return T;
}

RB_HDN int get_dudT(double nH, double nHe, double T, const PhotoRates &p,
                   double z, double Hz, const CaseA &fc, double tol, double &dudT,
                   const TeqTab &tb, int row) {
  rb_kchem_t kc;
  rb_kcool_t kl;
  kchem_kcool_memo(tb, row, T, fc, kc, kl);
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
                   double Hz, const CaseA &fc, double tol, double &dudT, rb_frac_t &x,
                   const TeqTab &tb, int row) {
  rb_kchem_t kc;
  rb_kcool_t kl;
  kchem_kcool_memo(tb, row, T, fc, kc, kl);
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
                     double &Tout, const TeqTab &tb) {
  double Tleft = RB_TFLOOR, Tright = RB_TMAX;
  double dudT, dudTmin, dudTmax;
  int rc = get_dudT(nH, nHe, Tleft, p, z, Hz, fc, tol, dudTmin, tb, 0);
  if (rc) return rc;
  if (dudTmin < 0.0) {  // :782-794: equilibrium below the floor
    const rb_kchem_t kc = kchem_memo(tb, 0, Tleft, fc);
    rc = solve_pce(nH, nHe, kc, p.H1i, p.He1i, p.He2i, tol, x);
    Tout = Tleft;
    return rc;
  }
  rc = get_dudT(nH, nHe, Tright, p, z, Hz, fc, tol, dudTmax, tb, teq_rows(tb));
  if (rc) return rc;
  if (dudTmin < 0.0 || dudTmax > 0.0) return ION_ERR_T_NOT_BRACKETED;

  double T = teq_root_T();   // :814-815
  int node = tb.rows ? 1 : -1;   // heap index of T in the bisection tree (memo row)
  double err = 1.7976931348623157e308;
  int iter = 0;
  while (err > tol) {
    rc = get_dudT(nH, nHe, T, p, z, Hz, fc, tol, dudT, tb, node);
    if (rc) return rc;
    const double Told = T;
    if (dudT < 0.0) {
      Tright = T;
      T = (Tleft + T) * 0.5;
    } else {
      Tleft = T;
      T = (T + Tright) * 0.5;
    }
    node = teq_child(tb, node, dudT < 0.0);
    err = fabs(1.0 - T / Told);
    iter = iter + 1;
    if (iter > RB_ION_MAX_ITER) return ION_ERR_MAX_ITER_PCTE;
  }
  const rb_kchem_t kc = kchem_memo(tb, node, T, fc);
  rc = solve_pce(nH, nHe, kc, p.H1i, p.He1i, p.He2i, tol, x);
  if (rc) return rc;
  zero_absent_species(nH, nHe, x);
  Tout = T;
  return ION_OK;
}

}  // namespace rb
