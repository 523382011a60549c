// rb_atomic.cuh -- atomic data evaluated on host and device:
//   Verner et al. 1996 photo-ionisation cross sections
//     (reference: rabacus/f2py/verner_96.f90:108-154, rows photo.dat:1-3)
//   Hui & Gnedin 1997 chemistry / cooling fits
//     (reference: rabacus/f2py/hui_gnedin_97.f90:21-322)
//   case A/B mixing and heating / cooling functions
//     (reference: rabacus/f2py/chem_cool_rates.f90:45-378)
//   E1 / E2 exponential-integral fits
//     (reference: rabacus/f2py/zhang_jin_f.f90:15-173)
#pragma once
#include "rb_common.cuh"
#include "rb_fastpow.cuh"

namespace rb {

// ---- Verner 96 ------------------------------------------------------------
struct V96Row {
  double Eth, E0, sigma0, ya, P, yw, y0, y1;
};

RB_HD V96Row v96_row(int species) {
  // species 0: H I (Z=1,N=1); 1: He I (Z=2,N=2); 2: He II (Z=2,N=1)
  switch (species) {
    case 0:
      return {13.60, 0.4298, 5.475e4, 32.88, 2.963, 0.0, 0.0, 0.0};
    case 1:
      return {24.59, 13.61, 949.2, 1.469, 3.188, 2.039, 0.4434, 2.136};
    default:
      return {54.42, 1.720, 1.369e4, 32.88, 2.963, 0.0, 0.0, 0.0};
  }
}

RB_HD double v96_sigma(int species, double E) {
  const V96Row r = v96_row(species);
  if (E < r.Eth) return 0.0;  // verner_96.f90:149-151
  const double x = E / r.E0 - r.y0;
  const double y = sqrt(x * x + r.y1 * r.y1);
  const double t1 = (x - 1.0) * (x - 1.0) + r.yw * r.yw;
  const double t2 = pow(y, 0.5 * r.P - 5.5);
  const double t3 = pow(1.0 + sqrt(y / r.ya), -r.P);
  return (r.sigma0 * 1.0e-18) * (t1 * t2 * t3);
}

// ---- Hui & Gnedin 97 --------------------------------------------------------
#define RB_T_H1 1.57807e5
#define RB_T_He1 2.85335e5
#define RB_T_He2 6.31515e5

// t1 * lam^t2 / (1 + (lam/t3)^t4)^t5
RB_HDN double hg_ferland(double lam, double t1, double t2, double t3, double t4,
                        double t5) {
  return t1 * pow(lam, t2) / pow(1.0 + pow(lam / t3, t4), t5);
}

// t1 * T^-1.5 * exp(-lam/2) * lam^t2 / (1 + (lam/t3)^t4)^t5   (Lotz)
RB_HDN double hg_lotz(double T, double lam, double t1, double t2, double t3,
                     double t4, double t5) {
  return t1 * pow(T, -1.5) * exp(-lam * 0.5) * pow(lam, t2) /
         pow(1.0 + pow(lam / t3, t4), t5);
}

RB_HD double hg_reHe2di(double T, double lamHe2) {
  return 1.90e-3 * pow(T, -1.5) * exp(-0.75 * lamHe2 * 0.5) *
         (1.0 + 0.3 * exp(-0.15 * lamHe2 * 0.5));
}

// 10^(log10(a) fcA + log10(b) (1-fcA))   chem_cool_rates.f90:67
RB_HD double mix_caseAB(double a, double b, double fcA) {
  return pow(10.0, log10(a) * fcA + log10(b) * (1.0 - fcA));
}

// chem_cool_rates.f90:45-88 (get_kchem_s), every power through libm / libdevice pow():
// the statement-by-statement form, used outside the validated range of rb_fastpow.cuh
RB_HDN rb_kchem_t get_kchem_libm(double T, double fcA_H2, double fcA_He2,
                                double fcA_He3) {
  const double lH1 = 2.0 * RB_T_H1 / T;
  const double lHe1 = 2.0 * RB_T_He1 / T;
  const double lHe2 = 2.0 * RB_T_He2 / T;
  rb_kchem_t k;
  k.reH2 = mix_caseAB(hg_ferland(lH1, 1.269e-13, 1.503, 0.522, 0.470, 1.923),
                      hg_ferland(lH1, 2.753e-14, 1.500, 2.740, 0.407, 2.242),
                      fcA_H2);
  k.reHe2 = mix_caseAB(3.0e-14 * pow(lHe1, 0.654), 1.26e-14 * pow(lHe1, 0.750),
                       fcA_He2);
  k.reHe2 = k.reHe2 + hg_reHe2di(T, lHe2);
  k.reHe3 =
      mix_caseAB(hg_ferland(lHe2, 2.0 * 1.269e-13, 1.503, 0.522, 0.470, 1.923),
                 hg_ferland(lHe2, 2.0 * 2.753e-14, 1.500, 2.740, 0.407, 2.242),
                 fcA_He3);
  k.ciH1 = hg_lotz(T, lH1, 21.11, -1.089, 0.354, 0.874, 1.101);
  k.ciHe1 = hg_lotz(T, lHe1, 32.38, -1.146, 0.416, 0.987, 1.056);
  k.ciHe2 = hg_lotz(T, lHe2, 19.95, -1.089, 0.553, 0.735, 1.275);
  return k;
}

// chem_cool_rates.f90:147-204 (get_kcool_s), libm form
RB_HDN rb_kcool_t get_kcool_libm(double T, double fcA_H2, double fcA_He2,
                                double fcA_He3) {
  const double lH1 = 2.0 * RB_T_H1 / T;
  const double lHe1 = 2.0 * RB_T_He1 / T;
  const double lHe2 = 2.0 * RB_T_He2 / T;
  rb_kcool_t k;
  // hui_gnedin_97.f90:162-187: t1 * T * lam^t2 / (...)^t5
  const double recH2a =
      1.778e-29 * T * pow(lH1, 1.965) / pow(1.0 + pow(lH1 / 0.541, 0.502), 2.697);
  const double recH2b =
      3.435e-30 * T * pow(lH1, 1.970) / pow(1.0 + pow(lH1 / 2.250, 0.376), 3.720);
  k.recH2 = mix_caseAB(recH2a, recH2b, fcA_H2);
  // :193-215
  const double recHe2a = RB_KB_ERG_K * T * (3.0e-14 * pow(lHe1, 0.654));
  const double recHe2b = RB_KB_ERG_K * T * (1.26e-14 * pow(lHe1, 0.750));
  k.recHe2 = mix_caseAB(recHe2a, recHe2b, fcA_He2);
  k.recHe2 = k.recHe2 + 0.75 * RB_KB_ERG_K * RB_T_He2 * hg_reHe2di(T, lHe2);
  // :220-245
  const double recHe3a = (8.0 * 1.778e-29) * T * pow(lHe2, 1.965) /
                         pow(1.0 + pow(lHe2 / 0.541, 0.502), 2.697);
  const double recHe3b = (8.0 * 3.435e-30) * T * pow(lHe2, 1.970) /
                         pow(1.0 + pow(lHe2 / 2.250, 0.376), 3.720);
  k.recHe3 = mix_caseAB(recHe3a, recHe3b, fcA_He3);
  // :252-277
  k.cicH1 = RB_KB_ERG_K * RB_T_H1 *
            hg_lotz(T, lH1, 21.11, -1.089, 0.354, 0.874, 1.101);
  k.cicHe1 = RB_KB_ERG_K * RB_T_He1 *
             hg_lotz(T, lHe1, 32.38, -1.146, 0.416, 0.987, 1.056);
  k.cicHe2 = RB_KB_ERG_K * RB_T_He2 *
             hg_lotz(T, lHe2, 19.95, -1.089, 0.553, 0.735, 1.275);
  // :283-301
  const double den = 1.0 + sqrt(T * 1.0e-5);
  k.cecH1 = 7.5e-19 * exp(-0.75 * lH1 * 0.5) / den;
  k.cecHe2 = 5.54e-17 * pow(1.0 / T, 0.397) * exp(-0.75 * lHe2 * 0.5) / den;
  // :307-322
  k.compton = 5.65e-36;
  const double d = 5.5 - log10(T);
  k.bremss = 1.43e-27 * sqrt(T) * (1.1 + 0.34 * exp(-(d * d) / 3.0));
  return k;
}

// ---------------------------------------------------------------------------
// The same fits with the powers of rb_fastpow.cuh (every statement keeps the reference's
// operation order; only x**y is evaluated by exp_dd(y log_dd(x)), < 0.53 ulp) and with the
// sub-expressions that get_kchem_s and get_kcool_s both evaluate computed once:
// T**-1.5, the three Lotz collisional-ionisation fits (ci* and cic* differ by kB T_X,
// hui_gnedin_97.f90:252-277), the dielectronic term, the HeII case A/B fits, the
// exp(-0.75 lam/2) factors.  One temperature probe of solve_pcte: 43 powers from 27
// logarithms instead of 61 pow() calls.
// ---------------------------------------------------------------------------
struct HgBase {  // lam = 2 T_X / T and its logarithm
  double lam;
  DD L;
};
RB_HD HgBase hg_base(double lam) {
  HgBase b;
  b.lam = lam;
  b.L = log_dd(lam);
  return b;
}
RB_HD double hg_pw(const HgBase &b, double y) { return pow_from_log_nf(b.L, y); }
// (1 + (lam/t3)^t4)^t5
RB_HD double hg_den(const HgBase &b, double t3, double t4, double t5) {
  return pow_fast_nf(1.0 + pow_fast_nf(b.lam / t3, t4), t5);
}
RB_HD double hg_ferland_f(const HgBase &b, double t1, double t2, double t3, double t4, double t5) {
  return t1 * hg_pw(b, t2) / hg_den(b, t3, t4, t5);
}
RB_HD double mix_caseAB_f(double a, double b, double fcA) {
  return pow10_fast_nf(log10_call(a) * fcA + log10_call(b) * (1.0 - fcA));
}

// valid range of the fast evaluation (all bases normal, all |y log x| < 700); `libm` != 0
// forces the libm statement form everywhere (a per-plan flag for A/B tests, rb_internal.h)
RB_HD bool hg_fast_ok(double T, int libm) { return !libm && T >= 1.0 && T <= 1.0e12; }

template <bool CHEM, bool COOL>
RB_HD void hg97_eval(double T, double fcA_H2, double fcA_He2, double fcA_He3, rb_kchem_t &kc,
                     rb_kcool_t &kl) {
  const HgBase H1 = hg_base(2.0 * RB_T_H1 / T);
  const HgBase He1 = hg_base(2.0 * RB_T_He1 / T);
  const HgBase He2 = hg_base(2.0 * RB_T_He2 / T);
  const double Tm15 = pow_fast_nf(T, -1.5);
  // Lotz fits, hui_gnedin_97.f90:101-155 (rates) = :252-277 (cooling) / (kB T_X)
  const double lotzH1 = 21.11 * Tm15 * exp_call(-H1.lam * 0.5) * hg_pw(H1, -1.089) /
                        hg_den(H1, 0.354, 0.874, 1.101);
  const double lotzHe1 = 32.38 * Tm15 * exp_call(-He1.lam * 0.5) * hg_pw(He1, -1.146) /
                         hg_den(He1, 0.416, 0.987, 1.056);
  const double lotzHe2 = 19.95 * Tm15 * exp_call(-He2.lam * 0.5) * hg_pw(He2, -1.089) /
                         hg_den(He2, 0.553, 0.735, 1.275);
  const double eHe2 = exp_call(-0.75 * He2.lam * 0.5);
  const double di = 1.90e-3 * Tm15 * eHe2 * (1.0 + 0.3 * exp_call(-0.15 * He2.lam * 0.5));
  const double he2a = 3.0e-14 * hg_pw(He1, 0.654);
  const double he2b = 1.26e-14 * hg_pw(He1, 0.750);
  if (CHEM) {  // chem_cool_rates.f90:45-88
    kc.reH2 = mix_caseAB_f(hg_ferland_f(H1, 1.269e-13, 1.503, 0.522, 0.470, 1.923),
                           hg_ferland_f(H1, 2.753e-14, 1.500, 2.740, 0.407, 2.242), fcA_H2);
    kc.reHe2 = mix_caseAB_f(he2a, he2b, fcA_He2);
    kc.reHe2 = kc.reHe2 + di;
    kc.reHe3 = mix_caseAB_f(hg_ferland_f(He2, 2.0 * 1.269e-13, 1.503, 0.522, 0.470, 1.923),
                            hg_ferland_f(He2, 2.0 * 2.753e-14, 1.500, 2.740, 0.407, 2.242),
                            fcA_He3);
    kc.ciH1 = lotzH1;
    kc.ciHe1 = lotzHe1;
    kc.ciHe2 = lotzHe2;
  }
  if (COOL) {  // chem_cool_rates.f90:147-204
    const double recH2a = 1.778e-29 * T * hg_pw(H1, 1.965) / hg_den(H1, 0.541, 0.502, 2.697);
    const double recH2b = 3.435e-30 * T * hg_pw(H1, 1.970) / hg_den(H1, 2.250, 0.376, 3.720);
    kl.recH2 = mix_caseAB_f(recH2a, recH2b, fcA_H2);
    const double recHe2a = RB_KB_ERG_K * T * he2a;
    const double recHe2b = RB_KB_ERG_K * T * he2b;
    kl.recHe2 = mix_caseAB_f(recHe2a, recHe2b, fcA_He2);
    kl.recHe2 = kl.recHe2 + 0.75 * RB_KB_ERG_K * RB_T_He2 * di;
    const double recHe3a =
        (8.0 * 1.778e-29) * T * hg_pw(He2, 1.965) / hg_den(He2, 0.541, 0.502, 2.697);
    const double recHe3b =
        (8.0 * 3.435e-30) * T * hg_pw(He2, 1.970) / hg_den(He2, 2.250, 0.376, 3.720);
    kl.recHe3 = mix_caseAB_f(recHe3a, recHe3b, fcA_He3);
    kl.cicH1 = RB_KB_ERG_K * RB_T_H1 * lotzH1;
    kl.cicHe1 = RB_KB_ERG_K * RB_T_He1 * lotzHe1;
    kl.cicHe2 = RB_KB_ERG_K * RB_T_He2 * lotzHe2;
    const double den = 1.0 + sqrt(T * 1.0e-5);
    kl.cecH1 = 7.5e-19 * exp_call(-0.75 * H1.lam * 0.5) / den;
    kl.cecHe2 = 5.54e-17 * pow_fast_nf(1.0 / T, 0.397) * eHe2 / den;
    kl.compton = 5.65e-36;
    const double d = 5.5 - log10_call(T);
    kl.bremss = 1.43e-27 * sqrt(T) * (1.1 + 0.34 * exp_call(-(d * d) / 3.0));
  }
}

RB_HDN rb_kchem_t get_kchem(double T, double fcA_H2, double fcA_He2, double fcA_He3,
                            int libm = 0) {
  if (!hg_fast_ok(T, libm)) return get_kchem_libm(T, fcA_H2, fcA_He2, fcA_He3);
  rb_kchem_t kc;
  rb_kcool_t kl;
  hg97_eval<true, false>(T, fcA_H2, fcA_He2, fcA_He3, kc, kl);
  return kc;
}
RB_HDN rb_kcool_t get_kcool(double T, double fcA_H2, double fcA_He2, double fcA_He3,
                            int libm = 0) {
  if (!hg_fast_ok(T, libm)) return get_kcool_libm(T, fcA_H2, fcA_He2, fcA_He3);
  rb_kchem_t kc;
  rb_kcool_t kl;
  hg97_eval<false, true>(T, fcA_H2, fcA_He2, fcA_He3, kc, kl);
  return kl;
}
// both sets at one temperature (one probe of solve_pcte, ion_solver.f90:1089-1134)
RB_HDN void get_kchem_kcool(double T, double fcA_H2, double fcA_He2, double fcA_He3,
                            rb_kchem_t &kc, rb_kcool_t &kl, int libm = 0) {
  if (!hg_fast_ok(T, libm)) {
    kc = get_kchem_libm(T, fcA_H2, fcA_He2, fcA_He3);
    kl = get_kcool_libm(T, fcA_H2, fcA_He2, fcA_He3);
    return;
  }
  hg97_eval<true, true>(T, fcA_H2, fcA_He2, fcA_He3, kc, kl);
}

// chem_cool_rates.f90:268-284
RB_HD double get_heat(double nH, double nHe, double H1h, double He1h,
                      double He2h, double xH1, double xHe1, double xHe2) {
  return xH1 * nH * H1h + xHe1 * nHe * He1h + xHe2 * nHe * He2h;
}

// chem_cool_rates.f90:289-332
RB_HD double get_cool(double nH, double nHe, double T, const rb_kcool_t &k,
                      const rb_frac_t &x, double z, double Hz) {
  const double ne = x.xH2 * nH + (x.xHe2 + 2.0 * x.xHe3) * nHe;
  double cool = k.recH2 * ne * x.xH2 * nH + k.recHe2 * ne * x.xHe2 * nHe +
                k.recHe3 * ne * x.xHe3 * nHe + k.cicH1 * ne * x.xH1 * nH +
                k.cicHe1 * ne * x.xHe1 * nHe + k.cicHe2 * ne * x.xHe2 * nHe +
                k.cecH1 * ne * x.xH1 * nH + k.cecHe2 * ne * x.xHe2 * nHe;
  const double Tcmb = 2.725 * (1.0 + z);
  const double zp1 = 1.0 + z;
  const double zp1_2 = zp1 * zp1;
  cool = cool + k.compton * (zp1_2 * zp1_2) * (T - Tcmb) * ne;
  cool = cool +
         k.bremss * ne * (nH * x.xH2 + nHe * x.xHe2 + 4.0 * nHe * x.xHe3);
  cool = cool + 3.0 * Hz * RB_KB_ERG_K * T * (nH + nHe + ne);
  return cool;
}

// ---- E1 / E2 low-order fits (these exact coefficients) ----------
RB_HD double e1xa(double x) {
  if (x == 0.0) return 1.0e300;
  if (x <= 1.0) {
    return -log(x) +
           ((((1.07857e-03 * x - 9.76004e-03) * x + 5.519968e-02) * x -
             0.24991055) * x + 0.99999193) * x -
           0.57721566;
  }
  const double es1 =
      (((x + 8.5733287401) * x + 18.059016973) * x + 8.6347608925) * x +
      0.2677737343;
  const double es2 =
      (((x + 9.5733223454) * x + 25.6329561486) * x + 21.0996530827) * x +
      3.9584969228;
  return exp(-x) / x * es1 / es2;
}

RB_HD double e2xa(double x) { return exp(-x) - x * e1xa(x); }

}  // namespace rb
