/*
 * rb_oracle_atomic.c -- oracle (TEST INFRASTRUCTURE ONLY, see rb_oracle.h):
 * physical constants, Verner+96 cross sections, Hui & Gnedin 97 rate fits,
 * chemistry/cooling rate sets, E1/E2 fits, trapezoid rule, Gauss-Legendre rule.
 *
 * Expression order (left-to-right for * and /, no FMA contraction) follows the
 * Fortran so that results are what `gfortran -O3` (no -ffast-math) computes.
 */
#include "rb_oracle_internal.h"

/* ------------------------------------------------------------------------
 * Verner et al. 1996 analytic fits.
 * Rows: rabacus/atomic/verner/photox/photo.dat:1-3
 * (columns Z N Eth Emax E0 sigma0 ya P yw y0 y1).
 * sigma0 is scaled by 1e-18 when the table is read: verner_96.f90:47.
 * ---------------------------------------------------------------------- */
typedef struct {
  double Eth, E0, sigma0, ya, P, yw, y0, y1;
} v96_row;

static const v96_row V96[3] = {
    /* H1  */ {1.360E+01, 4.298E-01, 5.475E+04, 3.288E+01, 2.963E+00, 0.000E+00,
               0.000E+00, 0.000E+00},
    /* He1 */ {2.459E+01, 1.361E+01, 9.492E+02, 1.469E+00, 3.188E+00, 2.039E+00,
               4.434E-01, 2.136E+00},
    /* He2 */ {5.442E+01, 1.720E+00, 1.369E+04, 3.288E+01, 2.963E+00, 0.000E+00,
               0.000E+00, 0.000E+00},
};

/* verner_96.f90:61-95 (v96_return_Eth), photo_xsections.f90:29-97 */
double orc_v96_Eth(int species) { return V96[species].Eth; }

/* verner_96.f90:108-154 (v96_return_sigma) */
void orc_v96_sigma(int species, const double *E, int nn, double *sigma) {
  const v96_row *r = &V96[species];
  const double sigma0 = r->sigma0 * 1.0e-18; /* verner_96.f90:47 */
  for (int i = 0; i < nn; i++) {
    double x = E[i] / r->E0 - r->y0;                    /* :139 */
    double y = sqrt(x * x + r->y1 * r->y1);             /* :140 */
    double t1 = (x - 1.0) * (x - 1.0) + r->yw * r->yw;  /* :141 */
    double t2 = pow(y, 0.5 * r->P - 5.5);               /* :142 */
    double t3 = pow(1.0 + sqrt(y / r->ya), -r->P);      /* :143 */
    double F = t1 * t2 * t3;                            /* :144 */
    sigma[i] = sigma0 * F;                              /* :145 */
    if (E[i] < r->Eth) sigma[i] = 0.0;                  /* :149-151 */
  }
}

/* ------------------------------------------------------------------------
 * Hui & Gnedin 1997 fits: hui_gnedin_97.f90
 * ---------------------------------------------------------------------- */
static const double T_H1 = 1.57807e5;  /* hui_gnedin_97.f90:6 */
static const double T_He1 = 2.85335e5; /* :7 */
static const double T_He2 = 6.31515e5; /* :8 */

/* the common shape  t1 * lambda**t2 / (1 + (lambda/t3)**t4)**t5  (:31,:44,...) */
static double hg97_G(double lambda, double t1, double t2, double t3, double t4,
                     double t5) {
  return t1 * pow(lambda, t2) / pow(1.0 + pow(lambda / t3, t4), t5);
}

/* hui_gnedin_97.f90:21-32 */
static double hg97_reH2a(double T) {
  double lambda = 2.0 * T_H1 / T;
  return hg97_G(lambda, 1.269e-13, 1.503, 0.522, 0.470, 1.923);
}
/* :34-45 */
static double hg97_reH2b(double T) {
  double lambda = 2.0 * T_H1 / T;
  return hg97_G(lambda, 2.753e-14, 1.500, 2.740, 0.407, 2.242);
}
/* :49-55 */
static double hg97_reHe2a(double T) {
  double lambda = 2.0 * T_He1 / T;
  return 3.0e-14 * pow(lambda, 0.654);
}
/* :57-63 */
static double hg97_reHe2b(double T) {
  double lambda = 2.0 * T_He1 / T;
  return 1.26e-14 * pow(lambda, 0.750);
}
/* :67-74 */
static double hg97_reHe2di(double T) {
  double lambda = 2.0 * T_He2 / T;
  return 1.90e-3 * pow(T, -1.5) * exp(-0.75 * lambda * 0.5) *
         (1.0 + 0.3 * exp(-0.15 * lambda * 0.5));
}
/* :78-89 ; t1 = 2.0d0 * 1.269d-13 is a compile-time product */
static double hg97_reHe3a(double T) {
  double lambda = 2.0 * T_He2 / T;
  return hg97_G(lambda, 2.0 * 1.269e-13, 1.503, 0.522, 0.470, 1.923);
}
/* :91-102 */
static double hg97_reHe3b(double T) {
  double lambda = 2.0 * T_He2 / T;
  return hg97_G(lambda, 2.0 * 2.753e-14, 1.500, 2.740, 0.407, 2.242);
}
/* collisional ionisation  t1*T**-1.5*exp(-lambda/2)*lambda**t2/(1+(lambda/t3)**t4)**t5
 * :106-150 */
static double hg97_ci(double T, double Tth, double t1, double t2, double t3,
                      double t4, double t5) {
  double lambda = 2.0 * Tth / T;
  return t1 * pow(T, -1.5) * exp(-lambda * 0.5) * pow(lambda, t2) /
         pow(1.0 + pow(lambda / t3, t4), t5);
}
static double hg97_ciH1(double T) {
  return hg97_ci(T, T_H1, 21.11, -1.089, 0.354, 0.874, 1.101);
}
static double hg97_ciHe1(double T) {
  return hg97_ci(T, T_He1, 32.38, -1.146, 0.416, 0.987, 1.056);
}
static double hg97_ciHe2(double T) {
  return hg97_ci(T, T_He2, 19.95, -1.089, 0.553, 0.735, 1.275);
}
/* recombination cooling  t1 * T * lambda**t2 / (...)**t5   :162-187, :220-245 */
static double hg97_recG(double T, double Tth, double t1, double t2, double t3,
                        double t4, double t5) {
  double lambda = 2.0 * Tth / T;
  return t1 * T * pow(lambda, t2) / pow(1.0 + pow(lambda / t3, t4), t5);
}
static double hg97_recH2a(double T) {
  return hg97_recG(T, T_H1, 1.778e-29, 1.965, 0.541, 0.502, 2.697);
}
static double hg97_recH2b(double T) {
  return hg97_recG(T, T_H1, 3.435e-30, 1.970, 2.250, 0.376, 3.720);
}
/* :193-206 */
static double hg97_recHe2a(double T) { return KB_ERG_K * T * hg97_reHe2a(T); }
static double hg97_recHe2b(double T) { return KB_ERG_K * T * hg97_reHe2b(T); }
/* :211-215 */
static double hg97_recHe2di(double T) {
  return 0.75 * KB_ERG_K * T_He2 * hg97_reHe2di(T);
}
static double hg97_recHe3a(double T) {
  return hg97_recG(T, T_He2, 8.0 * 1.778e-29, 1.965, 0.541, 0.502, 2.697);
}
static double hg97_recHe3b(double T) {
  return hg97_recG(T, T_He2, 8.0 * 3.435e-30, 1.970, 2.250, 0.376, 3.720);
}
/* :252-277 */
static double hg97_cicH1(double T) { return KB_ERG_K * T_H1 * hg97_ciH1(T); }
static double hg97_cicHe1(double T) { return KB_ERG_K * T_He1 * hg97_ciHe1(T); }
static double hg97_cicHe2(double T) { return KB_ERG_K * T_He2 * hg97_ciHe2(T); }
/* :283-288 */
static double hg97_cecH1(double T) {
  double lambda = 2.0 * T_H1 / T;
  return 7.5e-19 * exp(-0.75 * lambda * 0.5) / (1.0 + sqrt(T * 1.0e-5));
}
/* :293-301 */
static double hg97_cecHe2(double T) {
  double lambda = 2.0 * T_He2 / T;
  return 5.54e-17 * pow(1.0 / T, 0.397) * exp(-0.75 * lambda * 0.5) /
         (1.0 + sqrt(T * 1.0e-5));
}
/* :307-311 */
static double hg97_compton(double T) {
  (void)T;
  return 5.65e-36;
}
/* :316-322 ; (..)**2.0d0 is folded to a square by GCC */
static double hg97_bremss(double T) {
  double d = 5.5 - log10(T);
  double gff = 1.1 + 0.34 * exp(-(d * d) / 3.0);
  return 1.43e-27 * sqrt(T) * gff;
}

void orc_hg97_all(double T, double *o) {
  o[0] = hg97_reH2a(T);   o[1] = hg97_reH2b(T);
  o[2] = hg97_reHe2a(T);  o[3] = hg97_reHe2b(T);  o[4] = hg97_reHe2di(T);
  o[5] = hg97_reHe3a(T);  o[6] = hg97_reHe3b(T);
  o[7] = hg97_ciH1(T);    o[8] = hg97_ciHe1(T);   o[9] = hg97_ciHe2(T);
  o[10] = hg97_recH2a(T); o[11] = hg97_recH2b(T);
  o[12] = hg97_recHe2a(T); o[13] = hg97_recHe2b(T); o[14] = hg97_recHe2di(T);
  o[15] = hg97_recHe3a(T); o[16] = hg97_recHe3b(T);
  o[17] = hg97_cicH1(T);  o[18] = hg97_cicHe1(T); o[19] = hg97_cicHe2(T);
  o[20] = hg97_cecH1(T);  o[21] = hg97_cecHe2(T); o[22] = hg97_bremss(T);
}

/* case A/B log-interpolation  ten**( log10(a)*fcA + log10(b)*(one-fcA) )
 * chem_cool_rates.f90:67,71,76 */
static double mixAB(double a, double b, double fcA) {
  return pow(10.0, log10(a) * fcA + log10(b) * (1.0 - fcA));
}

/* Coefficient hook (tests only): when set, the 6 + 10 rate / cooling coefficients of a
 * temperature come from the caller's function instead of the fits below.  The parity tests
 * install the host build of the CUDA library's own evaluation here, so that the SOLVERS of
 * this oracle can be fed exactly the coefficients the device used (tests/helpers.py).
 * out16 = reH2 reHe2 reHe3 ciH1 ciHe1 ciHe2 recH2 recHe2 recHe3 cicH1 cicHe1 cicHe2
 *         cecH1 cecHe2 bremss compton.  NULL restores the fits.  Not thread safe to change. */
static orc_coef_hook g_coef_hook = 0;
void orc_set_coefficient_hook(orc_coef_hook fn) { g_coef_hook = fn; }

/* chem_cool_rates.f90:45-88 (get_kchem_s); the _v twin :91-140 is the same
 * arithmetic element-wise */
void orc_kchem1(double T, double fcA_H2, double fcA_He2, double fcA_He3,
                double *reH2, double *reHe2, double *reHe3, double *ciH1,
                double *ciHe1, double *ciHe2) {
  if (g_coef_hook) {
    double o[16];
    g_coef_hook(T, fcA_H2, fcA_He2, fcA_He3, o);
    *reH2 = o[0]; *reHe2 = o[1]; *reHe3 = o[2]; *ciH1 = o[3]; *ciHe1 = o[4]; *ciHe2 = o[5];
    return;
  }
  *reH2 = mixAB(hg97_reH2a(T), hg97_reH2b(T), fcA_H2);
  *reHe2 = mixAB(hg97_reHe2a(T), hg97_reHe2b(T), fcA_He2);
  *reHe2 = *reHe2 + hg97_reHe2di(T);
  *reHe3 = mixAB(hg97_reHe3a(T), hg97_reHe3b(T), fcA_He3);
  *ciH1 = hg97_ciH1(T);
  *ciHe1 = hg97_ciHe1(T);
  *ciHe2 = hg97_ciHe2(T);
}

/* chem_cool_rates.f90:147-204 (get_kcool_s) */
void orc_kcool1(double T, double fcA_H2, double fcA_He2, double fcA_He3,
                orc_kcool *k) {
  if (g_coef_hook) {
    double o[16];
    g_coef_hook(T, fcA_H2, fcA_He2, fcA_He3, o);
    k->recH2 = o[6]; k->recHe2 = o[7]; k->recHe3 = o[8]; k->cicH1 = o[9]; k->cicHe1 = o[10];
    k->cicHe2 = o[11]; k->cecH1 = o[12]; k->cecHe2 = o[13]; k->bremss = o[14];
    k->compton = o[15];
    return;
  }
  k->recH2 = mixAB(hg97_recH2a(T), hg97_recH2b(T), fcA_H2);
  k->recHe2 = mixAB(hg97_recHe2a(T), hg97_recHe2b(T), fcA_He2);
  k->recHe2 = k->recHe2 + hg97_recHe2di(T);
  k->recHe3 = mixAB(hg97_recHe3a(T), hg97_recHe3b(T), fcA_He3);
  k->cicH1 = hg97_cicH1(T);
  k->cicHe1 = hg97_cicHe1(T);
  k->cicHe2 = hg97_cicHe2(T);
  k->cecH1 = hg97_cecH1(T);
  k->cecHe2 = hg97_cecHe2(T);
  k->bremss = hg97_bremss(T);
  k->compton = hg97_compton(T);
}

void orc_get_kchem(const double *T, const double *fcA_H2, const double *fcA_He2,
                   const double *fcA_He3, int nn, double *reH2, double *reHe2,
                   double *reHe3, double *ciH1, double *ciHe1, double *ciHe2) {
  for (int i = 0; i < nn; i++)
    orc_kchem1(T[i], fcA_H2[i], fcA_He2[i], fcA_He3[i], &reH2[i], &reHe2[i],
               &reHe3[i], &ciH1[i], &ciHe1[i], &ciHe2[i]);
}

void orc_get_kcool(const double *T, const double *fcA_H2, const double *fcA_He2,
                   const double *fcA_He3, int nn, double *recH2, double *recHe2,
                   double *recHe3, double *cicH1, double *cicHe1,
                   double *cicHe2, double *cecH1, double *cecHe2,
                   double *bremss, double *compton) {
  for (int i = 0; i < nn; i++) {
    orc_kcool k;
    orc_kcool1(T[i], fcA_H2[i], fcA_He2[i], fcA_He3[i], &k);
    recH2[i] = k.recH2; recHe2[i] = k.recHe2; recHe3[i] = k.recHe3;
    cicH1[i] = k.cicH1; cicHe1[i] = k.cicHe1; cicHe2[i] = k.cicHe2;
    cecH1[i] = k.cecH1; cecHe2[i] = k.cecHe2;
    bremss[i] = k.bremss; compton[i] = k.compton;
  }
}

/* chem_cool_rates.f90:268-284 (get_heat_s) */
double orc_heat1(double nH, double nHe, double H1h, double He1h, double He2h,
                 double xH1, double xHe1, double xHe2) {
  return xH1 * nH * H1h + xHe1 * nHe * He1h + xHe2 * nHe * He2h;
}

/* chem_cool_rates.f90:289-332 (get_cool_s) */
double orc_cool1(double nH, double nHe, double T, const orc_kcool *k,
                 double xH1, double xH2, double xHe1, double xHe2, double xHe3,
                 double z, double Hz) {
  double ne = xH2 * nH + (xHe2 + 2.0 * xHe3) * nHe;
  double cool = k->recH2 * ne * xH2 * nH + k->recHe2 * ne * xHe2 * nHe +
                k->recHe3 * ne * xHe3 * nHe + k->cicH1 * ne * xH1 * nH +
                k->cicHe1 * ne * xHe1 * nHe + k->cicHe2 * ne * xHe2 * nHe +
                k->cecH1 * ne * xH1 * nH + k->cecHe2 * ne * xHe2 * nHe;
  double Tcmb = 2.725 * (1.0 + z);
  double zp1 = 1.0 + z;
  double zp1_2 = zp1 * zp1; /* (1+z)**4 expands to two squarings */
  cool = cool + k->compton * (zp1_2 * zp1_2) * (T - Tcmb) * ne;
  cool = cool + k->bremss * ne * (nH * xH2 + nHe * xHe2 + 4.0 * nHe * xHe3);
  cool = cool + 3.0 * Hz * KB_ERG_K * T * (nH + nHe + ne);
  return cool;
}

/* ------------------------------------------------------------------------
 * zhang_jin_f.f90:15-91 (e1xa), :94-173 (e2xa)
 * ---------------------------------------------------------------------- */
double orc_e1xa(double x) {
  double e1;
  if (x == 0.0) {
    e1 = 1.0e300;
  } else if (x <= 1.0) {
    e1 = -log(x) +
         ((((1.07857e-03 * x - 9.76004e-03) * x + 5.519968e-02) * x -
           0.24991055) * x + 0.99999193) * x -
         0.57721566;
  } else {
    double es1 = (((x + 8.5733287401) * x + 18.059016973) * x + 8.6347608925) * x +
                 0.2677737343;
    double es2 = (((x + 9.5733223454) * x + 25.6329561486) * x + 21.0996530827) * x +
                 3.9584969228;
    e1 = exp(-x) / x * es1 / es2;
  }
  return e1;
}

double orc_e2xa(double x) {
  double e1 = orc_e1xa(x);
  return exp(-x) - x * e1; /* zhang_jin_f.f90:171 */
}

/* utils.f90:11-21 (utils_integrate_trap): sequential sum */
double orc_integrate_trap(const double *xx, const double *yy, int nn) {
  double ss = 0.0;
  for (int i = 0; i < nn - 1; i++) {
    double dx = xx[i + 1] - xx[i];
    ss += (yy[i + 1] + yy[i]) * dx;
  }
  return 0.5 * ss;
}

/* ------------------------------------------------------------------------
 * Gauss-Legendre rule by the Golub-Welsch / implicit-QL method:
 * legendre_polynomial.f90:6-207 (imtqlx), :815-864 (p_quadrature_rule).
 * 1-based indexing kept (arrays offset by one) to follow the reference.
 * ---------------------------------------------------------------------- */
static int imtqlx(int n, double *d, double *e, double *z) {
  /* d,e,z are 1-based: valid indices 1..n */
  const int itn = 30;
  const double prec = 2.220446049250313e-16; /* epsilon(prec) */
  if (n == 1) return 0;
  e[n] = 0.0;
  for (int l = 1; l <= n; l++) {
    int j = 0;
    for (;;) {
      int m;
      for (m = l; m <= n; m++) {
        if (m == n) break;
        if (fabs(e[m]) <= prec * (fabs(d[m]) + fabs(d[m + 1]))) break;
      }
      double p = d[l];
      if (m == l) break;
      if (itn <= j) return -1;
      j = j + 1;
      double g = (d[l + 1] - p) / (2.0 * e[l]);
      double r = sqrt(g * g + 1.0);
      g = d[m] - p + e[l] / (g + copysign(r, g)); /* sign(r,g): |r| with sign of g */
      double s = 1.0, c = 1.0;
      p = 0.0;
      int mml = m - l;
      for (int ii = 1; ii <= mml; ii++) {
        int i = m - ii;
        double f = s * e[i];
        double b = c * e[i];
        if (fabs(g) <= fabs(f)) {
          c = g / f;
          r = sqrt(c * c + 1.0);
          e[i + 1] = f * r;
          s = 1.0 / r;
          c = c * s;
        } else {
          s = f / g;
          r = sqrt(s * s + 1.0);
          e[i + 1] = g * r;
          c = 1.0 / r;
          s = s * c;
        }
        g = d[i + 1] - p;
        r = (d[i] - g) * s + 2.0 * c * b;
        p = s * r;
        d[i + 1] = g + p;
        g = c * r - b;
        f = z[i + 1];
        z[i + 1] = s * z[i] + c * f;
        z[i] = c * z[i] - s * f;
      }
      d[l] = d[l] - p;
      e[l] = g;
      e[m] = 0.0;
    }
  }
  /* sorting: legendre_polynomial.f90:185-204 */
  for (int ii = 2; ii <= n; ii++) {
    int i = ii - 1;
    int k = i;
    double p = d[i];
    for (int j = ii; j <= n; j++) {
      if (d[j] < p) {
        k = j;
        p = d[j];
      }
    }
    if (k != i) {
      d[k] = d[i];
      d[i] = p;
      p = z[i];
      z[i] = z[k];
      z[k] = p;
    }
  }
  return 0;
}

void orc_p_quadrature_rule(int nt, double *t, double *wts) {
  double *bj = (double *)malloc(sizeof(double) * (size_t)(nt + 2));
  double *d = (double *)calloc((size_t)(nt + 2), sizeof(double));
  double *z = (double *)calloc((size_t)(nt + 2), sizeof(double));
  for (int i = 1; i <= nt; i++) {
    /* integer i*i as in the reference (kind=4) then converted */
    bj[i] = (double)(i * i) / (double)(4 * i * i - 1);
    bj[i] = sqrt(bj[i]);
  }
  z[1] = sqrt(2.0);
  imtqlx(nt, d, bj, z);
  for (int i = 1; i <= nt; i++) {
    t[i - 1] = d[i];
    wts[i - 1] = z[i] * z[i]; /* wts ** 2 */
  }
  free(bj);
  free(d);
  free(z);
}
