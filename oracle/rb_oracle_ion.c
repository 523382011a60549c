/*
 * rb_oracle_ion.c -- oracle (TEST INFRASTRUCTURE ONLY, see rb_oracle.h):
 * single-zone ionisation / temperature equilibrium solvers, restating
 * rabacus/f2py/ion_solver.f90.  Scalar (`_s`) and lock-step vector (`_v`)
 * twins are both kept because they are NOT equivalent.
 */
#include "rb_oracle_internal.h"

#define HUGE_ONE 1.7976931348623157e308 /* huge(one) */

/* gfortran MAX(): a NaN operand is ignored -> same as C fmax */
static double max3(double a, double b, double c) { return fmax(fmax(a, b), c); }

/* ion_solver.f90:68-80 (fpc_s) */
static void fpc1(double *xH1, double *xH2, double *xHe1, double *xHe2,
                 double *xHe3) {
  double total = *xH1 + *xH2;
  *xH1 = *xH1 / total;
  *xH2 = *xH2 / total;
  total = *xHe1 + *xHe2 + *xHe3;
  *xHe1 = *xHe1 / total;
  *xHe2 = *xHe2 / total;
  *xHe3 = *xHe3 / total;
}

/* ion_solver.f90:161-185 (analytic_H1_s) */
static double analytic_H1(double nH, double y, double reH2, double ciH1,
                          double H1i) {
  double RR = (ciH1 + reH2) * nH;
  double QQ = -(H1i + reH2 * nH + RR * (1.0 + y));
  double PP = reH2 * nH * (1.0 + y);
  double dd = QQ * QQ - 4.0 * RR * PP;
  if (dd < 0.0) dd = 0.0;
  double q = -0.5 * (QQ - sqrt(dd));
  return PP / q;
}

/* ion_solver.f90:238-276 (implicit_ionfracs_s) */
static void implicit_ionfracs(double ne, double reH2, double reHe2,
                              double reHe3, double ciH1, double ciHe1,
                              double ciHe2, double H1i, double He1i,
                              double He2i, double *xH1, double *xH2,
                              double *xHe1, double *xHe2, double *xHe3) {
  double R2 = reH2 * ne;
  double CG1 = ciH1 * ne + H1i;
  double DH = R2 + CG1;
  *xH1 = R2 / DH;
  *xH2 = CG1 / DH;

  R2 = reHe2 * ne;
  double R3 = reHe3 * ne;
  CG1 = ciHe1 * ne + He1i;
  double CG2 = ciHe2 * ne + He2i;
  double DHe = (R2 * R3) + (R3 * CG1) + (CG1 * CG2);
  *xHe1 = R2 * R3 / DHe;
  *xHe2 = R3 * CG1 / DHe;
  *xHe3 = CG1 * CG2 / DHe;
  fpc1(xH1, xH2, xHe1, xHe2, xHe3);
}

/* ion_solver.f90:339-374 (solve_ce_s) */
static int ce1(double reH2, double reHe2, double reHe3, double ciH1,
               double ciHe1, double ciHe2, double *xH1, double *xH2,
               double *xHe1, double *xHe2, double *xHe3) {
  double DD = reH2 + ciH1;
  *xH1 = reH2 / DD;
  *xH2 = ciH1 / DD;
  DD = reHe2 * reHe3 + reHe3 * ciHe1 + ciHe1 * ciHe2;
  *xHe1 = reHe2 * reHe3 / DD;
  *xHe2 = reHe3 * ciHe1 / DD;
  *xHe3 = ciHe1 * ciHe2 / DD;
  fpc1(xH1, xH2, xHe1, xHe2, xHe3);
  if (isnan(*xH1) || isnan(*xH2) || isnan(*xHe1) || isnan(*xHe2) ||
      isnan(*xHe3))
    return ORC_ERR_NAN; /* :369-372 `stop` */
  return ORC_OK;
}

int orc_solve_ce(const double *reH2, const double *reHe2, const double *reHe3,
                 const double *ciH1, const double *ciHe1, const double *ciHe2,
                 int nn, double *xH1, double *xH2, double *xHe1, double *xHe2,
                 double *xHe3) {
  for (int i = 0; i < nn; i++) {
    int rc = ce1(reH2[i], reHe2[i], reHe3[i], ciH1[i], ciHe1[i], ciHe2[i],
                 &xH1[i], &xH2[i], &xHe1[i], &xHe2[i], &xHe3[i]);
    if (rc) return rc;
  }
  return ORC_OK;
}

/* ion_solver.f90:440-562 (solve_pce_s) */
int orc_pce1(double nH, double nHe, double reH2, double reHe2, double reHe3,
             double ciH1, double ciHe1, double ciHe2, double H1i, double He1i,
             double He2i, double tol, double *pxH1, double *pxH2,
             double *pxHe1, double *pxHe2, double *pxHe3) {
  double xH1, xH2, xHe1, xHe2, xHe3;
  int rc = ce1(reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, &xH1, &xH2, &xHe1,
               &xHe2, &xHe3);
  if (rc) return rc;

  double ne_min = xH2 * nH + (xHe2 + 2.0 * xHe3) * nHe; /* :468 */
  double ne_max = nH + 2.0 * nHe;
  double ne_left = ne_min;
  double ne_right = ne_max;

  double y = (xHe2 + 2.0 * xHe3) * nHe / nH; /* :478 */
  xH1 = analytic_H1(nH, y, reH2, ciH1, H1i);
  xH2 = 1.0 - xH1;
  double ne = xH2 * nH + (xHe2 + 2.0 * xHe3) * nHe;

  double ne_old = ne;
  double xH1_old = xH1, xHe1_old = xHe1, xHe3_old = xHe3;

  int iter = 0;
  double err = HUGE_ONE;
  while (err > tol) { /* :495 */
    implicit_ionfracs(ne, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, H1i, He1i,
                      He2i, &xH1, &xH2, &xHe1, &xHe2, &xHe3);
    ne = xH2 * nH + (xHe2 + 2.0 * xHe3) * nHe;
    double ratio_ne = ne / ne_old;
    double ratio_x = max3(xH1 / xH1_old, xHe1 / xHe1_old, xHe3 / xHe3_old);
    err = fabs(ratio_x - 1.0);
    if (ratio_ne > 1.0) { /* :515-521 */
      ne = (ne_old + ne_right) * 0.5;
      ne_left = ne_old;
    } else {
      ne = (ne_left + ne_old) * 0.5;
      ne_right = ne_old;
    }
    ne_old = ne;
    xH1_old = xH1;
    xHe1_old = xHe1;
    xHe3_old = xHe3;
    iter = iter + 1;
    if (iter > ION_MAX_ITER) return ORC_ERR_MAX_ITER_PCE; /* :536-546 */
  }
  if (nH <= 0.0) { /* :550-553 */
    xH1 = 0.0;
    xH2 = 0.0;
  }
  if (nHe <= 0.0) {
    xHe1 = 0.0;
    xHe2 = 0.0;
    xHe3 = 0.0;
  }
  *pxH1 = xH1; *pxH2 = xH2; *pxHe1 = xHe1; *pxHe2 = xHe2; *pxHe3 = xHe3;
  return ORC_OK;
}

int orc_solve_pce_s(const double *nH, const double *nHe, const double *reH2,
                    const double *reHe2, const double *reHe3,
                    const double *ciH1, const double *ciHe1,
                    const double *ciHe2, const double *H1i, const double *He1i,
                    const double *He2i, double tol, int nn, double *xH1,
                    double *xH2, double *xHe1, double *xHe2, double *xHe3) {
  for (int i = 0; i < nn; i++) {
    int rc = orc_pce1(nH[i], nHe[i], reH2[i], reHe2[i], reHe3[i], ciH1[i],
                      ciHe1[i], ciHe2[i], H1i[i], He1i[i], He2i[i], tol,
                      &xH1[i], &xH2[i], &xHe1[i], &xHe2[i], &xHe3[i]);
    if (rc) return rc;
  }
  return ORC_OK;
}

/* ion_solver.f90:565-696 (solve_pce_v): all cells iterate in lock-step until
 * all have converged; ratio_ne == 1 leaves ne at the freshly computed value. */
int orc_solve_pce_v(const double *nH, const double *nHe, const double *reH2,
                    const double *reHe2, const double *reHe3,
                    const double *ciH1, const double *ciHe1,
                    const double *ciHe2, const double *H1i, const double *He1i,
                    const double *He2i, double tol, int nn, double *xH1,
                    double *xH2, double *xHe1, double *xHe2, double *xHe3) {
  size_t sz = sizeof(double) * (size_t)nn;
  double *buf = (double *)malloc(8 * sz);
  double *xH1_old = buf, *xHe1_old = buf + nn, *xHe3_old = buf + 2 * nn;
  double *ne_left = buf + 3 * nn, *ne_right = buf + 4 * nn;
  double *ne_old = buf + 5 * nn, *ne = buf + 6 * nn, *err = buf + 7 * nn;
  int rc = ORC_OK;

  for (int i = 0; i < nn; i++) {
    rc = ce1(reH2[i], reHe2[i], reHe3[i], ciH1[i], ciHe1[i], ciHe2[i], &xH1[i],
             &xH2[i], &xHe1[i], &xHe2[i], &xHe3[i]);
    if (rc) goto done;
  }
  for (int i = 0; i < nn; i++) {
    ne_left[i] = xH2[i] * nH[i] + (xHe2[i] + 2.0 * xHe3[i]) * nHe[i];
    ne_right[i] = nH[i] + 2.0 * nHe[i];
    double y = (xHe2[i] + 2.0 * xHe3[i]) * nHe[i] / nH[i];
    xH1[i] = analytic_H1(nH[i], y, reH2[i], ciH1[i], H1i[i]);
    xH2[i] = 1.0 - xH1[i];
    ne[i] = xH2[i] * nH[i] + (xHe2[i] + 2.0 * xHe3[i]) * nHe[i];
    ne_old[i] = ne[i];
    xH1_old[i] = xH1[i];
    xHe1_old[i] = xHe1[i];
    xHe3_old[i] = xHe3[i];
    err[i] = HUGE_ONE;
  }
  int iter = 0;
  for (;;) {
    int any = 0;
    for (int i = 0; i < nn; i++)
      if (err[i] > tol) any = 1;
    if (!any) break; /* :620 */
    for (int i = 0; i < nn; i++) {
      implicit_ionfracs(ne[i], reH2[i], reHe2[i], reHe3[i], ciH1[i], ciHe1[i],
                        ciHe2[i], H1i[i], He1i[i], He2i[i], &xH1[i], &xH2[i],
                        &xHe1[i], &xHe2[i], &xHe3[i]);
      ne[i] = xH2[i] * nH[i] + (xHe2[i] + 2.0 * xHe3[i]) * nHe[i];
      double ratio_ne = ne[i] / ne_old[i];
      double ratio_x = max3(xH1[i] / xH1_old[i], xHe1[i] / xHe1_old[i],
                            xHe3[i] / xHe3_old[i]);
      err[i] = fabs(ratio_x - 1.0);
      if (ratio_ne > 1.0) { /* :640-645 */
        ne[i] = (ne_old[i] + ne_right[i]) * 0.5;
        ne_left[i] = ne_old[i];
      }
      if (ratio_ne < 1.0) { /* :647-652 */
        ne[i] = (ne_left[i] + ne_old[i]) * 0.5;
        ne_right[i] = ne_old[i];
      }
      ne_old[i] = ne[i];
      xH1_old[i] = xH1[i];
      xHe1_old[i] = xHe1[i];
      xHe3_old[i] = xHe3[i];
    }
    iter = iter + 1;
    if (iter > ION_MAX_ITER) {
      rc = ORC_ERR_MAX_ITER_PCE;
      goto done;
    }
  }
  for (int i = 0; i < nn; i++) {
    if (nH[i] <= 0.0) {
      xH1[i] = 0.0;
      xH2[i] = 0.0;
    }
    if (nHe[i] <= 0.0) {
      xHe1[i] = 0.0;
      xHe2[i] = 0.0;
      xHe3[i] = 0.0;
    }
  }
done:
  free(buf);
  return rc;
}

/* ion_solver.f90:1089-1134 (get_dudT_s) */
static int dudT1(double nH, double nHe, double T, double H1i, double He1i,
                 double He2i, double H1h, double He1h, double He2h, double z,
                 double Hz, double fcA_H2, double fcA_He2, double fcA_He3,
                 double tol, double *dudT) {
  double reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2;
  double xH1, xH2, xHe1, xHe2, xHe3;
  orc_kcool k;
  orc_kchem1(T, fcA_H2, fcA_He2, fcA_He3, &reH2, &reHe2, &reHe3, &ciH1, &ciHe1,
             &ciHe2);
  orc_kcool1(T, fcA_H2, fcA_He2, fcA_He3, &k);
  int rc = orc_pce1(nH, nHe, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, H1i, He1i,
                    He2i, tol, &xH1, &xH2, &xHe1, &xHe2, &xHe3);
  if (rc) return rc;
  double heat = orc_heat1(nH, nHe, H1h, He1h, He2h, xH1, xHe1, xHe2);
  double cool = orc_cool1(nH, nHe, T, &k, xH1, xH2, xHe1, xHe2, xHe3, z, Hz);
  *dudT = heat - cool;
  return ORC_OK;
}

/* ion_solver.f90:729-880 (solve_pcte_s) */
int orc_pcte1(double nH, double nHe, double H1i, double He1i, double He2i,
              double H1h, double He1h, double He2h, double z, double Hz,
              double fcA_H2, double fcA_He2, double fcA_He3, double tol,
              double *xH1, double *xH2, double *xHe1, double *xHe2,
              double *xHe3, double *pT) {
  double Tleft = TFLOOR, Tright = TMAX, Told, T;
  double reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2;
  double dudT, dudTmin, dudTmax;
  int rc;

  rc = dudT1(nH, nHe, Tleft, H1i, He1i, He2i, H1h, He1h, He2h, z, Hz, fcA_H2,
             fcA_He2, fcA_He3, tol, &dudTmin);
  if (rc) return rc;
  if (dudTmin < 0.0) { /* :782-794 */
    orc_kchem1(Tleft, fcA_H2, fcA_He2, fcA_He3, &reH2, &reHe2, &reHe3, &ciH1,
               &ciHe1, &ciHe2);
    rc = orc_pce1(nH, nHe, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, H1i, He1i,
                  He2i, tol, xH1, xH2, xHe1, xHe2, xHe3);
    *pT = Tleft;
    return rc;
  }
  rc = dudT1(nH, nHe, Tright, H1i, He1i, He2i, H1h, He1h, He2h, z, Hz, fcA_H2,
             fcA_He2, fcA_He3, tol, &dudTmax);
  if (rc) return rc;
  if (dudTmin < 0.0 || dudTmax > 0.0) return ORC_ERR_T_NOT_BRACKETED; /* :804 */

  T = (log10(Tleft) + log10(Tright)) * 0.5; /* :814-815 */
  T = pow(10.0, T);

  double err = HUGE_ONE;
  int iter = 0;
  while (err > tol) { /* :822 */
    rc = dudT1(nH, nHe, T, H1i, He1i, He2i, H1h, He1h, He2h, z, Hz, fcA_H2,
               fcA_He2, fcA_He3, tol, &dudT);
    if (rc) return rc;
    Told = T;
    if (dudT < 0.0) {
      Tright = T;
      T = (Tleft + T) * 0.5;
    } else {
      Tleft = T;
      T = (T + Tright) * 0.5;
    }
    err = fabs(1.0 - T / Told);
    iter = iter + 1;
    if (iter > ION_MAX_ITER) return ORC_ERR_MAX_ITER_PCTE;
  }
  orc_kchem1(T, fcA_H2, fcA_He2, fcA_He3, &reH2, &reHe2, &reHe3, &ciH1, &ciHe1,
             &ciHe2);
  rc = orc_pce1(nH, nHe, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, H1i, He1i,
                He2i, tol, xH1, xH2, xHe1, xHe2, xHe3);
  if (rc) return rc;
  if (nH <= 0.0) {
    *xH1 = 0.0;
    *xH2 = 0.0;
  }
  if (nHe <= 0.0) {
    *xHe1 = 0.0;
    *xHe2 = 0.0;
    *xHe3 = 0.0;
  }
  *pT = T;
  return ORC_OK;
}

int orc_solve_pcte_s(const double *nH, const double *nHe, const double *H1i,
                     const double *He1i, const double *He2i, const double *H1h,
                     const double *He1h, const double *He2h, double z,
                     double Hz, const double *fcA_H2, const double *fcA_He2,
                     const double *fcA_He3, double tol, int nn, double *xH1,
                     double *xH2, double *xHe1, double *xHe2, double *xHe3,
                     double *T) {
  for (int i = 0; i < nn; i++) {
    int rc = orc_pcte1(nH[i], nHe[i], H1i[i], He1i[i], He2i[i], H1h[i],
                       He1h[i], He2h[i], z, Hz, fcA_H2[i], fcA_He2[i],
                       fcA_He3[i], tol, &xH1[i], &xH2[i], &xHe1[i], &xHe2[i],
                       &xHe3[i], &T[i]);
    if (rc) return rc;
  }
  return ORC_OK;
}

/* ion_solver.f90:1138-1183 (get_dudT_v): vector rates + lock-step solve_pce_v */
static int dudT_v(const double *nH, const double *nHe, const double *T,
                  const double *H1i, const double *He1i, const double *He2i,
                  const double *H1h, const double *He1h, const double *He2h,
                  double z, double Hz, const double *fcA_H2,
                  const double *fcA_He2, const double *fcA_He3, double tol,
                  int nn, double *dudT, double *work /* 11*nn */) {
  double *reH2 = work, *reHe2 = work + nn, *reHe3 = work + 2 * nn;
  double *ciH1 = work + 3 * nn, *ciHe1 = work + 4 * nn, *ciHe2 = work + 5 * nn;
  double *xH1 = work + 6 * nn, *xH2 = work + 7 * nn, *xHe1 = work + 8 * nn;
  double *xHe2 = work + 9 * nn, *xHe3 = work + 10 * nn;
  orc_get_kchem(T, fcA_H2, fcA_He2, fcA_He3, nn, reH2, reHe2, reHe3, ciH1,
                ciHe1, ciHe2);
  int rc = orc_solve_pce_v(nH, nHe, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2,
                           H1i, He1i, He2i, tol, nn, xH1, xH2, xHe1, xHe2,
                           xHe3);
  if (rc) return rc;
  for (int i = 0; i < nn; i++) {
    orc_kcool k;
    orc_kcool1(T[i], fcA_H2[i], fcA_He2[i], fcA_He3[i], &k);
    double heat =
        orc_heat1(nH[i], nHe[i], H1h[i], He1h[i], He2h[i], xH1[i], xHe1[i],
                  xHe2[i]);
    double cool = orc_cool1(nH[i], nHe[i], T[i], &k, xH1[i], xH2[i], xHe1[i],
                            xHe2[i], xHe3[i], z, Hz);
    dudT[i] = heat - cool;
  }
  return ORC_OK;
}

/* ion_solver.f90:884-1056 (solve_pcte_v) */
int orc_solve_pcte_v(const double *nH, const double *nHe, const double *H1i,
                     const double *He1i, const double *He2i, const double *H1h,
                     const double *He1h, const double *He2h, double z,
                     double Hz, const double *fcA_H2, const double *fcA_He2,
                     const double *fcA_He3, double tol, int nn, double *xH1,
                     double *xH2, double *xHe1, double *xHe2, double *xHe3,
                     double *T) {
  size_t n = (size_t)nn;
  double *buf = (double *)malloc(sizeof(double) * n * 24);
  char *b_floor = (char *)calloc(n, 1);
  double *Tleft = buf, *Tright = buf + n, *Told = buf + 2 * n;
  double *dudT = buf + 3 * n, *dudTmin = buf + 4 * n, *dudTmax = buf + 5 * n;
  double *err = buf + 6 * n;
  double *re_ci = buf + 7 * n;   /* 6*nn */
  double *work = buf + 13 * n;   /* 11*nn */
  int rc = ORC_OK;

  for (int i = 0; i < nn; i++) {
    Tleft[i] = TFLOOR;
    Tright[i] = TMAX;
  }
  rc = dudT_v(nH, nHe, Tleft, H1i, He1i, He2i, H1h, He1h, He2h, z, Hz, fcA_H2,
              fcA_He2, fcA_He3, tol, nn, dudTmin, work);
  if (rc) goto done;
  for (int i = 0; i < nn; i++)
    if (dudTmin[i] < 0.0) b_floor[i] = 1; /* :941-946 */
  rc = dudT_v(nH, nHe, Tright, H1i, He1i, He2i, H1h, He1h, He2h, z, Hz, fcA_H2,
              fcA_He2, fcA_He3, tol, nn, dudTmax, work);
  if (rc) goto done;
  for (int i = 0; i < nn; i++) {
    /* :958-969 ; the first mask can never be true */
    if (dudTmax[i] > 0.0 && !b_floor[i]) {
      rc = ORC_ERR_T_NOT_BRACKETED;
      goto done;
    }
  }
  for (int i = 0; i < nn; i++) { /* :974-979 */
    if (b_floor[i]) {
      T[i] = Tleft[i];
    } else {
      T[i] = (log10(Tleft[i]) + log10(Tright[i])) * 0.5;
      T[i] = pow(10.0, T[i]);
    }
    err[i] = HUGE_ONE;
  }
  int iter = 0;
  for (;;) {
    int any = 0;
    for (int i = 0; i < nn; i++)
      if (err[i] > tol) any = 1;
    if (!any) break; /* :988 */
    rc = dudT_v(nH, nHe, T, H1i, He1i, He2i, H1h, He1h, He2h, z, Hz, fcA_H2,
                fcA_He2, fcA_He3, tol, nn, dudT, work);
    if (rc) goto done;
    for (int i = 0; i < nn; i++) {
      Told[i] = T[i];
      if (dudT[i] < 0.0 && !b_floor[i]) { /* :1001-1005 */
        Tright[i] = T[i];
        T[i] = (Tleft[i] + T[i]) * 0.5;
      }
      /* second mask is evaluated on the same dudT: :1007-1011 */
      if (dudT[i] > 0.0 && !b_floor[i]) {
        Tleft[i] = Told[i];
        T[i] = (Told[i] + Tright[i]) * 0.5;
      }
      err[i] = fabs(1.0 - T[i] / Told[i]);
    }
    iter = iter + 1;
    if (iter > ION_MAX_ITER) {
      rc = ORC_ERR_MAX_ITER_PCTE;
      goto done;
    }
  }
  {
    double *reH2 = re_ci, *reHe2 = re_ci + n, *reHe3 = re_ci + 2 * n;
    double *ciH1 = re_ci + 3 * n, *ciHe1 = re_ci + 4 * n, *ciHe2 = re_ci + 5 * n;
    orc_get_kchem(T, fcA_H2, fcA_He2, fcA_He3, nn, reH2, reHe2, reHe3, ciH1,
                  ciHe1, ciHe2);
    rc = orc_solve_pce_v(nH, nHe, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, H1i,
                         He1i, He2i, tol, nn, xH1, xH2, xHe1, xHe2, xHe3);
    if (rc) goto done;
  }
  for (int i = 0; i < nn; i++) {
    if (nH[i] <= 0.0) {
      xH1[i] = 0.0;
      xH2[i] = 0.0;
    }
    if (nHe[i] <= 0.0) {
      xHe1[i] = 0.0;
      xHe2[i] = 0.0;
      xHe3[i] = 0.0;
    }
  }
done:
  free(buf);
  free(b_floor);
  return rc;
}
