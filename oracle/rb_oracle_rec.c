/*
 * rb_oracle_rec.c -- recombination-photon transfer of the geometric solvers
 * (rec_meth = outward / radial / isotropic for spheres, "ray" for slabs).
 * TEST INFRASTRUCTURE ONLY (see rb_oracle.h).
 *
 * Statement-by-statement restatements of
 *   set_recomb_photons_outward    rabacus/f2py/sphere_base.f90:79-346
 *   set_recomb_photons_radial     rabacus/f2py/sphere_base.f90:392-750
 *   set_recomb_photons_isotropic  rabacus/f2py/sphere_base.f90:800-1083
 *   set_recomb_photons (slabs)    rabacus/f2py/slab_base.f90:163-560
 * Arrays arrive in the orientation the CALLING driver uses; where the Fortran
 * routine reverses them on entry and back on exit (reverse_arrays,
 * sphere_base.f90:1404-1452) the same is done here through an index map, which
 * keeps every sum in the reference's order.
 */
#include "rb_oracle_internal.h"

typedef struct {
  double sH1_H1, sH1_He1, sH1_He2, sHe1_He1, sHe1_He2, sHe2_He2;
} xsec_t;

/* sphere_base.f90:224-236: cross sections at the three thresholds */
static void xsec_at_thresholds(const double sigma_th[3], const double E_th[3],
                               xsec_t *x) {
  double E, s;
  E = E_th[2];
  orc_v96_sigma(0, &E, 1, &s); x->sH1_He2 = s;
  orc_v96_sigma(1, &E, 1, &s); x->sHe1_He2 = s;
  x->sHe2_He2 = sigma_th[2];
  E = E_th[1];
  orc_v96_sigma(0, &E, 1, &s); x->sH1_He1 = s;
  x->sHe1_He1 = sigma_th[1];
  x->sH1_H1 = sigma_th[0];
}

/* emission coefficients [photons / (s cm^3 sr)]: sphere_base.f90:238-275 */
static void emission(const double *xH2, const double *xHe2, const double *xHe3,
                     const double *nH, const double *nHe, const double *T,
                     int Nl, double *em /* [3][Nl] */) {
  for (int i = 0; i < Nl; i++) {
    double aH2, aHe2, aHe3, bH2, bHe2, bHe3, c1, c2, c3;
    orc_kchem1(T[i], 1.0, 1.0, 1.0, &aH2, &aHe2, &aHe3, &c1, &c2, &c3);
    orc_kchem1(T[i], 0.0, 0.0, 0.0, &bH2, &bHe2, &bHe3, &c1, &c2, &c3);
    const double reH2_1 = aH2 - bH2, reHe2_1 = aHe2 - bHe2, reHe3_1 = aHe3 - bHe3;
    const double ne = xH2[i] * nH[i] + (xHe2[i] + 2.0 * xHe3[i]) * nHe[i];
    em[i] = (reH2_1 * nH[i] * xH2[i] * ne) / (4.0 * PI_);
    em[Nl + i] = (reHe2_1 * nHe[i] * xHe2[i] * ne) / (4.0 * PI_);
    em[2 * Nl + i] = (reHe3_1 * nHe[i] * xHe3[i] * ne) / (4.0 * PI_);
  }
}

/* rates from the three fluxes: sphere_base.f90:311-333 (outward, radial) */
static void rates_from_flux(const xsec_t *x, const double E_th[3], double F_H2,
                            double F_He2, double F_He3, double ion[3],
                            double heat[3]) {
  ion[0] = F_H2 * x->sH1_H1 + F_He2 * x->sH1_He1 + F_He3 * x->sH1_He2;
  ion[1] = F_He2 * x->sHe1_He1 + F_He3 * x->sHe1_He2;
  ion[2] = F_He3 * x->sHe2_He2;
  heat[0] = F_He2 * (E_th[1] - E_th[0]) * x->sH1_He1 +
            F_He3 * (E_th[2] - E_th[0]) * x->sH1_He2;
  heat[0] = heat[0] * EV_2_ERG;
  heat[1] = F_He2 * x->sHe1_He1 + F_He3 * x->sHe1_He2; /* sic: no energy factor */
  heat[1] = heat[1] * EV_2_ERG;
  heat[2] = 0.0;
  heat[2] = heat[2] * EV_2_ERG;
}

/* Shared set-up of outward / radial in the routine's own orientation.
 * `dec` = 1: index 0 is the outermost shell (outward); 0: innermost (radial).
 * `plan_dec`: orientation of the caller's arrays.  map(i) = caller index.    */
typedef struct {
  int Nl;
  double *r_c, *dtau[3], *F0[3], *buf;
} shell_t;

static void shells_init(shell_t *s, int want_dec, int have_dec,
                        const double *xH1, const double *xH2, const double *xHe1,
                        const double *xHe2, const double *xHe3,
                        const double *Edges, const double *nH, const double *nHe,
                        const double *T, const double sigma_th[3], int Nl) {
  size_t n = (size_t)Nl;
  s->Nl = Nl;
  s->buf = (double *)malloc(sizeof(double) * (10 * n));
  s->r_c = s->buf;
  for (int X = 0; X < 3; X++) {
    s->dtau[X] = s->buf + (1 + X) * n;
    s->F0[X] = s->buf + (4 + X) * n;
  }
  double *em = s->buf + 7 * n; /* [3][Nl], caller orientation */
  emission(xH2, xHe2, xHe3, nH, nHe, T, Nl, em);
  const int flip = (want_dec != have_dec);
  for (int i = 0; i < Nl; i++) {
    const int c = flip ? Nl - 1 - i : i;           /* caller cell index */
    const int e0 = flip ? Nl - i : i;              /* caller edge index of E(i) */
    const int e1 = flip ? Nl - i - 1 : i + 1;      /* ... of E(i+1) */
    const double Ei = Edges[e0], Ej = Edges[e1];
    const double dr = fabs(Ei - Ej);               /* sphere_base.f90:1534-1535 */
    s->r_c[i] = (Ei + Ej) * 0.5;
    const double dNH = dr * nH[c], dNHe = dr * nHe[c];
    s->dtau[0][i] = dNH * xH1[c] * sigma_th[0];    /* :216-218 */
    s->dtau[1][i] = dNHe * xHe1[c] * sigma_th[1];
    s->dtau[2][i] = dNHe * xHe2[c] * sigma_th[2];
    const double area = 4.0 * PI_ * s->r_c[i] * s->r_c[i];   /* :220 */
    const double vol = fabs(4.0 / 3.0 * PI_ * (Ei * Ei * Ei - Ej * Ej * Ej)); /* :221 */
    for (int X = 0; X < 3; X++)                    /* :281-283 */
      s->F0[X][i] = em[X * Nl + c] * 4.0 * PI_ * vol / area;
  }
}

/* one recombining layer's contribution: sphere_base.f90:275-307 */
static void add_layer(const xsec_t *x, const double sigma_th[3],
                      const double tau_th[3], double geo_fac, double fac,
                      const double F0[3], double F[3]) {
  double ra_H1, ra_He1, ra_He2, tau;
  ra_H1 = x->sH1_H1 / sigma_th[0];
  tau = tau_th[0] * ra_H1;
  F[0] = F[0] + fac * F0[0] * geo_fac * exp(-tau);
  ra_H1 = x->sH1_He1 / sigma_th[0];
  ra_He1 = x->sHe1_He1 / sigma_th[1];
  tau = tau_th[0] * ra_H1 + tau_th[1] * ra_He1;
  F[1] = F[1] + fac * F0[1] * geo_fac * exp(-tau);
  ra_H1 = x->sH1_He2 / sigma_th[0];
  ra_He1 = x->sHe1_He2 / sigma_th[1];
  ra_He2 = x->sHe2_He2 / sigma_th[2];
  tau = tau_th[0] * ra_H1 + tau_th[1] * ra_He1 + tau_th[2] * ra_He2;
  F[2] = F[2] + fac * F0[2] * geo_fac * exp(-tau);
}

/* sphere_base.f90:79-346.  fac == 1: the Fortran has no factor here; x*1.0 is exact. */
void orc_rec_outward(const double *xH1, const double *xH2, const double *xHe1,
                     const double *xHe2, const double *xHe3, const double *Edges,
                     const double *nH, const double *nHe, const double *T,
                     int have_dec, const double sigma_th[3], const double E_th[3],
                     int Nl, double *ion[3], double *heat[3]) {
  xsec_t x;
  xsec_at_thresholds(sigma_th, E_th, &x);
  shell_t s;
  shells_init(&s, 1, have_dec, xH1, xH2, xHe1, xHe2, xHe3, Edges, nH, nHe, T,
              sigma_th, Nl);
  const int flip = !have_dec;
#pragma omp parallel for schedule(dynamic, 8)
  for (int isl = 0; isl < Nl; isl++) {
    double tau[3] = {0.0, 0.0, 0.0}, F[3] = {0.0, 0.0, 0.0};
    for (int irl = isl; irl < Nl; irl++) { /* layers at smaller radii */
      for (int X = 0; X < 3; X++) {        /* :249-266 */
        if (irl == isl) {
          tau[X] = tau[X] + s.dtau[X][irl] * 0.5;
        } else if (irl == isl + 1) {
          tau[X] = tau[X] + s.dtau[X][irl] * 0.5;
        } else {
          tau[X] = tau[X] + s.dtau[X][irl - 1] * 0.5;
          tau[X] = tau[X] + s.dtau[X][irl] * 0.5;
        }
      }
      const double g = (s.r_c[irl] / s.r_c[isl]) * (s.r_c[irl] / s.r_c[isl]);
      const double F0[3] = {s.F0[0][irl], s.F0[1][irl], s.F0[2][irl]};
      add_layer(&x, sigma_th, tau, g, 1.0, F0, F);
    }
    double io[3], he[3];
    rates_from_flux(&x, E_th, F[0], F[1], F[2], io, he);
    const int c = flip ? Nl - 1 - isl : isl;
    for (int X = 0; X < 3; X++) {
      ion[X][c] = io[X];
      heat[X][c] = he[X];
    }
  }
  free(s.buf);
}

/* sphere_base.f90:392-750 (1-based in the Fortran; 0-based here: il runs over
 * isl, isl-1, .., 1, -1, .., -Nl in 1-based numbering, il == 0 skipped) */
void orc_rec_radial(const double *xH1, const double *xH2, const double *xHe1,
                    const double *xHe2, const double *xHe3, const double *Edges,
                    const double *nH, const double *nHe, const double *T,
                    int have_dec, const double sigma_th[3], const double E_th[3],
                    int Nl, double *ion[3], double *heat[3]) {
  xsec_t x;
  xsec_at_thresholds(sigma_th, E_th, &x);
  shell_t s;
  shells_init(&s, 0, have_dec, xH1, xH2, xHe1, xHe2, xHe3, Edges, nH, nHe, T,
              sigma_th, Nl);
  const int flip = have_dec;
#pragma omp parallel for schedule(dynamic, 8)
  for (int isl1 = 1; isl1 <= Nl; isl1++) { /* 1-based solve layer */
    double tau[3] = {0.0, 0.0, 0.0}, F[3] = {0.0, 0.0, 0.0};
    for (int il = isl1; il >= -Nl; il--) { /* :558-617 lower layers, through the centre */
      if (il == 0) continue;
      const int irl1 = abs(il), irl = irl1 - 1;
      for (int X = 0; X < 3; X++) {
        if (irl1 == isl1) {
          tau[X] = tau[X] + s.dtau[X][irl] * 0.5;
        } else if (irl1 == isl1 - 1) {
          tau[X] = tau[X] + s.dtau[X][irl] * 0.5;
        } else if (irl1 < isl1 - 1) {
          tau[X] = tau[X] + s.dtau[X][irl + 1] * 0.5;
          tau[X] = tau[X] + s.dtau[X][irl] * 0.5;
        }
      }
      const double g = (s.r_c[irl] / s.r_c[isl1 - 1]) * (s.r_c[irl] / s.r_c[isl1 - 1]);
      const double F0[3] = {s.F0[0][irl], s.F0[1][irl], s.F0[2][irl]};
      add_layer(&x, sigma_th, tau, g, 0.5, F0, F);
    }
    tau[0] = tau[1] = tau[2] = 0.0;
    for (int il = isl1; il <= Nl; il++) { /* :625-684 upper layers */
      const int irl1 = il, irl = irl1 - 1;
      for (int X = 0; X < 3; X++) {
        if (irl1 == isl1) {
          tau[X] = tau[X] + s.dtau[X][irl] * 0.5;
        } else if (irl1 == isl1 + 1) {
          tau[X] = tau[X] + s.dtau[X][irl] * 0.5;
        } else {
          tau[X] = tau[X] + s.dtau[X][irl] * 0.5;
          tau[X] = tau[X] + s.dtau[X][irl - 1] * 0.5;
        }
      }
      const double g = (s.r_c[irl] / s.r_c[isl1 - 1]) * (s.r_c[irl] / s.r_c[isl1 - 1]);
      const double F0[3] = {s.F0[0][irl], s.F0[1][irl], s.F0[2][irl]};
      add_layer(&x, sigma_th, tau, g, 0.5, F0, F);
    }
    double io[3], he[3];
    rates_from_flux(&x, E_th, F[0], F[1], F[2], io, he);
    const int c = flip ? Nl - isl1 : isl1 - 1;
    for (int X = 0; X < 3; X++) {
      ion[X][c] = io[X];
      heat[X][c] = he[X];
    }
  }
  free(s.buf);
}

/* sphere_base.f90:800-1083.  Needs decreasing Edges (index 0 = outermost). */
int orc_rec_isotropic(const double *xH1, const double *xH2, const double *xHe1,
                      const double *xHe2, const double *xHe3, const double *Edges,
                      const double *nH, const double *nHe, const double *T,
                      int have_dec, int Nmu, const double sigma_th[3],
                      const double E_th[3], const double em_fac[3], int Nl,
                      double *ion[3], double *heat[3]) {
  xsec_t x;
  xsec_at_thresholds(sigma_th, E_th, &x);
  size_t n = (size_t)Nl;
  double *buf = (double *)malloc(sizeof(double) * (10 * n + 1 + 2 * (size_t)Nmu));
  double *em_in = buf, *em[3] = {buf + 3 * n, buf + 4 * n, buf + 5 * n};
  double *dk[3] = {buf + 6 * n, buf + 7 * n, buf + 8 * n};
  double *E = buf + 9 * n; /* [Nl+1] decreasing */
  double *xi = buf + 10 * n + 1, *wi = xi + Nmu;
  emission(xH2, xHe2, xHe3, nH, nHe, T, Nl, em_in);
  const int flip = !have_dec;
  for (int i = 0; i <= Nl; i++) E[i] = Edges[flip ? Nl - i : i];
  for (int i = 0; i < Nl; i++) {
    const int c = flip ? Nl - 1 - i : i;
    /* :878-892 partial opacities at the three threshold energies */
    const double kH1_H1 = nH[c] * xH1[c] * x.sH1_H1;
    const double kHe1_H1 = nH[c] * xH1[c] * x.sH1_He1;
    const double kHe2_H1 = nH[c] * xH1[c] * x.sH1_He2;
    const double kHe1_He1 = nHe[c] * xHe1[c] * x.sHe1_He1;
    const double kHe2_He1 = nHe[c] * xHe1[c] * x.sHe1_He2;
    const double kHe2_He2 = nHe[c] * xHe2[c] * x.sHe2_He2;
    dk[0][i] = kH1_H1;
    dk[1][i] = kHe1_H1 + kHe1_He1;
    dk[2][i] = kHe2_H1 + kHe2_He1 + kHe2_He2;
    for (int X = 0; X < 3; X++) em[X][i] = em_in[X * Nl + c] * em_fac[X]; /* :935-937 */
  }
  orc_p_quadrature_rule(Nmu, xi, wi); /* :943 */
  int rc_all = ORC_OK;
#pragma omp parallel for schedule(dynamic, 1)
  for (int il = 0; il < Nl; il++) {
    double J[3] = {0.0, 0.0, 0.0};
    int bad = 0;
    for (int imu = 0; imu < Nmu; imu++) {
      const double mu = xi[imu];
      const int m = orc_m_for_ray(E, il, mu, Nl);
      if (m < 0) { bad = 1; break; }
      double tau[3] = {0.0, 0.0, 0.0}, I[3] = {0.0, 0.0, 0.0};
      const int j0 = (mu <= 0.0) ? il : 2 * m - il;
      int err = 0;
      /* ds_segment (geometry.f90:90-168) with its per-call ray set-up hoisted */
      const double r_ray = (E[il] + E[il + 1]) * 0.5;
      const double p_ray = r_ray * sqrt(1.0 - mu * mu);
      for (int j = j0; j <= 2 * m; j++) {
        const int jnl = m - abs(m - j);
        double ds = orc_ds_hoisted(E, p_ray, m, j);
        if (j == j0) ds = ds * 0.5;
        for (int X = 0; X < 3; X++) {
          tau[X] = tau[X] + ds * dk[X][jnl];
          const double dI = em[X][jnl] * ds * exp(-tau[X]);
          I[X] = I[X] + dI;
        }
      }
      if (err) { bad = 1; break; }
      /* J = sum(I * wi): the running sum over imu in index order (:1020-1022) */
      for (int X = 0; X < 3; X++) J[X] = J[X] + I[X] * wi[imu];
    }
    if (bad) {
#pragma omp critical
      rc_all = ORC_ERR_RAY_GEOMETRY;
      continue;
    }
    for (int X = 0; X < 3; X++) J[X] = J[X] * 0.5; /* :1030-1032 */
    const int c = flip ? Nl - 1 - il : il;
    /* :1037-1061 */
    ion[0][c] = (J[0] * x.sH1_H1 + J[1] * x.sH1_He1 + J[2] * x.sH1_He2) * 4.0 * PI_;
    ion[1][c] = (J[1] * x.sHe1_He1 + J[2] * x.sHe1_He2) * 4.0 * PI_;
    ion[2][c] = (J[2] * x.sHe2_He2) * 4.0 * PI_;
    double h = (J[1] * (E_th[1] - E_th[0]) * x.sH1_He1 +
                J[2] * (E_th[2] - E_th[0]) * x.sH1_He2) * 4.0 * PI_;
    heat[0][c] = h * EV_2_ERG;
    h = (J[1] * x.sHe1_He1 + J[2] * x.sHe1_He2) * 4.0 * PI_;
    heat[1][c] = h * EV_2_ERG;
    heat[2][c] = 0.0 * EV_2_ERG;
  }
  free(buf);
  return rc_all;
}

/* slab_base.f90:163-560: plane-parallel recombination radiation, E1 kernels
 * integrated with the trapezoid rule over the layer edges */
void orc_rec_slab(const double *xH1, const double *xH2, const double *xHe1,
                  const double *xHe2, const double *xHe3, const double *Edges,
                  const double *nH, const double *nHe, const double *T,
                  const double sigma_th[3], const double E_th[3], int Nl,
                  double *ion[3], double *heat[3]) {
  xsec_t x;
  xsec_at_thresholds(sigma_th, E_th, &x);
  size_t n = (size_t)Nl;
  double *buf = (double *)malloc(sizeof(double) * (4 * n + 6 * (n + 1)));
  double *em = buf;                 /* [3][Nl] */
  double *dl = buf + 3 * n;         /* [Nl] */
  double *tlo[3], *eme[3];          /* [Nl+1] each */
  for (int X = 0; X < 3; X++) {
    tlo[X] = buf + 4 * n + (size_t)X * (n + 1);
    eme[X] = buf + 4 * n + (size_t)(3 + X) * (n + 1);
  }
  emission(xH2, xHe2, xHe3, nH, nHe, T, Nl, em);
  /* :297-304 cross-section ratios */
  const double ra_H1_H1 = 1.0, ra_H1_He1 = x.sH1_He1 / x.sH1_H1,
               ra_H1_He2 = x.sH1_He2 / x.sH1_H1, ra_He1_He1 = 1.0,
               ra_He1_He2 = x.sHe1_He2 / x.sHe1_He1, ra_He2_He2 = 1.0;
  tlo[0][0] = tlo[1][0] = tlo[2][0] = 0.0;
  for (int k = 1; k <= Nl; k++) { /* :352-366 */
    dl[k - 1] = Edges[k] - Edges[k - 1];
    const double dNH = dl[k - 1] * nH[k - 1], dNHe = dl[k - 1] * nHe[k - 1];
    const double d1 = dNH * xH1[k - 1] * sigma_th[0];
    const double d2 = dNHe * xHe1[k - 1] * sigma_th[1];
    const double d3 = dNHe * xHe2[k - 1] * sigma_th[2];
    tlo[0][k] = tlo[0][k - 1] + d1 * ra_H1_H1;
    tlo[1][k] = tlo[1][k - 1] + d1 * ra_H1_He1 + d2 * ra_He1_He1;
    tlo[2][k] = tlo[2][k - 1] + d1 * ra_H1_He2 + d2 * ra_He1_He2 + d3 * ra_He2_He2;
  }
  for (int k = 0; k <= Nl; k++) /* :378-392 emission coefficients on edges */
    for (int X = 0; X < 3; X++) {
      if (k == 0) eme[X][k] = em[X * Nl + k] * 0.5;
      else if (k == Nl) eme[X][k] = em[X * Nl + k - 1] * 0.5;
      else eme[X][k] = (em[X * Nl + k - 1] + em[X * Nl + k]) * 0.5;
    }
#pragma omp parallel
  {
    double *Jint = (double *)malloc(sizeof(double) * (n + 1));
#pragma omp for schedule(dynamic, 8)
    for (int k = 0; k < Nl; k++) {
      double J[3];
      for (int X = 0; X < 3; X++) { /* :402-538, same code for the three species */
        double tau = (tlo[X][k + 1] - tlo[X][k]) * 0.5;
        double Jlo = dl[k] * em[X * Nl + k] * orc_e1xa(tau) * 0.5;
        double Jhi = Jlo;
        for (int i = 0; i <= Nl; i++) Jint[i] = 0.0;
        if (k != 0) {
          tau = (tlo[X][k + 1] - tlo[X][k]) * 0.5;
          Jint[k] = eme[X][k] * orc_e1xa(tau);
          for (int i = k - 1; i >= 0; i--) {
            tau = tau + (tlo[X][i + 1] - tlo[X][i]);
            Jint[i] = eme[X][i] * orc_e1xa(tau);
          }
          Jlo = Jlo + orc_integrate_trap(Edges, Jint, k + 1);
        }
        if (k != Nl - 1) {
          tau = (tlo[X][k + 1] - tlo[X][k]) * 0.5;
          Jint[k + 1] = eme[X][k + 1] * orc_e1xa(tau);
          for (int i = k + 1; i <= Nl - 1; i++) {
            tau = tau + (tlo[X][i + 1] - tlo[X][i]);
            Jint[i + 1] = eme[X][i + 1] * orc_e1xa(tau);
          }
          Jhi = Jhi + orc_integrate_trap(Edges + k + 1, Jint + k + 1, Nl - k);
        }
        J[X] = Jlo + Jhi;                /* :544-546 */
        J[X] = J[X] * 4.0 * PI_ * 0.5;   /* :550-552 */
      }
      /* :556-572 (J[0] = H2, J[1] = He2, J[2] = He3 recombinations) */
      ion[0][k] = J[2] * x.sH1_He2 + J[1] * x.sH1_He1 + J[0] * x.sH1_H1;
      ion[1][k] = J[2] * x.sHe1_He2 + J[1] * x.sHe1_He1;
      ion[2][k] = J[2] * x.sHe2_He2;
      heat[0][k] = J[2] * (E_th[2] - E_th[0]) * x.sH1_He2 * EV_2_ERG +
                   J[1] * (E_th[1] - E_th[0]) * x.sH1_He1 * EV_2_ERG;
      heat[1][k] = J[2] * (E_th[2] - E_th[0]) * x.sHe1_He2 * EV_2_ERG; /* sic: E_H1_th */
      heat[2][k] = 0.0;
    }
    free(Jint);
  }
  free(buf);
}
