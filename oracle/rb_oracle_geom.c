/*
 * rb_oracle_geom.c -- oracle (TEST INFRASTRUCTURE ONLY, see rb_oracle.h):
 * the four geometric drivers and their leaf numerics, restating
 *   rabacus/f2py/geometry.f90, source_background.f90, source_point.f90,
 *   source_plane.f90, sphere_base.f90 (helpers), slab_base.f90 (helpers),
 *   sphere_bgnd.f90, sphere_stromgren.f90, slab_bgnd.f90, slab_plane.f90.
 * Only rec_meth = "fixed" (i_rec_meth == 1) is implemented; the recombination
 * photon transfer routines are SURVEY 8(f) "next" rows and return BAD_ARG.
 */
#include "rb_oracle_internal.h"

#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* bench.py sets the team size explicitly (torchrun exports OMP_NUM_THREADS=1) */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n >= 1) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* wall time of the sweep loop of the last orc_sphere_bgnd_solve (thin start excluded):
 * what bench.py's CPU legs report.  Not thread safe (test infrastructure). */
static double g_sweep_seconds = 0.0;
double orc_last_sweep_seconds(void) { return g_sweep_seconds; }
static double wall_now(void) {
#ifdef _OPENMP
  return omp_get_wtime();
#else
  return 0.0;
#endif
}

enum { SRC_BGND = 0, SRC_POINT = 1, SRC_PLANE = 2 };

/* counters: 0 segment visits, 1 sqrt calls, 2 exp/e2xa calls, 3 cell solves */
typedef struct {
  long long seg, sq, ex, cell;
} cnt_t;

/* ---------------------------------------------------------------------- */
/* sphere_base.f90:1467-1509                                               */
static int monotonic_decreasing(const double *a, int N) {
  for (int i = 1; i < N; i++)
    if (a[i] > a[i - 1]) return 0;
  return 1;
}
static int monotonic_increasing(const double *a, int N) {
  for (int i = 1; i < N; i++)
    if (a[i] < a[i - 1]) return 0;
  return 1;
}
static void rev(double *a, int n) {
  for (int i = 0, j = n - 1; i < j; i++, j--) {
    double t = a[i];
    a[i] = a[j];
    a[j] = t;
  }
}

/* ---------------------------------------------------------------------- */
/* geometry.f90:37-73 (m_for_ray). Returns -1 where the reference would use
 * an undefined m (no Edges(i) <= p_ray).                                    */
int orc_m_for_ray(const double *Edges, int il, double mu, int Nl) {
  double r_ray = (Edges[il] + Edges[il + 1]) * 0.5;
  double p_ray = r_ray * sqrt(1.0 - mu * mu);
  for (int i = 1; i <= Nl; i++)
    if (Edges[i] <= p_ray) return i - 1;
  return -1;
}

/* geometry.f90:90-168 (ds_segment), including the per-call tangent-shell search
 * (SURVEY Q7). */
double orc_ds_segment(const double *Edges, int il, double mu, int j, int Nl,
                      int *err) {
  double r_ray = (Edges[il] + Edges[il + 1]) * 0.5;
  double p_ray = r_ray * sqrt(1.0 - mu * mu);
  int m = -1;
  for (int i = 1; i <= Nl; i++)
    if (Edges[i] <= p_ray) {
      m = i - 1;
      break;
    }
  int n = 2 * m + 1;
  if (m < 0 || j < 0 || j > n - 1) {
    if (err) *err = ORC_ERR_RAY_GEOMETRY;
    return 0.0;
  }
  double ds, s1, s2;
  if (j < m) {
    s1 = sqrt(Edges[j] * Edges[j] - p_ray * p_ray);
    s2 = sqrt(Edges[j + 1] * Edges[j + 1] - p_ray * p_ray);
    ds = fabs(s2 - s1);
  } else if (j == m) {
    s1 = sqrt(Edges[j] * Edges[j] - p_ray * p_ray);
    ds = 2 * s1;
  } else {
    int i1 = m - abs(m - j);
    int i2 = i1 + 1;
    s1 = sqrt(Edges[i1] * Edges[i1] - p_ray * p_ray);
    s2 = sqrt(Edges[i2] * Edges[i2] - p_ray * p_ray);
    ds = fabs(s2 - s1);
  }
  return ds;
}

/* same arithmetic as ds_segment with m and p hoisted out (flavour HOISTED) */
double orc_ds_hoisted(const double *Edges, double p_ray, int m, int j) {
  double s1, s2;
  if (j == m) {
    s1 = sqrt(Edges[j] * Edges[j] - p_ray * p_ray);
    return 2 * s1;
  }
  int i1 = (j < m) ? j : m - abs(m - j);
  int i2 = i1 + 1;
  s1 = sqrt(Edges[i1] * Edges[i1] - p_ray * p_ray);
  s2 = sqrt(Edges[i2] * Edges[i2] - p_ray * p_ray);
  return fabs(s2 - s1);
}

/* sphere_bgnd.f90:766-792: columns along one ray */
static int ray_columns(const double *Edges, const double *nH, const double *nHe,
                       const double *xH1, const double *xHe1,
                       const double *xHe2, int il, double mu, int Nl,
                       int flavour, double *pNH1, double *pNHe1, double *pNHe2,
                       cnt_t *c) {
  int m = orc_m_for_ray(Edges, il, mu, Nl);
  if (m < 0) return ORC_ERR_RAY_GEOMETRY;
  int j0 = (mu <= 0.0) ? il : 2 * m - il;
  double NH1 = 0.0, NHe1 = 0.0, NHe2 = 0.0;
  double r_ray = (Edges[il] + Edges[il + 1]) * 0.5;
  double p_ray = r_ray * sqrt(1.0 - mu * mu);
  for (int j = j0; j <= 2 * m; j++) {
    int jnl = m - abs(m - j);
    double ds;
    if (flavour == ORC_FLAVOUR_FAITHFUL) {
      int e = 0;
      ds = orc_ds_segment(Edges, il, mu, j, Nl, &e);
      if (e) return e;
    } else {
      ds = orc_ds_hoisted(Edges, p_ray, m, j);
    }
    if (j == j0) ds = ds * 0.5;
    NH1 = NH1 + ds * nH[jnl] * xH1[jnl];
    NHe1 = NHe1 + ds * nHe[jnl] * xHe1[jnl];
    NHe2 = NHe2 + ds * nHe[jnl] * xHe2[jnl];
    if (c) {
      c->seg += 1;
      c->sq += (j == m) ? 1 : 2;
    }
  }
  *pNH1 = NH1;
  *pNHe1 = NHe1;
  *pNHe2 = NHe2;
  return ORC_OK;
}

int orc_sphere_ray_columns(const double *Edges, const double *nH,
                           const double *nHe, const double *xH1,
                           const double *xHe1, const double *xHe2, int il,
                           double mu, int Nl, double *NH1, double *NHe1,
                           double *NHe2) {
  return ray_columns(Edges, nH, nHe, xH1, xHe1, xHe2, il, mu, Nl,
                     ORC_FLAVOUR_FAITHFUL, NH1, NHe1, NHe2, NULL);
}

/* ---------------------------------------------------------------------- */
/* spectrum context                                                        */
typedef struct {
  int Nnu;
  int src;
  const double *E, *shape;
  double *sigma[3];    /* sigma_X(nu) */
  double *sigma_ra[3]; /* sigma_X / sigma_X_th */
  double sigma_th[3];
  double E_th[3];
} spec_t;

static void spec_init(spec_t *s, int src, const double *E, const double *shape,
                      int Nnu) {
  s->Nnu = Nnu;
  s->src = src;
  s->E = E;
  s->shape = shape;
  for (int X = 0; X < 3; X++) {
    s->E_th[X] = orc_v96_Eth(X); /* sphere_bgnd.f90:206-208 */
    orc_v96_sigma(X, &s->E_th[X], 1, &s->sigma_th[X]); /* :213-215 */
    s->sigma[X] = (double *)malloc(sizeof(double) * (size_t)Nnu);
    s->sigma_ra[X] = (double *)malloc(sizeof(double) * (size_t)Nnu);
    orc_v96_sigma(X, E, Nnu, s->sigma[X]); /* :220-222 */
    for (int k = 0; k < Nnu; k++)
      s->sigma_ra[X][k] = s->sigma[X][k] / s->sigma_th[X]; /* :669-671 */
  }
}
static void spec_free(spec_t *s) {
  for (int X = 0; X < 3; X++) {
    free(s->sigma[X]);
    free(s->sigma_ra[X]);
  }
}

/* The un-attenuated photon-number integrand  yy  before `* sigma`:
 *   background: four_pi * Inu / (E_eV * h_erg_s)     source_background.f90:306
 *   plane:      dFu_over_dnu / (E_eV * h_erg_s)      source_plane.f90:296
 *   point:      dLu_over_dnu / (E_eV*h_erg_s*4.0d0*pi*radius*radius)
 *                                                    source_point.f90:348      */
static double yy_base(const spec_t *s, int k, double radius) {
  const double four_pi = 4.0 * PI_;
  switch (s->src) {
    case SRC_BGND:
      return four_pi * s->shape[k] / (s->E[k] * H_ERG_S);
    case SRC_PLANE:
      return s->shape[k] / (s->E[k] * H_ERG_S);
    default:
      return s->shape[k] /
             (s->E[k] * H_ERG_S * 4.0 * PI_ * radius * radius);
  }
}

/* monochromatic photon flux Fn (nn == 1):
 *   background: four_pi*Inu(0)/E_eV(0)*erg_2_eV      source_background.f90:112
 *   plane:      dFu(0)/E_eV(0)*erg_2_eV              source_plane.f90:101
 *   point:      Ln/(4 pi r^2), Ln = Lu(0)/E(0)*erg_2_eV  source_point.f90:104,126 */
static double Fn_mono(const spec_t *s, double radius) {
  const double four_pi = 4.0 * PI_;
  switch (s->src) {
    case SRC_BGND:
      return four_pi * s->shape[0] / s->E[0] * ERG_2_EV;
    case SRC_PLANE:
      return s->shape[0] / s->E[0] * ERG_2_EV;
    default: {
      double Ln = s->shape[0] / s->E[0] * ERG_2_EV;
      return Ln / (4.0 * PI_ * radius * radius);
    }
  }
}

/* X_src_return_ionrate_thin / heatrate_thin:
 * source_background.f90:186-241, source_point.f90:226-281, source_plane.f90:178-231 */
static void rates_thin(const spec_t *s, double radius, double *ion3,
                       double *heat3, double *yy /* scratch Nnu */) {
  int nn = s->Nnu;
  for (int X = 0; X < 3; X++) {
    if (nn == 1) {
      double Fn = Fn_mono(s, radius);
      ion3[X] = Fn * s->sigma[X][0];
      heat3[X] = Fn * s->sigma[X][0] * (s->E[0] - s->E_th[X]) * EV_2_ERG;
    } else {
      for (int k = 0; k < nn; k++) yy[k] = yy_base(s, k, radius) * s->sigma[X][k];
      ion3[X] = orc_integrate_trap(s->E, yy, nn);
      for (int k = 0; k < nn; k++)
        yy[k] = yy[k] * (s->E[k] - s->E_th[X]) * EV_2_ERG;
      heat3[X] = orc_integrate_trap(s->E, yy, nn);
    }
  }
}

/* attenuation kinds */
enum { ATT_EXP = 0, ATT_E2 = 1 };

/* The six shielded integrals for one (cell, direction):
 *   X_src_return_ionrate_shld / heatrate_shld
 *   (source_background.f90:287-354, source_point.f90:299-395,
 *    source_plane.f90:249-342) and the E2 variants `_shld_ei`
 *   (source_background.f90:364-446, incl. the sum(tau)==0 thin shortcut).
 * tau_th[X] are the threshold optical depths; tau_X(nu) = tau_th[X]*sigma_ra[X].
 * FAITHFUL flavour recomputes the attenuation vector in each of the 6 calls
 * (SURVEY Q8); the numbers are the same.                                      */
static void rates_shld(const spec_t *s, const double tau_th[3], double radius,
                       int att_kind, int flavour, double *ion3, double *heat3,
                       double *scr /* 3*Nnu scratch */, cnt_t *c) {
  int nn = s->Nnu;
  double *atten = scr, *yy = scr + nn, *tau = scr + 2 * nn;
  int ncalls = (flavour == ORC_FLAVOUR_FAITHFUL) ? 6 : 1;

  for (int call = 0; call < ncalls; call++) {
    /* tau = tauH1 + tauHe1 + tauHe2 ; source_background.f90:267,385 */
    for (int k = 0; k < nn; k++)
      tau[k] = tau_th[0] * s->sigma_ra[0][k] + tau_th[1] * s->sigma_ra[1][k] +
               tau_th[2] * s->sigma_ra[2][k];
    if (att_kind == ATT_E2) {
      double st = 0.0;
      for (int k = 0; k < nn; k++) st += tau[k];
      if (st == 0.0) { /* source_background.f90:380-383,425-428 */
        rates_thin(s, radius, ion3, heat3, yy);
        return;
      }
      for (int k = 0; k < nn; k++) atten[k] = orc_e2xa(tau[k]);
    } else {
      for (int k = 0; k < nn; k++) atten[k] = exp(-tau[k]);
    }
    if (c) c->ex += nn;
  }

  if (nn == 1) {
    double Fn = Fn_mono(s, radius);
    for (int X = 0; X < 3; X++) {
      double ion = Fn * s->sigma[X][0];
      double heat = Fn * s->sigma[X][0] * (s->E[0] - s->E_th[X]) * EV_2_ERG;
      ion3[X] = ion * atten[0];
      heat3[X] = heat * atten[0];
    }
    return;
  }
  for (int X = 0; X < 3; X++) {
    for (int k = 0; k < nn; k++)
      yy[k] = yy_base(s, k, radius) * s->sigma[X][k] * atten[k];
    ion3[X] = orc_integrate_trap(s->E, yy, nn);
    for (int k = 0; k < nn; k++)
      yy[k] = yy[k] * (s->E[k] - s->E_th[X]) * EV_2_ERG;
    heat3[X] = orc_integrate_trap(s->E, yy, nn);
  }
}

/* ---------------------------------------------------------------------- */
/* bundle of the 17 output arrays                                          */
typedef struct {
  double *xH1, *xH2, *xHe1, *xHe2, *xHe3;
  double *ion_src[3], *ion_rec[3], *heat_src[3], *heat_rec[3];
} out_t;

static void out_reverse(out_t *o, int Nl) {
  rev(o->xH1, Nl); rev(o->xH2, Nl); rev(o->xHe1, Nl); rev(o->xHe2, Nl);
  rev(o->xHe3, Nl);
  for (int X = 0; X < 3; X++) {
    rev(o->ion_src[X], Nl); rev(o->ion_rec[X], Nl);
    rev(o->heat_src[X], Nl); rev(o->heat_rec[X], Nl);
  }
}

/* cell solve used at the end of every sweep body:
 * sphere_bgnd.f90:886-907 (= sphere_stromgren.f90:847-868 = slab_bgnd.f90:814-837
 * = slab_plane.f90:752-776) */
static int cell_solve(int i_find_Teq, double nH, double nHe, const double ion[3],
                      const double heat[3], double z, double Hz, double fcA,
                      double tol, double *xH1, double *xH2, double *xHe1,
                      double *xHe2, double *xHe3, double *T) {
  if (i_find_Teq == 1) {
    return orc_pcte1(nH, nHe, ion[0], ion[1], ion[2], heat[0], heat[1], heat[2],
                     z, Hz, fcA, fcA, fcA, tol * TOL_FRAC, xH1, xH2, xHe1, xHe2,
                     xHe3, T);
  } else {
    double reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2;
    orc_kchem1(*T, fcA, fcA, fcA, &reH2, &reHe2, &reHe3, &ciH1, &ciHe1, &ciHe2);
    return orc_pce1(nH, nHe, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2, ion[0],
                    ion[1], ion[2], tol * TOL_FRAC, xH1, xH2, xHe1, xHe2, xHe3);
  }
}

/* the optically-thin initial state with the lock-step vector solvers:
 * sphere_bgnd.f90:379-499, sphere_stromgren.f90:385-516, slab_bgnd.f90:346-460,
 * slab_plane.f90:346-470.  `r_c` is only used for point sources.             */
static int thin_state(const spec_t *s, const double *r_c, const double *nH,
                      const double *nHe, double *T, double fcA, int i_find_Teq,
                      double z, double Hz, double tol, int Nl, out_t *o) {
  size_t n = (size_t)Nl;
  double *buf = (double *)malloc(sizeof(double) * (7 * n + (size_t)s->Nnu));
  double *fc = buf, *reH2 = buf + n, *reHe2 = buf + 2 * n, *reHe3 = buf + 3 * n;
  double *ciH1 = buf + 4 * n, *ciHe1 = buf + 5 * n, *ciHe2 = buf + 6 * n;
  double *yy = buf + 7 * n;
  int rc;
  for (int i = 0; i < Nl; i++) fc[i] = fcA;
  orc_get_kchem(T, fc, fc, fc, Nl, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2);

  if (s->src == SRC_POINT) {
    for (int i = 0; i < Nl; i++) {
      double ion[3], heat[3];
      rates_thin(s, r_c[i], ion, heat, yy);
      for (int X = 0; X < 3; X++) {
        o->ion_src[X][i] = ion[X];
        o->heat_src[X][i] = heat[X];
      }
    }
  } else {
    double ion[3], heat[3];
    rates_thin(s, 0.0, ion, heat, yy);
    for (int X = 0; X < 3; X++)
      for (int i = 0; i < Nl; i++) {
        o->ion_src[X][i] = ion[X];
        o->heat_src[X][i] = heat[X];
      }
  }
  if (i_find_Teq == 1) {
    rc = orc_solve_pcte_v(nH, nHe, o->ion_src[0], o->ion_src[1], o->ion_src[2],
                          o->heat_src[0], o->heat_src[1], o->heat_src[2], z, Hz,
                          fc, fc, fc, tol * TOL_FRAC, Nl, o->xH1, o->xH2,
                          o->xHe1, o->xHe2, o->xHe3, T);
  } else {
    rc = orc_solve_pce_v(nH, nHe, reH2, reHe2, reHe3, ciH1, ciHe1, ciHe2,
                         o->ion_src[0], o->ion_src[1], o->ion_src[2],
                         tol * TOL_FRAC, Nl, o->xH1, o->xH2, o->xHe1, o->xHe2,
                         o->xHe3);
  }
  for (int X = 0; X < 3; X++)
    for (int i = 0; i < Nl; i++) {
      o->ion_rec[X][i] = 0.0;
      o->heat_rec[X][i] = 0.0;
    }
  free(buf);
  return rc;
}

/* convergence test shared by the four drivers (sphere_bgnd.f90:289-296) */
static int ne_converged(const double *nH, const double *nHe, const out_t *o,
                        double *conv_old, double tol, int Nl) {
  int all = 1;
  for (int i = 0; i < Nl; i++) {
    double ne = o->xH2[i] * nH[i] + (o->xHe2[i] + 2.0 * o->xHe3[i]) * nHe[i];
    double change = ne / conv_old[i] - 1.0;
    if (!(fabs(change) < tol)) all = 0;
    conv_old[i] = ne;
  }
  return all;
}

static void add_counters(long long *counters, const cnt_t *c) {
  if (!counters) return;
  counters[0] += c->seg;
  counters[1] += c->sq;
  counters[2] += c->ex;
  counters[3] += c->cell;
}

/* ---------------------------------------------------------------------- */
/* sphere_bgnd.f90:578-918 (sphere_bgnd_sweep), rec_meth fixed.            */
/* recombination radiation at the current state, before the cell loop:
 * sphere_bgnd.f90:686-735, sphere_stromgren.f90:713-757 (i_rec_meth 2,3,4),
 * slab_bgnd.f90:653-672, slab_plane.f90:657-676,974-993 (i_rec_meth 2).
 * geom: 0 sphere with decreasing Edges, 1 sphere with increasing Edges, 2 slab */
static int rec_photons(int geom, int i_rec_meth, const spec_t *s,
                       const double *Edges, const double *nH, const double *nHe,
                       const double *T, int Nmu, const double em_fac[3], int Nl,
                       out_t *o) {
  if (i_rec_meth == 1) {
    for (int X = 0; X < 3; X++)
      for (int i = 0; i < Nl; i++) {
        o->ion_rec[X][i] = 0.0;
        o->heat_rec[X][i] = 0.0;
      }
    return ORC_OK;
  }
  if (geom == 2) {
    orc_rec_slab(o->xH1, o->xH2, o->xHe1, o->xHe2, o->xHe3, Edges, nH, nHe, T,
                 s->sigma_th, s->E_th, Nl, o->ion_rec, o->heat_rec);
    return ORC_OK;
  }
  const int dec = (geom == 0);
  if (i_rec_meth == 2)
    orc_rec_outward(o->xH1, o->xH2, o->xHe1, o->xHe2, o->xHe3, Edges, nH, nHe, T,
                    dec, s->sigma_th, s->E_th, Nl, o->ion_rec, o->heat_rec);
  else if (i_rec_meth == 3)
    orc_rec_radial(o->xH1, o->xH2, o->xHe1, o->xHe2, o->xHe3, Edges, nH, nHe, T,
                   dec, s->sigma_th, s->E_th, Nl, o->ion_rec, o->heat_rec);
  else
    return orc_rec_isotropic(o->xH1, o->xH2, o->xHe1, o->xHe2, o->xHe3, Edges, nH,
                             nHe, T, dec, Nmu, s->sigma_th, s->E_th, em_fac, Nl,
                             o->ion_rec, o->heat_rec);
  return ORC_OK;
}

static int sphere_bgnd_sweep(const spec_t *s, const double *Edges,
                             const double *nH, const double *nHe, double *T,
                             int Nmu, double fcA, int i_find_Teq, double z,
                             double Hz, double tol, int Nl, int order,
                             int flavour, int i_rec_meth, const double em_fac[3],
                             out_t *o, cnt_t *cnt) {
  int Nnu = s->Nnu;
  {
    int rrc = rec_photons(0, i_rec_meth, s, Edges, nH, nHe, T, Nmu, em_fac, Nl, o);
    if (rrc) return rrc;
  }
  double *xi = (double *)malloc(sizeof(double) * 2 * (size_t)Nmu);
  double *wi = xi + Nmu;
  orc_p_quadrature_rule(Nmu, xi, wi); /* :743 */

  /* Jacobi: rays read the pre-sweep state (what a cell-sharded all-gather
   * gives).  GS: rays read the live arrays, outer -> inner, one thread. */
  const double *rxH1 = o->xH1, *rxHe1 = o->xHe1, *rxHe2 = o->xHe2;
  double *frozen = NULL;
  if (order == ORC_ORDER_JACOBI) {
    frozen = (double *)malloc(sizeof(double) * 3 * (size_t)Nl);
    memcpy(frozen, o->xH1, sizeof(double) * (size_t)Nl);
    memcpy(frozen + Nl, o->xHe1, sizeof(double) * (size_t)Nl);
    memcpy(frozen + 2 * Nl, o->xHe2, sizeof(double) * (size_t)Nl);
    rxH1 = frozen;
    rxHe1 = frozen + Nl;
    rxHe2 = frozen + 2 * Nl;
  }
  int rc_all = ORC_OK;
  long long seg = 0, sq = 0, ex = 0, cell = 0;

#pragma omp parallel if (order == ORC_ORDER_JACOBI) reduction(+ : seg, sq, ex, cell)
  {
    double *scr = (double *)malloc(sizeof(double) * (3 * (size_t)Nnu + 6 * (size_t)Nmu));
    double *mu_rates = scr + 3 * Nnu; /* [6][Nmu] */
    cnt_t c = {0, 0, 0, 0};
#pragma omp for schedule(dynamic, 1)
    for (int il = 0; il < Nl; il++) {
      int rc = ORC_OK;
      for (int imu = 0; imu < Nmu; imu++) { /* :764-832 */
        double mu = xi[imu];
        double NH1, NHe1, NHe2;
        rc = ray_columns(Edges, nH, nHe, rxH1, rxHe1, rxHe2, il, mu, Nl,
                         flavour, &NH1, &NHe1, &NHe2, &c);
        if (rc) break;
        double tau_th[3];
        tau_th[0] = NH1 * s->sigma_th[0]; /* :794-796 */
        tau_th[1] = NHe1 * s->sigma_th[1];
        tau_th[2] = NHe2 * s->sigma_th[2];
        double ion[3], heat[3];
        rates_shld(s, tau_th, 0.0, ATT_EXP, flavour, ion, heat, scr, &c);
        for (int X = 0; X < 3; X++) {
          mu_rates[X * Nmu + imu] = ion[X];
          mu_rates[(3 + X) * Nmu + imu] = heat[X];
        }
      }
      if (rc) {
#pragma omp critical
        rc_all = rc;
        continue;
      }
      /* :837-856  accumulate onto the previous value, then halve */
      double ion[3], heat[3];
      for (int X = 0; X < 3; X++) {
        double a = o->ion_src[X][il], b = o->heat_src[X][il];
        for (int imu = 0; imu < Nmu; imu++) {
          a = a + mu_rates[X * Nmu + imu] * wi[imu];
          b = b + mu_rates[(3 + X) * Nmu + imu] * wi[imu];
        }
        o->ion_src[X][il] = a * 0.5;
        o->heat_src[X][il] = b * 0.5;
        /* :861-881 (i_rec_meth == 1: the rec arrays are zero) */
        ion[X] = o->ion_src[X][il] + o->ion_rec[X][il];
        heat[X] = o->heat_src[X][il] + o->heat_rec[X][il];
        if (i_rec_meth == 1) {
          ion[X] = o->ion_src[X][il];
          heat[X] = o->heat_src[X][il];
        }
      }
      rc = cell_solve(i_find_Teq, nH[il], nHe[il], ion, heat, z, Hz, fcA, tol,
                      &o->xH1[il], &o->xH2[il], &o->xHe1[il], &o->xHe2[il],
                      &o->xHe3[il], &T[il]);
      c.cell += 1;
      if (rc) {
#pragma omp critical
        rc_all = rc;
      }
    }
    seg += c.seg; sq += c.sq; ex += c.ex; cell += c.cell;
    free(scr);
  }
  if (cnt) {
    cnt->seg += seg; cnt->sq += sq; cnt->ex += ex; cnt->cell += cell;
  }
  free(frozen);
  free(xi);
  return rc_all;
}

static void bind_out(out_t *o, double *xH1, double *xH2, double *xHe1,
                     double *xHe2, double *xHe3, double *H1i_src,
                     double *He1i_src, double *He2i_src, double *H1i_rec,
                     double *He1i_rec, double *He2i_rec, double *H1h_src,
                     double *He1h_src, double *He2h_src, double *H1h_rec,
                     double *He1h_rec, double *He2h_rec) {
  o->xH1 = xH1; o->xH2 = xH2; o->xHe1 = xHe1; o->xHe2 = xHe2; o->xHe3 = xHe3;
  o->ion_src[0] = H1i_src; o->ion_src[1] = He1i_src; o->ion_src[2] = He2i_src;
  o->ion_rec[0] = H1i_rec; o->ion_rec[1] = He1i_rec; o->ion_rec[2] = He2i_rec;
  o->heat_src[0] = H1h_src; o->heat_src[1] = He1h_src; o->heat_src[2] = He2h_src;
  o->heat_rec[0] = H1h_rec; o->heat_rec[1] = He1h_rec; o->heat_rec[2] = He2h_rec;
}

/* sphere_bgnd.f90:107-314 (sphere_bgnd_solve) */
int orc_sphere_bgnd_solve(
    double *Edges, double *nH, double *nHe, double *T, int Nmu,
    const double *E_eV, const double *shape, int i_rec_meth, double fixed_fcA,
    int i_photo_fit, int i_rate_fit, int i_find_Teq, int i_thin,
    double em_H1_fac, double em_He1_fac, double em_He2_fac, double z, double Hz,
    double tol, int Nl, int Nnu, double *xH1, double *xH2, double *xHe1,
    double *xHe2, double *xHe3, double *H1i_src, double *He1i_src,
    double *He2i_src, double *H1i_rec, double *He1i_rec, double *He2i_rec,
    double *H1h_src, double *He1h_src, double *He2h_src, double *H1h_rec,
    double *He1h_rec, double *He2h_rec, const orc_opts *opts, int *iters,
    long long *counters) {
  const double em_fac[3] = {em_H1_fac, em_He1_fac, em_He2_fac};
  if (i_rec_meth < 1 || i_rec_meth > 4 || i_photo_fit != 1 || i_rate_fit != 1 ||
      Nl < 1 || Nnu < 1 || Nmu < 1)
    return ORC_ERR_BAD_ARG;
  int order = opts ? opts->order : ORC_ORDER_GS;
  int flavour = opts ? opts->flavour : ORC_FLAVOUR_HOISTED;
  int max_sweeps = opts ? opts->max_sweeps : 0;
  out_t o;
  bind_out(&o, xH1, xH2, xHe1, xHe2, xHe3, H1i_src, He1i_src, He2i_src, H1i_rec,
           He1i_rec, He2i_rec, H1h_src, He1h_src, He2h_src, H1h_rec, He1h_rec,
           He2h_rec);
  if (iters) *iters = 0;

  int reversed = 0; /* :182-202 */
  if (monotonic_increasing(Edges, Nl + 1)) {
    reversed = 1;
    rev(Edges, Nl + 1); rev(nH, Nl); rev(nHe, Nl); rev(T, Nl);
    out_reverse(&o, Nl);
  } else if (!monotonic_decreasing(Edges, Nl + 1)) {
    return ORC_ERR_EDGES_NOT_MONOTONIC;
  }
  spec_t s;
  spec_init(&s, SRC_BGND, E_eV, shape, Nnu);
  double fcA = (i_rec_meth == 1) ? fixed_fcA : 1.0; /* :435-443, :676-684 */
  cnt_t cnt = {0, 0, 0, 0};

  int rc = thin_state(&s, NULL, nH, nHe, T, fcA, i_find_Teq, z, Hz, tol, Nl, &o);
  if (rc == ORC_OK && i_thin != 1) {
    double *conv_old = (double *)malloc(sizeof(double) * (size_t)Nl);
    for (int i = 0; i < Nl; i++) /* :260-262 */
      conv_old[i] = xH2[i] * nH[i] + (xHe2[i] + 2.0 * xHe3[i]) * nHe[i];
    int iter = 0;
    int not_converged = 1;
    const double t_sweeps0 = wall_now();
    g_sweep_seconds = 0.0;
    while (not_converged) {
      rc = sphere_bgnd_sweep(&s, Edges, nH, nHe, T, Nmu, fcA, i_find_Teq, z, Hz,
                             tol, Nl, order, flavour, i_rec_meth, em_fac, &o, &cnt);
      g_sweep_seconds = wall_now() - t_sweeps0;
      if (rc) break;
      if (iter > OUTER_MAX_ITER) { /* :284-287 */
        rc = ORC_ERR_MAX_ITER_OUTER;
        break;
      }
      iter = iter + 1;
      if (ne_converged(nH, nHe, &o, conv_old, tol, Nl)) not_converged = 0;
      if (max_sweeps > 0 && iter >= max_sweeps) not_converged = 0;
    }
    if (iters) *iters = iter;
    free(conv_old);
  }
  add_counters(counters, &cnt);
  spec_free(&s);
  if (reversed) { /* :303-311 (also on the i_thin path :243-251) */
    rev(Edges, Nl + 1); rev(nH, Nl); rev(nHe, Nl); rev(T, Nl);
    out_reverse(&o, Nl);
  }
  return rc;
}

/* ---------------------------------------------------------------------- */
/* threshold optical depths below/above a layer:
 * slab_base.f90:63-115 and sphere_base.f90:1232-1367 (for increasing Edges the
 * sphere version reduces to the slab one: lo = sum(0..inl-1), hi = sum(inl+1..)).
 * Sequential sums in index order, like Fortran sum().                       */
static void optical_depths(const double *dtau, int inl, int Nl, double *lo,
                           double *hi) {
  double a = 0.0, b = 0.0;
  for (int k = 0; k < inl; k++) a += dtau[k];
  for (int k = inl + 1; k < Nl; k++) b += dtau[k];
  *lo = a;
  *hi = b;
}

/* A Jacobi sweep for the three "column" geometries:
 *   kind 0: SphereStromgren   sphere_stromgren.f90:595-878  (point, exp, 1 dir)
 *   kind 1: SlabPln one-sided slab_plane.f90:551-782        (plane, exp, 1 dir)
 *   kind 2: Slab2Pln          slab_plane.f90:856-1164       (plane, exp, lo+hi, mirror)
 *   kind 3: Slab2Bgnd         slab_bgnd.f90:532-883         (bgnd, E2, lo+hi, mirror)  */
static int column_sweep(int kind, const spec_t *s, const double *Edges,
                        const double *nH, const double *nHe, double *T,
                        double fcA, int i_find_Teq, double z, double Hz,
                        double tol, int Nl, int flavour, int i_rec_meth, int Nmu,
                        const double em_fac[3], out_t *o, cnt_t *cnt) {
  int Nnu = s->Nnu;
  {
    int rrc = rec_photons(kind == 0 ? 1 : 2, i_rec_meth, s, Edges, nH, nHe, T, Nmu,
                          em_fac, Nl, o);
    if (rrc) return rrc;
  }
  size_t n = (size_t)Nl;
  double *buf = (double *)malloc(sizeof(double) * 5 * n);
  double *r_c = buf, *dtau[3] = {buf + n, buf + 2 * n, buf + 3 * n};
  double *dl = buf + 4 * n;
  for (int i = 0; i < Nl; i++) {
    if (kind == 0) { /* get_dr_and_r_c: sphere_base.f90:1527-1537 */
      dl[i] = fabs(Edges[i] - Edges[i + 1]);
      r_c[i] = (Edges[i] + Edges[i + 1]) * 0.5;
    } else { /* slab_bgnd.f90:612-613 */
      dl[i] = Edges[i + 1] - Edges[i];
      r_c[i] = Edges[i] + dl[i] * 0.5;
    }
    double dNH = dl[i] * nH[i], dNHe = dl[i] * nHe[i];
    dtau[0][i] = dNH * o->xH1[i] * s->sigma_th[0]; /* frozen: Jacobi (Q3) */
    dtau[1][i] = dNHe * o->xHe1[i] * s->sigma_th[1];
    dtau[2][i] = dNHe * o->xHe2[i] * s->sigma_th[2];
  }
  int two_sided = (kind >= 2);
  int att = (kind == 3) ? ATT_E2 : ATT_EXP;
  int nloop = two_sided ? Nl / 2 : Nl;
  int rc_all = ORC_OK;
  long long ex = 0, cell = 0;

#pragma omp parallel reduction(+ : ex, cell)
  {
    double *scr = (double *)malloc(sizeof(double) * 3 * (size_t)Nnu);
    cnt_t c = {0, 0, 0, 0};
#pragma omp for schedule(dynamic, 1)
    for (int inl = 0; inl < nloop; inl++) {
      double lo[3], hi[3], tau_th[3], ion[3], heat[3];
      for (int X = 0; X < 3; X++)
        optical_depths(dtau[X], inl, Nl, &lo[X], &hi[X]);
      for (int X = 0; X < 3; X++) tau_th[X] = lo[X] + 0.5 * dtau[X][inl];
      rates_shld(s, tau_th, r_c[inl], att, flavour, ion, heat, scr, &c);
      if (two_sided) {
        double ion_hi[3], heat_hi[3];
        for (int X = 0; X < 3; X++) tau_th[X] = hi[X] + 0.5 * dtau[X][inl];
        rates_shld(s, tau_th, r_c[inl], att, flavour, ion_hi, heat_hi, scr, &c);
        for (int X = 0; X < 3; X++) { /* X_lo*half + X_hi*half */
          ion[X] = ion[X] * 0.5 + ion_hi[X] * 0.5;
          heat[X] = heat[X] * 0.5 + heat_hi[X] * 0.5;
        }
      }
      for (int X = 0; X < 3; X++) {
        o->ion_src[X][inl] = ion[X];
        o->heat_src[X][inl] = heat[X];
        if (i_rec_meth != 1) { /* totals: slab_bgnd.f90:752-760,800-808 and twins */
          ion[X] = o->ion_src[X][inl] + o->ion_rec[X][inl];
          heat[X] = o->heat_src[X][inl] + o->heat_rec[X][inl];
        }
      }
      int rc = cell_solve(i_find_Teq, nH[inl], nHe[inl], ion, heat, z, Hz, fcA,
                          tol, &o->xH1[inl], &o->xH2[inl], &o->xHe1[inl],
                          &o->xHe2[inl], &o->xHe3[inl], &T[inl]);
      c.cell += 1;
      if (rc) {
#pragma omp critical
        rc_all = rc;
      }
      if (two_sided) { /* mirror: slab_bgnd.f90:842-865 */
        int imir = Nl - 1 - inl;
        o->xH1[imir] = o->xH1[inl]; o->xH2[imir] = o->xH2[inl];
        o->xHe1[imir] = o->xHe1[inl]; o->xHe2[imir] = o->xHe2[inl];
        o->xHe3[imir] = o->xHe3[inl];
        T[imir] = T[inl];
        for (int X = 0; X < 3; X++) {
          o->ion_src[X][imir] = o->ion_src[X][inl];
          o->ion_rec[X][imir] = o->ion_rec[X][inl];
          o->heat_src[X][imir] = o->heat_src[X][inl];
          o->heat_rec[X][imir] = o->heat_rec[X][inl];
        }
      }
    }
    ex += c.ex; cell += c.cell;
    free(scr);
  }
  if (cnt) {
    cnt->ex += ex; cnt->cell += cell;
  }
  free(buf);
  return rc_all;
}

/* sphere_stromgren.f90:111-320 (sphere_stromgren_solve) */
int orc_sphere_stromgren_solve(
    double *Edges, double *nH, double *nHe, double *T, int Nmu,
    const double *E_eV, const double *shape, int i_rec_meth, double fixed_fcA,
    int i_photo_fit, int i_rate_fit, int i_find_Teq, int i_thin,
    double em_H1_fac, double em_He1_fac, double em_He2_fac, double z, double Hz,
    double tol, int Nl, int Nnu, double *xH1, double *xH2, double *xHe1,
    double *xHe2, double *xHe3, double *H1i_src, double *He1i_src,
    double *He2i_src, double *H1i_rec, double *He1i_rec, double *He2i_rec,
    double *H1h_src, double *He1h_src, double *He2h_src, double *H1h_rec,
    double *He1h_rec, double *He2h_rec, const orc_opts *opts, int *iters,
    long long *counters) {
  const double em_fac[3] = {em_H1_fac, em_He1_fac, em_He2_fac};
  if (i_rec_meth < 1 || i_rec_meth > 4 || i_photo_fit != 1 || i_rate_fit != 1 ||
      Nl < 1 || Nnu < 1 || (i_rec_meth == 4 && Nmu < 1))
    return ORC_ERR_BAD_ARG;
  int flavour = opts ? opts->flavour : ORC_FLAVOUR_HOISTED;
  int max_sweeps = opts ? opts->max_sweeps : 0;
  out_t o;
  bind_out(&o, xH1, xH2, xHe1, xHe2, xHe3, H1i_src, He1i_src, He2i_src, H1i_rec,
           He1i_rec, He2i_rec, H1h_src, He1h_src, He2h_src, H1h_rec, He1h_rec,
           He2h_rec);
  if (iters) *iters = 0;
  int reversed = 0; /* :186-204: this solver wants INCREASING Edges */
  if (monotonic_decreasing(Edges, Nl + 1)) {
    reversed = 1;
    rev(Edges, Nl + 1); rev(nH, Nl); rev(nHe, Nl); rev(T, Nl);
    out_reverse(&o, Nl);
  } else if (!monotonic_increasing(Edges, Nl + 1)) {
    return ORC_ERR_EDGES_NOT_MONOTONIC;
  }
  spec_t s;
  spec_init(&s, SRC_POINT, E_eV, shape, Nnu);
  double fcA = (i_rec_meth == 1) ? fixed_fcA : 1.0; /* sphere_stromgren.f90:450-458 */
  cnt_t cnt = {0, 0, 0, 0};
  double *r_c = (double *)malloc(sizeof(double) * (size_t)Nl);
  for (int i = 0; i < Nl; i++) r_c[i] = (Edges[i] + Edges[i + 1]) * 0.5;

  int rc = thin_state(&s, r_c, nH, nHe, T, fcA, i_find_Teq, z, Hz, tol, Nl, &o);
  if (rc == ORC_OK && i_thin != 1) {
    double *conv_old = (double *)malloc(sizeof(double) * (size_t)Nl);
    for (int i = 0; i < Nl; i++)
      conv_old[i] = xH2[i] * nH[i] + (xHe2[i] + 2.0 * xHe3[i]) * nHe[i];
    int iter = 0, not_converged = 1;
    while (not_converged) {
      rc = column_sweep(0, &s, Edges, nH, nHe, T, fcA, i_find_Teq, z, Hz, tol,
                        Nl, flavour, i_rec_meth, Nmu, em_fac, &o, &cnt);
      if (rc) break;
      if (iter > OUTER_MAX_ITER) {
        rc = ORC_ERR_MAX_ITER_OUTER;
        break;
      }
      iter = iter + 1;
      if (ne_converged(nH, nHe, &o, conv_old, tol, Nl)) not_converged = 0;
      if (max_sweeps > 0 && iter >= max_sweeps) not_converged = 0;
    }
    if (iters) *iters = iter;
    free(conv_old);
  }
  add_counters(counters, &cnt);
  free(r_c);
  spec_free(&s);
  if (reversed) {
    rev(Edges, Nl + 1); rev(nH, Nl); rev(nHe, Nl); rev(T, Nl);
    out_reverse(&o, Nl);
  }
  return rc;
}

/* slab_bgnd.f90:97-288 (slab_bgnd_solve): note the silent stop at 200 sweeps */
int orc_slab_bgnd_solve(
    const double *Edges, const double *nH, const double *nHe, double *T,
    const double *E_eV, const double *shape, int i_rec_meth, double fixed_fcA,
    int i_photo_fit, int i_rate_fit, int i_find_Teq, int i_thin, double z,
    double Hz, double tol, int Nl, int Nnu, double *xH1, double *xH2,
    double *xHe1, double *xHe2, double *xHe3, double *H1i_src, double *He1i_src,
    double *He2i_src, double *H1i_rec, double *He1i_rec, double *He2i_rec,
    double *H1h_src, double *He1h_src, double *He2h_src, double *H1h_rec,
    double *He1h_rec, double *He2h_rec, const orc_opts *opts, int *iters,
    long long *counters) {
  const double em_fac[3] = {1.0, 1.0, 1.0};
  if (i_rec_meth < 1 || i_rec_meth > 2 || i_photo_fit != 1 || i_rate_fit != 1 ||
      Nl < 1 || Nnu < 1)
    return ORC_ERR_BAD_ARG;
  int flavour = opts ? opts->flavour : ORC_FLAVOUR_HOISTED;
  int max_sweeps = opts ? opts->max_sweeps : 0;
  out_t o;
  bind_out(&o, xH1, xH2, xHe1, xHe2, xHe3, H1i_src, He1i_src, He2i_src, H1i_rec,
           He1i_rec, He2i_rec, H1h_src, He1h_src, He2h_src, H1h_rec, He1h_rec,
           He2h_rec);
  if (iters) *iters = 0;
  spec_t s;
  spec_init(&s, SRC_BGND, E_eV, shape, Nnu);
  double fcA = (i_rec_meth == 1) ? fixed_fcA : 1.0; /* slab_bgnd.f90:411-419 */
  cnt_t cnt = {0, 0, 0, 0};
  int rc = thin_state(&s, NULL, nH, nHe, T, fcA, i_find_Teq, z, Hz, tol, Nl, &o);
  if (rc == ORC_OK && i_thin != 1) {
    double *conv_old = (double *)malloc(sizeof(double) * (size_t)Nl);
    for (int i = 0; i < Nl; i++)
      conv_old[i] = xH2[i] * nH[i] + (xHe2[i] + 2.0 * xHe3[i]) * nHe[i];
    int iter = 0, not_converged = 1;
    while (not_converged) {
      rc = column_sweep(3, &s, Edges, nH, nHe, T, fcA, i_find_Teq, z, Hz, tol,
                        Nl, flavour, i_rec_meth, 0, em_fac, &o, &cnt);
      if (rc) break;
      iter = iter + 1;
      if (ne_converged(nH, nHe, &o, conv_old, tol, Nl)) not_converged = 0;
      if (iter >= 200) not_converged = 0; /* :278-282 */
      if (max_sweeps > 0 && iter >= max_sweeps) not_converged = 0;
    }
    if (iters) *iters = iter;
    free(conv_old);
  }
  add_counters(counters, &cnt);
  spec_free(&s);
  return rc;
}

/* slab_plane.f90:105-283 (slab_plane_solve) */
int orc_slab_plane_solve(
    const double *Edges, const double *nH, const double *nHe, double *T,
    const double *E_eV, const double *shape, int i_2side, int i_rec_meth,
    double fixed_fcA, int i_photo_fit, int i_rate_fit, int i_find_Teq,
    int i_thin, double z, double Hz, double tol, int Nl, int Nnu, double *xH1,
    double *xH2, double *xHe1, double *xHe2, double *xHe3, double *H1i_src,
    double *He1i_src, double *He2i_src, double *H1i_rec, double *He1i_rec,
    double *He2i_rec, double *H1h_src, double *He1h_src, double *He2h_src,
    double *H1h_rec, double *He1h_rec, double *He2h_rec, const orc_opts *opts,
    int *iters, long long *counters) {
  const double em_fac[3] = {1.0, 1.0, 1.0};
  if (i_rec_meth < 1 || i_rec_meth > 2 || i_photo_fit != 1 || i_rate_fit != 1 ||
      Nl < 1 || Nnu < 1)
    return ORC_ERR_BAD_ARG;
  int flavour = opts ? opts->flavour : ORC_FLAVOUR_HOISTED;
  int max_sweeps = opts ? opts->max_sweeps : 0;
  out_t o;
  bind_out(&o, xH1, xH2, xHe1, xHe2, xHe3, H1i_src, He1i_src, He2i_src, H1i_rec,
           He1i_rec, He2i_rec, H1h_src, He1h_src, He2h_src, H1h_rec, He1h_rec,
           He2h_rec);
  if (iters) *iters = 0;
  spec_t s;
  spec_init(&s, SRC_PLANE, E_eV, shape, Nnu);
  double fcA = (i_rec_meth == 1) ? fixed_fcA : 1.0; /* slab_plane.f90:411-419 */
  cnt_t cnt = {0, 0, 0, 0};
  int rc = thin_state(&s, NULL, nH, nHe, T, fcA, i_find_Teq, z, Hz, tol, Nl, &o);
  if (rc == ORC_OK && i_thin != 1) {
    double *conv_old = (double *)malloc(sizeof(double) * (size_t)Nl);
    for (int i = 0; i < Nl; i++)
      conv_old[i] = xH2[i] * nH[i] + (xHe2[i] + 2.0 * xHe3[i]) * nHe[i];
    int iter = 0, not_converged = 1;
    while (not_converged) {
      rc = column_sweep(i_2side == 1 ? 2 : 1, &s, Edges, nH, nHe, T, fcA,
                        i_find_Teq, z, Hz, tol, Nl, flavour, i_rec_meth, 0, em_fac,
                        &o, &cnt);
      if (rc) break;
      if (iter > OUTER_MAX_ITER) { /* :265-268 */
        rc = ORC_ERR_MAX_ITER_OUTER;
        break;
      }
      iter = iter + 1;
      if (ne_converged(nH, nHe, &o, conv_old, tol, Nl)) not_converged = 0;
      if (max_sweeps > 0 && iter >= max_sweeps) not_converged = 0;
    }
    if (iters) *iters = iter;
    free(conv_old);
  }
  add_counters(counters, &cnt);
  spec_free(&s);
  return rc;
}

/* sphere_base.f90:1115-1205 (set_column_vs_impact): mu = 0 ray per shell,
 * full chord j = 0 .. 2m */
int orc_set_column_vs_impact(double *xH1, double *xH2, double *xHe1,
                             double *xHe2, double *xHe3, double *Edges,
                             double *nH, double *nHe, double *T, int Nl,
                             double *NH, double *NHe, double *NH1, double *NHe1,
                             double *NHe2) {
  int reversed = 0;
  if (monotonic_increasing(Edges, Nl + 1)) {
    reversed = 1;
    rev(xH1, Nl); rev(xH2, Nl); rev(xHe1, Nl); rev(xHe2, Nl); rev(xHe3, Nl);
    rev(Edges, Nl + 1); rev(nH, Nl); rev(nHe, Nl); rev(T, Nl);
  }
  int rc = ORC_OK;
  if (!monotonic_decreasing(Edges, Nl + 1)) {
    rc = ORC_ERR_EDGES_NOT_MONOTONIC;
  } else {
    for (int il = 0; il < Nl; il++) {
      double mu = 0.0;
      int m = orc_m_for_ray(Edges, il, mu, Nl);
      if (m < 0) {
        rc = ORC_ERR_RAY_GEOMETRY;
        break;
      }
      double r_ray = (Edges[il] + Edges[il + 1]) * 0.5;
      double p_ray = r_ray * sqrt(1.0 - mu * mu);
      double a = 0.0, b = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
      for (int j = 0; j <= 2 * m; j++) {
        int jnl = m - abs(m - j);
        double ds = orc_ds_hoisted(Edges, p_ray, m, j);
        a = a + ds * nH[jnl];
        b = b + ds * nHe[jnl];
        c1 = c1 + ds * nH[jnl] * xH1[jnl];
        c2 = c2 + ds * nHe[jnl] * xHe1[jnl];
        c3 = c3 + ds * nHe[jnl] * xHe2[jnl];
      }
      NH[il] = a; NHe[il] = b; NH1[il] = c1; NHe1[il] = c2; NHe2[il] = c3;
    }
  }
  if (reversed) {
    rev(xH1, Nl); rev(xH2, Nl); rev(xHe1, Nl); rev(xHe2, Nl); rev(xHe3, Nl);
    rev(Edges, Nl + 1); rev(nH, Nl); rev(nHe, Nl); rev(T, Nl);
    rev(NH, Nl); rev(NHe, Nl); rev(NH1, Nl); rev(NHe1, Nl); rev(NHe2, Nl);
  }
  return rc;
}
