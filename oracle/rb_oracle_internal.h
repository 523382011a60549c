/* rb_oracle_internal.h -- shared declarations inside the oracle (TEST
 * INFRASTRUCTURE ONLY, see rb_oracle.h). */
#ifndef RB_ORACLE_INTERNAL_H
#define RB_ORACLE_INTERNAL_H

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "rb_oracle.h"

/* rabacus/f2py/physical_constants.f90:7-21 (values copied as literals; note
 * eV_2_erg is the older CODATA value -- do not "fix") */
#define PI_ 3.141592653589793e+00
#define KB_ERG_K 1.380648800000000e-16
#define C_CM_S 2.997924580000000e+10
#define H_EV_S 4.135667603369524e-15
#define H_ERG_S 6.626069570000000e-27
#define EV_2_ERG 1.602176530000000e-12
#define ERG_2_EV 6.241509479607719e+11

/* ion_solver.f90:13-15 */
#define ION_MAX_ITER 5000
#define TFLOOR 7.5e3
#define TMAX 1.0e5
/* sphere_bgnd.f90:45-46, sphere_stromgren.f90:44-45, slab_plane.f90:37-38 */
#define OUTER_MAX_ITER 500
#define TOL_FRAC 1.0e-2

typedef struct {
  double recH2, recHe2, recHe3, cicH1, cicHe1, cicHe2, cecH1, cecHe2, bremss,
      compton;
} orc_kcool;

void orc_kchem1(double T, double fcA_H2, double fcA_He2, double fcA_He3,
                double *reH2, double *reHe2, double *reHe3, double *ciH1,
                double *ciHe1, double *ciHe2);
void orc_kcool1(double T, double fcA_H2, double fcA_He2, double fcA_He3,
                orc_kcool *k);
double orc_heat1(double nH, double nHe, double H1h, double He1h, double He2h,
                 double xH1, double xHe1, double xHe2);
double orc_cool1(double nH, double nHe, double T, const orc_kcool *k,
                 double xH1, double xH2, double xHe1, double xHe2, double xHe3,
                 double z, double Hz);

/* scalar single-cell solvers used by the sweeps */
int orc_pce1(double nH, double nHe, double reH2, double reHe2, double reHe3,
             double ciH1, double ciHe1, double ciHe2, double H1i, double He1i,
             double He2i, double tol, double *xH1, double *xH2, double *xHe1,
             double *xHe2, double *xHe3);
int orc_pcte1(double nH, double nHe, double H1i, double He1i, double He2i,
              double H1h, double He1h, double He2h, double z, double Hz,
              double fcA_H2, double fcA_He2, double fcA_He3, double tol,
              double *xH1, double *xH2, double *xHe1, double *xHe2,
              double *xHe3, double *T);


/* geometry.f90:146-165 with m and p hoisted (rb_oracle_geom.c) */
double orc_ds_hoisted(const double *Edges, double p_ray, int m, int j);

/* recombination-photon transfer (rb_oracle_rec.c) */
void orc_rec_outward(const double *xH1, const double *xH2, const double *xHe1,
                     const double *xHe2, const double *xHe3, const double *Edges,
                     const double *nH, const double *nHe, const double *T,
                     int have_dec, const double sigma_th[3], const double E_th[3],
                     int Nl, double *ion[3], double *heat[3]);
void orc_rec_radial(const double *xH1, const double *xH2, const double *xHe1,
                    const double *xHe2, const double *xHe3, const double *Edges,
                    const double *nH, const double *nHe, const double *T,
                    int have_dec, const double sigma_th[3], const double E_th[3],
                    int Nl, double *ion[3], double *heat[3]);
int orc_rec_isotropic(const double *xH1, const double *xH2, const double *xHe1,
                      const double *xHe2, const double *xHe3, const double *Edges,
                      const double *nH, const double *nHe, const double *T,
                      int have_dec, int Nmu, const double sigma_th[3],
                      const double E_th[3], const double em_fac[3], int Nl,
                      double *ion[3], double *heat[3]);
void orc_rec_slab(const double *xH1, const double *xH2, const double *xHe1,
                  const double *xHe2, const double *xHe3, const double *Edges,
                  const double *nH, const double *nHe, const double *T,
                  const double sigma_th[3], const double E_th[3], int Nl,
                  double *ion[3], double *heat[3]);

#endif
