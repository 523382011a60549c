/*
 * rabacus_b200.h -- C ABI of librabacus_b200.so
 *
 * A B200-native (sm_100a, FP64 CUDA) replacement for the hot path behind
 * galtay/rabacus' f2py module `rabacus_fc`: ray-traced HI/HeI/HeII column
 * densities, the angle x frequency integrals giving photo-ionisation / heating
 * rates per cell, and the per-cell ionisation / temperature equilibrium solve,
 * iterated to convergence.  Every entry point cites the reference interface it
 * replaces (paths relative to the reference root, v0.9.5).  INTEGRATION.md
 * shows the Python binding a Rabacus maintainer would add.
 *
 * Conventions
 *  - plain C: pointers and sizes only; all arrays 1-D contiguous float64, cgs.
 *  - "host" entry points take HOST pointers (numpy buffers) and do their own
 *    host<->device copies; the rb_plan_* API keeps a batch of problems resident
 *    in HBM and exposes the sweep stages for multi-GPU cell sharding.
 *  - return value 0 = success; otherwise one of RB_ERR_* (the reference calls
 *    Fortran `stop`, i.e. kills the process: sphere_bgnd.f90:198-201,284-287;
 *    ion_solver.f90:536-546,804-809,852-856,369-372; geometry.f90:138-142).
 *    rb_strerror() gives the message.
 *  - there is no CPU fallback: without a CUDA device every compute entry
 *    returns RB_ERR_CUDA.
 */
#ifndef RABACUS_B200_H
#define RABACUS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB_OK 0
#define RB_ERR_EDGES_NOT_MONOTONIC 1
#define RB_ERR_MAX_ITER_OUTER 2
#define RB_ERR_MAX_ITER_PCE 3
#define RB_ERR_MAX_ITER_PCTE 4
#define RB_ERR_T_NOT_BRACKETED 5
#define RB_ERR_NAN 6
#define RB_ERR_RAY_GEOMETRY 7
#define RB_ERR_BAD_ARG 8
#define RB_ERR_UNSUPPORTED 9 /* fits other than verner / hg97, sizes beyond a kernel's limit */
#define RB_ERR_CUDA 10
#define RB_ERR_NOMEM 11
#define RB_ERR_PEER_TIMEOUT 12 /* cell-sharded sweep: a peer rank never reached the barrier */

/* geometry / driver selector for plans */
#define RB_GEOM_SPHERE_BGND 0      /* rabacus/f2py/sphere_bgnd.f90 */
#define RB_GEOM_SPHERE_STROMGREN 1 /* rabacus/f2py/sphere_stromgren.f90 */
#define RB_GEOM_SLAB_PLANE_1 2     /* rabacus/f2py/slab_plane.f90 sweep_1 */
#define RB_GEOM_SLAB_PLANE_2 3     /* rabacus/f2py/slab_plane.f90 sweep_2 */
#define RB_GEOM_SLAB_BGND 4        /* rabacus/f2py/slab_bgnd.f90 */

/* What a solve reports back across the boundary (the reference documents an
 * `itr` attribute, sphere_bgnd.py:118, but never returns it). */
typedef struct rb_info {
  int iters;          /* sweeps performed (max over the batch) */
  int n_launches;     /* CUDA kernels launched by this call */
  double resid;       /* last max |ne_new/ne_old - 1| (NaN if not computed) */
  double ms_total;    /* wall time of the call, host clock, incl. copies */
  double ms_device;   /* device time of the iteration loop (CUDA events) */
  double ms_columns;  /* summed device time of the column (ray / scan) kernel */
  double ms_nu;       /* summed device time of the frequency-integral kernel */
  double ms_cell;     /* summed device time of the cell-solver kernel */
  double ms_rec;      /* summed device time of the recombination-photon kernels */
  double ms_finish;   /* summed device time of the convergence / exchange kernel + read-back */
  long long evals;    /* cell*ray*frequency evaluations performed */
} rb_info;

const char *rb_strerror(int code);
/* number of visible CUDA devices (0 if none / no driver) */
int rb_device_count(void);
const char *rb_version(void);
/* The host-pointer entry points below keep up to 8 released plans (device and
 * pinned buffers, stream, events) and re-arm one when a later call has the same
 * geometry and grid shape on the same device; rb_cache_clear() frees them.  (The
 * reference allocates its work arrays per call, sphere_bgnd.f90:157-175.) */
void rb_cache_clear(void);

/* ------------------------------------------------------------------------
 * Host-pointer entry points == the f2py boundary.
 *
 * rb_sphere_bgnd_solve  <-  rabacus_fc.sphere_bgnd.sphere_bgnd_solve
 *   (rabacus/f2py/sphere_bgnd.py:216-244 ; sphere_bgnd.f90:107-155).
 *   Edges[Nl+1], nH/nHe/T[Nl] are intent(inout): T carries the equilibrium
 *   temperature on return when i_find_Teq == 1.  The 17 outputs are [Nl].
 *   SURVEY Q1 (the 1/2 (old + sum) recurrence of the mu quadrature) is
 *   reproduced.  Sweep order is Jacobi (DESIGN.md "Sweep order").
 * ---------------------------------------------------------------------- */
int rb_sphere_bgnd_solve(
    double *Edges, double *nH, double *nHe, double *T, int Nmu,
    const double *E_eV, const double *shape, int i_rec_meth, double fixed_fcA,
    int i_photo_fit, int i_rate_fit, int i_find_Teq, int i_thin,
    double em_H1_fac, double em_He1_fac, double em_He2_fac, double z, double Hz,
    double tol, int Nl, int Nnu, double *xH1, double *xH2, double *xHe1,
    double *xHe2, double *xHe3, double *H1i_src, double *He1i_src,
    double *He2i_src, double *H1i_rec, double *He1i_rec, double *He2i_rec,
    double *H1h_src, double *He1h_src, double *He2h_src, double *H1h_rec,
    double *He1h_rec, double *He2h_rec, rb_info *info);

/* rb_sphere_stromgren_solve <- rabacus_fc.sphere_stromgren.sphere_stromgren_solve
 *   (rabacus/f2py/sphere_stromgren.py:215-241 ; sphere_stromgren.f90:111-158) */
int rb_sphere_stromgren_solve(
    double *Edges, double *nH, double *nHe, double *T, int Nmu,
    const double *E_eV, const double *shape, int i_rec_meth, double fixed_fcA,
    int i_photo_fit, int i_rate_fit, int i_find_Teq, int i_thin,
    double em_H1_fac, double em_He1_fac, double em_He2_fac, double z, double Hz,
    double tol, int Nl, int Nnu, double *xH1, double *xH2, double *xHe1,
    double *xHe2, double *xHe3, double *H1i_src, double *He1i_src,
    double *He2i_src, double *H1i_rec, double *He1i_rec, double *He2i_rec,
    double *H1h_src, double *He1h_src, double *He2h_src, double *H1h_rec,
    double *He1h_rec, double *He2h_rec, rb_info *info);

/* rb_slab_bgnd_solve <- rabacus_fc.slab_bgnd.slab_bgnd_solve
 *   (rabacus/f2py/slab_bgnd.py:198-220 ; slab_bgnd.f90:97-129); only T is inout */
int rb_slab_bgnd_solve(
    const double *Edges, const double *nH, const double *nHe, double *T,
    const double *E_eV, const double *shape, int i_rec_meth, double fixed_fcA,
    int i_photo_fit, int i_rate_fit, int i_find_Teq, int i_thin, double z,
    double Hz, double tol, int Nl, int Nnu, double *xH1, double *xH2,
    double *xHe1, double *xHe2, double *xHe3, double *H1i_src, double *He1i_src,
    double *He2i_src, double *H1i_rec, double *He1i_rec, double *He2i_rec,
    double *H1h_src, double *He1h_src, double *He2h_src, double *H1h_rec,
    double *He1h_rec, double *He2h_rec, rb_info *info);

/* rb_slab_plane_solve <- rabacus_fc.slab_plane.slab_plane_solve
 *   (rabacus/f2py/slab_plane.py:178-200,419-441 ; slab_plane.f90:105-137) */
int rb_slab_plane_solve(
    const double *Edges, const double *nH, const double *nHe, double *T,
    const double *E_eV, const double *shape, int i_2side, int i_rec_meth,
    double fixed_fcA, int i_photo_fit, int i_rate_fit, int i_find_Teq,
    int i_thin, double z, double Hz, double tol, int Nl, int Nnu, double *xH1,
    double *xH2, double *xHe1, double *xHe2, double *xHe3, double *H1i_src,
    double *He1i_src, double *He2i_src, double *H1i_rec, double *He1i_rec,
    double *He2i_rec, double *H1h_src, double *He1h_src, double *He2h_src,
    double *H1h_rec, double *He1h_rec, double *He2h_rec, rb_info *info);

/* rb_set_column_vs_impact <- rabacus_fc.sphere_base.set_column_vs_impact
 *   (rabacus/f2py/sphere_base.py:182-193 ; sphere_base.f90:1115-1205) */
int rb_set_column_vs_impact(const double *xH1, const double *xH2,
                            const double *xHe1, const double *xHe2,
                            const double *xHe3, const double *Edges,
                            const double *nH, const double *nHe,
                            const double *T, int Nl, double *NH, double *NHe,
                            double *NH1, double *NHe1, double *NHe2);

/* Single-zone batch API <- rabacus_fc.ion_solver.solve_{ce,pce,pcte}_{s,v},
 *   rabacus_fc.chem_cool_rates.get_{kchem,kcool}_{s,v}
 *   (rabacus/f2py/ion_solver.py:150-159,371-380,633-643 ;
 *    rabacus/f2py/chem_cool_rates.py:104-106,332-334).
 *   `lockstep` != 0 selects the `_v` semantics (all zones iterate until all
 *   converge, ion_solver.f90:620,988); 0 runs the scalar solver per zone. */
int rb_get_kchem(const double *T, const double *fcA_H2, const double *fcA_He2,
                 const double *fcA_He3, int irate, int nn, double *reH2,
                 double *reHe2, double *reHe3, double *ciH1, double *ciHe1,
                 double *ciHe2);
int rb_get_kcool(const double *T, const double *fcA_H2, const double *fcA_He2,
                 const double *fcA_He3, int irate, int nn, double *recH2,
                 double *recHe2, double *recHe3, double *cicH1, double *cicHe1,
                 double *cicHe2, double *cecH1, double *cecHe2, double *bremss,
                 double *compton);
int rb_solve_ce(const double *reH2, const double *reHe2, const double *reHe3,
                const double *ciH1, const double *ciHe1, const double *ciHe2,
                int nn, double *xH1, double *xH2, double *xHe1, double *xHe2,
                double *xHe3);
int rb_solve_pce(const double *nH, const double *nHe, const double *reH2,
                 const double *reHe2, const double *reHe3, const double *ciH1,
                 const double *ciHe1, const double *ciHe2, const double *H1i,
                 const double *He1i, const double *He2i, double tol, int nn,
                 int lockstep, double *xH1, double *xH2, double *xHe1,
                 double *xHe2, double *xHe3);
int rb_solve_pcte(const double *nH, const double *nHe, const double *H1i,
                  const double *He1i, const double *He2i, const double *H1h,
                  const double *He1h, const double *He2h, double z, double Hz,
                  const double *fcA_H2, const double *fcA_He2,
                  const double *fcA_He3, int irate, double tol, int nn,
                  int lockstep, double *xH1, double *xH2, double *xHe1,
                  double *xHe2, double *xHe3, double *T);

/* Verner 96 cross sections  <- rabacus_fc.photo_xsections.return_sigma_*
 *   (rabacus/f2py/photo_xsections.f90:100-180); species 0=H1, 1=He1, 2=He2.
 *   Host-side helper (no GPU needed). */
int rb_photo_sigma(int species, const double *E_eV, int nn, double *sigma);
double rb_photo_Eth(int species);
/* Gauss-Legendre rule used for the polar angles
 *   <- p_quadrature_rule (rabacus/f2py/legendre_polynomial.f90:815-864).
 *   Host-side helper (no GPU needed). */
int rb_gauss_legendre(int n, double *x, double *w);

/* ------------------------------------------------------------------------
 * Device-resident plan API: a batch of `n_problems` independent problems of
 * identical shape (Nl, Nnu, Nmu) and solver flags, resident in HBM.  Used by
 * bench.py, by the parameter-sweep driver (problems sharded over
 * GPUs, no communication) and by the cell-sharded multi-GPU driver (each rank
 * sweeps cells {cell_start + j*cell_stride}, the host all-gathers the state
 * fields through torch.distributed/NCCL between rb_plan_sweep_local and
 * rb_plan_finish_sweep).
 * ---------------------------------------------------------------------- */
typedef struct rb_plan rb_plan;

typedef struct rb_plan_desc {
  int geometry;   /* RB_GEOM_* */
  int n_problems; /* batch size B */
  int Nl, Nnu, Nmu;
  int i_find_Teq;
  double fixed_fcA;
  double tol;
  int device;     /* CUDA device ordinal, -1 = current */
  int max_sweeps; /* 0: reference limits; >0: stop every problem after this many;
                     <0: fixed-work mode, exactly -max_sweeps sweeps, convergence ignored */
  int i_rec_meth; /* 0 or 1: fixed case-A fraction; spheres: 2 outward, 3 radial,
                     4 isotropic (sphere_base.f90:79-1083); slabs: 2 = "ray"
                     (slab_base.f90:163-560).  != 1 forces case-A rates. */
  double em_fac[3]; /* isotropic: multipliers of the H1/He1/He2 recombination emission
                       (sphere_bgnd.f90:75-77, applied at sphere_base.f90:955) */
  int em_fac_set;   /* 0: em_fac is ignored and (1,1,1) is used (a zeroed desc gives the
                       reference's Python defaults, sphere_bgnd.py:138-150); != 0: em_fac is
                       used as given, explicit zeros switch the emission off */
} rb_plan_desc;

int rb_plan_create(rb_plan **plan, const rb_plan_desc *desc);
void rb_plan_destroy(rb_plan *plan);
/* stage one problem's inputs (host pointers, copied into pinned staging) */
int rb_plan_set_problem(rb_plan *plan, int b, const double *Edges,
                        const double *nH, const double *nHe, const double *T,
                        const double *E_eV, const double *shape, double z,
                        double Hz);
/* the same for n problems b0 .. b0+n-1 in one call (parameter sweeps): row-major
 * Edges[n][Nl+1], nH/nHe/T[n][Nl]; problem i uses spectrum spec_index[i] of the nspec
 * rows E_eV/shape[nspec][Nnu]; z[n], Hz[n] (NULL: zeros) */
int rb_plan_set_batch(rb_plan *plan, int b0, int n, const double *Edges,
                      const double *nH, const double *nHe, const double *T, int nspec,
                      const double *E_eV, const double *shape, const int *spec_index,
                      const double *z, const double *Hz);
/* Photon-energy grid of rabacus/rad_src/source.py:178-259 (segmented=True): Nnu
 * log-spaced samples between q_min and q_max (units of 13.6 eV) with the samples nearest
 * the H I / He I / He II thresholds snapped onto them; E_eV [Nnu].  Host-side helper.
 * RB_ERR_BAD_ARG where the reference raises (q_min > 1, q_max < 54.42/13.6, Nnu < 2). */
int rb_photon_energies(double q_min, double q_max, int Nnu, double *E_eV);
/* Device-side input producers for parameter sweeps of SphereBgnd problems (replaces the
 * per-problem Python work of rabacus/rad_src/source.py:178-259 -- segmented photon-energy
 * grid --, rabacus/uv_bgnd/hm12.py:239-303 -- HM12 spectrum at redshift z: linear in z,
 * log-log in wavelength --, rabacus/f2py/verner_96.f90:108-154 -- cross sections -- and
 * of the density profile of doc/guide_bgnd_sphere_examples.rst:126-146).  Every problem b
 * of the plan is described by three numbers: Edges = linspace(0, R_cm[b], Nl+1),
 * nH = nH0[b] * (r < r0 ? 1 : (r/r0)^-slope) at the shell centres with r0 = R/core_div,
 * nHe = nH * nHe_over_nH, T = T0, spectrum = HM12 at z[b] on the Nnu-point grid
 * q_min..q_max (in units of 13.6 eV).  The HM12 table is passed as log10(lambda/cm)
 * [nlam] non-decreasing, z_tab [nzt] ascending and log10 I_nu [nlam][nzt].  rb_plan_upload then
 * generates all inputs ON THE DEVICE (k_family_profiles, k_family_spectra); only these
 * arguments cross PCIe.  E_eV_out (may be NULL) receives the energy grid.  Results come
 * back centre-first, like a call with increasing Edges. */
int rb_plan_set_sphere_family(rb_plan *plan, const double *nH0, const double *R_cm,
                              const double *z, const double *Hz, double core_div,
                              double slope, double nHe_over_nH, double T0, double q_min,
                              double q_max, int nlam, int nzt, const double *log_lam_cm,
                              const double *z_tab, const double *logInu, double *E_eV_out);
/* H2D copy of everything staged (or, after rb_plan_set_sphere_family, generation of the
 * inputs on the device); resets iteration state */
int rb_plan_upload(rb_plan *plan);
/* optically-thin initial state (sphere_bgnd.f90:379-499 and twins) */
int rb_plan_thin(rb_plan *plan);
/* Sweep order of SphereBgnd plans.  RB_ORDER_JACOBI (default): every cell's rays see the
 * pre-sweep state -- the order of the reference's other three drivers, and the one that
 * shards.  RB_ORDER_GAUSS_SEIDEL: rabacus/f2py/sphere_bgnd.f90:753-913 as it runs with
 * OMP_NUM_THREADS=1 (its only deterministic order): cells outermost first, each seeing the
 * cells before it already updated.  Serial in the cells (4-5 launches per cell): one GPU
 * only, no cell shards; meant for checks against the reference's iterates, not for speed.
 * The one-shot entry rb_sphere_bgnd_solve takes the order from the environment variable
 * RB_SPHERE_BGND_ORDER ("gs" | "jacobi"). */
#define RB_ORDER_JACOBI 0
#define RB_ORDER_GAUSS_SEIDEL 1
int rb_plan_set_sweep_order(rb_plan *plan, int order);
/* one Jacobi sweep over the local cells of the deal (cell_start, cell_stride) -- stride > 0:
 * il = cell_start + j*cell_stride; stride < 0: 32-cell blocks over -stride ranks (see
 * rb_plan_peer_deal) --: columns -> frequency/angle integrals -> cell solve.  Does NOT test
 * convergence. */
int rb_plan_sweep_local(rb_plan *plan, int cell_start, int cell_stride);
/* convergence test on n_e over all cells + iteration bookkeeping
 * (sphere_bgnd.f90:284-296); returns the number of still-active problems.
 * rb_plan_finish_sweep = _async + poll until drained.  With _async / _poll the
 * host may enqueue up to 6 sweeps ahead of the read-back: sweeps enqueued after
 * a problem converged are no-ops on the device (per-problem `active` flags). */
int rb_plan_finish_sweep(rb_plan *plan, int *n_active);
int rb_plan_finish_sweep_async(rb_plan *plan);
int rb_plan_poll_sweep(rb_plan *plan, int *n_active);
/* cell-shard exchange (single-problem plans): pack the 18 state fields of the
 * local cells into dev_send [18][n_local] / scatter an all-gathered buffer
 * dev_recv [world][18][n_local] (cell il = j*world + g) back into the plan.
 * Device pointers; enqueued on the plan's stream. */
int rb_plan_pack_cells(rb_plan *plan, int cell_start, int cell_stride, double *dev_send);
int rb_plan_unpack_cells(rb_plan *plan, int world, const double *dev_recv);
/* The deal of a shard's cells is a (start, stride) pair everywhere in this interface:
 * stride > 0: cyclic, il = start + j*stride; stride < 0: blocks of 32 consecutive cells
 * dealt cyclically over W = -stride ranks, il = ((j/32) W + start) 32 + j%32 (needs the
 * number of swept cells to be a multiple of 32 W).  rb_plan_unpack_cells takes the same code
 * as `world`.  Peer-connected plans use the block deal where the grid allows it (the 32
 * lanes of a ray-walk warp then hold consecutive shells); rb_plan_peer_deal returns the code
 * to pass to rb_plan_pack_cells / rb_plan_unpack_cells for the final gather. */
int rb_plan_peer_deal(const rb_plan *plan);
/* local cell index j of the deal (start, stride) -> cell (host-side helper, no GPU) */
int rb_cell_of(int j, int start, int stride);
/* Cell sharding over peer memory (NVLink/NVSwitch), one process per GPU, no
 * collective library on the per-sweep path.  Rank `rank` of `world` (<= 8) sweeps
 * the cells il = rank + j*world of a single-problem plan.
 *   rb_plan_peer_alloc   allocates this rank's staging + flag buffer and returns
 *                        its 64-byte CUDA IPC handle;
 *   rb_plan_peer_connect takes all ranks' handles ([world][64], exchanged by the
 *                        caller, e.g. torch.distributed.all_gather) and maps them;
 *   rb_plan_sweep_sharded enqueues one sweep: columns -> frequency integrals ->
 *                        cell solve, whose epilogue stores the new fractions into
 *                        every rank's staging buffer -> k_exchange_finish (flag
 *                        barrier across the GPUs, scatter, convergence test on
 *                        the full arrays).  Poll with rb_plan_poll_sweep.
 *   rb_plan_solve_sharded thin start + pipelined sweeps to convergence; every
 *                        rank takes the same decisions (Jacobi: the iterates and
 *                        the sweep count equal those of one GPU).  T and the six
 *                        rates of the other ranks' cells are NOT exchanged per
 *                        sweep: gather them once afterwards
 *                        (rb_plan_pack_cells / all-gather / rb_plan_unpack_cells). */
int rb_plan_peer_alloc(rb_plan *plan, int rank, int world, unsigned char handle[64]);
int rb_plan_peer_connect(rb_plan *plan, const unsigned char *handles);
int rb_plan_sweep_sharded(rb_plan *plan);
int rb_plan_solve_sharded(rb_plan *plan, rb_info *info);
/* thin state + sweeps to convergence on one GPU */
int rb_plan_solve(rb_plan *plan, rb_info *info);
/* D2H of one problem's results (un-reversed to the input order) */
int rb_plan_get_problem(rb_plan *plan, int b, double *T, double *xH1,
                        double *xH2, double *xHe1, double *xHe2, double *xHe3,
                        double *H1i_src, double *He1i_src, double *He2i_src,
                        double *H1h_src, double *He1h_src, double *He2h_src,
                        int *iters);
/* D2H of one problem's recombination-photon rates (zeros for rec_meth = fixed) */
int rb_plan_get_rec(rb_plan *plan, int b, double *H1i_rec, double *He1i_rec,
                    double *He2i_rec, double *H1h_rec, double *He1h_rec,
                    double *He2h_rec);
/* D2H of ALL problems' results in one pass (parameter sweeps): out12 is
 * [12][n_problems][Nl] = xH1,xH2,xHe1,xHe2,xHe3,T,H1i,He1i,He2i,H1h,He1h,He2h (source
 * rates), rec6 (may be NULL) [6][n_problems][Nl] the recombination-photon rates, iters
 * (may be NULL) [n_problems] the sweep counts; every problem in its input orientation. */
int rb_plan_get_all(rb_plan *plan, double *out12, double *rec6, int *iters);
/* The same results in a PINNED host buffer owned by the plan (no caller allocation, no
 * first-touch page faults inside the copy: the read-back of a large parameter sweep runs at
 * PCIe speed).  rb_plan_reserve_results allocates the buffer ahead of time (18 fields x
 * n_problems x Nl doubles); rb_plan_get_all_pinned fills it and returns pointers into it,
 * valid until the next call or rb_plan_destroy (any of the three may be NULL). */
int rb_plan_reserve_results(rb_plan *plan);
int rb_plan_get_all_pinned(rb_plan *plan, const double **out12, const double **rec6,
                           const int **iters);
/* device pointers for torch / NCCL interop: fields are [n_problems*Nl] float64.
 * 0..4 = xH1,xH2,xHe1,xHe2,xHe3 ; 5 = T ; 6..11 = H1i,He1i,He2i,H1h,He1h,He2h src */
void *rb_plan_field_ptr(rb_plan *plan, int field);
/* the CUDA stream (cudaStream_t) all plan work is enqueued on */
void *rb_plan_stream(rb_plan *plan);
int rb_plan_sync(rb_plan *plan);
long long rb_plan_evals_per_sweep(const rb_plan *plan);
int rb_plan_last_info(const rb_plan *plan, rb_info *info);

/* (measurement and diagnostic entry points: rabacus_b200/csrc/rb_internal.h) */

#ifdef __cplusplus
}
#endif
#endif
