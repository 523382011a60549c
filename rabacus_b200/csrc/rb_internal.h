/*
 * rb_internal.h -- measurement and diagnostic entry points of librabacus_b200.so.
 *
 * NOT part of the drop-in boundary (include/rabacus_b200.h): nothing here replaces an
 * interface of the reference.  These symbols exist for bench.py (timed steps, FP64-pipe
 * micro-benchmarks) and for the unit tests of the arithmetic in rb_fastpow.cuh /
 * rb_atomic.cuh; no solver entry point calls them.
 */
#ifndef RB_INTERNAL_H
#define RB_INTERNAL_H

#include "../../include/rabacus_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* out[i] = x[i]**y[i] by csrc/rb_fastpow.cuh.  on_device = 0 evaluates the same inline
 * source with the host compiler (unit tests without a GPU), 1 runs one CUDA thread per
 * element. */
int rbi_diag_pow(const double *x, const double *y, int n, double *out, int on_device);
/* The 6 + 10 coefficients of get_kchem_s / get_kcool_s (chem_cool_rates.f90:45-204) as
 * out16[16][n].  libm != 0: every power through libm / libdevice pow(), statement by
 * statement, instead of the table-driven form. */
int rbi_diag_hg97(const double *T, const double *fcA_H2, const double *fcA_He2,
                  const double *fcA_He3, int n, double *out16, int on_device, int libm);
/* One temperature's 6 + 10 coefficients from the HOST build of the device code (same bits
 * as the kernels): handed to the oracle as its coefficient hook by the parity tests, so
 * that the oracle's solvers can be required to reproduce the device's fractions exactly. */
void rbi_host_coefficients(double T, double fcA_H2, double fcA_He2, double fcA_He3,
                           double *out16);
/* Per-plan A/B switch: on != 0 makes every kernel of THIS plan evaluate the Hui & Gnedin
 * fits through libdevice pow() (no global state; other plans are unaffected). */
int rbi_plan_set_libm_pow(rb_plan *plan, int on);

/* Run `steps` Jacobi sweeps (rb_plan_sweep_local over all cells + rb_plan_finish_sweep;
 * rb_plan_sweep_sharded on a peer-connected plan) and return their summed device time in
 * ms, every step bracketed by its own pair of CUDA events on the plan's stream.  If
 * flush_bytes > 0 a scratch buffer of that size is overwritten between steps (L2 flush),
 * outside the timed spans.  graphs != 0: steps after the first are replayed as CUDA
 * graphs like in rb_plan_solve (single-problem plans). */
int rbi_plan_bench_steps(rb_plan *plan, int steps, size_t flush_bytes, int graphs,
                         double *ms_sum);

/* A/B switch of THIS plan: on = 0 makes the SphereBgnd column kernel walk every shell edge
 * of every ray (the round-1 path) instead of taking the far field from the moment table
 * (rb_kernels.cuh "Far-field evaluation"); RB_ERR_UNSUPPORTED if the plan has no table. */
int rbi_plan_set_far_field(rb_plan *plan, int on);

/* Rows [row0, row0 + n) of the memo table of the rate fits over the temperature bisection
 * tree (rb_ion.cuh: TeqTab) for case-A fraction fcA on the current device, built on demand:
 * out16[n][16] = {6 kchem, 8 kcool, bremss, T}; *depth = the table's depth (row 2^depth is
 * TMAX).  RB_ERR_UNSUPPORTED when the memo is disabled (RB_TEQ_DEPTH=0). */
int rbi_teq_table_fetch(double fcA, int row0, int n, double *out16, int *depth);

/* D2H of problem b's staged INPUTS as the kernels see them (solver orientation): Edges
 * [Nl+1], nH / nHe / T [Nl], the per-frequency table [Nnu][9] and the thin integrals [9];
 * any pointer may be NULL.  For the tests of the device-side input producers. */
int rbi_plan_fetch_inputs(rb_plan *plan, int b, double *Edges, double *nH, double *nHe,
                          double *T, double *spec, double *thin);

/* Checked build (-DRB_DEBUG_CHECKS, librabacus_b200_checked.so): how many of the guard bytes
 * behind this plan's device buffers were overwritten (0 = none); `where` receives the name of
 * the first damaged buffer.  Returns -1 in the release build (no guards). */
int rbi_plan_check_guards(rb_plan *plan, char *where, size_t wherelen);

/* The two decisions of a solve_pce step as the kernels take them (csrc/rb_ion.cuh: comparison
 * instead of ratio_ne, reciprocal-based certified error test) next to the IEEE divisions of
 * the reference, for n triples a[3n] / b[3n]: out[i] bit 0 = certified err > tol, bit 1 = by
 * division; bits 2/3 = a1 > b1 vs a1/b1 > 1; bits 4/5 = a1 < b1 vs a1/b1 < 1. */
int rbi_diag_ratio_err(const double *a, const double *b, int n, double tol, int *out);

/* device time [ms] of the kernel of this thread's last rb_solve_pce / rb_solve_pcte call
 * (scalar mode), for the single-zone batch line of bench.py */
double rbi_last_zone_kernel_ms(void);

/* FP64 pipe micro-benchmarks used as roofline denominators (bench.py):
 * kind 0 = DFMA (64 FMAs per loop trip on 32 independent chains, >= 50 ms per repetition),
 * 1 = exp, 2 = sqrt, 3 = pow (CUDA library functions).  Returns operations / second. */
int rbi_microbench(int kind, int device, double *ops_per_s);

#ifdef __cplusplus
}
#endif
#endif
