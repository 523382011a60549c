// rb_kernels.cuh -- sm_100a FP64 kernels of the Rabacus hot path.
//
//   k_sphere_columns_smem  ray-pair column densities in concentric shells: shell
//                        records staged in shared memory, LPT work list, summed by
//                        parts (reference hot loop A: sphere_bgnd.f90:764-792 with
//                        geometry.f90:37-168); k_columns_combine for chunked walks;
//                        k_pack_shells + k_sphere_columns: global-memory fallback
//   k_column_taus        threshold optical depths below/above each layer by a
//                        block-wide warp-shuffle scan (slab_base.f90:63-115,
//                         sphere_base.f90:1232-1367)
//   k_nu_rates           fused tau(nu) -> attenuation -> direction quadrature -> 6
//                        trapezoid integrals (reference hot loop B:
//                        sphere_bgnd.f90:794-856 with source_background.f90:259-446,
//                        source_point.f90:299-395, source_plane.f90:249-342)
//   k_cell_solve         batched per-cell solve_pce_s / solve_pcte_s
//                        (reference hot loop C: sphere_bgnd.f90:886-907);
//   k_cell_solve_pce_spec / k_cell_solve_spec: the same bisections (on n_e / on T)
//                        evaluated speculatively by a group of lanes per cell
//   k_converge_finish    n_e convergence test and iteration control
//                        (sphere_bgnd.f90:284-296); k_exchange_finish: the same after
//                        the cell-shard exchange over peer memory
//   k_thin_state         optically thin start with the lock-step `_v` solvers over a
//                        thread-block cluster
//                        (sphere_bgnd.f90:379-499; ion_solver.f90:565-696,884-1056)
// Recombination-photon transfer: rb_rec.cuh.
//
// None of this is a dense contraction: no tensor cores.  The governing
// resource is the FP64 pipe (DFMA + software exp / sqrt / pow), see DESIGN.md.
#pragma once
#include <cooperative_groups.h>

#ifdef RB_FAR_COEFFS_H   // another (RATIO, NT) pair: -DRB_FAR_COEFFS_H='"path"' (gen_far_coeffs.py)
#include RB_FAR_COEFFS_H
#else
#include "rb_far_coeffs.h"
#endif
#include "rb_ion.cuh"

namespace rb {

#ifndef RB_LS_UNROLL
#define RB_LS_UNROLL _Pragma("unroll")
#endif
#ifndef RB_FLAG_T
#define RB_FLAG_T bool
#endif
typedef RB_FLAG_T rb_flag_t;

constexpr int kPackPad = 8;    // padding records after the Nl+1 shell records
constexpr int kSpecStride = 9;  // per-frequency table: 3 sigma ratios, 3+3 weights
constexpr int kThinStride = 9;  // per problem: 6 optically thin integrals, 3 min sigma ratios

// Everything a kernel needs, passed by value.
struct PlanDev {
  int B, Nl, Nnu, Nmu, Ndir, geom, find_Teq, max_sweeps;
  int libm_pow, pad0_;     // libm_pow != 0: rate fits through libm pow() (diagnostic A/B)
  double fcA, tol, tol_inner;
  double sigma_th[3];
  const double *Edges;  // [B][Nl+1] in solver orientation
  const double *nH, *nHe;  // [B][Nl]
  double *T;               // [B][Nl]
  double *x[5];            // xH1,xH2,xHe1,xHe2,xHe3  [B][Nl]
  double *src[6];          // H1i,He1i,He2i,H1h,He1h,He2h  [B][Nl]  (source photons)
  double *rec[6];          // the same six rates from recombination photons (zero for
                           // rec_meth = fixed; rb_rec.cuh otherwise)
  double *conv_old;        // [B][Nl]
  double *kchem;           // [6][B][Nl] fixed-T runs: the six rate coefficients of every cell,
                           // evaluated once by the thin-state kernel (T never changes)
  double4 *pack;           // [B][Nl+1+kPackPad]
  const double *spec;      // [B][Nnu][9]
  const double *thin;      // [B][9] optically thin integrals (sum of weights), max_nu sigma ratios
  const double *mu, *wmu;  // [Nmu] Gauss-Legendre nodes / weights
  double4 *col;            // [B][Nl][Ndir] per ray {tau_HI,th, tau_HeI,th, tau_HeII,th, quadrature
                           // weight}: one aligned 32-byte sector per ray, written whole by the
                           // column kernels, and one contiguous Ndir*32 B run per cell for the
                           // frequency kernel
  const double *zHz;       // [B][2]
  int *active, *iters, *err;            // [B]
  unsigned long long *resid_bits;       // [B]
  int *nactive;                         // [kSlots] active-problem count per sweep slot
  TeqTab teq;              // find_Teq: memo of the rate coefficients over the temperature
                           // bisection tree (rb_ion.cuh), shared by every cell and sweep
};

// Cell-sharded multi-GPU exchange over peer memory (NVLink / NVSwitch), SURVEY 8e.
// Rank r owns the swept cells il = r + j*world.  The cell solver writes the five
// new fractions of its cells straight into EVERY rank's staging buffer (remote
// stores, `stage[r]` are cudaIpc-mapped peer pointers); the finish kernel then
// raises this rank's flag on every peer, waits for all peers' flags, scatters
// the staged fractions into the full state arrays and runs the convergence test.
// Two staging rows alternate: a rank can run at most one barrier ahead.
constexpr int kMaxPeers = 8;
// Layout of a rank's peer buffer: 256 control bytes, then the staging rows.
//   u64[0..7]   flags, slot r written by rank r ("my cells of sweep e are staged here")
//   u64[16]     epoch: sweeps this rank has completed (device-resident, so that a sharded
//               sweep has no per-sweep kernel argument and can be replayed as a CUDA graph)
//   u64[17..19] arrival counter / not-converged count / worst change of the multi-CTA finish
constexpr int kPeerCtl = 16;
constexpr unsigned long long kPeerPoison = ~0ull;   // flag value: "a rank gave up, abort"
struct PeerDev {
  int world, rank, n_local;
  int deal;                               // cell_of() stride code of the shard: +world cyclic,
                                          // -world 32-cell blocks
  unsigned long long timeout_ns;          // barrier timeout (RB_ERR_PEER_TIMEOUT)
  unsigned long long *ctl;                // this rank's control words (own buffer + kPeerCtl)
  double *stage[kMaxPeers];               // rank r's staging [2][world][6][n_local]
  unsigned long long *flags[kMaxPeers];   // rank r's flags [kMaxPeers], written by its peers
};

constexpr int kExchangeFields = 18;  // rb_plan_pack_cells / rb_plan_unpack_cells
constexpr int kPeerFields = 6;  // five fractions + T (the photon-transfer closures read all T)
// id of the sweep in flight = completed sweeps + 1; its staging row alternates
__device__ __forceinline__ unsigned long long peer_sweep_id(const PeerDev &Q) {
  return *(volatile const unsigned long long *)Q.ctl + 1ull;
}
__device__ __forceinline__ void peer_scatter(const PeerDev &Q, int j, const rb_frac_t &x,
                                             double T) {
  const size_t n = (size_t)Q.n_local;
  const int parity = (int)(peer_sweep_id(Q) & 1ull);
  RB_ASSERT(j >= 0 && j < Q.n_local && Q.rank < Q.world && Q.world <= kMaxPeers);
  const size_t off = ((size_t)(parity * Q.world + Q.rank) * kPeerFields) * n + j;
  for (int r = 0; r < Q.world; ++r) {
    double *d = Q.stage[r] + off;
    d[0] = x.xH1; d[n] = x.xH2; d[2 * n] = x.xHe1; d[3 * n] = x.xHe2; d[4 * n] = x.xHe3;
    d[5 * n] = T;
  }
}

__device__ __forceinline__ size_t col_index(const PlanDev &P, int b, int il, int idir) {
  return ((size_t)b * P.Nl + il) * (size_t)P.Ndir + idir;
}

__device__ __forceinline__ bool two_sided(int geom) {
  return geom == RB_GEOM_SLAB_PLANE_2 || geom == RB_GEOM_SLAB_BGND;
}

// absorber densities {n_HI, n_HeI, n_HeII} of shell k (zero outside 0 .. Nl-1)
__device__ __forceinline__ double3 shell_dens(const PlanDev &P, int b, int k) {
  if (k < 0 || k >= P.Nl) return make_double3(0.0, 0.0, 0.0);
  const size_t c = (size_t)b * P.Nl + k;
  return make_double3(P.nH[c] * P.x[0][c], P.nHe[c] * P.x[2][c], P.nHe[c] * P.x[3][c]);
}

// shell record k of problem b (k_pack_shells / the shared-memory fill):
// {E_k^2, c_k - c_{k-1}} for the three absorbers -- the column sums are taken in
// summed-by-parts form (see the column kernel), which needs the density STEPS at the
// shell edges rather than the densities.
__device__ __forceinline__ double4 shell_record(const PlanDev &P, int b, int k) {
  if (k > P.Nl) return make_double4(0.0, 0.0, 0.0, 0.0);  // padding read by the prefetch
  const double E = P.Edges[(size_t)b * (P.Nl + 1) + k];
  const double3 c1 = shell_dens(P, b, k), c0 = shell_dens(P, b, k - 1);
  // geometry.f90:148  Edges(i1)*Edges(i1)
  return make_double4(E * E, c1.x - c0.x, c1.y - c0.y, c1.z - c0.z);
}

// ---------------------------------------------------------------------------
// k_pack_shells: one thread per (problem, shell edge k = 0..Nl) -- only for the
// global-memory fallback of the column kernel
// ---------------------------------------------------------------------------
__global__ void k_pack_shells(PlanDev P) {
  const int b = blockIdx.y;
  if (!P.active[b]) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > P.Nl + kPackPad - 1) return;
  P.pack[(size_t)b * (P.Nl + 1 + kPackPad) + k] = shell_record(P, b, k);
}

// ---------------------------------------------------------------------------
// k_sphere_columns: one thread per (cell il, |mu| pair).  A warp holds 32
// neighbouring cells of the same direction so that ray lengths inside a warp
// are nearly equal and every lane reads the same shell record (broadcast).
//
// Geometry (geometry.f90:53-71,146-165; sphere_bgnd.f90:766-792),
// shells outer -> inner, s_k = sqrt(E_k^2 - p^2):
//   through shell k < m : ds_k = |s_{k+1} - s_k| ;  tangent shell m : 2 s_m
//   mu >  0 : shells il, il-1, .., 0          (first segment halved)
//   mu <= 0 : shells il, il+1, .., m, .., 0   (first segment halved)
// The +mu / -mu rays of a Gauss-Legendre pair share the impact parameter, so
// the s_k are computed once per pair (the reference computes them per ray and
// per segment).  With S = sum_{k<m} ds_k c_k, A = sum_{k<il} ds_k c_k,
// d = ds_il c_il, t = 2 s_m c_m:
//   il <  m : N(+mu) = A + d/2 ;  N(-mu) = 2 S - A - d/2 + t
//   il == m : N(+mu) = N(-mu) = A + t/2
// Summation by parts: with the density steps D_k = c_k - c_{k-1} (c_{-1} = 0) and
// P_j = sum_{k<=j} s_k D_k  (s_k decreases with k, so ds_k = s_k - s_{k+1}):
//   S + t/2 = P_m ,   A = P_{il-1} - s_il c_{il-1}
// so that per shell edge the loop does  x = E_k^2 - p^2, s = sqrt(x), P += s D_k
// (9 FP64 instructions: no ds, no carried s_{k-1}) and
//   il <  m : N(+mu) = A + d/2 ;  N(-mu) = 2 P_m - A - d/2
//   il == m : N(+mu) = N(-mu) = P_m
// with d = (s_il - s_{il+1}) c_il from the two roots caught on the way.  Same terms as
// the reference's running sums, associated differently: results agree to ~1e-14.
// ---------------------------------------------------------------------------
// sqrt(x) for finite x >= 0 in 5 FP64-pipe instructions.  y0 = MUFU.RSQ64H(x)
// (rsqrt.approx.ftz.f64, relative error <= 2^-22) gives g0 = x y0 ~ sqrt(x) and the
// exactly rounded residual e = 1 - g0 y0 = 1 - x y0^2; then
//     sqrt(x) = g0 (1 - e)^(-1/2) = g0 (1 + e/2 + 3 e^2/8 + O(e^3)),
// whose truncation 5 e^3/16 < 2^-63.  Result within ~0.5 ulp (CUDA's IEEE sqrt()
// costs ~14 instructions).  x = 0: the seed +inf is clamped to a finite value by
// one integer min on its high word, so g0 = 0 and the result is exactly 0.
template <bool SAFE = true>
__device__ __forceinline__ double sqrt_pos(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  // x = 0: the seed is +inf and 0 * inf = NaN.  SAFE clamps the seed (one integer min on
  // its high word) so that the result is exactly 0; callers that can prove x > 0 (the
  // column walk, see sphere_ray_pair) skip it.
  if (SAFE) y = __hiloint2double(min(__double2hiint(y), 0x7fe00000), 0);
  const double g0 = x * y;
  const double e = fma(-g0, y, 1.0);
  const double q = fma(e, 0.375, 0.5);
  const double ge = g0 * e;
  return fma(ge, q, g0);
}

// ---------------------------------------------------------------------------
// Far-field evaluation of the column sums.
//
// A ray pair with impact parameter p adds s_k D_k, s_k = sqrt(E_k^2 - p^2), over ALL shell
// edges above its tangent point -- 0.68 Nl edges on average, most of them far from the
// tangent point, where s_k is a smooth function of u_k = (p/E_k)^2:
//     s_k = E_k sqrt(1 - u_k) = E_k sum_n b_n u_k^n          (u_k <= 1/RATIO^2 = 0.64)
// (b_n: Chebyshev interpolant of degree NT-1, error 5e-18, rb_far_coeffs.h).  Summed over the
// edges k <= K that are far for this ray (E_k >= RATIO p) the p-dependence factors out:
//     sum_{k<=K} s_k D_k = E_K sum_n b_n u_K^n Mt_n[K],   Mt_n[K] = sum_{k<=K} D_k (E_K/E_k)^(2n-1)
// and the moments Mt_n obey  Mt_n[K] = Mt_n[K-1] (E_K/E_{K-1})^(2n-1) + D_K : one pass over
// the edges per sweep (k_far_moments), after which every ray pair gets its whole far field
// from ONE table row with NT fused multiply-adds per absorber -- and walks only the edges
// with p < E_k < RATIO p (11 % of the edge visits at C4).  The normalisation by E_K keeps
// every factor in (0, 1]: no overflow for any radial dynamic range.  Rounding: measured
// <= 2e-14 of a column for random 1e6 density jumps (scratch/farfield_proto.py), i.e. the
// same order as the summed-by-parts walk itself; the result depends on (cell, mu) only, not
// on how cells are dealt to warps or GPUs.
//
// Only every S-th row is kept (S = 2^slog, far_row_stride()): a ray pair takes the last kept
// row K' = Kf - Kf mod S at or above its far limit Kf and walks S/2 more edges on average.
// The kept rows -- (Nl/S + 1) x 3 NT doubles -- fit in SHARED memory next to the shell
// records, so the far field costs no memory traffic at all (with one row per edge in global
// memory a batch of 8192 problems moved 107 GB of rows through L2 per sweep, and that --
// not the arithmetic -- was the column stage: 1.81 ms per 1024 problems).
// ---------------------------------------------------------------------------
constexpr int kFarNT = RB_FAR_NT;
constexpr int kFarSeg = 32;           // edges per scan segment (fixed: part of the rounding)
constexpr int kMaxClusterFar = 8;     // CTAs per cluster of k_far_moments (= kMaxCluster)
constexpr int kFarRow = 3 * kFarNT;   // doubles per kept row: [NT][3 absorbers]
__constant__ double kFarB[kFarNT] = {RB_FAR_COEFFS};

// kept rows / row stride for a grid of Nl shells: the smallest power of two S >= 8 for which
// shell records + kept rows fit the shared memory of one CTA (three CTAs per SM where S = 8
// allows it)
__host__ __device__ inline int far_rows(int Nl, int slog) { return (Nl >> slog) + 1; }
__host__ __device__ inline size_t far_cols_smem(int Nl, int slog) {
  return (size_t)(Nl + 1 + kPackPad) * sizeof(double4) +
         (size_t)far_rows(Nl, slog) * kFarRow * sizeof(double);
}
__host__ inline int far_row_slog(int Nl) {
  int slog = 3;
  while (slog < 12 && far_cols_smem(Nl, slog) > (size_t)226 * 1024) ++slog;
  return slog;
}

// g_K = E_K / E_{K-1} (0 for K = 0 and wherever an edge is 0: the row then restarts at D_K)
__device__ __forceinline__ double far_g(const double *E, int K) {
  if (K == 0) return 0.0;
  const double a = E[K - 1];
  return a > 0.0 ? E[K] / a : 0.0;
}

// one step of the NT recurrences, f_n = g^(2n-1) generated as two interleaved chains
// (even n from 1/g, odd n from g, both stepped by g^4)
__device__ __forceinline__ void far_step(double (&M)[kFarNT], double g, double ginv, double D) {
  const double g2 = g * g, g4 = g2 * g2;
  double fe = ginv, fo = g;   // ginv = 1/g (0 where g = 0), formed once per edge by the fill
#pragma unroll
  for (int n = 0; n < kFarNT; n += 2) {
    M[n] = fma(M[n], fe, D);
    fe = fe * g4;
    if (n + 1 < kFarNT) {
      M[n + 1] = fma(M[n + 1], fo, D);
      fo = fo * g4;
    }
  }
}

// k_far_moments: one thread-block cluster of `csize` CTAs per problem; CTA `crank` owns the
// contiguous scan segments [seg0, seg1) of kFarSeg edges; thread <-> (segment, absorber).
//   pass 1   segment-local recurrences from zero -> B_s[X][n]; A_s[n] = prod g^(2n-1)
//   CTA composite (A, B) over its segments; exchanged through distributed shared memory
//   carry-in of the CTA from the CTAs before it, of the segment from the segments before it
//   pass 2   the recurrences again from the carry-in, every S-th row stored as b_n Mt_n[K]
// far: [B][Nl/S + 1][NT][3] doubles {HI, HeI, HeII}.
struct FarSmem {
  double4 *rec;     // [spc * kFarSeg] {g, D0, D1, D2}
  double *ginv;     // [spc * kFarSeg] 1/g
  double *sA;       // [spc][NT]
  double *sB;       // [spc][3][NT]
  double *cA;       // [NT]      CTA composite
  double *cB;       // [3][NT]
  double *cin;      // [3][NT]   CTA carry-in
};
__host__ __device__ inline size_t far_smem_bytes(int spc) {
  return (size_t)spc * kFarSeg * (sizeof(double4) + sizeof(double)) +
         ((size_t)spc * kFarNT * 4 + 7 * kFarNT) * sizeof(double);
}

constexpr int kFarThreads = 256;   // 3 * segments-per-CTA of them scan, all of them fill
__global__ void __launch_bounds__(kFarThreads) k_far_moments(PlanDev P, double *__restrict__ far,
                                                     int csize, int spc, int slog) {
  namespace cg = cooperative_groups;
  extern __shared__ double4 s_far4[];
  const int b = blockIdx.x / csize, crank = blockIdx.x % csize;
  const int Nl = P.Nl, nedge = Nl + 1;
  const int nseg = (nedge + kFarSeg - 1) / kFarSeg;
  const int seg0 = min(nseg, crank * spc), seg1 = min(nseg, seg0 + spc);
  const int nmine = seg1 - seg0;
  FarSmem S;
  S.rec = s_far4;
  S.ginv = reinterpret_cast<double *>(s_far4 + (size_t)spc * kFarSeg);
  S.sA = S.ginv + (size_t)spc * kFarSeg;
  S.sB = S.sA + (size_t)spc * kFarNT;
  S.cA = S.sB + (size_t)spc * 3 * kFarNT;
  S.cB = S.cA + kFarNT;
  S.cin = S.cB + 3 * kFarNT;
  const bool live = P.active[b] != 0;   // uniform over the cluster
  const double *E = P.Edges + (size_t)b * nedge;
  if (live) {
    const int k0 = seg0 * kFarSeg, k1 = min(nedge, seg1 * kFarSeg);
    for (int k = k0 + (int)threadIdx.x; k < k1; k += blockDim.x) {
      const double4 r = shell_record(P, b, k);
      const double g = far_g(E, k);
      S.rec[k - k0] = make_double4(g, r.y, r.z, r.w);
      S.ginv[k - k0] = g > 0.0 ? 1.0 / g : 0.0;
      // by-product: the shell records of the sweep in global memory, so that the column
      // kernel's CTAs copy them (16-byte loads) instead of each rebuilding all Nl + 1 of
      // them from six state arrays -- 148 CTAs x 4097 records x 11 loads at C4
      P.pack[(size_t)b * (nedge + kPackPad) + k] = r;
    }
    if (crank == csize - 1 && threadIdx.x < kPackPad)
      P.pack[(size_t)b * (nedge + kPackPad) + nedge + threadIdx.x] =
          make_double4(0.0, 0.0, 0.0, 0.0);
  }
  __syncthreads();
  const int t = threadIdx.x;
  const int sl = t / 3, X = t - 3 * sl;          // my segment (local index) and absorber
  const bool mine = live && sl < nmine;
  const int kb = (seg0 + sl) * kFarSeg, ke = min(nedge, kb + kFarSeg);   // [kb, ke)
  if (mine) {
    double M[kFarNT];
#pragma unroll
    for (int n = 0; n < kFarNT; ++n) M[n] = 0.0;
    double G = 1.0;
    for (int k = kb; k < ke; ++k) {
      const double4 r = S.rec[k - seg0 * kFarSeg];
      const double D = X == 0 ? r.y : (X == 1 ? r.z : r.w);
      far_step(M, r.x, S.ginv[k - seg0 * kFarSeg], D);
      G = G * r.x;
    }
#pragma unroll
    for (int n = 0; n < kFarNT; ++n) S.sB[(sl * 3 + X) * kFarNT + n] = M[n];
    if (X == 0) {   // A_s[n] = G^(2n-1)
      const double G2 = G * G;
      double f = G > 0.0 ? 1.0 / G : 0.0;
#pragma unroll
      for (int n = 0; n < kFarNT; ++n) { S.sA[sl * kFarNT + n] = f; f = f * G2; }
    }
  }
  __syncthreads();
  if (t < 3 * kFarNT) {   // CTA composite over its segments
    const int X2 = t / kFarNT, n = t - X2 * kFarNT;
    double ca = 1.0, cb = 0.0;
    for (int j = 0; j < nmine; ++j) {
      const double a = S.sA[j * kFarNT + n];
      cb = fma(cb, a, S.sB[(j * 3 + X2) * kFarNT + n]);
      ca = ca * a;
    }
    S.cB[X2 * kFarNT + n] = cb;
    if (X2 == 0) S.cA[n] = ca;
  }
  if (csize > 1) cg::this_cluster().sync(); else __syncthreads();
  if (t < 3 * kFarNT) {   // carry into this CTA from the CTAs before it
    const int X2 = t / kFarNT, n = t - X2 * kFarNT;
    // (the remote reads of all earlier CTAs are issued before the dependent chain starts)
    double ra[kMaxClusterFar], rb_[kMaxClusterFar];
#pragma unroll
    for (int c = 0; c < kMaxClusterFar; ++c) {
      ra[c] = 1.0; rb_[c] = 0.0;
      if (c < crank) {
        ra[c] = cg::this_cluster().map_shared_rank(S.cA, c)[n];
        rb_[c] = cg::this_cluster().map_shared_rank(S.cB, c)[X2 * kFarNT + n];
      }
    }
    double ci = 0.0;
#pragma unroll
    for (int c = 0; c < kMaxClusterFar; ++c)
      if (c < crank) ci = fma(ci, ra[c], rb_[c]);
    S.cin[X2 * kFarNT + n] = ci;
  }
  __syncthreads();
  if (mine) {
    double M[kFarNT];
#pragma unroll
    for (int n = 0; n < kFarNT; ++n) M[n] = S.cin[X * kFarNT + n];
    for (int j = 0; j < sl; ++j) {
#pragma unroll
      for (int n = 0; n < kFarNT; ++n)
        M[n] = fma(M[n], S.sA[j * kFarNT + n], S.sB[(j * 3 + X) * kFarNT + n]);
    }
    double *out = far + (size_t)b * far_rows(Nl, slog) * kFarRow + X;
    for (int k = kb; k < ke; ++k) {
      const double4 r = S.rec[k - seg0 * kFarSeg];
      const double D = X == 0 ? r.y : (X == 1 ? r.z : r.w);
      far_step(M, r.x, S.ginv[k - seg0 * kFarSeg], D);
      if ((k & ((1 << slog) - 1)) == 0) {
        RB_ASSERT((k >> slog) < far_rows(Nl, slog));
        double *row = out + (size_t)(k >> slog) * kFarRow;
#pragma unroll
        for (int n = 0; n < kFarNT; ++n) row[3 * n] = kFarB[n] * M[n];
      }
    }
  }
  if (csize > 1) cg::this_cluster().sync();   // peers may still read cA / cB
}

// k_far_moments_split: the same scan with the NT moments of a (segment, absorber) dealt over
// kFarParts threads.  In k_far_moments one thread carries all NT recurrences of its segment and
// absorber and regenerates the factors g^(2n-1) at every edge: 2 NT FP64 instructions per edge
// on 3 x segments threads -- two warps per CTA, 11 of the kernel's 24 us at C4 (ncu source
// page, profiles/r02z_far_moments_ncu.txt).  Here the fill forms the factors of every edge
// ONCE (the same two multiplication chains as far_step, so the same bits) into shared memory
// and thread (segment, absorber, part) advances NT / kFarParts recurrences by one fma each.
// Every M_n sees the same operations in the same order as in k_far_moments: the rows are
// bit-identical (tests/test_gpu_far_field.py compares the two kernels).
constexpr int kFarParts = 5;
constexpr int kFarPer = (kFarNT + kFarParts - 1) / kFarParts;   // moments per thread
constexpr int kFarFStride = kFarSeg * kFarNT + 2;   // doubles per segment of the factor table
__host__ __device__ inline size_t far_split_smem_bytes(int spc) {
  return far_smem_bytes(spc) + (size_t)spc * kFarFStride * sizeof(double);
}

__global__ void __launch_bounds__(kFarThreads) k_far_moments_split(PlanDev P,
                                                                  double *__restrict__ far,
                                                                  int csize, int spc, int slog) {
  namespace cg = cooperative_groups;
  extern __shared__ double4 s_far4[];
  const int b = blockIdx.x / csize, crank = blockIdx.x % csize;
  const int Nl = P.Nl, nedge = Nl + 1;
  const int nseg = (nedge + kFarSeg - 1) / kFarSeg;
  const int seg0 = min(nseg, crank * spc), seg1 = min(nseg, seg0 + spc);
  const int nmine = seg1 - seg0;
  FarSmem S;
  S.rec = s_far4;
  S.ginv = reinterpret_cast<double *>(s_far4 + (size_t)spc * kFarSeg);
  S.sA = S.ginv + (size_t)spc * kFarSeg;
  S.sB = S.sA + (size_t)spc * kFarNT;
  S.cA = S.sB + (size_t)spc * 3 * kFarNT;
  S.cB = S.cA + kFarNT;
  S.cin = S.cB + 3 * kFarNT;
  double *fT = S.cin + 3 * kFarNT;      // [spc][kFarSeg][NT] (+2 per segment: banks)
  const bool live = P.active[b] != 0;   // uniform over the cluster
  const double *E = P.Edges + (size_t)b * nedge;
  if (live) {
    const int k0 = seg0 * kFarSeg, k1 = min(nedge, seg1 * kFarSeg);
    for (int k = k0 + (int)threadIdx.x; k < k1; k += blockDim.x) {
      const double4 r = shell_record(P, b, k);
      const double g = far_g(E, k);
      const double ginv = g > 0.0 ? 1.0 / g : 0.0;
      S.rec[k - k0] = make_double4(g, r.y, r.z, r.w);
      P.pack[(size_t)b * (nedge + kPackPad) + k] = r;
      // the factors of far_step, generated by the same two chains
      double *f = fT + (size_t)((k - k0) / kFarSeg) * kFarFStride + ((k - k0) % kFarSeg) * kFarNT;
      const double g2 = g * g, g4 = g2 * g2;
      double fe = ginv, fo = g;
#pragma unroll
      for (int n = 0; n < kFarNT; n += 2) {
        f[n] = fe;
        fe = fe * g4;
        if (n + 1 < kFarNT) {
          f[n + 1] = fo;
          fo = fo * g4;
        }
      }
    }
    if (crank == csize - 1 && threadIdx.x < kPackPad)
      P.pack[(size_t)b * (nedge + kPackPad) + nedge + threadIdx.x] =
          make_double4(0.0, 0.0, 0.0, 0.0);
  }
  __syncthreads();
  const int t = threadIdx.x;
  const int nunit = live ? nmine * 3 * kFarParts : 0;   // (segment, absorber, part)
  for (int u = t; u < nunit; u += blockDim.x) {
    const int sl = u / (3 * kFarParts), rem = u - sl * 3 * kFarParts;
    const int X = rem / kFarParts, n0 = (rem - X * kFarParts) * kFarPer;
    const int kb = (seg0 + sl) * kFarSeg, ke = min(nedge, kb + kFarSeg);
    double M[kFarPer];
#pragma unroll
    for (int i = 0; i < kFarPer; ++i) M[i] = 0.0;
    const double *f = fT + (size_t)sl * kFarFStride + n0;
    for (int k = kb; k < ke; ++k, f += kFarNT) {
      const double4 r = S.rec[k - seg0 * kFarSeg];
      const double D = X == 0 ? r.y : (X == 1 ? r.z : r.w);
#pragma unroll
      for (int i = 0; i < kFarPer; ++i)
        if (n0 + i < kFarNT) M[i] = fma(M[i], f[i], D);
    }
#pragma unroll
    for (int i = 0; i < kFarPer; ++i)
      if (n0 + i < kFarNT) S.sB[(sl * 3 + X) * kFarNT + n0 + i] = M[i];
    if (X == 0 && n0 == 0) {   // A_s[n] = G^(2n-1)
      double G = 1.0;
      for (int k = kb; k < ke; ++k) G = G * S.rec[k - seg0 * kFarSeg].x;
      const double G2 = G * G;
      double a = G > 0.0 ? 1.0 / G : 0.0;
#pragma unroll
      for (int n = 0; n < kFarNT; ++n) { S.sA[sl * kFarNT + n] = a; a = a * G2; }
    }
  }
  __syncthreads();
  if (t < 3 * kFarNT) {   // CTA composite over its segments
    const int X2 = t / kFarNT, n = t - X2 * kFarNT;
    double ca = 1.0, cb = 0.0;
    for (int j = 0; j < nmine; ++j) {
      const double a = S.sA[j * kFarNT + n];
      cb = fma(cb, a, S.sB[(j * 3 + X2) * kFarNT + n]);
      ca = ca * a;
    }
    S.cB[X2 * kFarNT + n] = cb;
    if (X2 == 0) S.cA[n] = ca;
  }
  if (csize > 1) cg::this_cluster().sync(); else __syncthreads();
  if (t < 3 * kFarNT) {   // carry into this CTA from the CTAs before it
    const int X2 = t / kFarNT, n = t - X2 * kFarNT;
    double ra[kMaxClusterFar], rb_[kMaxClusterFar];
#pragma unroll
    for (int c = 0; c < kMaxClusterFar; ++c) {
      ra[c] = 1.0; rb_[c] = 0.0;
      if (c < crank) {
        ra[c] = cg::this_cluster().map_shared_rank(S.cA, c)[n];
        rb_[c] = cg::this_cluster().map_shared_rank(S.cB, c)[X2 * kFarNT + n];
      }
    }
    double ci = 0.0;
#pragma unroll
    for (int c = 0; c < kMaxClusterFar; ++c)
      if (c < crank) ci = fma(ci, ra[c], rb_[c]);
    S.cin[X2 * kFarNT + n] = ci;
  }
  __syncthreads();
  for (int u = t; u < nunit; u += blockDim.x) {
    const int sl = u / (3 * kFarParts), rem = u - sl * 3 * kFarParts;
    const int X = rem / kFarParts, n0 = (rem - X * kFarParts) * kFarPer;
    const int kb = (seg0 + sl) * kFarSeg, ke = min(nedge, kb + kFarSeg);
    double M[kFarPer], bn[kFarPer];
#pragma unroll
    for (int i = 0; i < kFarPer; ++i) {
      const bool on = n0 + i < kFarNT;
      M[i] = on ? S.cin[X * kFarNT + n0 + i] : 0.0;
      bn[i] = on ? kFarB[n0 + i] : 0.0;
    }
    for (int j = 0; j < sl; ++j) {
#pragma unroll
      for (int i = 0; i < kFarPer; ++i)
        if (n0 + i < kFarNT)
          M[i] = fma(M[i], S.sA[j * kFarNT + n0 + i], S.sB[(j * 3 + X) * kFarNT + n0 + i]);
    }
    double *out = far + (size_t)b * far_rows(Nl, slog) * kFarRow + X;
    const double *f = fT + (size_t)sl * kFarFStride + n0;
    for (int k = kb; k < ke; ++k, f += kFarNT) {
      const double4 r = S.rec[k - seg0 * kFarSeg];
      const double D = X == 0 ? r.y : (X == 1 ? r.z : r.w);
#pragma unroll
      for (int i = 0; i < kFarPer; ++i)
        if (n0 + i < kFarNT) M[i] = fma(M[i], f[i], D);
      if ((k & ((1 << slog) - 1)) == 0) {
        RB_ASSERT((k >> slog) < far_rows(Nl, slog));
        double *row = out + (size_t)(k >> slog) * kFarRow;
#pragma unroll
        for (int i = 0; i < kFarPer; ++i)
          if (n0 + i < kFarNT) row[3 * (n0 + i)] = bn[i] * M[i];
      }
    }
  }
  if (csize > 1) cg::this_cluster().sync();   // peers may still read cA / cB
}

// sum_{k<=K} s_k D_k for the three absorbers from table row K (E_K >= RATIO p)
__device__ __forceinline__ void far_eval(const double *__restrict__ row, double u, double EK,
                                         double (&out)[3]) {
  double h0 = 0.0, h1 = 0.0, h2 = 0.0;
#pragma unroll
  for (int n = kFarNT - 1; n >= 0; --n) {
    h0 = fma(h0, u, row[3 * n]);
    h1 = fma(h1, u, row[3 * n + 1]);
    h2 = fma(h2, u, row[3 * n + 2]);
  }
  out[0] = EK * h0; out[1] = EK * h1; out[2] = EK * h2;
}

struct ColAcc {
  double P0, P1, P2;     // running sum of s_k D_k (HI, HeI, HeII)
  double A0, A1, A2;     // its value before the own shell's edge (P_{il-1})
  double s_il, s_il1;    // roots at the own shell's two edges
};

// add shell edge k (record `rec`); SNAP: k may be the own shell's outer or inner edge
template <bool SNAP>
__device__ __forceinline__ void col_close(ColAcc &a, const double4 &rec, double s, int k, int il) {
  if (SNAP) {
    if (k == il) {
      a.A0 = a.P0; a.A1 = a.P1; a.A2 = a.P2;
      a.s_il = s;
    }
    if (k == il + 1) a.s_il1 = s;
  }
  a.P0 = fma(s, rec.y, a.P0);
  a.P1 = fma(s, rec.z, a.P1);
  a.P2 = fma(s, rec.w, a.P2);
}

// advance over shell edges k = kb .. ke (inclusive).  U records per trip (their U square
// roots are independent chains) with the next U prefetched.  The record array is padded by
// kPackPad entries, so the prefetch needs no bounds clamp.
template <bool SNAP, int U, bool SAFE>
__device__ __forceinline__ void col_advance(ColAcc &a, const double4 *__restrict__ pk, int kb,
                                            int ke, int il, double p2) {
  if (kb > ke) return;
  RB_ASSERT(kb >= 0);
  int k = kb;
  double4 n[U];
#pragma unroll
  for (int u = 0; u < U; ++u) n[u] = pk[k + u];
  for (; k + U - 1 <= ke; k += U) {
    double4 f[U];
#pragma unroll
    for (int u = 0; u < U; ++u) f[u] = pk[k + U + u];
    double sq[U];
#pragma unroll
    for (int u = 0; u < U; ++u) sq[u] = sqrt_pos<SAFE>(n[u].x - p2);
#pragma unroll
    for (int u = 0; u < U; ++u) col_close<SNAP>(a, n[u], sq[u], k + u, il);
#pragma unroll
    for (int u = 0; u < U; ++u) n[u] = f[u];
  }
#pragma unroll
  for (int u = 0; u < U - 1; ++u) {
    if (k + u <= ke) col_close<SNAP>(a, n[u], sqrt_pos<SAFE>(n[u].x - p2), k + u, il);
  }
}

// the walk over shell edges kb .. ke of one ray pair (see sphere_ray_pair); [kA, kB] is
// the warp-uniform window of edges that can be a lane's own-shell edge
template <int U, bool SAFE>
__device__ __forceinline__ void col_walk_from(ColAcc &a, const double4 *__restrict__ pk, int kb,
                                              int ke, int kA, int kB, int il, double p2) {
  col_advance<false, U, SAFE>(a, pk, kb, min(ke, kA - 1), il, p2);
  col_advance<true, U, SAFE>(a, pk, max(kb, kA), min(ke, kB), il, p2);
  col_advance<false, U, SAFE>(a, pk, max(kb, kB + 1), ke, il, p2);
}

template <int U, bool SAFE>
__device__ __forceinline__ void col_walk(ColAcc &a, const double4 *__restrict__ pk, int kb, int ke,
                                         int kA, int kB, int il, double p2) {
  a.P0 = a.P1 = a.P2 = 0.0;
  a.A0 = a.A1 = a.A2 = 0.0;
  a.s_il = a.s_il1 = 0.0;
  col_advance<false, U, SAFE>(a, pk, kb, min(ke, kA - 1), il, p2);
  col_advance<true, U, SAFE>(a, pk, max(kb, kA), min(ke, kB), il, p2);
  col_advance<false, U, SAFE>(a, pk, max(kb, kB + 1), ke, il, p2);
}

// +mu / -mu assembly of one ray pair from the sums of its walk (header comment of
// this section) and the store of its two ray records.  Pm = P_m, Pa = P_{il-1}.
// tau_X_th = N_X * sigma_X_th   (sphere_bgnd.f90:794-796); 4th slot: GL weight
__device__ __forceinline__ void store_ray_pair(const PlanDev &P, int b, int il, int ip, int m,
                                               double mu, const double (&Pm)[3],
                                               const double (&Pa)[3], double s_il,
                                               double s_il1, const double3 &c0,
                                               const double3 &c1) {
  const int imu_in = ip, imu_out = P.Nmu - 1 - ip;
  double Nout[3], Nin[3];
  if (il == m) {
#pragma unroll
    for (int X = 0; X < 3; ++X) Nout[X] = Nin[X] = Pm[X];
  } else {
    const double cm[3] = {c0.x, c0.y, c0.z}, cc[3] = {c1.x, c1.y, c1.z};
    const double ds = s_il - s_il1;
#pragma unroll
    for (int X = 0; X < 3; ++X) {
      const double A = Pa[X] - s_il * cm[X];
      const double d = ds * cc[X];
      Nout[X] = A + 0.5 * d;
      Nin[X] = 2.0 * Pm[X] - A - 0.5 * d;
    }
  }
  if (imu_in == imu_out) {  // odd Nmu: the mu ~ 0 node, both formulas agree
    const bool inward = (mu <= 0.0);
    P.col[col_index(P, b, il, imu_in)] =
        make_double4((inward ? Nin[0] : Nout[0]) * P.sigma_th[0],
                     (inward ? Nin[1] : Nout[1]) * P.sigma_th[1],
                     (inward ? Nin[2] : Nout[2]) * P.sigma_th[2], P.wmu[imu_in]);
  } else {
    P.col[col_index(P, b, il, imu_in)] = make_double4(
        Nin[0] * P.sigma_th[0], Nin[1] * P.sigma_th[1], Nin[2] * P.sigma_th[2], P.wmu[imu_in]);
    P.col[col_index(P, b, il, imu_out)] = make_double4(
        Nout[0] * P.sigma_th[0], Nout[1] * P.sigma_th[1], Nout[2] * P.sigma_th[2], P.wmu[imu_out]);
  }
}

// One (cell il, |mu| pair ip) of SphereBgnd: both rays' threshold optical depths.
// `pk` = shell records (global or shared), `il_w0` = cell of lane 0 of the warp.
// nchunk == 1: the whole walk, result stored as two ray records.  nchunk > 1: the
// walk is cut at the shell records c*Q (Q = ceil(Nl/nchunk)) and this call does
// chunk `chunk` only, leaving its partial sums {P, P before the own shell, the two
// own-shell roots, flags} in `part` for k_columns_combine -- finer work items for
// launches with few rays.
// The two searches of a ray pair on the monotone (non-increasing) edges.  They run on the
// squared edges of the shell records -- shared memory -- and are then VERIFIED on the edges
// themselves with two independent loads (rounding is monotone, so E_i <= p implies
// fl(E_i^2) <= fl(p^2), but two edges within an ulp of p can square to the same double); a
// failed check repeats the search on the edges.  Result: the reference's index, always, with
// 2 instead of ~12 dependent global loads per search (only ~20 KB of L1 remain beside the
// kernel's shared memory, so those loads mostly went to L2: the latency floor of a launch
// with few rays per warp).
// smallest i in [lo0, Nl] with E_i <= p, given E_Nl <= p (m_for_ray, geometry.f90:61-71)
__device__ __forceinline__ int first_edge_le(const double *__restrict__ E,
                                             const double4 *__restrict__ pk, int lo0, int Nl,
                                             double p, double p2) {
  int lo = lo0, hi = Nl;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (pk[mid].x <= p2) hi = mid; else lo = mid + 1;
  }
  const double e1 = E[lo], e0 = E[lo > lo0 ? lo - 1 : lo0];
  if ((lo == Nl || e1 <= p) && (lo == lo0 || e0 > p)) return lo;
  lo = lo0; hi = Nl;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (E[mid] <= p) hi = mid; else lo = mid + 1;
  }
  return lo;
}
// largest a in [0, m] with E_a >= rp, given E_0 >= rp
__device__ __forceinline__ int last_edge_ge(const double *__restrict__ E,
                                            const double4 *__restrict__ pk, int m, double rp,
                                            double rp2) {
  int a0 = 0, a1 = m;
  while (a0 < a1) {
    const int mid = (a0 + a1 + 1) >> 1;
    if (pk[mid].x >= rp2) a0 = mid; else a1 = mid - 1;
  }
  const double e0 = E[a0], e1 = E[a0 < m ? a0 + 1 : m];
  if (e0 >= rp && (a0 == m || e1 < rp)) return a0;
  a0 = 0; a1 = m;
  while (a0 < a1) {
    const int mid = (a0 + a1 + 1) >> 1;
    if (E[mid] >= rp) a0 = mid; else a1 = mid - 1;
  }
  return a0;
}

constexpr int kPartStride = 9;
template <int U>
__device__ __forceinline__ void sphere_ray_pair(const PlanDev &P, int b, int il, int il_w0,
                                                int cell_stride, int ip,
                                                const double4 *__restrict__ pk, bool store,
                                                int nchunk = 1, int chunk = 0,
                                                double *part = nullptr,
                                                const double *__restrict__ far = nullptr,
                                                int slog = 0) {
  const int Nl = P.Nl;
  const double *E = P.Edges + (size_t)b * (Nl + 1);

  const double mu = P.mu[ip];  // xi ascending: negative mu first
  const double r_ray = (E[il] + E[il + 1]) * 0.5;
  const double p_ray = r_ray * sqrt(1.0 - mu * mu);
  const double p2 = p_ray * p_ray;

  // m_for_ray (geometry.f90:61-71): smallest i >= 1 with E_i <= p, m = i-1.
  // E is non-increasing, so a bisection finds the same i as the linear search.
  const bool hollow = E[Nl] > p_ray;  // reference would use an undefined m
  if (hollow) {
    atomicMax(&P.err[b], RB_ERR_RAY_GEOMETRY);
    if (far == nullptr) return;
    store = false;   // (the far path is warp-collective: the lane stays, with an empty walk)
  }
  // the own shell's absorber densities (store_ray_pair): loaded now, used after the walk
  const double3 c_prev = shell_dens(P, b, il - 1), c_own = shell_dens(P, b, il);
  int lo = il + 1;  // E_il >= r_ray >= p: the answer is > il
  if (E[il] <= p_ray) lo = 1;  // degenerate (zero-width shells): full search
  lo = first_edge_le(E, pk, lo, Nl, p_ray, p2);
  const int m = hollow ? -1 : lo - 1;
  RB_ASSERT(m >= -1 && m < Nl && (hollow || m >= il) && il >= 0 && il < Nl);

  if (far != nullptr) {
    // ---- far field from the kept moment rows (shared memory), the rest walked.
    // Kf = last edge k <= m with E_k >= RATIO p (-1: none); Kc = last KEPT row at or above
    // it.  Called by all 32 lanes.
    const double rp = p_ray * RB_FAR_RATIO;
    int Kf = -1;
    if (m >= 0 && E[0] >= rp) Kf = last_edge_ge(E, pk, m, rp, rp * rp);
    const int smask = (1 << slog) - 1;
    const int Kc = Kf >= 0 ? (Kf & ~smask) : -1;
    RB_ASSERT(Kf <= m && Kc <= Kf && (Kc < 0 || (Kc >> slog) < far_rows(Nl, slog)));
    RB_ASSERT(Kc < 0 || E[Kc] >= rp);   // the series is only valid for u <= 1 / RATIO^2
    ColAcc a;
    a.P0 = a.P1 = a.P2 = 0.0;
    a.A0 = a.A1 = a.A2 = 0.0;
    a.s_il = a.s_il1 = 0.0;
    if (Kc >= 0) {
      double f[3];
      far_eval(far + (size_t)(Kc >> slog) * kFarRow, p2 / pk[Kc].x, E[Kc], f);
      a.P0 = f[0]; a.P1 = f[1]; a.P2 = f[2];
      if (il <= Kc) {
        // the own shell lies above the walk: its two roots directly, and the sum before it
        // from the kept row at or above edge il-1 plus the (< S) edges down to il-1
        a.s_il = sqrt_pos<true>(pk[il].x - p2);
        if (il + 1 <= Kc) a.s_il1 = sqrt_pos<true>(pk[il + 1].x - p2);
        if (il >= 1) {
          const int Ks = (il - 1) & ~smask;
          RB_ASSERT(Ks >= 0 && Ks <= il - 1 && E[Ks] >= rp);
          far_eval(far + (size_t)(Ks >> slog) * kFarRow, p2 / pk[Ks].x, E[Ks], f);
          for (int k = Ks + 1; k <= il - 1; ++k) {
            const double4 rec = pk[k];
            const double sq = sqrt_pos<true>(rec.x - p2);
            f[0] = fma(sq, rec.y, f[0]);
            f[1] = fma(sq, rec.z, f[1]);
            f[2] = fma(sq, rec.w, f[2]);
          }
          a.A0 = f[0]; a.A1 = f[1]; a.A2 = f[2];
        }
      }
    }
    // lanes start at different edges: the warp walks from the smallest start with the
    // others masked until their own (a short ragged zone), then all lanes together
    const int kmin = __reduce_min_sync(0xffffffffu, Kc), kmax = __reduce_max_sync(0xffffffffu, Kc);
    {
      int k = kmin + 1;
      for (; k + U - 1 <= kmax; k += U) {   // U independent roots per trip, closed in edge order
        double4 rec[U];
        double sq[U];
#pragma unroll
        for (int u = 0; u < U; ++u) rec[u] = pk[k + u];
#pragma unroll
        for (int u = 0; u < U; ++u) sq[u] = sqrt_pos<true>(rec[u].x - p2);   // (k > m: unused)
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (k + u > Kc && k + u <= m) col_close<true>(a, rec[u], sq[u], k + u, il);
      }
      for (; k <= kmax; ++k) {
        const double4 rec = pk[k];
        if (k > Kc && k <= m) col_close<true>(a, rec, sqrt_pos<true>(rec.x - p2), k, il);
      }
    }
    const int kA2 = il_w0, kB2 = il_w0 + 31 * cell_stride + 1;
    col_walk_from<U, false>(a, pk, kmax + 1, m, kA2, kB2, il, p2);
    if (!store) return;
    const double Pm2[3] = {a.P0, a.P1, a.P2}, Pa2[3] = {a.A0, a.A1, a.A2};
    store_ray_pair(P, b, il, ip, m, mu, Pm2, Pa2, a.s_il, a.s_il1, c_prev, c_own);
    return;
  }

  // shell edges of this call: k = kb .. ke, all of 0 .. m without chunks
  int kb = 0, ke = m;
  if (nchunk > 1) {
    const int Q = (Nl + nchunk) / nchunk;   // ceil((Nl + 1) / nchunk)
    kb = chunk * Q;
    ke = min(m, (chunk + 1) * Q - 1);
  }
  // The lanes of a warp hold cells il_w .. il_w + 31*stride: only edges in
  // [il_w, il_w + 31*stride + 1] can be a lane's own-shell edge, so the snapshot
  // test is confined to a warp-uniform sub-range of the loop.
  const int kA = il_w0;
  const int kB = il_w0 + 31 * cell_stride + 1;
  // The walk only takes square roots of E_k^2 - p^2 for k <= m, i.e. E_k >= E_m > p (m is
  // the last edge above the impact parameter).  Two distinct doubles a > b >= 0 differ by at
  // least 2^-52 relative, their squares by 2^-51 >= 2 ulp, so fl(a^2) > fl(b^2): the
  // argument is strictly positive and the seed needs no clamp inside the loop.
  ColAcc a;
  col_walk<U, false>(a, pk, kb, ke, kA, kB, il, p2);
  if (!store) return;
  const double Pm[3] = {a.P0, a.P1, a.P2}, Pa[3] = {a.A0, a.A1, a.A2};
  if (nchunk == 1) {
    store_ray_pair(P, b, il, ip, m, mu, Pm, Pa, a.s_il, a.s_il1, c_prev, c_own);
    return;
  }
  // chunk partials: the chunk sum, and what the own shell's edges left in this chunk
  const bool has_il = (il >= kb) && (il <= ke), has_il1 = (il + 1 >= kb) && (il + 1 <= ke);
#pragma unroll
  for (int X = 0; X < 3; ++X) { part[X] = Pm[X]; part[3 + X] = Pa[X]; }
  part[6] = a.s_il; part[7] = a.s_il1;
  part[8] = (has_il ? 1.0 : 0.0) + (has_il1 ? 2.0 : 0.0);
}

// Fallback for grids whose shell records do not fit in shared memory
// (Nl > ~7000): records read from global memory (k_pack_shells), one thread per
// (cell, |mu| pair).
template <int BLOCK, int U, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
k_sphere_columns(PlanDev P, int cell_start, int cell_stride, int n_local) {
  const int b = blockIdx.z;
  if (!P.active[b]) return;
  const int j = blockIdx.x * BLOCK + threadIdx.x;
  if (j >= n_local) return;
  const int il = cell_of(j, cell_start, cell_stride);
  const int lane = threadIdx.x & 31;
  const int ls = lane_stride_of(cell_stride);
  sphere_ray_pair<U>(P, b, il, il - lane * ls, ls, blockIdx.y,
                     P.pack + (size_t)b * (P.Nl + 1 + kPackPad), true);
}

// ---------------------------------------------------------------------------
// k_sphere_columns_smem: the ray-segment column kernel (reference hot loop A).
// A CTA first stages the problem's Nl+1 shell records {E^2, n_HI, n_HeI, n_HeII}
// in shared memory (coalesced loads of Edges, nH, nHe and the three fractions;
// this replaces the separate pack kernel), then its warps walk a list of work
// items -- (block of 32 consecutive local cells, |mu| pair) -- sorted by
// descending ray length on the host (longest-processing-time-first: the grid
// drains evenly although ray lengths span 1..Nl shells).  In the segment loop
// every lane of a warp reads the same record: a shared-memory broadcast instead
// of an L1 round trip per segment.
// grid = (CTAs per problem, B); items[it] = jblk | chunk << 12 | ip << 16.  With
// nchunk > 1 (few rays per launch: cell shards over many GPUs, small grids) every
// walk is cut into nchunk pieces that are separate items; k_columns_combine adds
// the pieces up in shell order and stores the ray records.
// ---------------------------------------------------------------------------
template <int THREADS, int U, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_sphere_columns_smem(PlanDev P, int cell_start, int cell_stride, int n_local,
                      const unsigned *__restrict__ items, int n_items, int *work, int nchunk,
                      double *__restrict__ part, const double *__restrict__ far, int slog) {
  const int b = blockIdx.y;
  if (!P.active[b]) return;
  extern __shared__ double4 s_pk[];
  const int nrec = P.Nl + 1 + kPackPad;
  if (far != nullptr) {   // records of this sweep as left by k_far_moments
    const double4 *g = P.pack + (size_t)b * nrec;
    for (int k = threadIdx.x; k < nrec; k += THREADS) s_pk[k] = g[k];
  } else {
    for (int k = threadIdx.x; k < nrec; k += THREADS) s_pk[k] = shell_record(P, b, k);
  }
  // kept far-field rows of this problem (k_far_moments), behind the shell records
  double *s_far = nullptr;
  if (far != nullptr) {
    s_far = reinterpret_cast<double *>(s_pk + nrec);
    const int nd = far_rows(P.Nl, slog) * kFarRow;
    const double *g = far + (size_t)b * nd;
    for (int i = threadIdx.x; i < nd; i += THREADS) s_far[i] = g[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // items are handed out by a per-problem counter (zeroed before the launch):
  // longest rays first, whichever warp is free takes the next one
  int it = 0, nxt = 0;
  if (lane == 0) it = atomicAdd(&work[b], 1);
  it = __shfl_sync(0xffffffffu, it, 0);
  while (it < n_items) {
    if (lane == 0) nxt = atomicAdd(&work[b], 1);  // in flight while this item is traced
    const unsigned item = items[it];
    const int jw = (int)(item & 0xfffu) * 32;
    const int chunk = (int)((item >> 12) & 0xfu);
    const int ip = (int)(item >> 16);
    const int j = jw + lane;
    RB_ASSERT(jw < n_local && ip < (P.Nmu + 1) / 2 && chunk < nchunk);
    const bool valid = j < n_local;
    const int il = cell_of(valid ? j : n_local - 1, cell_start, cell_stride);
    double *pp = nullptr;
    if (nchunk > 1)
      pp = part + ((((size_t)b * ((P.Nmu + 1) / 2) + ip) * n_local + (valid ? j : 0)) * nchunk +
                   chunk) * kPartStride;
    sphere_ray_pair<U>(P, b, il, cell_of(jw, cell_start, cell_stride),
                       lane_stride_of(cell_stride), ip, s_pk, valid,
                       nchunk, chunk, pp, s_far, slog);
    it = __shfl_sync(0xffffffffu, nxt, 0);
  }
}

// k_columns_combine: one thread per (local cell, |mu| pair): sums the chunk partials in
// shell order.  P_{il-1} = chunks before the one that holds the own shell's outer edge +
// that chunk's snapshot.
__global__ void k_columns_combine(PlanDev P, int cell_start, int cell_stride, int n_local,
                                  int nchunk, const double *__restrict__ part) {
  const int b = blockIdx.z;
  if (!P.active[b]) return;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int ip = blockIdx.y;
  if (j >= n_local) return;
  const int il = cell_of(j, cell_start, cell_stride);
  const int Nl = P.Nl;
  const double *E = P.Edges + (size_t)b * (Nl + 1);
  const double mu = P.mu[ip];
  const double r_ray = (E[il] + E[il + 1]) * 0.5;
  const double p_ray = r_ray * sqrt(1.0 - mu * mu);
  if (E[Nl] > p_ray) return;  // the walk kernel has raised RB_ERR_RAY_GEOMETRY
  int lo = il + 1, hi = Nl;
  if (E[il] <= p_ray) lo = 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (E[mid] <= p_ray) hi = mid; else lo = mid + 1;
  }
  const int m = lo - 1;
  const double *pp = part + (((size_t)b * ((P.Nmu + 1) / 2) + ip) * n_local + j) * nchunk *
                                kPartStride;
  double Pm[3] = {0.0, 0.0, 0.0}, Pa[3] = {0.0, 0.0, 0.0}, s_il = 0.0, s_il1 = 0.0;
  for (int c = 0; c < nchunk; ++c) {
    const double *q = pp + c * kPartStride;
    const int flags = (int)q[8];
    if (flags & 1) {   // P_{il-1} = chunks before + this chunk's sum up to the own shell
#pragma unroll
      for (int X = 0; X < 3; ++X) Pa[X] = Pm[X] + q[3 + X];
      s_il = q[6];
    }
    if (flags & 2) s_il1 = q[7];
#pragma unroll
    for (int X = 0; X < 3; ++X) Pm[X] = Pm[X] + q[X];
  }
  store_ray_pair(P, b, il, ip, m, mu, Pm, Pa, s_il, s_il1, shell_dens(P, b, il - 1),
                 shell_dens(P, b, il));
}

// ---------------------------------------------------------------------------
// warp / block scan helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_incl_scan(double v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// ---------------------------------------------------------------------------
// k_column_taus: one block per problem.  dtau_X[k] = dl_k n_k x_X,k sigma_X,th
// frozen from the pre-sweep state (Jacobi), then for every layer
//   lo = sum_{k<i} dtau_k ,  hi = sum_{k>i} dtau_k      (slab_base.f90:63-115)
// and the "rays" of the layer: tau_th = lo + dtau_i/2 (and hi + dtau_i/2 for the
// two-sided slabs).  The exclusive prefix is a block-wide shuffle scan over
// per-thread sequential chunks; hi = total - lo - dtau_i.
// ---------------------------------------------------------------------------
// exclusive block scan of three running sums; `wsum` is [3][BLOCK/32] scratch.  The
// exclusive value of a lane is the INCLUSIVE value of its left neighbour (shuffle), never
// "inclusive minus own": one optically thick layer (own value 1e6 against a prefix of 1e-2)
// would cancel 1e-16 x 1e6 into every thin layer behind it in the same warp / block.
template <int BLOCK>
__device__ __forceinline__ void block_excl_scan3(const double (&part)[3], double (&excl)[3],
                                                 double (*wsum)[BLOCK / 32]) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double inw[3];
#pragma unroll
  for (int X = 0; X < 3; ++X) {
    const double inc = warp_incl_scan(part[X]);
    const double left = __shfl_up_sync(0xffffffffu, inc, 1);
    inw[X] = (lane == 0) ? 0.0 : left;
    if (lane == 31) wsum[X][wid] = inc;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int X = 0; X < 3; ++X) {
      const double v = (lane < BLOCK / 32) ? wsum[X][lane] : 0.0;
      const double inc = warp_incl_scan(v);
      const double left = __shfl_up_sync(0xffffffffu, inc, 1);
      if (lane < BLOCK / 32) wsum[X][lane] = (lane == 0) ? 0.0 : left;
    }
  }
  __syncthreads();
#pragma unroll
  for (int X = 0; X < 3; ++X) excl[X] = inw[X] + wsum[X][wid];
  __syncthreads();  // wsum may be reused by the next scan
}

__device__ __forceinline__ void layer_dtau(const PlanDev &P, const double *E, int b, int k,
                                           double (&dt)[3]) {
  const size_t c = (size_t)b * P.Nl + k;
  const double dl = (P.geom == RB_GEOM_SPHERE_STROMGREN)
                        ? fabs(E[k] - E[k + 1])   // sphere_base.f90:1535
                        : (E[k + 1] - E[k]);      // slab_bgnd.f90:612
  const double dNH = dl * P.nH[c], dNHe = dl * P.nHe[c];
  dt[0] = dNH * P.x[0][c] * P.sigma_th[0];   // slab_bgnd.f90:632-634
  dt[1] = dNHe * P.x[2][c] * P.sigma_th[1];
  dt[2] = dNHe * P.x[3][c] * P.sigma_th[2];
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_column_taus(PlanDev P) {
  const int b = blockIdx.x;
  if (!P.active[b]) return;
  const int Nl = P.Nl;
  __shared__ double wsum[3][BLOCK / 32];
  const double *E = P.Edges + (size_t)b * (Nl + 1);
  const int per = (Nl + BLOCK - 1) / BLOCK;
  const int k0 = min(Nl, (int)threadIdx.x * per);
  const int k1 = min(Nl, k0 + per);
  const int npass = two_sided(P.geom) ? 2 : 1;
  // pass 0: ascending layer index  -> "lo" sums (layers below);
  // pass 1: descending layer index -> "hi" sums (layers above)
  for (int pass = 0; pass < npass; ++pass) {
    double part[3] = {0.0, 0.0, 0.0}, run[3], dt[3];
    for (int q = k0; q < k1; ++q) {
      const int k = pass ? (Nl - 1 - q) : q;
      layer_dtau(P, E, b, k, dt);
      part[0] += dt[0]; part[1] += dt[1]; part[2] += dt[2];
    }
    block_excl_scan3<BLOCK>(part, run, wsum);
    for (int q = k0; q < k1; ++q) {
      const int k = pass ? (Nl - 1 - q) : q;
      layer_dtau(P, E, b, k, dt);
      // slab_bgnd.f90:701-713; 4th slot: weight of the direction (1/2 + 1/2 for the
      // two-sided slabs, slab_bgnd.f90:718-747)
      P.col[col_index(P, b, k, pass)] = make_double4(run[0] + 0.5 * dt[0], run[1] + 0.5 * dt[1],
                                                     run[2] + 0.5 * dt[2], npass == 2 ? 0.5 : 1.0);
#pragma unroll
      for (int X = 0; X < 3; ++X) run[X] += dt[X];
    }
  }
}

// ---------------------------------------------------------------------------
// exp() of the hot loops.
// ---------------------------------------------------------------------------
constexpr int kExpTabCopies = 16;
// copies of the 2^L-entry table: 16 (a half-warp's lookups never conflict) up to L = 8;
// the 1024-entry table is kept 4 times (32 KB, ~2-way conflicts) -- see exp_fast
template <int L> struct ExpTab {
  static constexpr int copies = (L >= 10) ? 4 : kExpTabCopies;
  static constexpr int shift = (L >= 10) ? 5 : 7;   // byte offset of entry j: j * copies * 8
};

// exp(x) for -708 <= x <= 0 with a 2^L-entry table (L = 6: degree-5 polynomial,
// L = 8: degree 4), 8 / 7 FP64 instructions + 6 integer / load instructions:
//   t = x 2^L/ln2 + 1.5 2^52 ; n = low word of t = 2^L e + j ; r = x - n ln2/2^L
//   exp(x) = 2^e 2^(j/2^L) (1 + r + r^2/2 + ...)
// The reduction uses the rounded constant ln2/2^L in one fma (error |x| 1.1e-16
// relative to exp(x), i.e. below 1 ulp wherever exp(x) > 1e-3 and weighted by
// exp(-|x|) in every sum this feeds); the table is a static __shared__ array (its
// address is an immediate of the LDS), entry j copy c at double index 16 j + c,
// `lane_off` = 8 (lane % 16): a half-warp always hits 16 distinct 8-byte banks.
// Range: x is clamped to >= -708.0 by one unsigned integer min on its high word (for
// negative doubles the bit pattern grows with the magnitude); exp(-708) = 3e-308 then
// stands in for values the reference's libm returns as denormals or zero -- neither is
// visible in any sum this feeds.
// constants in constant memory: the loop then takes them as c[bank][offset] / uniform
// register operands instead of rebuilding 64-bit immediates with two moves each
__constant__ double kExpF6[7] = {92.33248261689366,          // 64/ln2
                                 -1.0830424696249145e-02,    // -ln2/64
                                 8.3333333333333332e-03, 4.1666666666666664e-02,
                                 1.6666666666666666e-01, 0.5, 6755399441055744.0};
// L = 10: |r| <= ln2/2048 = 3.4e-4, cubic 1 + r + r^2 (c2 + r/6) with c2 = 1/2 + (sqrt2 - 1)/12 a^2
// (equal-ripple against the dropped r^4/24: max error 9.4e-17), one FP64 instruction less
__constant__ double kExpF10[7] = {1477.3197218702985,        // 1024/ln2
                                  -0.0006769015435155716,    // -ln2/1024
                                  0.0, 0.0, 1.6666666666666666e-01, 0.5000000039539765,
                                  6755399441055744.0};
__constant__ double kExpF8[7] = {369.32993046757463,         // 256/ln2
                                 -2.7076061740622863e-03,    // -ln2/256
                                 0.0, 4.1666666666666664e-02,
                                 1.6666666666666666e-01, 0.5, 6755399441055744.0};

template <int L>
__device__ __forceinline__ double exp_fast(double x, const double *s_tab, unsigned lane_off) {
  const double *C = (L >= 10) ? kExpF10 : ((L >= 8) ? kExpF8 : kExpF6);
  x = __hiloint2double((int)min((unsigned)__double2hiint(x), 0xC0862000u), __double2loint(x));
  const double t = fma(x, C[0], C[6]);
  const int n = __double2loint(t);
  const double nd = t - C[6];
  const double r = fma(nd, C[1], x);
  double q;
  if (L >= 10) {
    q = C[4];
  } else if (L >= 8) {
    q = fma(r, C[3], C[4]);
  } else {
    q = fma(r, C[2], C[3]);
    q = fma(r, q, C[4]);
  }
  q = fma(r, q, C[5]);
  const double r2 = r * r;
  const double pl = fma(r2, q, r);
  constexpr int SH = ExpTab<L>::shift;
  const unsigned off = ((unsigned)(n << SH) & (unsigned)(((1 << L) - 1) << SH)) | lane_off;
  RB_ASSERT(off + 8u <= (unsigned)((1 << L) * ExpTab<L>::copies * 8) && x <= 0.0 && x >= -708.5);
  // The table holds 2^(j/2^L) with j 2^(20-L) SUBTRACTED from its high word, so that adding
  // n 2^(20-L) = (e << 20) + j 2^(20-L) (one imad, no mask) leaves the entry scaled by 2^e;
  // the scaling is applied BEFORE the last fma -- exact, x >= -708 keeps 2^e 2^(j/2^L) normal --
  // so the result is the bits of 2^e fma(T, pl, T), and nothing but that fma follows the
  // polynomial on the dependent chain (the lop3 + imad on the RESULT's high word that this
  // replaces cost 9 % of the frequency kernel: 0.499 -> 0.454 ms at C4).
  const double Ts = *reinterpret_cast<const double *>(reinterpret_cast<const char *>(s_tab) + off);
  int hi;
  asm("mad.lo.s32 %0, %1, %2, %3;"
      : "=r"(hi)
      : "r"(n), "n"(1 << (20 - L)), "r"(__double2hiint(Ts)));
  const double T = __hiloint2double(hi, __double2loint(Ts));
  return fma(T, pl, T);
}

template <int L>
__device__ __forceinline__ void fill_exp_table_L(double *s_exp) {
  for (int j = threadIdx.x; j < (1 << L); j += blockDim.x) {
    double v = exp2((double)j * (1.0 / (double)(1 << L)));
    v = __hiloint2double(__double2hiint(v) - (j << (20 - L)), __double2loint(v));   // see exp_fast
#pragma unroll
    for (int c = 0; c < ExpTab<L>::copies; ++c) s_exp[j * ExpTab<L>::copies + c] = v;
  }
}

// ---------------------------------------------------------------------------
// k_nu_rates: fused optical depth -> attenuation -> direction quadrature -> 6
// frequency integrals, for the cells il = cell_start + j*cell_stride
// (reference hot loop B: sphere_bgnd.f90:794-856, source_background.f90:259-446).
//
// Thread <-> KU frequencies, kept in registers as -sigma_X(nu)/sigma_X,th; the
// CTA (BLOCK threads = BLOCK*KU frequencies per pass) owns one cell at a time
// and walks its rays, whose records {tau_HI,th, tau_HeI,th, tau_HeII,th, w} are
// staged in shared memory (one contiguous run per cell) and read as broadcasts,
// one read serving all KU frequencies.  Per (ray, nu):
//     -tau = sum_X tau_X,th * (-ra_X)      (sphere_bgnd.f90:798-800, s_b.f90:267)
//     a    = exp(-tau)   or   E2(tau)      (s_b.f90:268 ; :385-388)
//     A(nu) += w_ray * a
// and once per (cell, nu):  rate_j += W_j(nu) * A(nu)   (s_b.f90:306-308,348-351 +
// utils.f90:19-20, trapezoid folded into W_j).  The reference integrates over
// nu per ray and then sums the rays (sphere_bgnd.f90:837-856); both sums are
// finite and linear, so they are exchanged here: the six weighted accumulations
// are paid per (cell, nu) instead of per (cell, ray, nu).  Then for SphereBgnd
//     src <- (src_old + sum)/2             (sphere_bgnd.f90:837-856).
// Algorithmic FP64 work per (ray, nu): 3 + 10 + 1 = 14 instructions.
// grid = a few waves of CTAs striding over the flat list of (problem, unit of cells).
// ---------------------------------------------------------------------------
// CPB = 1: the whole CTA works on one cell.  CPB > 1 (narrow frequency grids): every WARP of
// the CTA works on its own cell (BLOCK = 32 CPB) -- the warps only share the exp table, so a
// 1024-entry table serves 4 cells and the CTA count per SM stays low.
template <int ATT_E2, int KU, int RU, int BLOCK, int MINB, int LTAB, int CPB = 1>
__global__ void __launch_bounds__(BLOCK, MINB)
k_nu_rates(PlanDev P, int cell_start, int cell_stride, int n_local, int cpu) {
  static_assert(CPB == 1 || BLOCK == 32 * CPB, "one warp per cell");
  constexpr int CT = (CPB > 1) ? 32 : BLOCK;   // threads working on one cell
  extern __shared__ double4 s_dyn4[];
  const int Nnu = P.Nnu, Nl = P.Nl, ndir = P.Ndir;
  const int ndir_pad = (ndir + RU - 1) / RU * RU;
  __shared__ double s_exp[ATT_E2 ? 1 : (1 << LTAB) * ExpTab<LTAB>::copies];   // [2^LTAB][copies]
  __shared__ double s_red[6][CT / 32];
  if (!ATT_E2) fill_exp_table_L<LTAB>(s_exp);
  if (CPB > 1) __syncthreads();   // (CPB = 1: the first barrier of the cell loop covers it)
  const unsigned lane_off = (threadIdx.x & (ExpTab<LTAB>::copies - 1)) << 3;

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int tid = (CPB > 1) ? lane : (int)threadIdx.x;        // thread index within the cell
  double4 *s_ray = s_dyn4 + ((CPB > 1) ? wid * ndir_pad : 0);   // [ndir_pad] per cell in flight
  auto cell_sync = [&]() { if (CPB > 1) __syncwarp(); else __syncthreads(); };

  // the grid is one wave of CTAs striding over the flat list of (problem, local cell):
  // balanced to one cell whatever the batch size, also while problems drop out
  // (in batches a work unit is `cpu` consecutive cells of one problem, so that a CTA
  // re-reads one problem's frequency table from L1 instead of a new one from L2 per cell)
  const int upp = (n_local + cpu - 1) / cpu;   // units per problem
  const int n_units = P.B * upp;
  for (int u = blockIdx.x * CPB + ((CPB > 1) ? wid : 0); u < n_units; u += gridDim.x * CPB)
  for (int q = 0; q < cpu; ++q) {
    const int b = u / upp, j = (u % upp) * cpu + q;
    if (j >= n_local || !P.active[b]) continue;
    const double *spec = P.spec + (size_t)b * Nnu * kSpecStride;
    const int il = cell_of(j, cell_start, cell_stride);
    RB_ASSERT(il >= 0 && il < Nl && b < P.B);
    const size_t c = (size_t)b * Nl + il;
    cell_sync();  // previous cell's rays / reduction scratch are no longer read
    {
      const double4 *g = P.col + col_index(P, b, il, 0);
      // A ray whose optical depth exceeds 708 at EVERY frequency (lower bound: each species'
      // smallest cross-section ratio on the grid) contributes exp(-tau) < 3e-308 to every
      // integral: a denormal or zero in the reference's libm.  exp_fast clamps its argument
      // at -708 (3e-308 stands in), which is invisible next to any normal term but would be
      // the whole answer behind a wall of neutral gas (monochromatic sources): such rays
      // get weight 0, i.e. exactly the reference's zero.
      const double *th = P.thin + (size_t)b * kThinStride;
      const double m0 = th[6], m1 = th[7], m2 = th[8];
      for (int i = tid; i < ndir_pad; i += CT) {
        double4 r = (i < ndir) ? g[i] : make_double4(0.0, 0.0, 0.0, 0.0);  // pad: weight 0
        if (!ATT_E2 && r.x * m0 + r.y * m1 + r.z * m2 > 708.0) r.w = 0.0;
        s_ray[i] = r;
      }
    }
    cell_sync();
    double part[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int k0 = 0; k0 < Nnu; k0 += CT * KU) {
      double n0[KU], n1[KU], n2[KU], A[KU];
#pragma unroll
      for (int u = 0; u < KU; ++u) {
        const int k = k0 + u * CT + tid;
        const bool has = k < Nnu;
        const double *row = spec + (size_t)(has ? k : 0) * kSpecStride;
        n0[u] = has ? -row[0] : 0.0;  // -ra_X: the loop forms -tau
        n1[u] = has ? -row[1] : 0.0;
        n2[u] = has ? -row[2] : 0.0;
        A[u] = 0.0;
      }
#pragma unroll 1
      for (int i = 0; i < ndir_pad; i += RU) {
        double4 ry[RU];
#pragma unroll
        for (int r = 0; r < RU; ++r) ry[r] = s_ray[i + r];
#pragma unroll
        for (int r = 0; r < RU; ++r) {
#pragma unroll
          for (int u = 0; u < KU; ++u) {
            const double x = fma(ry[r].z, n2[u], fma(ry[r].y, n1[u], ry[r].x * n0[u]));  // -tau
            const double a = ATT_E2 ? e2xa(-x) : exp_fast<LTAB>(x, s_exp, lane_off);
            A[u] = fma(ry[r].w, a, A[u]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < KU; ++u) {
        const int k = k0 + u * CT + tid;
        if (k < Nnu) {
          const double *row = spec + (size_t)k * kSpecStride;
#pragma unroll
          for (int q = 0; q < 6; ++q) part[q] = fma(row[3 + q], A[u], part[q]);
        }
      }
    }
    // reduce over the CTA: shuffles inside a warp, shared memory across warps
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      double v = part[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      part[q] = v;
    }
    if (CT > 32) {
      if (lane == 0)
#pragma unroll
        for (int q = 0; q < 6; ++q) s_red[q][wid] = part[q];
      __syncthreads();
    }
    if (tid == 0) {
      double scale = 1.0;
      if (P.geom == RB_GEOM_SPHERE_STROMGREN) {
        // point source dilution 1/(4 pi r_c^2): source_point.f90:348
        const double *E = P.Edges + (size_t)b * (Nl + 1);
        const double r_c = (E[il] + E[il + 1]) * 0.5;
        scale = 1.0 / (4.0 * RB_PI * r_c * r_c);
      }
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        double v = part[q];
        if (CT > 32) {
          v = 0.0;
          for (int w = 0; w < CT / 32; ++w) v += s_red[q][w];
        }
        v *= scale;
        if (P.geom == RB_GEOM_SPHERE_BGND) v = (P.src[q][c] + v) * 0.5;
        P.src[q][c] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// k_cell_solve: one thread per local cell.  i_rec_meth == 1: total rates are
// the source rates (sphere_bgnd.f90:861-881).  Two-sided slabs compute the
// first half and mirror (slab_bgnd.f90:842-865).
// ---------------------------------------------------------------------------
// (128 registers / 4 CTAs per SM: capping at 96 / 80 / 64 registers for 5 / 6 / 8 CTAs
// measured 18.1 / 17.3 / 18.1 ms against 17.3 ms on a 2048-problem find_Teq batch)
__global__ void __launch_bounds__(128)
k_cell_solve(PlanDev P, PeerDev Q, int cell_start, int cell_stride, int n_local) {
  pow_tables_to_smem();
  const int b = blockIdx.y;
  if (!P.active[b]) return;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_local) return;
  const int il = cell_of(j, cell_start, cell_stride);
  const size_t c = (size_t)b * P.Nl + il;
  PhotoRates p;  // total = source + recombination photons (sphere_bgnd.f90:861-881)
  p.H1i = P.src[0][c] + P.rec[0][c]; p.He1i = P.src[1][c] + P.rec[1][c];
  p.He2i = P.src[2][c] + P.rec[2][c]; p.H1h = P.src[3][c] + P.rec[3][c];
  p.He1h = P.src[4][c] + P.rec[4][c]; p.He2h = P.src[5][c] + P.rec[5][c];
  const double nH = P.nH[c], nHe = P.nHe[c];
  rb_frac_t x;
  double T = P.T[c];
  int rc;
  if (P.find_Teq) {
    const CaseA fc = {P.fcA, P.fcA, P.fcA, P.libm_pow};
    rc = solve_pcte(nH, nHe, p, P.zHz[2 * b], P.zHz[2 * b + 1], fc, P.tol_inner, x, T, P.teq);
  } else {
    // get_kchem(T) of sphere_bgnd.f90:897-899: T is fixed, the coefficients were stored
    // by k_thin_state
    const size_t NC = (size_t)P.B * P.Nl;
    rb_kchem_t k;
    k.reH2 = P.kchem[c]; k.reHe2 = P.kchem[NC + c]; k.reHe3 = P.kchem[2 * NC + c];
    k.ciH1 = P.kchem[3 * NC + c]; k.ciHe1 = P.kchem[4 * NC + c]; k.ciHe2 = P.kchem[5 * NC + c];
    rc = solve_pce(nH, nHe, k, p.H1i, p.He1i, p.He2i, P.tol_inner, x);
  }
  if (rc) {
    atomicMax(&P.err[b], rc);
    return;
  }
  P.x[0][c] = x.xH1; P.x[1][c] = x.xH2; P.x[2][c] = x.xHe1;
  P.x[3][c] = x.xHe2; P.x[4][c] = x.xHe3;
  if (P.find_Teq) P.T[c] = T;
  if (Q.world > 1) peer_scatter(Q, j, x, T);
  if (two_sided(P.geom)) {
    const size_t cm = (size_t)b * P.Nl + (P.Nl - 1 - il);
    P.x[0][cm] = x.xH1; P.x[1][cm] = x.xH2; P.x[2][cm] = x.xHe1;
    P.x[3][cm] = x.xHe2; P.x[4][cm] = x.xHe3;
    P.T[cm] = T;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      P.src[q][cm] = P.src[q][c];
      P.rec[q][cm] = P.rec[q][c];  // slab_bgnd.f90:842-865
    }
  }
}

// ---------------------------------------------------------------------------
// k_cell_solve_pce_spec: solve_pce_s (ion_solver.f90:440-562) with a group of G
// lanes per cell (fixed-T runs).  The reference bisects n_e serially: ~25
// dependent evaluations of implicit_ionfracs (a dozen FP64 divisions each) --
// 35 us of pure latency per sweep whatever the number of cells.  As in
// k_cell_solve_spec the lanes of a group evaluate every n_e the bisection can
// reach in its next L = log2(G) steps (each lane replays the reference's
// mid-point arithmetic along its own path) and the group then walks the tree:
// a step needs the fractions of the node (for the stopping rule
// |max(x/x_old) - 1| <= tol, :507-512) and its direction bit ratio_ne > 1.
// Brackets, mid-points, stopping rule and result are the serial algorithm's bit
// for bit.
// ---------------------------------------------------------------------------
struct NeBisect {
  double left, right, ne;
};
__device__ __forceinline__ void nebisect_step(NeBisect &s, bool up) {
  if (up) {  // :515-517
    s.left = s.ne;
    s.ne = (s.ne + s.right) * 0.5;
  } else {   // :519-521
    s.right = s.ne;
    s.ne = (s.left + s.ne) * 0.5;
  }
}

template <int G>
__global__ void __launch_bounds__(128)
k_cell_solve_pce_spec(PlanDev P, PeerDev Q, int cell_start, int cell_stride, int n_local) {
  const int b = blockIdx.y;
  if (!P.active[b]) return;
  constexpr int L = (G == 32) ? 5 : (G == 16) ? 4 : (G == 8) ? 3 : 2;
  const int t = blockIdx.x * 128 + threadIdx.x;
  const int j = t / G, sub = t % G;
  const bool valid = j < n_local;
  const int il = cell_of(valid ? j : n_local - 1, cell_start, cell_stride);
  const int lane = threadIdx.x & 31, gbase = lane & ~(G - 1);
  const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
  const size_t c = (size_t)b * P.Nl + il;
  const size_t NC = (size_t)P.B * P.Nl;
  const double H1i = P.src[0][c] + P.rec[0][c], He1i = P.src[1][c] + P.rec[1][c],
               He2i = P.src[2][c] + P.rec[2][c];
  const double nH = P.nH[c], nHe = P.nHe[c];
  rb_kchem_t k;
  k.reH2 = P.kchem[c]; k.reHe2 = P.kchem[NC + c]; k.reHe3 = P.kchem[2 * NC + c];
  k.ciH1 = P.kchem[3 * NC + c]; k.ciHe1 = P.kchem[4 * NC + c]; k.ciHe2 = P.kchem[5 * NC + c];

  rb_frac_t xlast;   // fractions of the last executed iteration (the result)
  PceState st;
  int rc = pce_init(nH, nHe, k, H1i, xlast, st);   // uniform over the group
  NeBisect root = {st.ne_left, st.ne_right, st.ne};
  double o1 = st.xH1_old, o2 = st.xHe1_old, o3 = st.xHe3_old, err = st.err;
  int iter = 0;
  bool done = (rc != 0);
  while (!done) {
    // ---- my node of this round's tree (heap index, root = 1); lane 0 doubles the root
    const int node = sub == 0 ? 1 : sub;
    NeBisect mine = root;
    // the direction taken at an ancestor is that ancestor's own result: a lane can only
    // replay a path, so it evaluates the ancestors' bits as assumed by its heap index
    const int depth = 31 - __clz(node);
    for (int i = depth - 1; i >= 0; --i) nebisect_step(mine, ((node >> i) & 1) != 0);
    rb_frac_t x;
    implicit_ionfracs(mine.ne, k, H1i, He1i, He2i, x);
    const double ne_new = ne_of(nH, nHe, x);
    const bool up = ne_new > mine.ne;   // ratio_ne = ne_new / ne > 1 of :503 (rb_ion.cuh)
    const unsigned ups = __ballot_sync(mask, up) >> gbase;
    // ---- walk
    int n = 1, last = -1;
    for (int lev = 0; lev < L; ++lev) {
      if (!(err > P.tol_inner)) { done = true; break; }   // loop test of :495
      const double a1 = __shfl_sync(mask, x.xH1, gbase + n);
      const double a2 = __shfl_sync(mask, x.xHe1, gbase + n);
      const double a3 = __shfl_sync(mask, x.xHe3, gbase + n);
      err = ratio_err_exceeds(a1, o1, a2, o2, a3, o3, P.tol_inner)   // :507-512
                ? 1.7976931348623157e308 : 0.0;
      o1 = a1; o2 = a2; o3 = a3;
      const bool u = (ups >> n) & 1u;
      nebisect_step(root, u);
      last = n;
      n = 2 * n + (u ? 1 : 0);
      iter = iter + 1;
      if (iter > RB_ION_MAX_ITER) { rc = ION_ERR_MAX_ITER_PCE; done = true; break; }
    }
    if (last >= 0) {   // keep the fractions of the last executed iteration
      xlast.xH1 = __shfl_sync(mask, x.xH1, gbase + last);
      xlast.xH2 = __shfl_sync(mask, x.xH2, gbase + last);
      xlast.xHe1 = __shfl_sync(mask, x.xHe1, gbase + last);
      xlast.xHe2 = __shfl_sync(mask, x.xHe2, gbase + last);
      xlast.xHe3 = __shfl_sync(mask, x.xHe3, gbase + last);
    }
    if (!done && !(err > P.tol_inner)) done = true;
  }
  if (rc) {
    if (valid && sub == 0) atomicMax(&P.err[b], rc);
    return;
  }
  if (!valid || sub != 0) return;
  zero_absent_species(nH, nHe, xlast);
  P.x[0][c] = xlast.xH1; P.x[1][c] = xlast.xH2; P.x[2][c] = xlast.xHe1;
  P.x[3][c] = xlast.xHe2; P.x[4][c] = xlast.xHe3;
  if (Q.world > 1) peer_scatter(Q, j, xlast, P.T[c]);
  if (two_sided(P.geom)) {
    const size_t cm = (size_t)b * P.Nl + (P.Nl - 1 - il);
    P.x[0][cm] = xlast.xH1; P.x[1][cm] = xlast.xH2; P.x[2][cm] = xlast.xHe1;
    P.x[3][cm] = xlast.xHe2; P.x[4][cm] = xlast.xHe3;
    P.T[cm] = P.T[c];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      P.src[q][cm] = P.src[q][c];
      P.rec[q][cm] = P.rec[q][c];  // slab_bgnd.f90:842-865
    }
  }
}

// ---------------------------------------------------------------------------
// k_cell_solve_spec: solve_pcte_s (ion_solver.f90:729-880) with a group of G
// lanes per cell.  The reference bisects T serially: ~22 probes of
// get_dudT (each = 16 rate fits + a solve_pce), i.e. a dependent chain of
// ~3e5 instructions per cell -- pure latency when a sweep has only Nl cells.
// Here the lanes of a group evaluate, in one pass, every temperature the
// bisection can reach in its next L = log2(G) steps (the G-1 nodes of a binary
// tree, each lane replaying the reference's mid-point arithmetic along its own
// path), then the group walks the tree with the signs it has found.  Brackets,
// mid-points, stopping rule and the final fractions are those of the serial
// algorithm bit for bit; the bisection is ~L times shorter in wall time.
// Round 0 also carries the two bracket probes (lanes 0, 1).
// ---------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(128)
k_cell_solve_spec(PlanDev P, PeerDev Q, int cell_start, int cell_stride, int n_local) {
  pow_tables_to_smem();
  const int b = blockIdx.y;
  if (!P.active[b]) return;
  constexpr int L = (G == 32) ? 5 : (G == 16) ? 4 : (G == 8) ? 3 : 2;
  const int t = blockIdx.x * 128 + threadIdx.x;
  const int j = t / G, sub = t % G;
  const bool valid = j < n_local;
  const int il = cell_of(valid ? j : n_local - 1, cell_start, cell_stride);
  const int lane = threadIdx.x & 31, gbase = lane & ~(G - 1);
  const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
  const size_t c = (size_t)b * P.Nl + il;
  PhotoRates p;  // total = source + recombination photons (sphere_bgnd.f90:861-881)
  p.H1i = P.src[0][c] + P.rec[0][c]; p.He1i = P.src[1][c] + P.rec[1][c];
  p.He2i = P.src[2][c] + P.rec[2][c]; p.H1h = P.src[3][c] + P.rec[3][c];
  p.He1h = P.src[4][c] + P.rec[4][c]; p.He2h = P.src[5][c] + P.rec[5][c];
  const double nH = P.nH[c], nHe = P.nHe[c];
  const double z = P.zHz[2 * b], Hz = P.zHz[2 * b + 1];
  const CaseA fc = {P.fcA, P.fcA, P.fcA, P.libm_pow};
  const double tol = P.tol_inner;

  TBisect root;
  root.Tl = RB_TFLOOR; root.Tr = RB_TMAX;
  root.T = teq_root_T();   // :814-815
  root.err = 1.7976931348623157e308;
  const TeqTab tb = P.teq;
  int groot = tb.rows ? 1 : -1;   // heap index of root.T in the whole bisection tree (memo row)
  int rc = 0, iter = 0, src = 0;   // src: lane of the group holding the answer
  bool done = false;
  rb_frac_t x;
  double dudT = 0.0, Tfinal = root.Tl;
  int first = 1;
  while (!done) {
    // ---- my temperature for this round
    int node;          // heap index in this round's tree (root = 1), 0 = bracket probe
    const int nlev = first ? L - 1 : L;
    if (first) node = (sub >= 2 && sub - 1 < (1 << (L - 1))) ? sub - 1 : (sub < 2 ? 0 : 1);
    else node = sub == 0 ? 1 : sub;
    TBisect mine = root;
    if (node > 0) {
      const int depth = 31 - __clz(node);
      for (int i = depth - 1; i >= 0; --i) tbisect_step(mine, ((node >> i) & 1) == 0);
    }
    const double myT = (first && sub == 0) ? root.Tl : (first && sub == 1) ? root.Tr : mine.T;
    int row = -1;   // memo row of myT: the bracket rows, or my node below the round's root
    if (first && sub < 2) {
      row = sub == 0 ? 0 : teq_rows(tb);
    } else if (groot > 0) {
      const int depth = 31 - __clz(node);
      const long long gl = ((long long)groot << depth) | (long long)(node - (1 << depth));
      if (gl < (long long)teq_rows(tb)) row = (int)gl;
    }
    int prc = probe_T(nH, nHe, myT, p, z, Hz, fc, tol, dudT, x, tb, row);
    // ---- brackets (round 0): :771-812
    if (first) {
      const double dmin = __shfl_sync(mask, dudT, gbase + 0);
      const int rmin = __shfl_sync(mask, prc, gbase + 0);
      const double dmax = __shfl_sync(mask, dudT, gbase + 1);
      const int rmax = __shfl_sync(mask, prc, gbase + 1);
      if (rmin) { rc = rmin; done = true; }
      else if (dmin < 0.0) { done = true; src = 0; Tfinal = root.Tl; }   // :782-794 floor
      else if (rmax) { rc = rmax; done = true; }
      else if (dmax > 0.0) { rc = ION_ERR_T_NOT_BRACKETED; done = true; }
    }
    // ---- walk the tree with the signs found (uniform over the group)
    const int lane0 = first ? 1 : 0;   // node n was probed by lane n + lane0 (node 1 also by lane 0)
    int n = 1;
    for (int lev = 0; lev < nlev && !done; ++lev) {
      if (!(root.err > tol)) {   // loop test of :820: T = root.T has been probed by node n
        done = true; src = n + lane0; Tfinal = root.T;
        break;
      }
      const double d = __shfl_sync(mask, dudT, gbase + n + lane0);
      const int r = __shfl_sync(mask, prc, gbase + n + lane0);
      if (r) { rc = r; done = true; break; }
      const bool neg = d < 0.0;
      tbisect_step(root, neg);
      groot = teq_child(tb, groot, neg);
      n = 2 * n + (neg ? 0 : 1);
      iter = iter + 1;
      if (iter > RB_ION_MAX_ITER) { rc = ION_ERR_MAX_ITER_PCTE; done = true; }
    }
    first = 0;
  }
  if (rc) {
    if (valid && sub == 0) atomicMax(&P.err[b], rc);
    return;
  }
  rb_frac_t xo;
  xo.xH1 = __shfl_sync(mask, x.xH1, gbase + src);
  xo.xH2 = __shfl_sync(mask, x.xH2, gbase + src);
  xo.xHe1 = __shfl_sync(mask, x.xHe1, gbase + src);
  xo.xHe2 = __shfl_sync(mask, x.xHe2, gbase + src);
  xo.xHe3 = __shfl_sync(mask, x.xHe3, gbase + src);
  if (!valid || sub != 0) return;
  P.x[0][c] = xo.xH1; P.x[1][c] = xo.xH2; P.x[2][c] = xo.xHe1;
  P.x[3][c] = xo.xHe2; P.x[4][c] = xo.xHe3;
  P.T[c] = Tfinal;
  if (Q.world > 1) peer_scatter(Q, j, xo, Tfinal);
  if (two_sided(P.geom)) {
    const size_t cm = (size_t)b * P.Nl + (P.Nl - 1 - il);
    P.x[0][cm] = xo.xH1; P.x[1][cm] = xo.xH2; P.x[2][cm] = xo.xHe1;
    P.x[3][cm] = xo.xHe2; P.x[4][cm] = xo.xHe3;
    P.T[cm] = Tfinal;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      P.src[q][cm] = P.src[q][c];
      P.rec[q][cm] = P.rec[q][c];  // slab_bgnd.f90:842-865
    }
  }
}

// ---------------------------------------------------------------------------
// k_converge_finish: one block per problem.  n_e = xH2 nH + (xHe2 + 2 xHe3) nHe
// over ALL cells; converged when all(|ne/ne_old - 1| < tol)
// (sphere_bgnd.f90:289-296); then the iteration bookkeeping of the drivers
// (sphere_bgnd.f90:284-288; slab_bgnd.f90:278-282).  The number of problems still
// active is accumulated in nactive[slot]; the next slot is cleared for the
// following sweep, so the host can enqueue sweeps ahead and poll later: once a
// problem is inactive every kernel returns immediately for it.
// ---------------------------------------------------------------------------
constexpr int kSlots = 8;

// change of n_e in one cell (sphere_bgnd.f90:289-296); updates the reference value
__device__ __forceinline__ void cell_change(const PlanDev &P, size_t c, int &notconv,
                                            double &worst) {
  const double ne = P.x[1][c] * P.nH[c] + (P.x[3][c] + 2.0 * P.x[4][c]) * P.nHe[c];
  const double change = fabs(ne / P.conv_old[c] - 1.0);
  if (!(change < P.tol)) notconv = 1;
  if (change > worst) worst = change;  // NaN never wins: it only sets notconv
  P.conv_old[c] = ne;
}

// iteration bookkeeping of the drivers for problem b, by one thread
__device__ __forceinline__ void finish_bookkeeping(const PlanDev &P, int b, int slot, int notconv,
                                                   double worst) {
  P.resid_bits[b] = (unsigned long long)__double_as_longlong(worst);
  int act = 1;
  if (P.err[b]) act = 0;
  // sphere_bgnd.f90:284-288: the limit is tested before the increment
  if (P.geom != RB_GEOM_SLAB_BGND && P.iters[b] > RB_OUTER_MAX_ITER && P.max_sweeps >= 0) {
    atomicMax(&P.err[b], RB_ERR_MAX_ITER_OUTER);
    act = 0;
  }
  P.iters[b] += 1;
  // max_sweeps < 0: fixed-work mode (benchmarks): exactly -max_sweeps sweeps
  if (!notconv && P.max_sweeps >= 0) act = 0;
  if (P.geom == RB_GEOM_SLAB_BGND && P.iters[b] >= RB_SLAB_BGND_MAX_SWEEPS &&
      P.max_sweeps >= 0) act = 0;
  if (P.max_sweeps > 0 && P.iters[b] >= P.max_sweeps) act = 0;
  if (P.max_sweeps < 0 && P.iters[b] >= -P.max_sweeps) act = 0;
  P.active[b] = act;
  if (act) atomicAdd(&P.nactive[slot], 1);
}

template <int BLOCK>
__device__ __forceinline__ void converge_finish(const PlanDev &P, int slot) {
  const int b = blockIdx.x;
  if (b == 0 && threadIdx.x == 0) P.nactive[(slot + 1) % kSlots] = 0;
  if (!P.active[b]) return;
  __shared__ double s_max[BLOCK / 32];
  int notconv = 0;
  double worst = 0.0;
  for (int il = threadIdx.x; il < P.Nl; il += BLOCK)
    cell_change(P, (size_t)b * P.Nl + il, notconv, worst);
  notconv = __syncthreads_or(notconv);
  for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_down_sync(0xffffffffu, worst, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = worst;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < BLOCK / 32; ++w) worst = fmax(worst, s_max[w]);
    finish_bookkeeping(P, b, slot, notconv, worst);
  }
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_converge_finish(PlanDev P, int slot) {
  converge_finish<BLOCK>(P, slot);
}

// k_exchange_finish: the cell-sharded sweep's barrier + scatter + convergence test
// (single-problem plans), spread over a few CTAs.  CTA 0 publishes "my cells of sweep e
// are staged everywhere" (system-scope release after the cell kernel's remote stores);
// every CTA acquires the same from every peer, scatters its slice of the staged fractions
// into the full state arrays and tests the cells it scattered; the CTA that arrives last
// does the bookkeeping and advances the device-resident sweep counter.  A peer that never
// arrives trips the timeout: the rank raises RB_ERR_PEER_TIMEOUT, skips the scatter and
// poisons its flag on every peer, so that all ranks abort the solve instead of iterating
// on half-written staging (the plan is unusable afterwards).
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_exchange_finish(PlanDev P, PeerDev Q, int slot) {
  __shared__ int s_bad;
  __shared__ double s_max[BLOCK / 32];
  const unsigned long long e = peer_sweep_id(Q);   // read by every CTA before it arrives
  if (threadIdx.x == 0) {
    int bad = 0;
    if (blockIdx.x == 0) {
      P.nactive[(slot + 1) % kSlots] = 0;
      __threadfence_system();
      for (int r = 0; r < Q.world; ++r)
        if (r != Q.rank)
          asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(Q.flags[r] + Q.rank), "l"(e)
                       : "memory");
    }
    const unsigned long long t0 = global_ns();
    for (int r = 0; r < Q.world && !bad; ++r) {
      if (r == Q.rank) continue;
      for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(Q.flags[Q.rank] + r)
                     : "memory");
        if (v == kPeerPoison) { bad = 1; break; }
        if (v >= e) break;
        if (global_ns() - t0 > Q.timeout_ns) { bad = 1; break; }
      }
    }
    if (bad) {
      atomicMax(&P.err[0], RB_ERR_PEER_TIMEOUT);
      for (int r = 0; r < Q.world; ++r)
        if (r != Q.rank)
          asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(Q.flags[r] + Q.rank),
                       "l"(kPeerPoison) : "memory");
    }
    s_bad = bad;
  }
  __syncthreads();
  int notconv = 0;
  double worst = 0.0;
  if (P.active[0] && !s_bad) {
    const size_t n = (size_t)Q.n_local;
    const double *st = Q.stage[Q.rank] + (size_t)(e & 1ull) * Q.world * kPeerFields * n;
    const int ncell = Q.world * Q.n_local;
    const bool two = two_sided(P.geom);
    for (int i = blockIdx.x * BLOCK + threadIdx.x; i < ncell; i += gridDim.x * BLOCK) {
      const int g = i / Q.n_local, j = i - g * Q.n_local;
      const int il = cell_of(j, g, Q.deal);
#pragma unroll
      for (int f = 0; f < kPeerFields; ++f) {
        const double v = __ldcv(st + ((size_t)g * kPeerFields + f) * n + j);
        double *dst = (f < 5) ? P.x[f] : P.T;
        dst[il] = v;
        if (two) dst[P.Nl - 1 - il] = v;  // mirrored half
      }
      cell_change(P, (size_t)il, notconv, worst);
      if (two) cell_change(P, (size_t)(P.Nl - 1 - il), notconv, worst);
    }
    // the never-solved middle layer of a two-sided slab with odd Nl
    if (two && (P.Nl & 1) && blockIdx.x == 0 && threadIdx.x == 0)
      cell_change(P, (size_t)(P.Nl / 2), notconv, worst);
  }
  notconv = __syncthreads_or(notconv);
  for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_down_sync(0xffffffffu, worst, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = worst;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < BLOCK / 32; ++w) worst = fmax(worst, s_max[w]);
    unsigned long long *ctl = Q.ctl;
    if (notconv) atomicAdd(ctl + 2, 1ull);
    atomicMax(ctl + 3, (unsigned long long)__double_as_longlong(worst));  // worst >= 0
    __threadfence();
    const unsigned long long arrived = atomicAdd(ctl + 1, 1ull);
    if (arrived == (unsigned long long)gridDim.x - 1ull) {   // last CTA of this rank
      __threadfence();
      const unsigned long long nc = atomicAdd(ctl + 2, 0ull);
      const unsigned long long wb = atomicAdd(ctl + 3, 0ull);
      if (P.active[0])
        finish_bookkeeping(P, 0, slot, nc != 0ull, __longlong_as_double((long long)wb));
      ctl[1] = 0ull; ctl[2] = 0ull; ctl[3] = 0ull;
      __threadfence();
      *(volatile unsigned long long *)ctl = e;   // sweep e is complete on this rank
    }
  }
}

// ---------------------------------------------------------------------------
// cell-shard exchange helpers (multi-GPU Jacobi): 18 state fields
// (5 fractions, T, 6 source rates, 6 recombination rates) of the local cells
// il = start + j*stride are packed as [18][n_local]; after the all-gather the buffer is
// [G][18][n_local]
// and cell il = j*G + g comes from rank g.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double *plan_field(const PlanDev &P, int f) {
  return f < 5 ? P.x[f] : (f == 5 ? P.T : (f < 12 ? P.src[f - 6] : P.rec[f - 12]));
}

__global__ void k_pack_cells(PlanDev P, int cell_start, int cell_stride, int n_local,
                             double *send) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  if (j >= n_local) return;
  send[(size_t)f * n_local + j] = plan_field(P, f)[cell_of(j, cell_start, cell_stride)];
}

__global__ void k_unpack_cells(PlanDev P, int deal, int n_local, const double *recv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y, g = blockIdx.z;
  if (j >= n_local) return;
  const int il = cell_of(j, g, deal);
  const double v = recv[((size_t)g * gridDim.y + f) * n_local + j];
  double *dst = plan_field(P, f);
  dst[il] = v;
  if (two_sided(P.geom)) dst[P.Nl - 1 - il] = v;  // mirrored half
}

// k_gather_outputs: all problems' results in INPUT orientation, field-major
// out[f][b][i], f = xH1, xH2, xHe1, xHe2, xHe3, T, 6 source rates (, 6 recombination rates):
// one device pass + one copy instead of 18 small copies per problem (rb_plan_get_all)
__global__ void k_gather_outputs(PlanDev P, const unsigned char *__restrict__ reversed,
                                 int nfields, double *__restrict__ out) {
  const int b = blockIdx.y, f = blockIdx.z;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.Nl || f >= nfields) return;
  const double *src = plan_field(P, f) + (size_t)b * P.Nl;
  const int k = reversed[b] ? P.Nl - 1 - i : i;
  out[((size_t)f * P.B + b) * P.Nl + i] = src[k];
}

// ---------------------------------------------------------------------------
// lock-step ("_v") solvers, one block per problem, CPT cells per thread.
// ---------------------------------------------------------------------------
// any() of the lock-step loops (`any(err > tol)`, ion_solver.f90:620,988) over all
// the threads that share one problem.  BlockAny: one CTA (on the host, where the
// loop logic is unit-tested, a "block" is one thread).  ClusterAny: the CTAs of a
// thread-block cluster exchange their block results through distributed shared
// memory -- every CTA writes its flag into every peer's `flags` and the cluster
// barrier orders the writes; the two flag rows alternate so that a CTA that is
// already in the next call cannot overwrite a row a slower peer still reads.
struct BlockAny {
  __host__ __device__ __forceinline__ int operator()(int v) {
#ifdef __CUDA_ARCH__
    return __syncthreads_or(v);
#else
    return v;
#endif
  }
};

constexpr int kMaxCluster = 8;
struct ClusterAny {
  int *flags;  // shared memory, [2][kMaxCluster]
  int size, rank, par;
  __device__ __forceinline__ int operator()(int v) {
    int any = __syncthreads_or(v);
    if (size == 1) return any;
    cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
    if ((int)threadIdx.x < size)
      *cl.map_shared_rank(flags + par * kMaxCluster + rank, threadIdx.x) = any;
    cl.sync();
    any = 0;
    for (int q = 0; q < size; ++q) any |= flags[par * kMaxCluster + q];
    par ^= 1;
    return any;
  }
};

template <int CPT, class ANY>
__host__ __device__ int pce_lockstep(const double (&nH)[CPT], const double (&nHe)[CPT],
                            const rb_kchem_t (&k)[CPT], const PhotoRates (&p)[CPT],
                            const rb_flag_t (&valid)[CPT], double tol,
                            rb_frac_t (&x)[CPT], ANY &block_any) {
  PceState s[CPT];
  rb_flag_t ok[CPT];
  int rc = 0;
RB_LS_UNROLL
  for (int i = 0; i < CPT; ++i) {
    ok[i] = valid[i];
    if (ok[i]) {
      const int r = pce_init(nH[i], nHe[i], k[i], p[i].H1i, x[i], s[i]);
      if (r) { rc = r; ok[i] = false; }
    }
  }
  int iter = 0;
  for (;;) {
    int any = 0;
RB_LS_UNROLL
    for (int i = 0; i < CPT; ++i)
      if (ok[i] && s[i].err > tol) any = 1;
    if (!block_any(any)) break;  // ion_solver.f90:620
RB_LS_UNROLL
    for (int i = 0; i < CPT; ++i)
      if (ok[i]) pce_step(nH[i], nHe[i], k[i], p[i].H1i, p[i].He1i, p[i].He2i, x[i], s[i], true, tol);
    iter = iter + 1;
    if (iter > RB_ION_MAX_ITER) { rc = ION_ERR_MAX_ITER_PCE; break; }
  }
RB_LS_UNROLL
  for (int i = 0; i < CPT; ++i)
    if (ok[i]) zero_absent_species(nH[i], nHe[i], x[i]);
  return rc;
}

template <int CPT, class ANY>
__host__ __device__ int dudT_lockstep(const double (&nH)[CPT], const double (&nHe)[CPT],
                             const double (&T)[CPT], const PhotoRates (&p)[CPT],
                             const rb_flag_t (&valid)[CPT], double z, double Hz,
                             const CaseA &fc, double tol, double (&dudT)[CPT], ANY &block_any,
                             const TeqTab &tb, const int (&row)[CPT]) {
  rb_kchem_t kc[CPT];
  rb_frac_t x[CPT];
RB_LS_UNROLL
  for (int i = 0; i < CPT; ++i)
    if (valid[i]) kc[i] = kchem_memo(tb, row[i], T[i], fc);
  const int rc = pce_lockstep<CPT>(nH, nHe, kc, p, valid, tol, x, block_any);
RB_LS_UNROLL
  for (int i = 0; i < CPT; ++i) {
    if (valid[i]) {
      const rb_kcool_t kl = kcool_memo(tb, row[i], T[i], fc);
      const double heat = get_heat(nH[i], nHe[i], p[i].H1h, p[i].He1h, p[i].He2h,
                                   x[i].xH1, x[i].xHe1, x[i].xHe2);
      const double cool = get_cool(nH[i], nHe[i], T[i], kl, x[i], z, Hz);
      dudT[i] = heat - cool;
    } else {
      dudT[i] = 0.0;
    }
  }
  return rc;
}

// ion_solver.f90:884-1056 (solve_pcte_v)
template <int CPT, class ANY>
__host__ __device__ int pcte_lockstep(const double (&nH)[CPT], const double (&nHe)[CPT],
                             const PhotoRates (&p)[CPT], const rb_flag_t (&valid)[CPT],
                             double z, double Hz, const CaseA &fc, double tol,
                             rb_frac_t (&x)[CPT], double (&T)[CPT], ANY &block_any,
                             const TeqTab &tb) {
  double Tleft[CPT], Tright[CPT], dudT[CPT], err[CPT];
  rb_flag_t floor_[CPT];
  int node[CPT];   // memo row of the temperature in flight (rb_ion.cuh: TeqTab)
  int rc = 0, r;
RB_LS_UNROLL
  for (int i = 0; i < CPT; ++i) { Tleft[i] = RB_TFLOOR; Tright[i] = RB_TMAX; node[i] = 0; }
  r = dudT_lockstep<CPT>(nH, nHe, Tleft, p, valid, z, Hz, fc, tol, dudT, block_any, tb, node);
  if (r) rc = r;
RB_LS_UNROLL
  for (int i = 0; i < CPT; ++i) floor_[i] = valid[i] && (dudT[i] < 0.0);
RB_LS_UNROLL
  for (int i = 0; i < CPT; ++i) node[i] = teq_rows(tb);
  r = dudT_lockstep<CPT>(nH, nHe, Tright, p, valid, z, Hz, fc, tol, dudT, block_any, tb, node);
  if (r) rc = r;
RB_LS_UNROLL
  for (int i = 0; i < CPT; ++i) {
    if (valid[i] && !floor_[i] && dudT[i] > 0.0) rc = ION_ERR_T_NOT_BRACKETED;
    if (floor_[i]) {
      T[i] = Tleft[i];
      node[i] = 0;
    } else {
      T[i] = teq_root_T();   // ion_solver.f90:978-979
      node[i] = tb.rows ? 1 : -1;
    }
    err[i] = 1.7976931348623157e308;
  }
  int iter = 0;
  for (;;) {
    int any = 0;
RB_LS_UNROLL
    for (int i = 0; i < CPT; ++i)
      if (valid[i] && err[i] > tol) any = 1;
    if (!block_any(any)) break;  // ion_solver.f90:988
    r = dudT_lockstep<CPT>(nH, nHe, T, p, valid, z, Hz, fc, tol, dudT, block_any, tb, node);
    if (r) rc = r;
RB_LS_UNROLL
    for (int i = 0; i < CPT; ++i) {
      const double Told = T[i];
      if (dudT[i] < 0.0 && !floor_[i]) {
        Tright[i] = T[i];
        T[i] = (Tleft[i] + T[i]) * 0.5;
        node[i] = teq_child(tb, node[i], true);
      } else if (dudT[i] > 0.0 && !floor_[i]) {
        Tleft[i] = T[i];
        T[i] = (T[i] + Tright[i]) * 0.5;
        node[i] = teq_child(tb, node[i], false);
      }
      err[i] = fabs(1.0 - T[i] / Told);
    }
    iter = iter + 1;
    if (iter > RB_ION_MAX_ITER) { rc = ION_ERR_MAX_ITER_PCTE; break; }
  }
  rb_kchem_t kc[CPT];
RB_LS_UNROLL
  for (int i = 0; i < CPT; ++i)
    if (valid[i]) kc[i] = kchem_memo(tb, node[i], T[i], fc);
  r = pce_lockstep<CPT>(nH, nHe, kc, p, valid, tol, x, block_any);
  if (r) rc = r;
  return rc;
}

// k_thin_state: optically thin rates, lock-step solve, n_e reference for the
// convergence test (sphere_bgnd.f90:260-262).  One thread-block cluster of
// `csize` CTAs per problem (grid = B * csize); the `_v` solvers' any() spans the
// cluster (ClusterAny), so that up to 8 x 512 cells iterate in lock step with
// one cell per thread.
template <int CPT>
__global__ void __launch_bounds__(512) k_thin_state(PlanDev P, int csize) {
  pow_tables_to_smem();
  const int b = blockIdx.x / csize, crank = blockIdx.x % csize;
  const int Nl = P.Nl;
  __shared__ int s_flags[2 * kMaxCluster];
  ClusterAny any_fn = {s_flags, csize, crank, 0};
  const int tid = crank * blockDim.x + threadIdx.x;   // thread index within the problem
  const int tstride = csize * blockDim.x;
  double nH[CPT], nHe[CPT], T[CPT];
  PhotoRates p[CPT];
  rb_flag_t valid[CPT];
  rb_frac_t x[CPT];
  const double *E = P.Edges + (size_t)b * (Nl + 1);
RB_LS_UNROLL
  for (int i = 0; i < CPT; ++i) {
    const int il = tid + i * tstride;
    valid[i] = il < Nl;
    const size_t c = (size_t)b * Nl + (valid[i] ? il : 0);
    nH[i] = P.nH[c]; nHe[i] = P.nHe[c]; T[i] = P.T[c];
    double scale = 1.0;
    if (P.geom == RB_GEOM_SPHERE_STROMGREN) {
      const int q = valid[i] ? il : 0;
      const double r_c = (E[q] + E[q + 1]) * 0.5;
      scale = 1.0 / (4.0 * RB_PI * r_c * r_c);
    }
    const double *th = P.thin + (size_t)b * kThinStride;
    p[i].H1i = th[0] * scale; p[i].He1i = th[1] * scale; p[i].He2i = th[2] * scale;
    p[i].H1h = th[3] * scale; p[i].He1h = th[4] * scale; p[i].He2h = th[5] * scale;
  }
  int rc;
  const CaseA fc = {P.fcA, P.fcA, P.fcA, P.libm_pow};
  if (P.find_Teq) {
    rc = pcte_lockstep<CPT>(nH, nHe, p, valid, P.zHz[2 * b], P.zHz[2 * b + 1], fc,
                            P.tol_inner, x, T, any_fn, P.teq);
  } else {
    rb_kchem_t k[CPT];
RB_LS_UNROLL
    for (int i = 0; i < CPT; ++i)
      if (valid[i]) k[i] = get_kchem(T[i], fc.H2, fc.He2, fc.He3, fc.libm);
    rc = pce_lockstep<CPT>(nH, nHe, k, p, valid, P.tol_inner, x, any_fn);
    const size_t NC = (size_t)P.B * Nl;
RB_LS_UNROLL
    for (int i = 0; i < CPT; ++i) {
      if (!valid[i]) continue;
      const size_t c = (size_t)b * Nl + tid + i * tstride;
      P.kchem[c] = k[i].reH2; P.kchem[NC + c] = k[i].reHe2; P.kchem[2 * NC + c] = k[i].reHe3;
      P.kchem[3 * NC + c] = k[i].ciH1; P.kchem[4 * NC + c] = k[i].ciHe1;
      P.kchem[5 * NC + c] = k[i].ciHe2;
    }
  }
  if (rc) atomicMax(&P.err[b], rc);
RB_LS_UNROLL
  for (int i = 0; i < CPT; ++i) {
    if (!valid[i]) continue;
    const int il = tid + i * tstride;
    const size_t c = (size_t)b * Nl + il;
    P.x[0][c] = x[i].xH1; P.x[1][c] = x[i].xH2; P.x[2][c] = x[i].xHe1;
    P.x[3][c] = x[i].xHe2; P.x[4][c] = x[i].xHe3;
    if (P.find_Teq) P.T[c] = T[i];
    P.src[0][c] = p[i].H1i; P.src[1][c] = p[i].He1i; P.src[2][c] = p[i].He2i;
    P.src[3][c] = p[i].H1h; P.src[4][c] = p[i].He1h; P.src[5][c] = p[i].He2h;
    P.conv_old[c] = x[i].xH2 * nH[i] + (x[i].xHe2 + 2.0 * x[i].xHe3) * nHe[i];
  }
}

// ---------------------------------------------------------------------------
// set_column_vs_impact (sphere_base.f90:1115-1205): mu = 0 chord per shell.
// p = r_c, m = il; N = 2 sum_{k<il} ds_k c_k + 2 s_il c_il  for 5 densities.
// One thread per shell.
// ---------------------------------------------------------------------------
__global__ void k_column_vs_impact(const double *E, const double *nH, const double *nHe,
                                   const double *xH1, const double *xHe1,
                                   const double *xHe2, int Nl, double *NH, double *NHe,
                                   double *NH1, double *NHe1, double *NHe2, int *err) {
  const int il = blockIdx.x * blockDim.x + threadIdx.x;
  if (il >= Nl) return;
  const double mu = 0.0;
  const double r_ray = (E[il] + E[il + 1]) * 0.5;
  const double p_ray = r_ray * sqrt(1.0 - mu * mu);
  const double p2 = p_ray * p_ray;
  if (E[Nl] > p_ray) { atomicMax(err, RB_ERR_RAY_GEOMETRY); return; }
  int lo = 1, hi = Nl;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (E[mid] <= p_ray) hi = mid; else lo = mid + 1;
  }
  const int m = lo - 1;
  double S[5] = {0, 0, 0, 0, 0};
  double s_prev = sqrt(E[0] * E[0] - p2);
  for (int k = 1; k <= m; ++k) {
    const double s = sqrt(E[k] * E[k] - p2);
    const double ds = fabs(s - s_prev);
    const int q = k - 1;
    S[0] = fma(ds, nH[q], S[0]);
    S[1] = fma(ds, nHe[q], S[1]);
    S[2] = fma(ds, nH[q] * xH1[q], S[2]);
    S[3] = fma(ds, nHe[q] * xHe1[q], S[3]);
    S[4] = fma(ds, nHe[q] * xHe2[q], S[4]);
    s_prev = s;
  }
  const double t = 2.0 * s_prev;
  NH[il] = 2.0 * S[0] + t * nH[m];
  NHe[il] = 2.0 * S[1] + t * nHe[m];
  NH1[il] = 2.0 * S[2] + t * (nH[m] * xH1[m]);
  NHe1[il] = 2.0 * S[3] + t * (nHe[m] * xHe1[m]);
  NHe2[il] = 2.0 * S[4] + t * (nHe[m] * xHe2[m]);
}

// ---------------------------------------------------------------------------
// single-zone batch kernels
// ---------------------------------------------------------------------------
__global__ void k_kchem(const double *T, const double *a, const double *b_, const double *c,
                        int nn, double *o0, double *o1, double *o2, double *o3, double *o4,
                        double *o5) {
  pow_tables_to_smem();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const rb_kchem_t k = get_kchem(T[i], a[i], b_[i], c[i]);
  o0[i] = k.reH2; o1[i] = k.reHe2; o2[i] = k.reHe3;
  o3[i] = k.ciH1; o4[i] = k.ciHe1; o5[i] = k.ciHe2;
}

__global__ void k_kcool(const double *T, const double *a, const double *b_, const double *c,
                        int nn, double *o) {  // o: [10][nn]
  pow_tables_to_smem();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const rb_kcool_t k = get_kcool(T[i], a[i], b_[i], c[i]);
  o[0 * (size_t)nn + i] = k.recH2; o[1 * (size_t)nn + i] = k.recHe2;
  o[2 * (size_t)nn + i] = k.recHe3; o[3 * (size_t)nn + i] = k.cicH1;
  o[4 * (size_t)nn + i] = k.cicHe1; o[5 * (size_t)nn + i] = k.cicHe2;
  o[6 * (size_t)nn + i] = k.cecH1; o[7 * (size_t)nn + i] = k.cecHe2;
  o[8 * (size_t)nn + i] = k.bremss; o[9 * (size_t)nn + i] = k.compton;
}

// in: [11][nn] = nH,nHe,reH2,reHe2,reHe3,ciH1,ciHe1,ciHe2,H1i,He1i,He2i ; out [5][nn]
__global__ void k_zone_pce(const double *in, int nn, double tol, double *out, int *err,
                           int ce_only) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const size_t n = (size_t)nn;
  rb_kchem_t k;
  rb_frac_t x;
  int rc;
  if (ce_only) {  // in: [6][nn] rates only
    k.reH2 = in[0 * n + i]; k.reHe2 = in[1 * n + i]; k.reHe3 = in[2 * n + i];
    k.ciH1 = in[3 * n + i]; k.ciHe1 = in[4 * n + i]; k.ciHe2 = in[5 * n + i];
    rc = solve_ce(k, x);
  } else {
    k.reH2 = in[2 * n + i]; k.reHe2 = in[3 * n + i]; k.reHe3 = in[4 * n + i];
    k.ciH1 = in[5 * n + i]; k.ciHe1 = in[6 * n + i]; k.ciHe2 = in[7 * n + i];
    rc = solve_pce(in[0 * n + i], in[1 * n + i], k, in[8 * n + i], in[9 * n + i],
                   in[10 * n + i], tol, x);
  }
  if (rc) atomicMax(err, rc);
  out[0 * n + i] = x.xH1; out[1 * n + i] = x.xH2; out[2 * n + i] = x.xHe1;
  out[3 * n + i] = x.xHe2; out[4 * n + i] = x.xHe3;
}

// in: [11][nn] = nH,nHe,H1i,He1i,He2i,H1h,He1h,He2h,fcA_H2,fcA_He2,fcA_He3 ; out [6][nn]
__global__ void k_zone_pcte(const double *in, int nn, double z, double Hz, double tol,
                            double *out, int *err, TeqTab tb) {
  pow_tables_to_smem();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const size_t n = (size_t)nn;
  PhotoRates p = {in[2 * n + i], in[3 * n + i], in[4 * n + i],
                  in[5 * n + i], in[6 * n + i], in[7 * n + i]};
  const CaseA fc = {in[8 * n + i], in[9 * n + i], in[10 * n + i]};
  rb_frac_t x;
  double T = 0.0;
  const int rc = solve_pcte(in[0 * n + i], in[1 * n + i], p, z, Hz, fc, tol, x, T, tb);
  if (rc) atomicMax(err, rc);
  out[0 * n + i] = x.xH1; out[1 * n + i] = x.xH2; out[2 * n + i] = x.xHe1;
  out[3 * n + i] = x.xHe2; out[4 * n + i] = x.xHe3; out[5 * n + i] = T;
}

// k_teq_table: the memo rows of TeqTab (rb_ion.cuh) -- one thread per row replays the
// bisection's mid-point arithmetic down to its node and evaluates the 16 coefficients with
// the solvers' own get_kchem_kcool.
__global__ void __launch_bounds__(128) k_teq_table(double *rows, int depth, double fcA, int libm) {
  pow_tables_to_smem();
  const int last = 1 << depth;   // row of TMAX
  const int r = blockIdx.x * 128 + threadIdx.x;
  if (r > last) return;
  double T;
  if (r == 0) {
    T = RB_TFLOOR;
  } else if (r == last) {
    T = RB_TMAX;
  } else {
    TBisect s;
    s.Tl = RB_TFLOOR; s.Tr = RB_TMAX; s.T = teq_root_T(); s.err = 0.0;
    const int d = 31 - __clz(r);
    for (int i = d - 1; i >= 0; --i) tbisect_step(s, ((r >> i) & 1) == 0);
    T = s.T;
  }
  rb_kchem_t kc;
  rb_kcool_t kl;
  get_kchem_kcool(T, fcA, fcA, fcA, kc, kl, libm);
  teq_store(rows + (size_t)r * kTeqStride, T, kc, kl);
}

// lock-step single-zone solvers: ONE block, CPT zones per thread (nn <= 512*CPT)
template <int CPT>
__global__ void __launch_bounds__(512)
k_zone_lockstep(const double *in, int nn, double z, double Hz, double tol, double *out,
                int *err, int pcte, TeqTab tb) {
  pow_tables_to_smem();
  const size_t n = (size_t)nn;
  double nH[CPT], nHe[CPT], T[CPT];
  PhotoRates p[CPT];
  rb_flag_t valid[CPT];
  rb_frac_t x[CPT];
  rb_kchem_t k[CPT];
  CaseA fc = {1.0, 1.0, 1.0};
RB_LS_UNROLL
  for (int q = 0; q < CPT; ++q) {
    const int i = threadIdx.x + q * blockDim.x;
    valid[q] = i < nn;
    const int ii = valid[q] ? i : 0;
    nH[q] = in[0 * n + ii]; nHe[q] = in[1 * n + ii];
    T[q] = 0.0;
    if (pcte) {
      p[q] = {in[2 * n + ii], in[3 * n + ii], in[4 * n + ii],
              in[5 * n + ii], in[6 * n + ii], in[7 * n + ii]};
      // the reference's _v solver takes per-zone fcA arrays; the lock-step
      // kernel supports the (universal in practice) uniform case: zone 0's.
      fc = {in[8 * n], in[9 * n], in[10 * n]};
    } else {
      k[q].reH2 = in[2 * n + ii]; k[q].reHe2 = in[3 * n + ii]; k[q].reHe3 = in[4 * n + ii];
      k[q].ciH1 = in[5 * n + ii]; k[q].ciHe1 = in[6 * n + ii]; k[q].ciHe2 = in[7 * n + ii];
      p[q] = {in[8 * n + ii], in[9 * n + ii], in[10 * n + ii], 0.0, 0.0, 0.0};
    }
  }
  int rc;
  BlockAny any_fn;
  if (pcte) rc = pcte_lockstep<CPT>(nH, nHe, p, valid, z, Hz, fc, tol, x, T, any_fn, tb);
  else rc = pce_lockstep<CPT>(nH, nHe, k, p, valid, tol, x, any_fn);
  if (rc) atomicMax(err, rc);
RB_LS_UNROLL
  for (int q = 0; q < CPT; ++q) {
    const int i = threadIdx.x + q * blockDim.x;
    if (i >= nn) continue;
    out[0 * n + i] = x[q].xH1; out[1 * n + i] = x[q].xH2; out[2 * n + i] = x[q].xHe1;
    out[3 * n + i] = x[q].xHe2; out[4 * n + i] = x[q].xHe3;
    if (pcte) out[5 * n + i] = T[q];
  }
}

// ---------------------------------------------------------------------------
// FP64-pipe micro-benchmarks (roofline denominators)
// ---------------------------------------------------------------------------
// KIND 0 (DFMA): 32 independent accumulators, 64 FMAs per loop trip fully unrolled, so that
// the loop's own integer add / compare / branch are ~3 of 67 instructions (each non-FP64
// instruction costs a dispatch slot: with 8 FMAs per trip the figure read 8 % low).
template <int KIND>
__global__ void __launch_bounds__(256) k_microbench(double *out, int iters, double seed) {
  if (KIND == 0) {
    double a[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) a[q] = seed + threadIdx.x * 1e-9 + 0.01 * q;
    const double m = 0.999999, c = 1e-7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int q = 0; q < 32; ++q) a[q] = fma(a[q], m, c);
    }
    double sum = 0.0;
#pragma unroll
    for (int q = 0; q < 32; ++q) sum += a[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    return;
  }
  double a0 = seed + threadIdx.x * 1e-9, a1 = a0 + 0.1, a2 = a0 + 0.2, a3 = a0 + 0.3;
  double a4 = a0 + 0.4, a5 = a0 + 0.5, a6 = a0 + 0.6, a7 = a0 + 0.7;
  for (int i = 0; i < iters; ++i) {
    if (KIND == 1) {
      a0 = exp(-a0) + 0.5; a1 = exp(-a1) + 0.5; a2 = exp(-a2) + 0.5; a3 = exp(-a3) + 0.5;
      a4 = exp(-a4) + 0.5; a5 = exp(-a5) + 0.5; a6 = exp(-a6) + 0.5; a7 = exp(-a7) + 0.5;
    } else if (KIND == 2) {
      a0 = sqrt(a0) + 0.5; a1 = sqrt(a1) + 0.5; a2 = sqrt(a2) + 0.5; a3 = sqrt(a3) + 0.5;
      a4 = sqrt(a4) + 0.5; a5 = sqrt(a5) + 0.5; a6 = sqrt(a6) + 0.5; a7 = sqrt(a7) + 0.5;
    } else {
      a0 = pow(a0, 0.47) + 0.5; a1 = pow(a1, 0.47) + 0.5; a2 = pow(a2, 0.47) + 0.5;
      a3 = pow(a3, 0.47) + 0.5; a4 = pow(a4, 0.47) + 0.5; a5 = pow(a5, 0.47) + 0.5;
      a6 = pow(a6, 0.47) + 0.5; a7 = pow(a7, 0.47) + 0.5;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace rb
