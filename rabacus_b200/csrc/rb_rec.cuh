// rb_rec.cuh -- recombination-photon transfer (rec_meth != fixed), SURVEY 8f-1.
//
// Every sweep of the reference first evaluates, from the pre-sweep state, the
// photo-ionisation / heating rates that ground-state recombination photons of
// H II, He II and He III add in every cell, then solves the cells with case-A
// rates and (source + recombination) photo-rates:
//   sphere_bgnd.f90:686-735,861-881; sphere_stromgren.f90:713-757,805-843;
//   slab_bgnd.f90:653-672,752-808; slab_plane.f90:657-676,716-753,974-993.
// Kernels here (each cites the routine it restates):
//   k_rec_prep        emission coefficients, shell fluxes, opacities per cell
//   k_rec_outward     sphere_base.f90:79-346   one warp per solve layer
//   k_rec_radial      sphere_base.f90:392-750  one warp per solve layer
//   k_rec_iso_rays    sphere_base.f90:800-1083 one lane per (cell, ray) (_smem: shell data in shared memory)
//   k_rec_iso_finish  angle average + rates
//   k_rec_slab        slab_base.f90:163-560    one warp per (layer, line)
// Index orientation: the sphere routines work on a fixed orientation (outward
// and isotropic: index 0 = outermost shell; radial: index 0 = innermost) and the
// reference reverses the arrays on entry when the caller's differ
// (reverse_arrays, sphere_base.f90:1404-1452); here routine index i maps to plan
// cell (flip ? Nl-1-i : i), which keeps every running sum in the reference's
// order.  SphereBgnd plans are stored outermost-first, SphereStromgren plans
// innermost-first (rb_plan_set_problem).
#pragma once
#include "rb_kernels.cuh"

namespace rb {

// cross sections at the three thresholds (sphere_base.f90:224-236) + thresholds
struct RecConst {
  double sH1_H1, sH1_He1, sH1_He2, sHe1_He1, sHe1_He2, sHe2_He2;
  double E_th[3];
  double em_fac[3];
};

// per-cell scratch, all [B][Nl] in PLAN orientation
struct RecDev {
  double *em[3];    // emission coefficients (isotropic: already times em_fac)
  double *F0[3];    // flux leaving the shell (outward / radial)
  double *dtau[3];  // threshold optical depths through the shell (outward / radial / slab)
  double *r_c;      // shell centre
  double4 *recA;    // isotropic: [B][Nl+1] {E^2, dk_EH1, dk_EHe1, dk_EHe2}, index 0 = outermost
  double4 *recB;    // isotropic: [B][Nl+1] {em_EH1, em_EHe1, em_EHe2, 0}
  double4 *Iray;    // isotropic: [B][Nl][Nmu] specific intensities per ray
  double *tlo;      // slabs: [B][3][Nl+1] optical depth below each edge at the 3 line energies
  double *rec[6];   // OUTPUT rates H1i, He1i, He2i, H1h, He1h, He2h  [B][Nl]
};

__device__ __forceinline__ bool plan_outer_first(int geom) { return geom == RB_GEOM_SPHERE_BGND; }

// The closures run for the solve layers of one cell shard: plan cells
// start + j*stride, j < n_local (all cells on one GPU: 0, 1, Nl).  Slabs sharded over
// their first half add the never-solved middle layer of an odd Nl as entry n_local.
struct RecCells {
  int start, stride, n_local, n_total;
  __device__ __forceinline__ int cell(int j, int Nl) const {
    return j < n_local ? cell_of(j, start, stride) : Nl / 2;
  }
};

// ---------------------------------------------------------------------------
// k_rec_prep: one thread per (problem, cell).
// ---------------------------------------------------------------------------
__global__ void k_rec_prep(PlanDev P, RecDev R, RecConst K, int rec_meth) {
  pow_tables_to_smem();
  const int b = blockIdx.y;
  if (!P.active[b]) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int Nl = P.Nl;
  const bool sphere = (P.geom == RB_GEOM_SPHERE_BGND || P.geom == RB_GEOM_SPHERE_STROMGREN);
  const double *E = P.Edges + (size_t)b * (Nl + 1);
  if (sphere && rec_meth == 4 && i == Nl) {  // innermost edge record (routine index Nl)
    const double Ee = plan_outer_first(P.geom) ? E[Nl] : E[0];
    R.recA[(size_t)b * (Nl + 1) + Nl] = make_double4(Ee * Ee, 0.0, 0.0, 0.0);
    R.recB[(size_t)b * (Nl + 1) + Nl] = make_double4(0.0, 0.0, 0.0, 0.0);
  }
  if (i >= Nl) return;
  const size_t c = (size_t)b * Nl + i;
  const double nH = P.nH[c], nHe = P.nHe[c], T = P.T[c];
  const double xH1 = P.x[0][c], xH2 = P.x[1][c], xHe1 = P.x[2][c], xHe2 = P.x[3][c],
               xHe3 = P.x[4][c];
  // case A - case B = recombinations to the ground state (sphere_base.f90:238-262)
  const rb_kchem_t kA = get_kchem(T, 1.0, 1.0, 1.0, P.libm_pow);
  const rb_kchem_t kB = get_kchem(T, 0.0, 0.0, 0.0, P.libm_pow);
  const double reH2_1 = kA.reH2 - kB.reH2, reHe2_1 = kA.reHe2 - kB.reHe2,
               reHe3_1 = kA.reHe3 - kB.reHe3;
  const double ne = xH2 * nH + (xHe2 + 2.0 * xHe3) * nHe;  // :267
  double em[3];
  em[0] = (reH2_1 * nH * xH2 * ne) / (4.0 * RB_PI);   // :269-271
  em[1] = (reHe2_1 * nHe * xHe2 * ne) / (4.0 * RB_PI);
  em[2] = (reHe3_1 * nHe * xHe3 * ne) / (4.0 * RB_PI);
  if (sphere && rec_meth == 4) {
    // :878-892, :935-937 ; records in the routine's orientation (0 = outermost)
    const int k = plan_outer_first(P.geom) ? i : Nl - 1 - i;
    const double Ek = plan_outer_first(P.geom) ? E[i] : E[i + 1];  // outer edge of the shell
    const double kH1_H1 = nH * xH1 * K.sH1_H1, kHe1_H1 = nH * xH1 * K.sH1_He1,
                 kHe2_H1 = nH * xH1 * K.sH1_He2, kHe1_He1 = nHe * xHe1 * K.sHe1_He1,
                 kHe2_He1 = nHe * xHe1 * K.sHe1_He2, kHe2_He2 = nHe * xHe2 * K.sHe2_He2;
    R.recA[(size_t)b * (Nl + 1) + k] =
        make_double4(Ek * Ek, kH1_H1, kHe1_H1 + kHe1_He1, kHe2_H1 + kHe2_He1 + kHe2_He2);
    R.recB[(size_t)b * (Nl + 1) + k] =
        make_double4(em[0] * K.em_fac[0], em[1] * K.em_fac[1], em[2] * K.em_fac[2], 0.0);
    return;
  }
  const double Ei = E[i], Ej = E[i + 1];
  double dl;
  if (sphere) {
    dl = fabs(Ei - Ej);                       // sphere_base.f90:1534
    const double r_c = (Ei + Ej) * 0.5;       // :1535
    const double area = 4.0 * RB_PI * r_c * r_c;                                    // :220
    const double vol = fabs(4.0 / 3.0 * RB_PI * (Ei * Ei * Ei - Ej * Ej * Ej));     // :221
    R.r_c[c] = r_c;
    for (int X = 0; X < 3; ++X) R.F0[X][c] = em[X] * 4.0 * RB_PI * vol / area;      // :281-283
  } else {
    dl = Ej - Ei;                             // slab_base.f90:250
    R.r_c[c] = dl;
  }
  const double dNH = dl * nH, dNHe = dl * nHe;
  R.dtau[0][c] = dNH * xH1 * P.sigma_th[0];   // :216-218
  R.dtau[1][c] = dNHe * xHe1 * P.sigma_th[1];
  R.dtau[2][c] = dNHe * xHe2 * P.sigma_th[2];
  for (int X = 0; X < 3; ++X) R.em[X][c] = em[X];
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one recombining layer (sphere_base.f90:275-307): the three fluxes
struct RecRatios {
  double a0, b0, b1, c0, c1, c2;  // sigma_X(E_Y,th) / sigma_X,th
};
__device__ __forceinline__ RecRatios rec_ratios(const PlanDev &P, const RecConst &K) {
  RecRatios q;
  q.a0 = K.sH1_H1 / P.sigma_th[0];
  q.b0 = K.sH1_He1 / P.sigma_th[0]; q.b1 = K.sHe1_He1 / P.sigma_th[1];
  q.c0 = K.sH1_He2 / P.sigma_th[0]; q.c1 = K.sHe1_He2 / P.sigma_th[1];
  q.c2 = K.sHe2_He2 / P.sigma_th[2];
  return q;
}
__device__ __forceinline__ void rec_add_layer(const RecRatios &q, const double (&tau)[3], double g,
                                              double fac, double F0a, double F0b, double F0c,
                                              double (&F)[3]) {
  double t = tau[0] * q.a0;
  F[0] = F[0] + fac * F0a * g * exp(-t);
  t = tau[0] * q.b0 + tau[1] * q.b1;
  F[1] = F[1] + fac * F0b * g * exp(-t);
  t = tau[0] * q.c0 + tau[1] * q.c1 + tau[2] * q.c2;
  F[2] = F[2] + fac * F0c * g * exp(-t);
}
// sphere_base.f90:311-333
__device__ __forceinline__ void rec_store_flux_rates(const RecDev &R, const RecConst &K, size_t c,
                                                     const double (&F)[3]) {
  R.rec[0][c] = F[0] * K.sH1_H1 + F[1] * K.sH1_He1 + F[2] * K.sH1_He2;
  R.rec[1][c] = F[1] * K.sHe1_He1 + F[2] * K.sHe1_He2;
  R.rec[2][c] = F[2] * K.sHe2_He2;
  double h = F[1] * (K.E_th[1] - K.E_th[0]) * K.sH1_He1 + F[2] * (K.E_th[2] - K.E_th[0]) * K.sH1_He2;
  R.rec[3][c] = h * RB_EV_2_ERG;
  h = F[1] * K.sHe1_He1 + F[2] * K.sHe1_He2;  // sic (no energy factor in the reference)
  R.rec[4][c] = h * RB_EV_2_ERG;
  R.rec[5][c] = 0.0;
}

// ---------------------------------------------------------------------------
// Outward / radial closures: for every solve layer the reference walks the other layers
// in sequence, adding half threshold optical depths to three running tau and
// F_X += fac F0_X(layer) (r/r_s)^2 exp(-tau ratios)  (sphere_base.f90:238-307, 548-684).
// One WARP per solve layer: the walk of Nq positions is cut into 32 contiguous pieces;
// a first pass sums each piece's tau increments, an exclusive warp scan gives every
// lane its starting tau, a second pass does the exponentials, a shuffle tree adds the
// 32 partial fluxes.  Same terms as the reference; the running sums are re-associated
// at the 32 piece boundaries (1e-16-level differences; O(Nl^2) work, ~50 us at C4
// instead of 2-6 ms with one thread per layer).
// `irl_of(q)` = plan cell of position q, `add_inc(q, tau)` = the increment rule.
// ---------------------------------------------------------------------------
template <class IRL, class INC>
__device__ __forceinline__ void rec_leg(int Nq, IRL irl_of, INC add_inc, double fac, double rs,
                                        const RecDev &R, const RecRatios &q6, double (&F)[3]) {
  const int lane = threadIdx.x & 31;
  const int C = (Nq + 31) / 32;
  const int q0 = min(Nq, lane * C), q1 = min(Nq, q0 + C);
  double t[3] = {0.0, 0.0, 0.0};
  for (int q = q0; q < q1; ++q) add_inc(q, t);
  double tau[3];
#pragma unroll
  for (int X = 0; X < 3; ++X) {
    const double incl = warp_incl_scan(t[X]);
    const double prev = __shfl_up_sync(0xffffffffu, incl, 1);
    tau[X] = (lane == 0) ? 0.0 : prev;
  }
  for (int q = q0; q < q1; ++q) {
    add_inc(q, tau);
    const size_t ci = irl_of(q);
    const double rr = R.r_c[ci] / rs;
    rec_add_layer(q6, tau, rr * rr, fac, R.F0[0][ci], R.F0[1][ci], R.F0[2][ci], F);
  }
}

// k_rec_outward (sphere_base.f90:79-346): routine index 0 = outermost shell; each solve
// layer collects the flux of the layers at smaller radii.  One warp per solve layer.
__global__ void __launch_bounds__(128) k_rec_outward(PlanDev P, RecDev R, RecConst K,
                                                     RecCells L) {
  const int b = blockIdx.y;
  if (!P.active[b]) return;
  const int j = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int Nl = P.Nl;
  if (j >= L.n_total) return;
  const bool flip = !plan_outer_first(P.geom);
  const int isl = flip ? Nl - 1 - L.cell(j, Nl) : L.cell(j, Nl);
  const size_t base = (size_t)b * Nl;
  auto rix = [=](int i) { return base + (flip ? Nl - 1 - i : i); };
  const RecRatios q6 = rec_ratios(P, K);
  double F[3] = {0.0, 0.0, 0.0};
  const double rs = R.r_c[rix(isl)];
  rec_leg(
      Nl - isl, [=](int q) { return rix(isl + q); },
      [=](int q, double (&tau)[3]) {  // :249-266
        const int irl = isl + q;
#pragma unroll
        for (int X = 0; X < 3; ++X) {
          if (q <= 1) {
            tau[X] = tau[X] + R.dtau[X][rix(irl)] * 0.5;
          } else {
            tau[X] = tau[X] + R.dtau[X][rix(irl - 1)] * 0.5;
            tau[X] = tau[X] + R.dtau[X][rix(irl)] * 0.5;
          }
        }
      },
      1.0, rs, R, q6, F);
#pragma unroll
  for (int X = 0; X < 3; ++X) F[X] = warp_sum(F[X]);
  if ((threadIdx.x & 31) == 0) rec_store_flux_rates(R, K, rix(isl), F);
}

// k_rec_radial (sphere_base.f90:392-750): routine index 1 = innermost shell (1-based as
// in the Fortran); "lower" leg runs isl, isl-1, .., 1, then through the centre -1, ..,
// -Nl; "upper" leg isl .. Nl; each contribution weighs one half.  One warp per layer.
__global__ void __launch_bounds__(128) k_rec_radial(PlanDev P, RecDev R, RecConst K,
                                                    RecCells L) {
  const int b = blockIdx.y;
  if (!P.active[b]) return;
  const int j = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int Nl = P.Nl;
  if (j >= L.n_total) return;
  const bool flip = plan_outer_first(P.geom);
  const int isl1 = flip ? Nl - L.cell(j, Nl) : L.cell(j, Nl) + 1;
  const size_t base = (size_t)b * Nl;
  auto rix1 = [=](int i1) { return base + (flip ? Nl - i1 : i1 - 1); };
  const RecRatios q6 = rec_ratios(P, K);
  double F[3] = {0.0, 0.0, 0.0};
  const double rs = R.r_c[rix1(isl1)];
  // lower leg (:558-617): positions q = 0 .. isl1-1 -> layer isl1 - q ; then the far side
  // q = isl1 .. isl1+Nl-1 -> layer q - isl1 + 1
  auto lower_layer = [=](int q) { return q < isl1 ? isl1 - q : q - isl1 + 1; };
  rec_leg(
      isl1 + Nl, [=](int q) { return rix1(lower_layer(q)); },
      [=](int q, double (&tau)[3]) {
        const int irl1 = lower_layer(q);
#pragma unroll
        for (int X = 0; X < 3; ++X) {
          if (irl1 == isl1 || irl1 == isl1 - 1) {
            tau[X] = tau[X] + R.dtau[X][rix1(irl1)] * 0.5;
          } else if (irl1 < isl1 - 1) {
            tau[X] = tau[X] + R.dtau[X][rix1(irl1 + 1)] * 0.5;
            tau[X] = tau[X] + R.dtau[X][rix1(irl1)] * 0.5;
          }
        }
      },
      0.5, rs, R, q6, F);
  // upper leg (:625-684): layers isl1 .. Nl
  rec_leg(
      Nl - isl1 + 1, [=](int q) { return rix1(isl1 + q); },
      [=](int q, double (&tau)[3]) {
        const int irl1 = isl1 + q;
#pragma unroll
        for (int X = 0; X < 3; ++X) {
          if (q <= 1) {
            tau[X] = tau[X] + R.dtau[X][rix1(irl1)] * 0.5;
          } else {
            tau[X] = tau[X] + R.dtau[X][rix1(irl1)] * 0.5;
            tau[X] = tau[X] + R.dtau[X][rix1(irl1 - 1)] * 0.5;
          }
        }
      },
      0.5, rs, R, q6, F);
#pragma unroll
  for (int X = 0; X < 3; ++X) F[X] = warp_sum(F[X]);
  if ((threadIdx.x & 31) == 0) rec_store_flux_rates(R, K, rix1(isl1), F);
}

// ---------------------------------------------------------------------------
// k_rec_iso_rays (sphere_base.f90:966-1016): one thread per (cell, ray).  Lanes
// are consecutive cells of one direction (near-equal ray lengths, broadcast
// record reads).  Along the ray, from the cell outwards:
//     tau_E += ds dk_E(shell) ;  I_E += em_E(shell) ds exp(-tau_E)
// for the three threshold energies.  Segment lengths as in the column kernel
// (geometry.f90:146-165): shell k < m: |s_{k+1} - s_k|, tangent shell m: 2 s_m,
// first segment halved.  mu <= 0: shells il, il+1, .., m, m-1, .., 0;
// mu > 0: il, il-1, .., 0.
// ---------------------------------------------------------------------------
struct IsoAcc {
  double tau[3], I[3];
};
__device__ __forceinline__ void iso_step(IsoAcc &a, double ds, const double4 &ka, const double4 &kb,
                                         const double *s_exp, unsigned lane_off) {
  a.tau[0] = a.tau[0] + ds * ka.y;
  a.tau[1] = a.tau[1] + ds * ka.z;
  a.tau[2] = a.tau[2] + ds * ka.w;
  a.I[0] = a.I[0] + kb.x * ds * exp_fast<6>(-a.tau[0], s_exp, lane_off);
  a.I[1] = a.I[1] + kb.y * ds * exp_fast<6>(-a.tau[1], s_exp, lane_off);
  a.I[2] = a.I[2] + kb.z * ds * exp_fast<6>(-a.tau[2], s_exp, lane_off);
}

__global__ void __launch_bounds__(128) k_rec_iso_rays(PlanDev P, RecDev R, int Nmu, RecCells L) {
  const int b = blockIdx.z;
  if (!P.active[b]) return;
  __shared__ double s_exp[64 * kExpTabCopies];
  fill_exp_table_L<6>(s_exp);
  __syncthreads();
  const unsigned lane_off = (threadIdx.x & (kExpTabCopies - 1)) << 3;
  const int jl = blockIdx.x * blockDim.x + threadIdx.x;
  const int imu = blockIdx.y;
  const int Nl = P.Nl;
  if (jl >= L.n_total) return;
  // routine index, 0 = outermost
  const int il = plan_outer_first(P.geom) ? L.cell(jl, Nl) : Nl - 1 - L.cell(jl, Nl);
  const double4 *ka = R.recA + (size_t)b * (Nl + 1);
  const double4 *kb = R.recB + (size_t)b * (Nl + 1);
  const double mu = P.mu[imu];
  // the edges themselves (not their squares) define r and p: read them from the plan
  const double *E = P.Edges + (size_t)b * (Nl + 1);
  const bool of = plan_outer_first(P.geom);
  const double Ea = of ? E[il] : E[Nl - il], Eb = of ? E[il + 1] : E[Nl - il - 1];
  const double r_ray = (Ea + Eb) * 0.5;
  const double p_ray = r_ray * sqrt(1.0 - mu * mu);
  const double p2 = p_ray * p_ray;
  // m_for_ray (geometry.f90:61-71) on the outermost-first edges
  const double E_in = of ? E[Nl] : E[0];
  if (E_in > p_ray) {
    atomicMax(&P.err[b], RB_ERR_RAY_GEOMETRY);
    return;
  }
  int lo = 1, hi = Nl;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const double Em = of ? E[mid] : E[Nl - mid];
    if (Em <= p_ray) hi = mid; else lo = mid + 1;
  }
  const int m = lo - 1;
  IsoAcc a;
  a.tau[0] = a.tau[1] = a.tau[2] = 0.0;
  a.I[0] = a.I[1] = a.I[2] = 0.0;
  int k_out;          // first shell of the outward leg
  double s_hi;        // s at the inner edge of that shell (s_{k_out+1}), unused when k_out == m
  bool first = true;
  if (mu <= 0.0) {
    // inward leg: shells il .. m
    double s_k = sqrt_pos(ka[il].x - p2);
    for (int k = il; k <= m; ++k) {
      double ds, s_n = 0.0;
      if (k < m) {
        s_n = sqrt_pos(ka[k + 1].x - p2);
        ds = fabs(s_n - s_k);
      } else {
        ds = 2.0 * s_k;
      }
      if (first) { ds = ds * 0.5; first = false; }
      iso_step(a, ds, ka[k], kb[k], s_exp, lane_off);
      if (k < m) s_k = s_n;
    }
    k_out = m - 1;
    s_hi = s_k;  // = s_m
  } else {
    k_out = il;
    s_hi = (il < m) ? sqrt_pos(ka[il + 1].x - p2) : 0.0;
  }
  // outward leg: shells k_out .. 0
  for (int k = k_out; k >= 0; --k) {
    const double s_k = sqrt_pos(ka[k].x - p2);
    double ds = (k == m) ? 2.0 * s_k : fabs(s_hi - s_k);
    if (first) { ds = ds * 0.5; first = false; }
    iso_step(a, ds, ka[k], kb[k], s_exp, lane_off);
    s_hi = s_k;
  }
  R.Iray[((size_t)b * Nl + il) * Nmu + imu] = make_double4(a.I[0], a.I[1], a.I[2], 0.0);
}

// ---------------------------------------------------------------------------
// k_rec_iso_rays_smem: the same rays with the shell data of the problem staged in
// shared memory (56 B per shell: E^2 | dk_EH1, dk_EHe1, dk_EHe2, em_EH1 | em_EHe1,
// em_EHe2 -- the 4097 edges of C4 fill 224 KB, one CTA per SM) and a 512-byte exp table
// (exp_small).  Work items = (block of 32 consecutive cells, ray) handed out
// longest path first by an atomic counter; unclamped square root inside the walk
// (E_k > p for k <= m, see sphere_ray_pair).  grid = (CTAs per problem, B);
// items[it] = jblk | imu << 16, jblk = block of 32 consecutive entries of the cell list.
// ---------------------------------------------------------------------------
struct IsoSmem {
  const double *E2;    // [nrec]
  const double4 *K;    // [nrec] {dk0, dk1, dk2, em0}
  const double2 *M;    // [nrec] {em1, em2}
};

// exp(x), -708 <= x <= 0, with a 32-entry 2^(j/32) table of two copies (512 B: all that
// is left of shared memory beside the shell data of a 4096-shell problem; even / odd lanes
// read their own copy) and a degree-6 polynomial (|r| <= ln2/64, truncation
// r^7/5040 < 4e-18); otherwise exp_fast.  10 FP64 instructions.
__constant__ double kExpS[8] = {46.16624130844683,            // 32/ln2
                                -2.1660849392498290e-02,      // -ln2/32
                                1.3888888888888889e-03, 8.3333333333333332e-03,
                                4.1666666666666664e-02, 1.6666666666666666e-01, 0.5,
                                6755399441055744.0};
__device__ __forceinline__ double exp_small(double x, const double *s_tab, unsigned lane_off) {
  x = __hiloint2double((int)min((unsigned)__double2hiint(x), 0xC0862000u), __double2loint(x));
  const double t = fma(x, kExpS[0], kExpS[7]);
  const int n = __double2loint(t);
  const double nd = t - kExpS[7];
  const double r = fma(nd, kExpS[1], x);
  double q = fma(r, kExpS[2], kExpS[3]);
  q = fma(r, q, kExpS[4]);
  q = fma(r, q, kExpS[5]);
  q = fma(r, q, kExpS[6]);
  const double r2 = r * r;
  const double pl = fma(r2, q, r);
  const unsigned off = ((unsigned)(n << 4) & (31u << 4)) | lane_off;
  // (entries carry -j 2^15 in their high word: one imad on the entry, before the last fma --
  // see exp_fast in rb_kernels.cuh)
  const double Ts = *reinterpret_cast<const double *>(reinterpret_cast<const char *>(s_tab) + off);
  int hi;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(hi) : "r"(n), "n"(1 << 15), "r"(__double2hiint(Ts)));
  const double T = __hiloint2double(hi, __double2loint(Ts));
  return fma(T, pl, T);
}

__device__ __forceinline__ void iso_step_s(IsoAcc &a, double ds, const IsoSmem &S, int k,
                                           const double *s_tab, unsigned lane_off) {
  const double4 kk = S.K[k];
  const double2 mm = S.M[k];
  // fused multiply-adds (the reference rounds the products separately: 1e-16 per step)
  a.tau[0] = fma(ds, kk.x, a.tau[0]);
  a.tau[1] = fma(ds, kk.y, a.tau[1]);
  a.tau[2] = fma(ds, kk.z, a.tau[2]);
  a.I[0] = fma(kk.w * ds, exp_small(-a.tau[0], s_tab, lane_off), a.I[0]);
  a.I[1] = fma(mm.x * ds, exp_small(-a.tau[1], s_tab, lane_off), a.I[1]);
  a.I[2] = fma(mm.y * ds, exp_small(-a.tau[2], s_tab, lane_off), a.I[2]);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
k_rec_iso_rays_smem(PlanDev P, RecDev R, int Nmu, const unsigned *__restrict__ items, int n_items,
                    int *work, RecCells L) {
  const int b = blockIdx.y;
  if (!P.active[b]) return;
  extern __shared__ double4 s_iso[];
  const int Nl = P.Nl;
  const int nrec = Nl + 1;
  double4 *sK = s_iso;
  double2 *sM = reinterpret_cast<double2 *>(sK + nrec);
  double *sE2 = reinterpret_cast<double *>(sM + nrec);
  {
    const double4 *ka = R.recA + (size_t)b * (Nl + 1);
    const double4 *kb = R.recB + (size_t)b * (Nl + 1);
    for (int k = threadIdx.x; k < nrec; k += THREADS) {
      const double4 A = ka[k], Bv = kb[k];
      sE2[k] = A.x;
      sK[k] = make_double4(A.y, A.z, A.w, Bv.x);
      sM[k] = make_double2(Bv.y, Bv.z);
    }
  }
  __syncthreads();
  const IsoSmem S = {sE2, sK, sM};
  __shared__ double s_tab[64];   // [32 entries][2 copies]
  if (threadIdx.x < 64) {
    const int j = threadIdx.x >> 1;
    const double v = exp2((double)j * (1.0 / 32.0));
    s_tab[threadIdx.x] = __hiloint2double(__double2hiint(v) - (j << 15), __double2loint(v));
  }
  __syncthreads();
  const unsigned lane_off = (threadIdx.x & 1) << 3;
  const int lane = threadIdx.x & 31;
  const double *E = P.Edges + (size_t)b * (Nl + 1);
  const bool of = plan_outer_first(P.geom);
  const double E_in = of ? E[Nl] : E[0];
  int it = 0, nxt = 0;
  if (lane == 0) it = atomicAdd(&work[b], 1);
  it = __shfl_sync(0xffffffffu, it, 0);
  while (it < n_items) {
    if (lane == 0) nxt = atomicAdd(&work[b], 1);
    const unsigned item = items[it];
    const int imu = (int)(item >> 16);
    const int j_raw = (int)(item & 0xffffu) * 32 + lane;
    const bool valid = j_raw < L.n_total;
    const int pc = L.cell(valid ? j_raw : L.n_total - 1, Nl);
    const int il = of ? pc : Nl - 1 - pc;   // routine index, 0 = outermost
    const double mu = P.mu[imu];
    const double Ea = of ? E[il] : E[Nl - il], Eb = of ? E[il + 1] : E[Nl - il - 1];
    const double r_ray = (Ea + Eb) * 0.5;
    const double p_ray = r_ray * sqrt(1.0 - mu * mu);
    const double p2 = p_ray * p_ray;
    if (E_in > p_ray) {  // m_for_ray undefined (geometry.f90:61-71)
      atomicMax(&P.err[b], RB_ERR_RAY_GEOMETRY);
    } else {
      int lo = 1, hi = Nl;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double Em = of ? E[mid] : E[Nl - mid];
        if (Em <= p_ray) hi = mid; else lo = mid + 1;
      }
      const int m = lo - 1;
      IsoAcc a;
      a.tau[0] = a.tau[1] = a.tau[2] = 0.0;
      a.I[0] = a.I[1] = a.I[2] = 0.0;
      int k_out;
      double s_hi;
      // the halved first segment and the tangent shell are peeled off so that the two long
      // loops are branch-free (the compiler interleaves two shells: 6 exp chains in flight)
      double s_k = sqrt_pos<false>(S.E2[il] - p2);
      if (mu <= 0.0) {
        if (il < m) {
          double s_n = sqrt_pos<false>(S.E2[il + 1] - p2);
          iso_step_s(a, fabs(s_n - s_k) * 0.5, S, il, s_tab, lane_off);
          s_k = s_n;
#pragma unroll 2
          for (int k = il + 1; k < m; ++k) {   // inward leg
            s_n = sqrt_pos<false>(S.E2[k + 1] - p2);
            iso_step_s(a, fabs(s_n - s_k), S, k, s_tab, lane_off);
            s_k = s_n;
          }
          iso_step_s(a, 2.0 * s_k, S, m, s_tab, lane_off);   // tangent shell
        } else {
          iso_step_s(a, (2.0 * s_k) * 0.5, S, m, s_tab, lane_off);  // own shell is the tangent shell
        }
        k_out = m - 1;
        s_hi = s_k;  // = s_m
      } else {
        double ds;
        if (il < m) ds = fabs(sqrt_pos<false>(S.E2[il + 1] - p2) - s_k) * 0.5;
        else ds = (2.0 * s_k) * 0.5;
        iso_step_s(a, ds, S, il, s_tab, lane_off);
        k_out = il - 1;
        s_hi = s_k;
      }
#pragma unroll 2
      for (int k = k_out; k >= 0; --k) {   // outward leg
        const double s_o = sqrt_pos<false>(S.E2[k] - p2);
        iso_step_s(a, fabs(s_hi - s_o), S, k, s_tab, lane_off);
        s_hi = s_o;
      }
      if (valid)
        R.Iray[((size_t)b * Nl + il) * Nmu + imu] = make_double4(a.I[0], a.I[1], a.I[2], 0.0);
    }
    it = __shfl_sync(0xffffffffu, nxt, 0);
  }
}

// angle average J = (1/2) sum_mu I w (sphere_base.f90:1020-1032) and the rates (:1037-1061)
__global__ void k_rec_iso_finish(PlanDev P, RecDev R, RecConst K, int Nmu, RecCells L) {
  const int b = blockIdx.y;
  if (!P.active[b]) return;
  const int jl = blockIdx.x * blockDim.x + threadIdx.x;
  const int Nl = P.Nl;
  if (jl >= L.n_total) return;
  const int il = plan_outer_first(P.geom) ? L.cell(jl, Nl) : Nl - 1 - L.cell(jl, Nl);
  double J[3] = {0.0, 0.0, 0.0};
  const double4 *I = R.Iray + ((size_t)b * Nl + il) * Nmu;
  for (int imu = 0; imu < Nmu; ++imu) {
    const double4 v = I[imu];
    const double w = P.wmu[imu];
    J[0] = J[0] + v.x * w; J[1] = J[1] + v.y * w; J[2] = J[2] + v.z * w;
  }
  J[0] = J[0] * 0.5; J[1] = J[1] * 0.5; J[2] = J[2] * 0.5;
  const size_t c = (size_t)b * Nl + (plan_outer_first(P.geom) ? il : Nl - 1 - il);
  R.rec[0][c] = (J[0] * K.sH1_H1 + J[1] * K.sH1_He1 + J[2] * K.sH1_He2) * 4.0 * RB_PI;
  R.rec[1][c] = (J[1] * K.sHe1_He1 + J[2] * K.sHe1_He2) * 4.0 * RB_PI;
  R.rec[2][c] = (J[2] * K.sHe2_He2) * 4.0 * RB_PI;
  double h = (J[1] * (K.E_th[1] - K.E_th[0]) * K.sH1_He1 +
              J[2] * (K.E_th[2] - K.E_th[0]) * K.sH1_He2) * 4.0 * RB_PI;
  R.rec[3][c] = h * RB_EV_2_ERG;
  h = (J[1] * K.sHe1_He1 + J[2] * K.sHe1_He2) * 4.0 * RB_PI;
  R.rec[4][c] = h * RB_EV_2_ERG;
  R.rec[5][c] = 0.0;
}

// slab_base.f90:344-366: tau_X_lo(k) = tau_X_lo(k-1) + sum_Y dtau_Y,th(k-1) ra_Y(E_X), a
// sequential running sum per line energy (one thread each, Nl dependent adds)
__global__ void k_rec_slab_prefix(PlanDev P, RecDev R, RecConst K) {
  const int b = blockIdx.x, X = threadIdx.x;
  if (!P.active[b] || X > 2) return;
  const int Nl = P.Nl;
  const size_t base = (size_t)b * Nl;
  double *t = R.tlo + ((size_t)b * 3 + X) * (Nl + 1);
  const double r0 = (X == 0) ? 1.0 : (X == 1 ? K.sH1_He1 / K.sH1_H1 : K.sH1_He2 / K.sH1_H1);
  const double r1 = (X == 1) ? 1.0 : K.sHe1_He2 / K.sHe1_He1;
  double acc = 0.0;
  t[0] = acc;
  for (int k = 1; k <= Nl; ++k) {
    if (X == 0) acc = acc + R.dtau[0][base + k - 1] * r0;
    else if (X == 1) acc = acc + R.dtau[0][base + k - 1] * r0 + R.dtau[1][base + k - 1] * r1;
    else acc = acc + R.dtau[0][base + k - 1] * r0 + R.dtau[1][base + k - 1] * r1 +
               R.dtau[2][base + k - 1] * 1.0;
    t[k] = acc;
  }
}

// ---------------------------------------------------------------------------
// k_rec_slab (slab_base.f90:163-560): one warp per (layer k, recombination line X).
// Optical depths at the line energy from the layer centre to every edge, E1 kernels
// on the edges, trapezoid rule over the edges below and above (:402-538).  Each of the
// two walks is cut into 32 pieces: a first pass + exclusive warp scan gives every lane
// the optical depth at the start of its piece, the second pass evaluates the E1 kernels
// and its part of the trapezoid sum, a shuffle tree adds the parts (the reference
// fills J_int edge by edge and sums from edge 0 upwards: same terms, re-associated).
// The three lines' J are combined into the six rates by k_rec_slab_finish.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_rec_slab(PlanDev P, RecDev R, double *Jout, RecCells L) {
  const int b = blockIdx.z;
  if (!P.active[b]) return;
  const int jl = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int X = blockIdx.y;
  const int Nl = P.Nl;
  if (jl >= L.n_total) return;
  const int k = L.cell(jl, Nl);
  const int lane = threadIdx.x & 31;
  const size_t base = (size_t)b * Nl;
  const double *E = P.Edges + (size_t)b * (Nl + 1);
  const double *tl = R.tlo + ((size_t)b * 3 + X) * (Nl + 1);
  const double *em = R.em[X] + base;
  auto dtau_at = [&](int i) { return tl[i + 1] - tl[i]; };   // as the reference forms it
  auto em_edge = [&](int e) {  // :378-392
    if (e == 0) return em[0] * 0.5;
    if (e == Nl) return em[Nl - 1] * 0.5;
    return (em[e - 1] + em[e]) * 0.5;
  };
  const double tau0 = dtau_at(k) * 0.5;
  const double self = R.r_c[base + k] * em[k] * e1xa(tau0) * 0.5;   // :404-405 (r_c holds dl)
  double ss_lo = 0.0, ss_hi = 0.0;
  {  // lower integral (:411-421): edges i = k-1 .. 0, position q <-> i = k-1-q
    const int Nq = k;
    const int C = (Nq + 31) / 32;
    const int q0 = min(Nq, lane * C), q1 = min(Nq, q0 + C);
    double t = 0.0;
    for (int q = q0; q < q1; ++q) t = t + dtau_at(k - 1 - q);
    const double incl = warp_incl_scan(t);
    const double prev = __shfl_up_sync(0xffffffffu, incl, 1);
    double tau = tau0 + ((lane == 0) ? 0.0 : prev);
    if (q0 < q1) {
      double y_hi = em_edge(k - q0) * e1xa(tau);
      for (int q = q0; q < q1; ++q) {
        const int i = k - 1 - q;
        tau = tau + dtau_at(i);
        const double y_lo = em_edge(i) * e1xa(tau);
        ss_lo = ss_lo + (y_hi + y_lo) * (E[i + 1] - E[i]);
        y_hi = y_lo;
      }
    }
  }
  {  // upper integral (:425-435): layers i = k+1 .. Nl-1, position q <-> i = k+1+q
    const int Nq = Nl - 1 - k;
    const int C = (Nq + 31) / 32;
    const int q0 = min(Nq, lane * C), q1 = min(Nq, q0 + C);
    double t = 0.0;
    for (int q = q0; q < q1; ++q) t = t + dtau_at(k + 1 + q);
    const double incl = warp_incl_scan(t);
    const double prev = __shfl_up_sync(0xffffffffu, incl, 1);
    double tau = tau0 + ((lane == 0) ? 0.0 : prev);
    if (q0 < q1) {
      double y_lo = em_edge(k + 1 + q0) * e1xa(tau);
      for (int q = q0; q < q1; ++q) {
        const int i = k + 1 + q;
        tau = tau + dtau_at(i);
        const double y_hi = em_edge(i + 1) * e1xa(tau);
        ss_hi = ss_hi + (y_hi + y_lo) * (E[i + 1] - E[i]);
        y_lo = y_hi;
      }
    }
  }
  ss_lo = warp_sum(ss_lo);
  ss_hi = warp_sum(ss_hi);
  if (lane == 0) {
    const double Jlo = self + 0.5 * ss_lo, Jhi = self + 0.5 * ss_hi;
    Jout[((size_t)b * 3 + X) * Nl + k] = (Jlo + Jhi) * 4.0 * RB_PI * 0.5;  // :544-552
  }
}

// slab_base.f90:556-572: the six rates from the three lines' J
__global__ void k_rec_slab_finish(PlanDev P, RecDev R, RecConst K, const double *Jin,
                                  RecCells L) {
  const int b = blockIdx.y;
  if (!P.active[b]) return;
  const int jl = blockIdx.x * blockDim.x + threadIdx.x;
  const int Nl = P.Nl;
  if (jl >= L.n_total) return;
  const int k = L.cell(jl, Nl);
  const double J0 = Jin[((size_t)b * 3 + 0) * Nl + k], J1 = Jin[((size_t)b * 3 + 1) * Nl + k],
               J2 = Jin[((size_t)b * 3 + 2) * Nl + k];
  const size_t c = (size_t)b * Nl + k;
  R.rec[0][c] = J2 * K.sH1_He2 + J1 * K.sH1_He1 + J0 * K.sH1_H1;
  R.rec[1][c] = J2 * K.sHe1_He2 + J1 * K.sHe1_He1;
  R.rec[2][c] = J2 * K.sHe2_He2;
  R.rec[3][c] = J2 * (K.E_th[2] - K.E_th[0]) * K.sH1_He2 * RB_EV_2_ERG +
                J1 * (K.E_th[1] - K.E_th[0]) * K.sH1_He1 * RB_EV_2_ERG;
  R.rec[4][c] = J2 * (K.E_th[2] - K.E_th[0]) * K.sHe1_He2 * RB_EV_2_ERG;  // sic
  R.rec[5][c] = 0.0;
}

}  // namespace rb
