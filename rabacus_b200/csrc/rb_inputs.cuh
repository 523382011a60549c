// rb_inputs.cuh -- device-side input producers for parameter sweeps.
//
// A parameter sweep of SphereBgnd problems (BASELINE.json configs[4]: 8192 solves over an
// (nH0, R, z) grid) needs, per problem, what the reference builds in Python before it calls
// the solver:
//   * the photon-energy grid and the HM12 spectrum at the problem's redshift
//       rad_src/source.py:178-259   log-spaced grid, thresholds snapped (host, once: Nnu values)
//       uv_bgnd/hm12.py:239-303     I_nu(z, E): linear in z between the two table columns,
//                                   log-log in wavelength (numpy.interp semantics)
//   * the Verner cross sections on the grid     verner_96.f90:108-154
//   * the density profile on the radial grid    doc/guide_bgnd_sphere_examples.rst:126-146
//       Edges = linspace(0, R, Nl+1), r = shell centres, r0 = R / core_div,
//       nH = nH0 (r < r0 ? 1 : (r/r0)^-slope), nHe = nH * he_ratio, T = T0
// and from those the plan's per-frequency table [Nnu][9] with the trapezoid rule folded into
// six weights, and the optically thin integrals (build_spec_table in rb_api.cu, same
// statements).  Here all of that is produced ON THE DEVICE from the three numbers per
// problem: nothing but (nH0, R, z)[B], the HM12 table and the Nnu-point energy grid crosses
// PCIe, and no per-problem host work remains.
#pragma once
#include "rb_kernels.cuh"

namespace rb {

struct Hm12Dev {
  int nlam, nz;
  const double *log_lam;   // [nlam] non-decreasing log10(lambda / cm)
  const double *z;         // [nz] ascending
  const double *logI;      // [nlam][nz] log10 I_nu
};

struct FamilyDev {
  int B, Nl, Nnu, pad_;
  const double *nH0, *R, *z, *Hz;   // [B]
  const double *E;                  // [Nnu] photon energies (eV), shared by the family
  double core_div, slope, he_ratio, T0;
  Hm12Dev hm;
  // outputs: the plan's input buffers
  double *Edges, *nH, *nHe, *T, *spec, *thin, *zHz;
  unsigned char *reversed;
};

// numpy.linspace(0, R, Nl+1)[k]: k * step with step = R / Nl, the end point exact
__device__ __forceinline__ double family_edge(double R, int Nl, int k) {
  return k == Nl ? R : (double)k * (R / (double)Nl);
}

// Edges / nH / nHe / T of every problem in SphereBgnd solver orientation (outermost shell
// first: sphere_bgnd.f90:182-202); grid (ceil((Nl+1)/128), B)
__global__ void __launch_bounds__(128) k_family_profiles(FamilyDev F) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 128 + threadIdx.x;   // solver index (0 = outermost)
  const int Nl = F.Nl;
  if (i > Nl) return;
  const double R = F.R[b];
  const int k = Nl - i;                           // input index (0 = centre)
  F.Edges[(size_t)b * (Nl + 1) + i] = family_edge(R, Nl, k);
  if (i == 0) {
    F.zHz[2 * b] = F.z[b];
    F.zHz[2 * b + 1] = F.Hz ? F.Hz[b] : 0.0;
    F.reversed[b] = 1;   // results go back in the input orientation (centre first)
  }
  if (i == Nl) return;
  // shell i spans input edges k-1 .. k
  const double r = 0.5 * (family_edge(R, Nl, k) + family_edge(R, Nl, k - 1));
  const double r0 = R / F.core_div;
  const double prof = r < r0 ? 1.0 : pow(r / r0, -F.slope);
  const double nH = F.nH0[b] * prof;
  const size_t c = (size_t)b * Nl + i;
  F.nH[c] = nH;
  F.nHe[c] = nH * F.he_ratio;
  F.T[c] = F.T0;
}

// log10 I_nu(z, E) of uv_bgnd/hm12.py:239-303 (columns i_lo / i_hi bracket z, z_frac linear),
// interpolated in log10(lambda) like numpy.interp (clamped at both ends, exact at the nodes)
__device__ __forceinline__ double hm12_logI(const Hm12Dev &H, int i_lo, int i_hi, double z_frac,
                                            double E_eV) {
  const double lam_cm = 2.99792458e10 / (E_eV / 4.135667603369524e-15);
  const double x = log10(lam_cm);
  auto f = [&](int j) {
    const double a = H.logI[(size_t)j * H.nz + i_lo], c = H.logI[(size_t)j * H.nz + i_hi];
    return a + z_frac * (c - a);
  };
  const int n = H.nlam;
  if (x <= H.log_lam[0]) return f(0);
  if (x >= H.log_lam[n - 1]) return f(n - 1);
  int lo = 0, hi = n - 1;   // log_lam[lo] <= x < log_lam[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (H.log_lam[mid] <= x) lo = mid; else hi = mid;
  }
  if (x == H.log_lam[lo]) return f(lo);   // numpy.interp: exact at a node (the table holds
                                          // repeated wavelengths at the spectral edges)
  RB_ASSERT(lo >= 0 && lo + 1 < n && H.log_lam[lo] <= x && x < H.log_lam[lo + 1]);
  const double f0 = f(lo), f1 = f(lo + 1);
  const double slope = (f1 - f0) / (H.log_lam[lo + 1] - H.log_lam[lo]);
  return slope * (x - H.log_lam[lo]) + f0;
}

// per-frequency table + thin integrals of every problem; one block per problem, one thread
// per frequency, the nine thin sums taken in frequency order (as build_spec_table does)
__global__ void __launch_bounds__(128) k_family_spectra(FamilyDev F, double s_th0, double s_th1,
                                                        double s_th2) {
  const int b = blockIdx.x;
  const int Nnu = F.Nnu;
  const Hm12Dev &H = F.hm;
  // hm12.py:262-275: the table column nearest z, and its neighbour on the other side of z
  const double z = F.z[b];
  int i_lo = 0;
  {
    double best = fabs(H.z[0] - z);
    for (int i = 1; i < H.nz; ++i) {
      const double d = fabs(H.z[i] - z);
      if (d < best) { best = d; i_lo = i; }
    }
  }
  int i_hi;
  if (H.z[i_lo] < z) { i_hi = i_lo + 1; } else { i_hi = i_lo; i_lo = i_lo - 1; }
  if (i_lo < 0) { i_lo = 0; i_hi = 1; }
  if (i_hi > H.nz - 1) { i_hi = H.nz - 1; i_lo = i_hi - 1; }
  RB_ASSERT(i_lo >= 0 && i_hi == i_lo + 1 && i_hi < H.nz);
  const double z_frac = (z - H.z[i_lo]) / (H.z[i_hi] - H.z[i_lo]);
  const double sth[3] = {s_th0, s_th1, s_th2};
  double *tab = F.spec + (size_t)b * Nnu * kSpecStride;
  for (int k = threadIdx.x; k < Nnu; k += blockDim.x) {
    const double E = F.E[k];
    const double shape = pow(10.0, hm12_logI(H, i_lo, i_hi, z_frac, E));
    const double dlo = (k > 0) ? (E - F.E[k - 1]) : 0.0;
    const double dhi = (k < Nnu - 1) ? (F.E[k + 1] - E) : 0.0;
    const double wk = 0.5 * (dlo + dhi);
    const double base = (4.0 * RB_PI * shape) / (E * RB_H_ERG_S);
    double *row = tab + (size_t)k * kSpecStride;
#pragma unroll
    for (int X = 0; X < 3; ++X) {
      const double sig = v96_sigma(X, E);
      row[X] = sig / sth[X];
      const double yi = base * sig;
      const double yh = yi * (E - v96_row(X).Eth) * RB_EV_2_ERG;
      row[3 + X] = yi * wk;
      row[6 + X] = yh * wk;
    }
  }
  __syncthreads();
  if (threadIdx.x < kThinStride) {
    const int q = threadIdx.x;
    double acc = 0.0;
    if (q < 6) {
      for (int k = 0; k < Nnu; ++k) acc += tab[(size_t)k * kSpecStride + 3 + q];
    } else {
      acc = tab[q - 6];
      for (int k = 1; k < Nnu; ++k) acc = fmin(acc, tab[(size_t)k * kSpecStride + q - 6]);
    }
    F.thin[(size_t)b * kThinStride + q] = acc;
  }
}

}  // namespace rb
