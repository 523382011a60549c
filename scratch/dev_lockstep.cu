#include <cstdio>
#include "../rabacus_b200/csrc/rb_kernels.cuh"
using namespace rb;
template <int C>
__global__ void kern(double *out) {
  double nH[C], nHe[C], T[C];
  PhotoRates p[C]; rb_flag_t valid[C]; rb_frac_t x[C];
  const double base[4] = {1e-4, 3e-4, 1e-3, 1e-2};
  for (int i = 0; i < C; ++i) { nH[i] = base[i % 4] * (1.0 + threadIdx.x); nHe[i] = nH[i]*0.08; p[i] = {8e-13,5e-13,6e-15,5e-24,6e-24,2e-25}; valid[i] = threadIdx.x < 3; T[i]=0; }
  CaseA fc = {1,1,1};
  double dudT[C], Tl[C];
  for (int i = 0; i < C; ++i) Tl[i] = 7500.0;
  int r0 = dudT_lockstep<C>(nH, nHe, Tl, p, valid, 3.0, 0.0, fc, 1e-7, dudT);
  int rc = pcte_lockstep<C>(nH, nHe, p, valid, 3.0, 0.0, fc, 1e-7, x, T);
  if (threadIdx.x == 0) {
    for (int i = 0; i < C; ++i) { out[i] = T[i]; out[8 + i] = dudT[i]; }
    out[16] = rc; out[17] = r0;
  }
}
int main() {
  double *d; cudaMalloc(&d, 32 * 8); double h[32];
  kern<4><<<1, 32>>>(d); cudaMemcpy(h, d, 32 * 8, cudaMemcpyDeviceToHost);
  printf("C=4 rc=%g r0=%g T = %.6f %.6f %.6f %.6f dudT(7500) = %.4e %.4e %.4e %.4e\n", h[16], h[17], h[0], h[1], h[2], h[3], h[8], h[9], h[10], h[11]);
  kern<1><<<1, 32>>>(d); cudaMemcpy(h, d, 32 * 8, cudaMemcpyDeviceToHost);
  printf("C=1 rc=%g T = %.6f dudT(7500) = %.4e  err=%s\n", h[16], h[0], h[8], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
