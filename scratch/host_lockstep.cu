#include <cstdio>
#include "../rabacus_b200/csrc/rb_kernels.cuh"
using namespace rb;
int main() {
  const int C = 4;
  double nH[C] = {1e-4, 3e-4, 1e-3, 1e-2}, nHe[C], T[C];
  PhotoRates p[C]; bool valid[C]; rb_frac_t x[C];
  for (int i = 0; i < C; ++i) { nHe[i] = nH[i]*0.08; p[i] = {8e-13,5e-13,6e-15,5e-24,6e-24,2e-25}; valid[i] = true; T[i]=0; }
  CaseA fc = {1,1,1};
  int rc = pcte_lockstep<C>(nH, nHe, p, valid, 3.0, 0.0, fc, 1e-7, x, T);
  printf("rc=%d T = %.6f %.6f %.6f %.6f\n", rc, T[0], T[1], T[2], T[3]);
  double T1[1]; rb_frac_t x1[1]; double a[1]={nH[0]}, b[1]={nHe[0]}; PhotoRates p1[1]={p[0]}; bool v1[1]={true};
  rc = pcte_lockstep<1>(a, b, p1, v1, 3.0, 0.0, fc, 1e-7, x1, T1);
  printf("rc=%d T1 = %.6f\n", rc, T1[0]);
  return 0;
}
