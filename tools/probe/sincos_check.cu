// accuracy of bmpc_sincos_core (csrc/bmpc_portable.h) against the CUDA library's sincos: maximum difference in ulps over a sweep
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I belief-planning_b200/csrc tools/probe/sincos_check.cu -o build/sincos_check
#include <cstdio>
#include <cmath>
#include "bmpc_portable.h"

__global__ void check(int n, double lo, double hi, double* out) {
  double worst_s = 0, worst_c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double a = lo + (hi - lo) * ((double)i / n);
    double s, c, s2, c2;
    bmpc_sincos_core(a, s, c);
    sincos(a, &s2, &c2);
    // ulp of the result's binade
    const double us = fabs(s - s2) / fmax(ldexp(1.0, ilogb(s2) - 52), 1e-300), uc = fabs(c - c2) / fmax(ldexp(1.0, ilogb(c2) - 52), 1e-300);
    worst_s = fmax(worst_s, us);
    worst_c = fmax(worst_c, uc);
  }
  atomicMax((unsigned long long*)&out[0], __double_as_longlong(worst_s));
  atomicMax((unsigned long long*)&out[1], __double_as_longlong(worst_c));
}

int main() {
  double* d;
  cudaMalloc(&d, 16);
  const double ranges[][2] = {{-0.8, 0.8}, {-3.2, 3.2}, {-100.0, 100.0}, {-9.9e4, 9.9e4}, {-1e-8, 1e-8}};
  for (auto& r : ranges) {
    cudaMemset(d, 0, 16);
    check<<<296, 256>>>(1 << 26, r[0], r[1], d);
    double h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("[%g, %g]: max |sin - lib| = %.2f ulp, max |cos - lib| = %.2f ulp\n", r[0], r[1], h[0], h[1]);
  }
  return 0;
}
