// Portability layer of libbranchmpc: the solver text in bmpc_models.h / bmpc_solver.h is written once
// against these few primitives.  nvcc builds it as the sm_100a kernel (one warp = one problem,
// BMPC_LANES = 32); tests/hostsim builds the SAME text with g++ as a single-lane program
// (BMPC_LANES = 1, barriers and reductions degenerate) so that the algorithm can be checked against
// the oracle without a GPU.  The host-sim build is test infrastructure: the product library never
// contains or calls it.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define BMPC_HD __host__ __device__
#define BMPC_D __device__ __forceinline__
#define BMPC_DN __device__ __noinline__
#define BMPC_LANES 32
#else
#define BMPC_HD
#define BMPC_D inline
#define BMPC_DN inline
#define BMPC_LANES 1
#endif

typedef double real;

#define BMPC_FULL_MASK 0xffffffffu

BMPC_D void lanes_sync() {
#if defined(__CUDA_ARCH__)
  __syncwarp();
#endif
}
BMPC_D real lanes_max(real v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(BMPC_FULL_MASK, v, o));
#endif
  return v;
}
BMPC_D real lanes_sum(real v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BMPC_FULL_MASK, v, o);
#endif
  return v;
}
BMPC_D int lanes_sum_int(int v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BMPC_FULL_MASK, v, o);
#endif
  return v;
}
BMPC_D int lanes_or_int(int v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(BMPC_FULL_MASK, v, o);
#endif
  return v;
}
// hides a constant from the optimiser: out-of-line phase functions that take a runtime selector must not be cloned per
// call site (code size; nvcc 12.9 also emits unparsable '.specialized' clone names with -lineinfo)
BMPC_D int bmpc_opaque(int v) {
#if defined(__CUDA_ARCH__)
  asm volatile("" : "+r"(v));
#endif
  return v;
}
// num / den for a ratio test (den > 0): single precision on the device
BMPC_D real bmpc_ratio(real num, real den) {
#if defined(__CUDA_ARCH__)
  return (real)__fdividef((float)num, (float)den);
#else
  return num / den;
#endif
}
// a / b through the hardware reciprocal (one ulp from the IEEE quotient): the full double division is a long sequence with
// a slow path, and the expansion phase does ~200 of them per solve
BMPC_D real bmpc_div(real a, real b) {
#if defined(__CUDA_ARCH__)
  return a * __drcp_rn(b);
#else
  return a / b;
#endif
}
BMPC_D real bmpc_min(real a, real b) { return fmin(a, b); }
BMPC_D real bmpc_max(real a, real b) { return fmax(a, b); }
BMPC_D real bmpc_clamp(real v, real lo, real hi) { return fmin(fmax(v, lo), hi); }
// Transcendentals are called through one out-of-line copy each: the inlined double-precision exp / sincos bodies
// (~60-150 instructions per call site) otherwise dominate the code size of the tree-expansion phase.
BMPC_DN void bmpc_sincos(real a, real* s, real* c) {
#if defined(__CUDA_ARCH__)
  sincos(a, s, c);
#else
  *s = sin(a);
  *c = cos(a);
#endif
}
BMPC_DN real bmpc_exp(real a) { return exp(a); }
