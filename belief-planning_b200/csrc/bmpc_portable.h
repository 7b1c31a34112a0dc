// Portability layer of libbranchmpc: the solver text in bmpc_models.h / bmpc_solver.h is written once
// against these few primitives.  nvcc builds it as the sm_100a kernel: one TEAM of warps (one thread block: three warps
// for the default trees, up to eight for trees that leave the SM emptier) owns one problem, BMPC_LANES = blockDim.x lanes;
// the node-parallel passes spread the tree's nodes
// over all lanes of the team (97 highway nodes = one round of 96 lanes + the root), the tree sweeps use one lane per
// branch of a level.  tests/hostsim builds the SAME text with g++ as a single-lane program
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
#ifndef BMPC_TEAM_WARPS
#define BMPC_TEAM_WARPS 3
#endif
#define BMPC_MAX_TEAM_WARPS 8
// lanes of a team = threads of its block: 96 by default, more for trees that leave room on the SM (chosen per handle by the
// host, bmpc_api.cu configure_instance); device code only
#define BMPC_LANES ((int)blockDim.x)
#define BMPC_BLANES 32
#else
#define BMPC_HD
#define BMPC_D inline
#define BMPC_DN inline
#define BMPC_TEAM_WARPS 1
#define BMPC_LANES 1
#define BMPC_BLANES 1
#endif

typedef double real;

#define BMPC_FULL_MASK 0xffffffffu

// team_sync(): every lane of the team (block).  The tree sweeps (lanes = branches of a level) are run by the team's first
// warp alone (team_leader(), lane stride BMPC_BLANES, bsync() between levels) while the other warps wait at the next
// team_sync(): followers that walked through the sweep code as well cost 30 % of the sweep time (measured).
BMPC_D void team_sync() {
#if defined(__CUDA_ARCH__)
#if BMPC_TEAM_WARPS > 1
  __syncthreads();
#else
  __syncwarp();
#endif
#endif
}
BMPC_D void lanes_sync() { team_sync(); }
BMPC_D void bsync() {
#if defined(__CUDA_ARCH__)
  __syncwarp();
#endif
}
// ---- asynchronous staging of the next episode's state: 1-D bulk copies (TMA, cp.async.bulk) completing on an mbarrier ----
#if defined(__CUDACC__)
__device__ __forceinline__ unsigned bmpc_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bmpc_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bmpc_smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// orders this thread's earlier generic-proxy accesses of shared memory before later async-proxy (bulk copy) writes
__device__ __forceinline__ void bmpc_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bmpc_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bmpc_smem_addr(bar)), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes `bytes` on the mbarrier
__device__ __forceinline__ void bmpc_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(bmpc_smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(bmpc_smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ bool bmpc_mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bmpc_smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
#endif

BMPC_D bool team_leader() {
#if defined(__CUDA_ARCH__)
  return threadIdx.x < 32;
#else
  return true;
#endif
}
// Team-wide reductions: shuffles inside a warp, then one shared-memory slot per warp.  Every lane of the team gets the
// same value (the solver branches on these results, so they must be uniform over the block).
#if defined(__CUDA_ARCH__) && BMPC_TEAM_WARPS > 1
#define BMPC_TEAM_COMBINE(T, v, OP)                                  \
  {                                                                  \
    __shared__ T red_[BMPC_MAX_TEAM_WARPS];                          \
    __syncthreads();                                                 \
    if ((threadIdx.x & 31) == 0) red_[threadIdx.x >> 5] = v;         \
    __syncthreads();                                                 \
    v = red_[0];                                                     \
    /* unrolled with an early exit: a loop with a runtime trip count here cost 6 % of the whole kernel (measured) */ \
    _Pragma("unroll") for (int w_ = 1; w_ < BMPC_MAX_TEAM_WARPS; ++w_) { \
      if (w_ >= (BMPC_LANES >> 5)) break;                            \
      const T o_ = red_[w_];                                         \
      v = OP;                                                        \
    }                                                                \
  }
#else
#define BMPC_TEAM_COMBINE(T, v, OP)
#endif
BMPC_D real lanes_max(real v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(BMPC_FULL_MASK, v, o));
  BMPC_TEAM_COMBINE(real, v, fmax(v, o_))
#endif
  return v;
}
BMPC_D real lanes_sum(real v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BMPC_FULL_MASK, v, o);
  BMPC_TEAM_COMBINE(real, v, v + o_)
#endif
  return v;
}
BMPC_D int lanes_sum_int(int v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BMPC_FULL_MASK, v, o);
  BMPC_TEAM_COMBINE(int, v, v + o_)
#endif
  return v;
}
BMPC_D int lanes_or_int(int v) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(BMPC_FULL_MASK, v, o);
  BMPC_TEAM_COMBINE(int, v, v | o_)
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
BMPC_D real bmpc_nan() { return nan(""); }
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
// sin and cos of one angle as straight-line code (device): three-term Cody-Waite reduction by pi/2 with fused multiply-adds and
// the degree-13 / degree-14 minimax kernels of fdlibm (k_sin.c, k_cos.c), within 1 ulp for |a| < 1e5.  No branch and no call, so
// several of them written one after the other are scheduled into each other - the three rollouts of a tree level (bmpc_sincos3)
// pay one latency instead of three out-of-line calls.  Larger or non-finite arguments take the library routine.
#if defined(__CUDACC__)
__device__ __forceinline__ void bmpc_sincos_core(real a, real& s, real& c) {
  const real magic = 6755399441055744.0;   // 1.5 * 2^52: adding it rounds to the nearest integer and leaves it in the low word
  const real t = fma(a, 0.6366197723675814, magic);
  const int q = __double2loint(t);
  const real j = t - magic;
  real r = fma(-j, 1.5707963267948966, a);
  r = fma(-j, 6.123233995736766e-17, r);
  r = fma(-j, -1.4973849048591698e-33, r);
  const real z = r * r;
  real ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  real pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const real sr = fma(ps * z, r, r);
  const real cr = fma(z * z, pc, fma(z, -0.5, 1.0));
  const real s0 = (q & 1) ? cr : sr, c0 = (q & 1) ? sr : cr;
  s = (q & 2) ? -s0 : s0;
  c = ((q + 1) & 2) ? -c0 : c0;
}
#endif
BMPC_D void bmpc_sincos_inline(real a, real& s, real& c) {
#if defined(__CUDA_ARCH__)
  if (fabs(a) < 1.0e5) {
    bmpc_sincos_core(a, s, c);
    return;
  }
#endif
  bmpc_sincos(a, &s, &c);
}
// sc = {sin a0, cos a0, sin a1, cos a1, sin a2, cos a2}
BMPC_D void bmpc_sincos3(real a0, real a1, real a2, real* sc) {
#if defined(__CUDA_ARCH__)
  if (fmax(fabs(a0), fmax(fabs(a1), fabs(a2))) < 1.0e5) {
    bmpc_sincos_core(a0, sc[0], sc[1]);
    bmpc_sincos_core(a1, sc[2], sc[3]);
    bmpc_sincos_core(a2, sc[4], sc[5]);
    return;
  }
#endif
  bmpc_sincos(a0, &sc[0], &sc[1]);
  bmpc_sincos(a1, &sc[2], &sc[3]);
  bmpc_sincos(a2, &sc[4], &sc[5]);
}
