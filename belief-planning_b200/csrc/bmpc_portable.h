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
