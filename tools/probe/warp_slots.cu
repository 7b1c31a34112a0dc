// Diagnostic: which hardware warp slots (%warpid) and SM does each warp of a 96/64/32-thread block with 46 KB of dynamic
// shared memory land on?  And what does one block barrier cost when only warp 0 does work?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(int* out, long long* cyc, int iters) {
  extern __shared__ double sm[];
  unsigned wid, sid;
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(sid));
  if ((threadIdx.x & 31) == 0) {
    out[(blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) * 2] = sid;
    out[(blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) * 2 + 1] = wid;
  }
  // barrier cost: warp 0 does a short dependent chain between barriers
  double a = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (threadIdx.x < 32) { a = a * 1.0000001 + 1.0; a = a * 1.0000001 + 1.0; a = a * 1.0000001 + 1.0; a = a * 1.0000001 + 1.0; }
    __syncthreads();
  }
  long long t1 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (threadIdx.x < 32) { a = a * 1.0000001 + 1.0; a = a * 1.0000001 + 1.0; a = a * 1.0000001 + 1.0; a = a * 1.0000001 + 1.0; }
    __syncwarp();
  }
  long long t2 = clock64();
  sm[threadIdx.x] = a;
  if (threadIdx.x == 0) { cyc[blockIdx.x * 2] = t1 - t0; cyc[blockIdx.x * 2 + 1] = t2 - t1; }
}
int main() {
  for (int threads : {32, 64, 96, 128}) {
    int blocks = 148 * 4, wpb = threads / 32;
    int* out; long long* cyc;
    cudaMallocManaged(&out, blocks * wpb * 2 * sizeof(int));
    cudaMallocManaged(&cyc, blocks * 2 * sizeof(long long));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 46392);
    probe<<<blocks, threads, 46392>>>(out, cyc, 1000);
    cudaDeviceSynchronize();
    printf("threads %d: %s\n", threads, cudaGetErrorString(cudaGetLastError()));
    // print the blocks that landed on SM of block 0
    int sm0 = out[0];
    for (int b = 0; b < blocks; ++b)
      if (out[b * wpb * 2] == sm0) {
        printf("  block %d sm %d warp slots:", b, sm0);
        for (int w = 0; w < wpb; ++w) printf(" %d", out[(b * wpb + w) * 2 + 1]);
        printf("   syncthreads loop %.1f cyc/iter, syncwarp loop %.1f cyc/iter\n", cyc[b * 2] / 1000.0, cyc[b * 2 + 1] / 1000.0);
      }
    cudaFree(out); cudaFree(cyc);
  }
  return 0;
}
