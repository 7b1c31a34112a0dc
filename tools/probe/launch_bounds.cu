// Diagnostic: does __launch_bounds__(96, 4) make a launch with 148 KB of dynamic shared memory fail?
#include <cstdio>
#include <cuda_runtime.h>
__shared__ double red_a[3];
template <int MINB>
__global__ void __launch_bounds__(96, MINB) k(double* out) {
  extern __shared__ double sm[];
  __shared__ double red_b[3];
  sm[threadIdx.x] = threadIdx.x;
  if (threadIdx.x < 3) { red_a[threadIdx.x] = 1; red_b[threadIdx.x] = 2; }
  __syncthreads();
  out[blockIdx.x * 96 + threadIdx.x] = sm[95 - threadIdx.x] + red_a[0] + red_b[1];
}
template <int MINB>
void run(size_t smem) {
  double* out; cudaMalloc(&out, 148 * 96 * 8);
  cudaError_t e1 = cudaFuncSetAttribute(k<MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per = -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k<MINB>, 96, smem);
  k<MINB><<<148, 96, smem>>>(out);
  cudaError_t e2 = cudaGetLastError();
  cudaError_t e3 = cudaDeviceSynchronize();
  printf("minblocks %d smem %zu: setattr %s, occupancy %d, launch %s, sync %s\n", MINB, smem, cudaGetErrorString(e1), per,
         cudaGetErrorString(e2), cudaGetErrorString(e3));
  cudaFree(out);
}
int main() {
  for (size_t smem : {46392ul, 124792ul, 148560ul, 200000ul}) { run<1>(smem); run<4>(smem); }
  return 0;
}
