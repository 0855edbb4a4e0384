// Does warp-group register reallocation work from a 512-thread / 128-register launch?  (development probe)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k(float* out, int mode) {
  float acc[96];
  if (threadIdx.x < 128) {
    if (mode == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    out[threadIdx.x] = 1.f;
  } else {
    if (mode == 1) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
#pragma unroll
    for (int i = 0; i < 96; ++i) acc[i] = out[1024 + i] * threadIdx.x;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 96; ++i) acc[i] = fmaf(acc[i], acc[(i + 1) % 96], 1.f);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 96; ++i) s += acc[i];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}
int main() {
  float* d; cudaMalloc(&d, 4096 * 4); cudaMemset(d, 0, 4096 * 4);
  for (int mode = 0; mode < 2; ++mode) {
    k<<<148, 512>>>(d, mode);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d: %s\n", mode, cudaGetErrorString(e)); fflush(stdout);
  }
  return 0;
}
