// Microbenchmark: throughput of 3-register FFMA against the packed FFMA2 (fma.rn.f32x2) on sm_100a, at the warp counts
// of the cell-update phase (3 or 4 warps per SM sub-partition).   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(int iters, float* out) {
  float x[16], y[16], z[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = 1.0f + 1e-6f * (i + 1); z[i] = 1e-7f * (i + threadIdx.x); }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], y[i], z[i]);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float2 r = __ffma2_rn(make_float2(x[2 * i], x[2 * i + 1]), make_float2(y[2 * i], y[2 * i + 1]), make_float2(z[2 * i], z[2 * i + 1]));
        x[2 * i] = r.x; x[2 * i + 1] = r.y;
      }
    } else if (MODE == 2) {   // mul + add separately, scalar
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = x[i] * y[i];
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = x[i] + z[i];
    } else if (MODE == 3) {   // packed
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float2 r = __fmul2_rn(make_float2(x[2 * i], x[2 * i + 1]), make_float2(y[2 * i], y[2 * i + 1]));
        r = __fadd2_rn(r, make_float2(z[2 * i], z[2 * i + 1]));
        x[2 * i] = r.x; x[2 * i + 1] = r.y;
      }
    } else if (MODE == 4) {   // scalar FMA + MUFU mix (6 FFMA : 1 MUFU)
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], y[i], z[i]);
      float e;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x[0]));
      x[1] += e;
      asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x[2]));
      x[3] += e;
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 1234.5f) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1 + MODE] = (float)(t1 - t0) / iters;
}

int main() {
  float* d; cudaMalloc(&d, 64 * 4);
  const char* names[] = {"16 x FFMA (3 registers)", "8 x FFMA2 (= 16 FMA)", "16 x FMUL + 16 x FADD", "8 x FMUL2 + 8 x FADD2", "16 FFMA + 2 MUFU + 2 FADD"};
  for (int threads : {384, 512}) {
    k<0><<<148, threads>>>(2000, d); k<1><<<148, threads>>>(2000, d); k<2><<<148, threads>>>(2000, d); k<3><<<148, threads>>>(2000, d); k<4><<<148, threads>>>(2000, d);
    cudaDeviceSynchronize();
    float h[8]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    for (int m = 0; m < 5; ++m) printf("%d threads/CTA (%d warps per sub-partition): %-28s %.1f cycles per iteration\n", threads, threads / 128, names[m], h[1 + m]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
