// Micro-benchmark: cost (SM cycles per warp-instruction) of shared-memory loads for different
// per-lane address patterns on sm_100a.  Used to decide the operand layout of the FFMA GEMM loops.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bench lds_bench.cu && ./lds_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int WIDTH>  // bytes per lane: 4, 8, 16
__global__ void __launch_bounds__(256) k(const int* __restrict__ lane_chunk, int iters, float* out, long long* cyc) {
  __shared__ __align__(16) float sm[8192];
  for (int i = threadIdx.x; i < 8192; i += 256) sm[i] = i * 1e-3f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  unsigned base = (unsigned)__cvta_generic_to_shared(sm) + lane_chunk[lane] * 16;
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      unsigned addr = base + ((it + u) & 7) * 2048;   // 8 slabs of 2 KB: same pattern, varying slab
      if (WIDTH == 16) {
        float x, y, z, w;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(addr));
        a0 += x; a1 += y; a2 += z; a3 += w;
      } else if (WIDTH == 8) {
        float x, y;
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(addr));
        a0 += x; a1 += y;
      } else {
        float x;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr));
        a0 += x;
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * 256 + threadIdx.x] = a0 + a1 + a2 + a3;
}

int main() {
  const char* names[] = {"all lanes same chunk", "chunk=lane (512B distinct)", "chunk=lane%8 (128B)", "chunk=lane%4 (64B)",
                         "chunk=lane/8 (4 distinct, per quarter)", "chunk=5*(lane%10) stride80B x10", "chunk=5*(lane%5) stride80B x5",
                         "chunk=lane/10 (3 distinct)", "chunk=lane%16 (256B)", "chunk=lane%2", "chunk=lane/16 (2, per half)",
                         "chunk=(lane%8)*5 stride 80B x8", "chunk=lane/5 (6 distinct+)", "chunk=lane%10 (160B contiguous)"};
  const int NP = 14;
  int pat[NP][32];
  for (int l = 0; l < 32; ++l) {
    pat[0][l] = 0; pat[1][l] = l; pat[2][l] = l % 8; pat[3][l] = l % 4; pat[4][l] = l / 8; pat[5][l] = 5 * (l % 10);
    pat[6][l] = 5 * (l % 5); pat[7][l] = l / 10; pat[8][l] = l % 16; pat[9][l] = l % 2; pat[10][l] = l / 16;
    pat[11][l] = (l % 8) * 5; pat[12][l] = l / 5; pat[13][l] = l % 10;
  }
  int* d_pat; float* d_out; long long* d_cyc;
  cudaMalloc(&d_pat, 32 * sizeof(int)); cudaMalloc(&d_out, 148 * 256 * sizeof(float)); cudaMalloc(&d_cyc, 148 * sizeof(long long));
  const int iters = 2000;
  for (int w = 0; w < 3; ++w) {
    int width = w == 0 ? 16 : (w == 1 ? 8 : 4);
    printf("---- ld.shared %d bytes/lane: SM cycles per warp-instruction (8 warps/SM, 1 CTA/SM) ----\n", width);
    for (int p = 0; p < NP; ++p) {
      cudaMemcpy(d_pat, pat[p], 32 * sizeof(int), cudaMemcpyHostToDevice);
      for (int rep = 0; rep < 2; ++rep) {
        if (width == 16) k<16><<<148, 256>>>(d_pat, iters, d_out, d_cyc);
        else if (width == 8) k<8><<<148, 256>>>(d_pat, iters, d_out, d_cyc);
        else k<4><<<148, 256>>>(d_pat, iters, d_out, d_cyc);
      }
      cudaDeviceSynchronize();
      long long c[148];
      cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += c[i]; avg /= 148;
      printf("  %-42s %.2f cycles/instr (SM-wide)\n", names[p], avg / (iters * 16.0 * 8.0));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
