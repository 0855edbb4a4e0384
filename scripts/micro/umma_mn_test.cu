// Feasibility test for the weight-gradient product of the surrogate training path on tcgen05:
//   D[gate m (128 of 208) x n (112)] = sum over samples s of dG[s][m] * act[s][n]      (K = samples)
// with BOTH operands MN-major in shared memory, no swizzle: image [mn / 8][S samples][8 halves], i.e. exactly the K-major
// operand images the roll-out kernels already write (row = sample), read "transposed".  Descriptor: SBO = stride between
// groups of 8 along M/N = S * 16 bytes, LBO = stride between core matrices of 8 samples along K = 128 bytes; instruction
// descriptor bits 15 / 16 = MN-major A / B.  variant 1 swaps LBO and SBO (to find out which reading the hardware uses).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
constexpr int S = 64, MG = 208, N = 112, M = 128;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__global__ void __launch_bounds__(128, 1) kern(const float* __restrict__ G, const float* __restrict__ A, float* __restrict__ out, int variant, int m0, int reps, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __half* g_img = reinterpret_cast<__half*>(smem);            // [MG/8][S][8]
  __half* a_img = g_img + MG * S;                             // [N/8][S][8]
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
  for (int i = tid; i < (MG + 48) * S; i += 128) g_img[i] = __float2half_rn(0.f);   // rows beyond 208 read as zero
  __syncthreads();
  for (int i = tid; i < S * MG; i += 128) { int s = i / MG, g = i % MG; g_img[(g / 8) * (S * 8) + s * 8 + (g % 8)] = __float2half_rn(G[i]); }
  for (int i = tid; i < S * N; i += 128) { int s = i / N, n = i % N; a_img[(n / 8) * (S * 8) + s * 8 + (n % 8)] = __float2half_rn(A[i]); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_addr = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
  const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const uint32_t group_stride = S * 16, k_stride = 128;
  const uint32_t lbo = variant == 0 ? k_stride : group_stride, sbo = variant == 0 ? group_stride : k_stride;
  uint32_t parity = 0;
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    if (tid == 0) {
      const uint32_t ga = smem_u32(g_img) + (uint32_t)(m0 / 8) * group_stride, aa = smem_u32(a_img);
      for (int ks = 0; ks < S / 16; ++ks)
        mma_f16_ss(tbase, make_desc(ga + ks * 2 * k_stride, lbo, sbo), make_desc(aa + ks * 2 * k_stride, lbo, sbo), idesc, ks > 0);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(parity) : "memory");
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  long long t1 = clock64();
  if (tid == 0) cycles[0] = t1 - t0;
  const int row = 32 * warp + lane;
  for (int c = 0; c < N; c += 8) {
    float v[8];
    tmem_ld8(lane_addr + c, v);
    for (int i = 0; i < 8; ++i) out[row * N + c + i] = v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}
static float h_round(float x) { return __half2float(__float2half_rn(x)); }
int main() {
  std::vector<float> G(S * MG), A(S * N), out(M * N);
  srand(2);
  for (auto& v : G) v = h_round((rand() / (float)RAND_MAX) * 2 - 1);
  for (auto& v : A) v = h_round((rand() / (float)RAND_MAX) * 2 - 1);
  float *dG, *dA, *dO; long long* dC;
  cudaMalloc(&dG, G.size() * 4); cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dO, out.size() * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  const int smem_bytes = ((MG + 48) * S + N * S) * 2 + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  for (int variant = 0; variant < 2; ++variant)
    for (int m0 : {0, 128})
      for (int reps : {1, 200}) {
        kern<<<1, 128, smem_bytes>>>(dG, dA, dO, variant, m0, reps, dC);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: CUDA error: %s\n", variant, cudaGetErrorString(e)); return 1; }
        long long cyc; cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
          double s = 0; const int g = m0 + m;
          if (g < MG) for (int k = 0; k < S; ++k) s += (double)G[k * MG + g] * (double)A[k * N + n];
          maxerr = fmax(maxerr, fabs(out[m * N + n] - s)); maxref = fmax(maxref, fabs(s));
        }
        printf("variant %d (lbo %s) m0 %3d reps %3d: max err %.3e (max |ref| %.3f) cycles/rep %.0f (%d MMAs)\n", variant, variant == 0 ? "= K stride 128 B" : "= group stride",
               m0, reps, maxerr, maxref, (double)cyc / reps, S / 16);
      }
  return 0;
}
