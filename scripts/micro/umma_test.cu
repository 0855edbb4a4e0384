// Bring-up test for the tcgen05 building blocks of the planned 3xTF32 gate contraction:
//   D[128 x N] (TMEM, fp32) = A[128 x K] (TMEM, tf32 hi/lo) * B[N x K]^T (smem, K-major, no swizzle, tf32 hi/lo)
// Checks descriptor encodings, tcgen05.st/ld row ownership, commit/mbarrier and the 3xTF32 error,
// and times the MMA sequence.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_test umma_test.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128;
#ifndef NN
#define NN 208
#endif
#ifndef KK
#define KK 104
#endif
constexpr int N = NN, K = KK;
constexpr int COL_D = 0, COL_AHI = (K > 104 ? 112 : 256), COL_ALO = (K > 104 ? 312 : 384);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;   // descriptor version (Blackwell)
  return d;          // layout_type = 0 (no swizzle), base_offset = 0
}

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

template <int NCOL>
__device__ __forceinline__ void tmem_st(uint32_t taddr, const float* v) {
  if constexpr (NCOL == 4) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
                 : "memory");
  } else {
    static_assert(NCOL % 4 == 0, "");
    tmem_st<4>(taddr, v);
    tmem_st<NCOL - 4>(taddr + 4, v + 4);
  }
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(256, 1) umma_test_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out,
                                                           int mode, int reps, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // B operand, canonical K-major no-swizzle: [k_chunk (K/4)][N][16 bytes]; hi then lo
  float* b_hi = reinterpret_cast<float*>(smem);
  float* b_lo = b_hi + N * K;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // B -> smem
  for (int i = tid; i < N * K; i += 256) {
    int n = i / K, k = i % K;
    float w = B[i];
    float hi = tf32_rna(w);
    int off = (k / 4) * (N * 4) + n * 4 + (k % 4);
    b_hi[off] = hi;
    b_lo[off] = tf32_rna(w - hi);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const int row = 32 * (warp & 3) + lane, half = warp >> 2;
  const uint32_t lane_addr = tbase + ((uint32_t)(32 * (warp & 3)) << 16);

  // A -> TMEM (each thread: its row, its half of the K columns: 52 = 32 + 16 + 4)
  {
    constexpr int HK = K / 2;
    float hi[HK], lo[HK];
#pragma unroll
    for (int j = 0; j < HK; ++j) {
      float a = A[row * K + half * HK + j];
      hi[j] = tf32_rna(a);
      lo[j] = tf32_rna(a - hi[j]);
    }
    tmem_st<HK>(lane_addr + COL_AHI + half * HK, hi);
    tmem_st<HK>(lane_addr + COL_ALO + half * HK, lo);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const uint32_t lbo = N * 16, sbo = 128;
  uint32_t parity = 0;
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    if (tid == 0) {
      uint32_t acc = 0;
      // term list as (a_lo?, b_lo?) pairs
      int ta[4], tb[4], nterm = 0;
      if (mode == 1) { ta[0] = 0; tb[0] = 0; nterm = 1; }
      if (mode == 3) { ta[0] = 1; tb[0] = 0; ta[1] = 0; tb[1] = 1; ta[2] = 0; tb[2] = 0; nterm = 3; }
      if (mode == 4) { ta[0] = 1; tb[0] = 1; ta[1] = 1; tb[1] = 0; ta[2] = 0; tb[2] = 1; ta[3] = 0; tb[3] = 0; nterm = 4; }
      if (mode == 5) { ta[0] = 0; tb[0] = 0; ta[1] = 1; tb[1] = 0; ta[2] = 0; tb[2] = 1; nterm = 3; }
      if (mode == 12 || mode == 13) {
        const int nch = mode == 13 ? 4 : 2;
        const int c0[4] = {0, mode == 13 ? 48 : 96, 96, 160};
        const int cn[4] = {mode == 13 ? 48 : 96, mode == 13 ? 48 : 112, 64, 48};
        for (int ch = 0; ch < nch; ++ch) {
          const uint32_t idc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(cn[ch] >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
          acc = 0;
          for (int term = 0; term < 3; ++term) {
            const uint32_t a_col = term == 0 ? COL_ALO : COL_AHI;
            const uint32_t b_addr = smem_u32(term == 1 ? b_lo : b_hi) + (c0[ch] / 8) * 128;
#pragma unroll 1
            for (int ks = 0; ks < K / 8; ++ks) {
              uint64_t bd = make_desc(b_addr + ks * 2 * lbo, lbo, sbo);
              mma_tf32_ts(tbase + COL_D + c0[ch], tbase + a_col + ks * 8, bd, idc, acc);
              acc = 1;
            }
          }
        }
      } else
      for (int term = 0; term < nterm; ++term) {
        const uint32_t a_col = ta[term] ? COL_ALO : COL_AHI;
        const uint32_t b_addr = smem_u32(tb[term] ? b_lo : b_hi);
#pragma unroll 1
        for (int ks = 0; ks < K / 8; ++ks) {
          uint64_t bd = make_desc(b_addr + ks * 2 * lbo, lbo, sbo);
          mma_tf32_ts(tbase + COL_D, tbase + a_col + ks * 8, bd, idesc, acc);
          acc = 1;
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    // everybody waits for the accumulator
    {
      uint32_t done = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done)
                     : "r"(smem_u32(&mbar)), "r"(parity)
                     : "memory");
      }
      parity ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  long long t1 = clock64();
  if (tid == 0) cycles[0] = t1 - t0;

  // D -> global (thread: its row, its half of the N columns: 104 = 13 x 8)
  for (int c = 0; c < N / 2; c += 8) {
    float v[8];
    tmem_ld8(lane_addr + COL_D + half * (N / 2) + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[row * N + half * (N / 2) + c + i] = v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

int main() {
  std::vector<float> A(M * K), B(N * K), out(M * N);
  srand(1);
  for (auto& v : A) { float r = (rand() / (float)RAND_MAX) * 2 - 1; v = r * r * r; }   // mostly small, like h
  for (auto& v : B) v = ((rand() / (float)RAND_MAX) * 2 - 1) * 0.3f;
  std::vector<double> ref(M * N);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * (double)B[n * K + k];
      ref[m * N + n] = s;
    }
  float *dA, *dB, *dO;
  long long* dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, out.size() * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem_bytes = 2 * N * K * 4;
  cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  for (int dist = 0; dist < 2; ++dist) {
  for (auto& v : A) { float r = (rand() / (float)RAND_MAX) * 2 - 1; v = dist ? r : r * r * r; }
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s2 = 0; for (int k = 0; k < K; ++k) s2 += (double)A[m * K + k] * (double)B[n * K + k]; ref[m * N + n] = s2; }
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  for (int mode : {3, 12, 13}) {
    for (int reps : {1, 100}) {
      cudaMemset(dO, 0, out.size() * 4);
      umma_test_kernel<<<1, 256, smem_bytes>>>(dA, dB, dO, mode, reps, dC);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
      long long cyc; cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0, maxref = 0, sq = 0, bias = 0, bias2 = 0; int bad = -1;
      for (int i = 0; i < M * N; ++i) {
        double e2 = fabs(out[i] - ref[i]);
        sq += e2 * e2;
        bias += (out[i] - ref[i]) * (ref[i] > 0 ? 1.0 : -1.0);
        bias2 += (out[i] - ref[i]);
        if (e2 > maxerr) { maxerr = e2; bad = i; }
        maxref = fmax(maxref, fabs(ref[i]));
      }
      double mabs = 0; for (int i = 0; i < M * N; ++i) mabs += fabs(ref[i]); mabs /= (M * N);
      printf("K=%d N=%d dist=%d: rms err %.3e bias %+.3e = %+.3f ulp of mean|ref| ", K, N, dist, sqrt(sq / (M * N)), bias / (M * N), bias / (M * N) / mabs / 1.1920929e-7);
      printf("mode %dxTF32 reps %3d: max|err| %.3e (max|ref| %.3f, rel %.2e, worst at row %d col %d: got %.6f want %.6f)  cycles/rep %.0f\n",
             mode, reps, maxerr, maxref, maxerr / maxref, bad / N, bad % N, out[bad], ref[bad], (double)cyc / reps);
    }
  }
  }
  // fp32 FFMA reference error for comparison
  double maxerr = 0, sq = 0, cb = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s = fmaf(A[m * K + k], B[n * K + k], s);
      double e = fabs((double)s - ref[m * N + n]);
      maxerr = fmax(maxerr, e); sq += e * e; cb += ((double)s - ref[m * N + n]) * (ref[m * N + n] > 0 ? 1.0 : -1.0);
    }
  printf("plain fp32 fmaf chain: rms err %.3e mean(err*sign) %+.3e max|err| %.3e\n", sqrt(sq / (M * N)), cb / (M * N), maxerr);
  return 0;
}
