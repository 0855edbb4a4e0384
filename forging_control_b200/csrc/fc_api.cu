// C ABI of libforging_b200.so (see include/forging_b200.h) + sm_100a kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/forging_b200.h"

#define FC_HD_CTX __device__ __forceinline__
#include "fc_mpc_kernel.inl"
#include "fc_mpc_tc_kernel.inl"
#include "fc_mpc_pair_kernel.inl"
#include "fc_plant.cuh"

namespace fc {

// ---------------------------------------------------------------------------------------------------
// device execution context for the kernel body in fc_mpc_kernel.inl
// ---------------------------------------------------------------------------------------------------
extern __shared__ __align__(16) float fc_dyn_smem[];

struct DevCtx {
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int bid() const { return blockIdx.x; }
  __device__ __forceinline__ int nblk() const { return gridDim.x; }
  __device__ __forceinline__ float* smem() const { return fc_dyn_smem; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  static __device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  }
  static __device__ __forceinline__ float rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  }
  static __device__ __forceinline__ F4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void sts4(float* p, F4 v) { *reinterpret_cast<float4*>(p) = v; }
  static __device__ __forceinline__ F4 ldg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ F4 ldg4_stream(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ void stg4(float* p, F4 v) { *reinterpret_cast<float4*>(p) = v; }
  static __device__ __forceinline__ void stg4_stream(float* p, F4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
  static __device__ __forceinline__ void stg2_stream(float* p, float a, float b) { __stcs(reinterpret_cast<float2*>(p), make_float2(a, b)); }
  static __device__ __forceinline__ void ldg2_stream(const float* p, float& a, float& b) {
    const float2 v = __ldcs(reinterpret_cast<const float2*>(p));
    a = v.x; b = v.y;
  }
  static __device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }
  static __device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  }
  static __device__ __forceinline__ void cp_async8(float* dst, const float* src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
  }
  static __device__ __forceinline__ void cp_async4(float* dst, const float* src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
  }
  static __device__ __forceinline__ void sts1(float* p, float v) { *p = v; }
  static __device__ __forceinline__ void stg2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
  static __device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
  template <int N>
  static __device__ __forceinline__ void cp_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
  }
};

__global__ void __launch_bounds__(kThreads, 1) mpc_loss_kernel(const MpcParams p) {
  DevCtx ctx;
  MpcTile<DevCtx> k(ctx, p);
  k.run();
}

// ---------------------------------------------------------------------------------------------------
// execution context of the tcgen05 variant: TMEM, UMMA issue, mbarriers, bulk copies
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

}  // namespace fc
#include "fc_lstm_train.cuh"
#include "fc_fnn.cuh"
namespace fc {

template <int N> struct TmemIO;
template <> struct TmemIO<1> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(a) : "memory");
  }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(a), "r"(r[0]) : "memory");
  }
};
template <> struct TmemIO<2> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a) : "memory");
  }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(a), "r"(r[0]), "r"(r[1]) : "memory");
  }
};
template <> struct TmemIO<4> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
  }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
  }
};
template <> struct TmemIO<8> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
  }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
  }
};
template <> struct TmemIO<16> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(a) : "memory");
  }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
                 "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
  }
};
template <> struct TmemIO<32> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t* r) {
    TmemIO<16>::ld(a, r);
    TmemIO<16>::ld(a + 16, r + 16);
  }
};

#ifdef FC_TC_TRACE
__device__ long long g_trace[3][32768];
__device__ int g_trace_n[3];
#endif

struct DevCtxTC : DevCtx {
#ifdef FC_TC_TRACE
  static constexpr int kTraceMax = 32768;
  static __device__ __forceinline__ void trace_put(int slot, int i, long long v) { g_trace[slot][i] = v; }
  static __device__ __forceinline__ void trace_count(int slot, int n) { g_trace_n[slot] = n; }
#endif
  static __device__ __forceinline__ void prefetch_l2(const float* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
  // one instruction for a contiguous, 16-byte aligned region (bytes a multiple of 16)
  static __device__ __forceinline__ void prefetch_l2_bulk(const float* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
  }
  static __device__ __forceinline__ long long clock() { return clock64(); }
  static __device__ void report(const long long* tm) {
    printf("[fc timing, CTA 0 thread 0, cycles] other %lld | fwd: prologue %lld issue+input %lld mma-wait %lld pointwise %lld store+sync %lld glue %lld | bwd: glue %lld prologue %lld finish+sync %lld issue+shadow %lld mma-wait %lld post %lld\n",
           tm[0], tm[9], tm[1], tm[2], tm[3], tm[4], tm[10], tm[11], tm[12], tm[5], tm[6], tm[7], tm[8]);
  }
  static __device__ void report_pair(int tid, const long long* tm) {
    printf("[fc pair timing, CTA 0 thread %d, cycles] other %lld | fwd: wait-full %lld update %lld arrive %lld wait-ready %lld issue %lld | bwd: wait-full %lld update %lld arrive %lld wait-ready %lld issue %lld | swap %lld loads %lld fwd-t9 %lld prologue/tail %lld glue %lld | before: fwd-item %lld fwd-wait %lld prologue %lld fwd-sync %lld bwd-item %lld bwd-window %lld bwd-tail %lld service-item %lld\n",
           tid, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], tm[7], tm[8], tm[9], tm[10], tm[11], tm[12], tm[13], tm[14], tm[15],
           tm[16], tm[17], tm[18], tm[19], tm[20], tm[21], tm[22], tm[23]);
  }
  __device__ __forceinline__ void warp_sync() const { __syncwarp(); }
  static constexpr bool kAccTruncates = true;   // tcgen05 accumulates with truncation, see tc::acc_correction
  uint32_t tbase;        // TMEM base address of this CTA's 512-column allocation
  uint32_t lane_addr;    // tbase + first lane of this warp's quadrant
  uint32_t bar0;         // shared address of mbarrier 0

  __device__ __forceinline__ void tc_setup(float* bar_area) {
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar_area) + 16;
    bar0 = smem_u32(bar_area);
    if ((threadIdx.x >> 5) == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
      for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + i * 8) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tbase = *slot;
    lane_addr = tbase + ((uint32_t)(32 * ((threadIdx.x >> 5) & 3)) << 16);
  }
  __device__ __forceinline__ void tc_teardown() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
  }
  __device__ __forceinline__ void tc_sync() const {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  template <int N>
  __device__ __forceinline__ void tmem_ld(int col, float* v) const {
    uint32_t r[N];
    TmemIO<N>::ld(lane_addr + col, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = __uint_as_float(r[i]);
  }
  // split-phase load: the destination registers must not be read before tmem_ld_wait()
  template <int N>
  __device__ __forceinline__ void tmem_ld_nowait(int col, float* v) const {
    TmemIO<N>::ld(lane_addr + col, reinterpret_cast<uint32_t*>(v));
  }
  __device__ __forceinline__ void tmem_ld_wait() const { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
  template <int N>
  __device__ __forceinline__ void tmem_st(int col, const float* v) const {
    uint32_t r[N];
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = __float_as_uint(v[i]);
    TmemIO<N>::st(lane_addr + col, r);
  }
  __device__ __forceinline__ void tmem_st_wait() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
  // fp16 hi/lo split of two pre-scaled values, packed (first value in the low half) as 32-bit operand words.
  // The conversions saturate to the largest finite fp16 (F2FP.SATFINITE): no infinities reach the tensor core.
  static __device__ __forceinline__ void split_h2(float x0, float x1, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(x1 - hf.y), "f"(x0 - hf.x));
    hi = __uint_as_float(h);
    lo = __uint_as_float(l);
  }
  // D[128 x n] (+)= A[128 x 16*ksteps] * B[n x 16*ksteps]^T, one hi/lo term (kind::f16, fp32 accumulate);
  // A: 8 TMEM columns per k-step (two fp16 per column); B image = [k/8][n_img][8 halves]
  __device__ __forceinline__ void mma(int d_col, int n, int a_col, const float* b_img, int n_img, int row0, int ksteps, bool accumulate) const {
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t lbo = (uint32_t)n_img * 16u, sbo = 128u;
    const uint32_t b0 = smem_u32(b_img) + (uint32_t)(row0 >> 3) * 128u;
    // The issuing thread is on the critical path: keep the per-MMA instruction count minimal (descriptor and
    // TMEM address advance by constants) and unroll.
    uint64_t desc = (uint64_t)((b0 >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
    const uint64_t dstep = (uint64_t)((2u * lbo) >> 4);
    const uint32_t d_addr = tbase + d_col;
    uint32_t a_addr = tbase + a_col;
    if (!accumulate) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_addr),
                   "r"(a_addr), "l"(desc), "r"(idesc) : "memory");
      desc += dstep; a_addr += 8; --ksteps;
    }
#pragma unroll 4
    for (int ks = 0; ks < ksteps; ++ks) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_addr),
                   "r"(a_addr), "l"(desc), "r"(idesc) : "memory");
      desc += dstep; a_addr += 8;
    }
  }
  __device__ __forceinline__ void commit(int bar) const {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + bar * 8) : "memory");
  }
  __device__ __forceinline__ void bar_wait(int bar, unsigned phase) const {
    const uint32_t addr = bar0 + bar * 8, parity = phase & 1u;
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // ---- pair kernel: per-tile hand-shakes and shared-memory A operands ----
  __device__ __forceinline__ void bar_init(int bar, int count) const {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + bar * 8), "r"(count) : "memory");
  }
  __device__ __forceinline__ void bar_init_fence() const { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __device__ __forceinline__ void bar_arrive(int bar) const {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + bar * 8) : "memory");
  }
  // generic-proxy shared-memory writes -> visible to the tensor core (async proxy); TMEM accesses ordered before the arrive
  __device__ __forceinline__ void operand_fence() const {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  static __device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }
  __device__ __forceinline__ void tmem_fence() const { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
  static __device__ __forceinline__ void sts2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
  // D[128 x n] (+)= A[128 x 16*ksteps] * B[n x 16*ksteps]^T with BOTH operands in shared memory:
  // A image = [k/8][128][8 halves], B image = [k/8][n_img][8 halves]
  __device__ __forceinline__ void mma_ss(int d_col, int n, const float* a_img, const float* b_img, int n_img, int ksteps, bool accumulate) const {
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t lbo_b = (uint32_t)n_img * 16u, lbo_a = 128u * 16u, sbo = 128u;
    uint64_t bdesc = (uint64_t)((smem_u32(b_img) >> 4) & 0x3FFF) | ((uint64_t)((lbo_b >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
    uint64_t adesc = (uint64_t)((smem_u32(a_img) >> 4) & 0x3FFF) | ((uint64_t)((lbo_a >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
    const uint64_t bstep = (uint64_t)((2u * lbo_b) >> 4), astep = (uint64_t)((2u * lbo_a) >> 4);
    const uint32_t d_addr = tbase + d_col;
    if (!accumulate) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_addr),
                   "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
      adesc += astep; bdesc += bstep; --ksteps;
    }
#pragma unroll 4
    for (int ks = 0; ks < ksteps; ++ks) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_addr),
                   "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
      adesc += astep; bdesc += bstep;
    }
  }
  // wait of a warp that is off the critical path: back off between polls (frees issue slots, saves power)
  __device__ __forceinline__ void bar_wait_relaxed(int bar, unsigned phase) const {
    const uint32_t addr = bar0 + bar * 8, parity = phase & 1u;
    uint32_t done = 0;
    while (true) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(addr), "r"(parity) : "memory");
      if (done) break;
      __nanosleep(1000);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // the same in two parts: announce the bytes once, then any number of copies onto the barrier
  __device__ __forceinline__ void bulk_expect(int bar, int bytes) const {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + bar * 8), "r"((uint32_t)bytes) : "memory");
  }
  __device__ __forceinline__ void bulk_copy(float* dst, const float* src, int bytes, int bar) const {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"((uint32_t)bytes), "r"(bar0 + bar * 8) : "memory");
  }
  // one thread: bulk (TMA) copy global -> shared, completion on an mbarrier
  __device__ __forceinline__ void bulk_load(float* dst, const float* src, int nfloats, int bar) const {
    const uint32_t baddr = bar0 + bar * 8;
    const uint32_t bytes = (uint32_t)nfloats * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(baddr), "r"(bytes) : "memory");
    const uint32_t chunk = 32768u;
    for (uint32_t off = 0; off < bytes; off += chunk) {
      const uint32_t sz = bytes - off < chunk ? bytes - off : chunk;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst) + off),
                   "l"(reinterpret_cast<const char*>(src) + off), "r"(sz), "r"(baddr)
                   : "memory");
    }
  }
};

__global__ void __launch_bounds__(tc::kThreadsTC, 1) mpc_loss_tc_kernel(const MpcParams p) {
  DevCtxTC ctx;
  tc::MpcTileTC<DevCtxTC> k(ctx, p);
  k.run();
}

__global__ void __launch_bounds__(pr::kThreadsP, 1) mpc_loss_pair_kernel(const MpcParams p) {
  DevCtxTC ctx;
  pr::MpcPair<DevCtxTC> k(ctx, p);
  k.run();
}

// surrogate training mode of the same source (MpcParams::train), tanh with its small-argument polynomial
__global__ void __launch_bounds__(pr::kThreadsP, 1) lstm_train_pair_kernel(const MpcParams p) {
  DevCtxTC ctx;
  pr::MpcPair<DevCtxTC, 1, true> k(ctx, p);
  k.run();
}

// ... and its replica mode (32-sample tiles) for small training batches
__global__ void __launch_bounds__(pr::kThreadsP, 1) lstm_train_replica_kernel(const MpcParams p) {
  DevCtxTC ctx;
  pr::MpcPair<DevCtxTC, 4, true> k(ctx, p);
  k.run();
}

}  // namespace fc
#include "fc_lstm_train_tc.cuh"
namespace fc {

// replica mode of the same source: one 32-trajectory tile per CTA, the 50 hidden units of a trajectory split over
// twelve threads (small / mid-size batches, fc_mpc_pair_kernel.inl)
__global__ void __launch_bounds__(pr::kThreadsP, 1) mpc_loss_replica_kernel(const MpcParams p) {
  DevCtxTC ctx;
  pr::MpcPair<DevCtxTC, 4> k(ctx, p);
  k.run();
}

// pair-kernel operand images behind the two others in the packed buffer
__global__ void pack_weights_pair_kernel(RawWeights w, float* out, long base_off) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;        // half index
  const long n_halves = 2L * pr::kSmallOff;
  float* base = out + base_off;
  __half* oh = reinterpret_cast<__half*>(base);
  if (i < n_halves) {
    const pr::PrSlot s = pr::decode_half(i);
    const float v = (s.kind == 0 ? pr::fwd_weight(w, s.l, s.h) : pr::bwd_weight(w, s.l, s.h)) * pr::kScaleW;
    const __half hi = __float2half_rn(v);
    oh[i] = s.lo ? __float2half_rn(v - __half2float(hi)) : hi;
  } else if (i < n_halves + kSmallFloats) {
    const int j = (int)(i - n_halves);
    base[pr::kSmallOff + j] = packed_value(w, kFCW + j);
  }
}

// packed buffer = [FFMA layouts (kPackFloats) | tcgen05 operand images, fp16 hi/lo (tc::kPackFloatsTC floats) |
//                  pair-kernel operand images (pr::kPackFloatsP floats)]
__global__ void pack_weights_tc_kernel(RawWeights w, float* out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;        // half index
  const long n_halves = 2L * tc::kSmallOff;
  __half* oh = reinterpret_cast<__half*>(out + kPackFloats);
  if (i < n_halves) {
    const tc::TcSlot s = tc::tc_decode_half(i);
    const float v = (s.kind == 0 ? tc::fwd_weight(w, s.l, s.h) : tc::bwd_weight(w, s.l, s.h)) * tc::kScaleW;
    const __half hi = __float2half_rn(v);
    oh[i] = s.lo ? __float2half_rn(v - __half2float(hi)) : hi;
  } else if (i < n_halves + kSmallFloats) {
    const int j = (int)(i - n_halves);
    out[kPackFloats + tc::kSmallOff + j] = packed_value(w, kFCW + j);
  }
}

__global__ void pack_weights_kernel(RawWeights w, float* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kPackFloats) out[i] = packed_value(w, i);
}

// sums the per-CTA partials: one warp per output, lanes stride over CTAs, warp-shuffle tree.
__global__ void __launch_bounds__(1024) mpc_finalize_kernel(const double* __restrict__ partial, int grid, double loss_scale,
                                                            float* __restrict__ gl) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i <= kNumFnnGrad; i += 32) {
    double a = 0.0;
    for (int b = lane; b < grid; b += 32) a += partial[(size_t)b * kPartialStride + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) gl[i] = (float)(i == kNumFnnGrad ? a * loss_scale : a);
  }
  if (threadIdx.x > kNumFnnGrad && threadIdx.x < 256) gl[threadIdx.x] = 0.f;
}

// Training-sample construction on the device: SequenceDataset.__getitem__ (UL/Functions.py:109-132) over the
// per-trajectory slices of Data.get_individual_dataset (:479-516), for a batch of global sample indices.
// Pure gather (bit-exact); one thread per output float, outputs written coalesced.
// I = index type of the element loop (unsigned when everything fits 32 bits: the divisions are the cost here),
// LB = compile-time look-back (10 in the reference) or 0 = run-time value.
template <typename I, int LB>
__global__ void __launch_bounds__(256) build_windows_kernel(const float* __restrict__ Xtab, const float* __restrict__ ytab,
                                                            const float* __restrict__ Ztab, I t_traj, int lookback_rt,
                                                            const long long* __restrict__ idx, I B,
                                                            float* __restrict__ X, float* __restrict__ y, float* __restrict__ Z) {
  const I lookback = LB ? (I)LB : (I)lookback_rt;
  const I zper = lookback * 5;
  const I nz = B * zper, total = nz + B * 4;
  // element order: all Z floats first (the bulk, contiguous [B][lookback][5]), then X [B][3], then y [B]
  for (I e = (I)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (I)gridDim.x * blockDim.x) {
    if (e < nz) {
      const I s = e / zper, w = e - s * zper, r = w / 5, f = w - r * 5;
      const I g = (I)idx[s], k = g / t_traj, i = g - k * t_traj;
      const I back = lookback - 1 - r;                               // rows before the current one
      const I loc = i >= back ? i - back : 0;                        // :117-123: front padding with the trajectory's row 0
      Z[e] = __ldg(Ztab + ((size_t)k * t_traj + loc) * 5 + f);
    } else if (e < nz + B * 3) {
      const I q = e - nz, s = q / 3;
      X[q] = __ldg(Xtab + (size_t)idx[s] * 3 + (q - s * 3));
    } else {
      const I s = e - nz - B * 3, g = (I)idx[s], k = g / t_traj, i = g - k * t_traj;
      y[s] = __ldg(ytab + (size_t)k * t_traj + (i + 1 < t_traj ? i + 1 : t_traj - 1));   // :126-129: target of the next step
    }
  }
}

// out[i] = sum over CTAs of partial[b][i], i < n (one warp per output)
__global__ void __launch_bounds__(1024) sum_partials_kernel(const double* __restrict__ partial, int grid, int stride, int n,
                                                            float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 32 + warp; i < n; i += gridDim.x * 32) {
    double a = 0.0;
    for (int b = lane; b < grid; b += 32) a += partial[(size_t)b * stride + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) out[i] = (float)a;
  }
}

// register-resident FFMA loop: 8 independent chains per thread
__global__ void __launch_bounds__(256) ffma_peak_kernel(int iters, float* sink) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-6f + i;
  const float m = 1.0000001f, c = 1e-7f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], m, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678f) sink[0] = s;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, const char* a = "", long long v0 = 0, long long v1 = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, v0, v1);
  return code;
}

static int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "CUDA error in %s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
  return FC_ERR_CUDA;
}

#define FC_CUDA(call, where)                         \
  do {                                               \
    cudaError_t e__ = (call);                        \
    if (e__ != cudaSuccess) return cuda_fail(e__, where); \
  } while (0)

static int sm_count(int* out) {
  int dev = 0;
  FC_CUDA(cudaGetDevice(&dev), "cudaGetDevice");
  static int cached[64] = {0};
  if (dev >= 0 && dev < 64 && cached[dev]) { *out = cached[dev]; return FC_OK; }
  int n = 0;
  FC_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
  if (dev >= 0 && dev < 64) cached[dev] = n;
  *out = n;
  return FC_OK;
}

// cudaFuncSetAttribute is per device: remember which devices have been prepared (several GPUs in one process)
static int ensure_smem_attributes();

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// kernel selection: 0 = auto, 1 = FP32 FFMA kernel, 2 = tcgen05 kernel (one tile per CTA), 3 = tcgen05 pair kernel,
// 4 = replica mode of the pair kernel (32-trajectory tiles), 5 = pair kernel with the small-argument tanh polynomial (for a
// surrogate whose cell values are tiny, e.g. freshly initialised weights: relative instead of absolute tanh accuracy, -7 %).
// Thread-local: fc_mpc_select_kernel affects the calling thread's subsequent fc_mpc_loss* / workspace queries only
// (no process-global state; the library stays re-entrant across threads and devices).
static thread_local int g_mpc_mode = -1;
static int mpc_mode() {
  if (g_mpc_mode < 0) {
    const char* e = getenv("FC_MPC_KERNEL");
    g_mpc_mode = 0;
    if (e && !strcmp(e, "ffma")) g_mpc_mode = 1;
    if (e && !strcmp(e, "tc")) g_mpc_mode = 2;
    if (e && !strcmp(e, "pair")) g_mpc_mode = 3;
    if (e && !strcmp(e, "replica")) g_mpc_mode = 4;
    if (e && !strcmp(e, "pair-precise")) g_mpc_mode = 5;
  }
  return g_mpc_mode;
}

struct MpcPlan {
  int kind;             // 0 = FFMA, 1 = tcgen05 (one tile per CTA), 2 = tcgen05 pair, 3 = replica mode of the pair kernel
  int grid, tiles;
  bool precise;         // pair kernel instantiation with the tanh polynomial
  bool single;          // pair kernel with one tile per CTA (#tiles <= #SMs)
  size_t work_stride;   // floats per CTA
  size_t bytes;
};
// Automatic choice (measured on B200, profiles/r02b_midbatch.jsonl, profiles/r01_bench_configs_three_kernels.jsonl):
//   B <= 32 * #SMs   replica mode of the pair kernel: 32-trajectory tiles, one per SM, the hidden units of a trajectory
//                    split over twelve threads (B = 15: 1.33 ms against 1.97 ms, N = 5 / B = 4096: 0.61 ms against 0.90 ms)
//   .. <= 128 * #SMs the one-tile tcgen05 kernel: every 128-trajectory tile has an SM of its own
//   beyond           the pair kernel, which keeps two tiles per CTA in flight (B = 524288: 89 M against 77 M
//                    trajectory-steps/s; the FFMA kernel reaches 22 M).
// All four stay selectable (fc_mpc_select_kernel / FC_MPC_KERNEL).
static int mpc_plan(int B, int N, int with_grad, MpcPlan* pl, int width_dim = 1) {
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  const int mode = mpc_mode();
  pl->kind = mode == 1 ? 0 : (mode == 3 || mode == 5 ? 2 : (mode == 4 ? 3 : 1));
  pl->precise = mode == 5;
  {
    const int t128 = (B + tc::kTileTC - 1) / tc::kTileTC;
    // more tiles than SMs: two tiles per CTA overlap tensor and cell-update work; a single tile: the two-tile kernels'
    // dedicated issuer warp alone is worth 5 % (B=15: 2.11 ms against 2.23 ms)
    if (mode == 0 && (t128 > sms || t128 == 1)) pl->kind = 2;
    if (mode == 0 && B <= (pr::kTileP / 4) * sms) pl->kind = 3;
  }
  if (width_dim > 1) { pl->kind = 1; pl->precise = false; }   // hidden-layer repeats of the controller live in the one-tile tcgen05 kernel only
  const int tile = pl->kind == 3 ? pr::kTileP / 4 : (pl->kind ? tc::kTileTC : kTile);
  pl->tiles = (B + tile - 1) / tile;
  pl->single = pl->kind == 2 && pl->tiles <= sms;
  const int units = pl->kind == 2 && !pl->single ? (pl->tiles + pr::kTiles - 1) / pr::kTiles : pl->tiles;   // CTA work items
  pl->grid = units < sms ? units : sms;
  pl->work_stride = pl->kind == 2 ? pr::kTiles * pr::work_layout_p(N, with_grad).total
                    : pl->kind == 3 ? pr::work_layout_p(N, with_grad).total
                    : (pl->kind == 1 ? tc::work_layout_tc(N, with_grad, width_dim).total : work_layout(N, with_grad).total);
  pl->bytes = (size_t)pl->grid * kPartialStride * sizeof(double) + (size_t)pl->grid * pl->work_stride * sizeof(float);
  if (width_dim > 1) pl->bytes += (size_t)pl->grid * kWidePartialStride * sizeof(double);
  return FC_OK;
}

template <typename R>
static int closed_loop_launch(const R* x0, const R* ref, int n_ref, int steps_per_ref, int B, int T, R ts, int substeps,
                              const R* scale_in, const R* scale_out, const float* inp_w, const float* inp_b,
                              const float* out_w, R* meas, R* u, R* x_final, void* stream,
                              const float* process_std = nullptr, const float* meas_std = nullptr, unsigned long long seed = 0,
                              const float* int_w = nullptr, const float* int_b = nullptr, int width_dim = 1) {
  if (B <= 0 || T < 0 || n_ref <= 0 || steps_per_ref <= 0 || substeps <= 0 || !(ts > 0))
    return fail(FC_ERR_BAD_SHAPE, "fc_closed_loop_rk4: bad shape%s B=%lld T=%lld", "", B, T);
  if (!x0 || !ref || !scale_in || !scale_out || !inp_w || !inp_b || !out_w)
    return fail(FC_ERR_NULL_POINTER, "fc_closed_loop_rk4: null pointer%s");
  const int threads = 128;
  ClosedLoopNoise nz;
  memset(&nz, 0, sizeof(nz));
  nz.seed = seed;
  for (int i = 0; i < 5; ++i) {                         // HOST arrays of 5 floats (may be NULL = no noise)
    nz.process_std[i] = process_std ? process_std[i] : 0.f;
    nz.meas_std[i] = meas_std ? meas_std[i] : 0.f;
    if (!(nz.process_std[i] >= 0.f) || !(nz.meas_std[i] >= 0.f))
      return fail(FC_ERR_BAD_SHAPE, "fc_closed_loop_rk4: noise standard deviations%s must be >= 0");
    if (nz.process_std[i] > 0.f || nz.meas_std[i] > 0.f) nz.on = 1;
  }
  if (width_dim > 1 && (!int_w || !int_b)) return fail(FC_ERR_NULL_POINTER, "fc_closed_loop_rk4: width_dim > 1 needs fc_int%s");
  if (width_dim > 64) return fail(FC_ERR_UNSUPPORTED, "fc_closed_loop_rk4: width_dim%s=%lld out of range", "", width_dim);
  ClosedLoopWide wd;
  wd.int_w = int_w; wd.int_b = int_b; wd.width_dim = width_dim;
  const dim3 grid((B + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)stream;
#define FC_CL_LAUNCH(W, Z)                                                                                              \
  closed_loop_kernel<R, W, Z><<<grid, threads, 0, st>>>(x0, ref, n_ref, steps_per_ref, B, T, ts, substeps, scale_in, scale_out, \
                                                        inp_w, inp_b, out_w, meas, u, x_final, nz, wd)
  const bool wide = width_dim > 1, noisy = nz.on != 0;
  if (wide && noisy) FC_CL_LAUNCH(true, true);
  else if (wide) FC_CL_LAUNCH(true, false);
  else if (noisy) FC_CL_LAUNCH(false, true);
  else FC_CL_LAUNCH(false, false);
#undef FC_CL_LAUNCH
  FC_CUDA(cudaGetLastError(), "closed_loop_kernel launch");
  return FC_OK;
}

static int ensure_smem_attributes() {
  int dev = 0;
  FC_CUDA(cudaGetDevice(&dev), "cudaGetDevice");
  static bool done[64] = {false};
  if (dev >= 0 && dev < 64 && done[dev]) return FC_OK;
  FC_CUDA(cudaFuncSetAttribute(mpc_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmBytes),
          "cudaFuncSetAttribute(smem)");
  FC_CUDA(cudaFuncSetAttribute(mpc_loss_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmBytesWide),
          "cudaFuncSetAttribute(smem, tc)");
  FC_CUDA(cudaFuncSetAttribute(mpc_loss_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pr::kSmBytesP),
          "cudaFuncSetAttribute(smem, pair)");
  FC_CUDA(cudaFuncSetAttribute(mpc_loss_replica_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pr::kSmBytesP),
          "cudaFuncSetAttribute(smem, replica)");
  FC_CUDA(cudaFuncSetAttribute(lstm_train_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pr::kSmBytesP),
          "cudaFuncSetAttribute(smem, training)");
  FC_CUDA(cudaFuncSetAttribute(lstm_train_replica_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pr::kSmBytesP),
          "cudaFuncSetAttribute(smem, training replica)");
  FC_CUDA(cudaFuncSetAttribute(lt2::dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lt2::kDwSmem),
          "cudaFuncSetAttribute(smem, dw)");
  FC_CUDA(cudaFuncSetAttribute(lt::lstm_window_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lt::kSmemFwd),
          "cudaFuncSetAttribute(smem, lstm fwd)");
  FC_CUDA(cudaFuncSetAttribute(lt::lstm_window_fwd80_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lt::kSmemFwd80),
          "cudaFuncSetAttribute(smem, lstm fwd80)");
  FC_CUDA(cudaFuncSetAttribute(lt::lstm_window_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lt::kSmemBwd),
          "cudaFuncSetAttribute(smem, lstm bwd)");
  if (dev >= 0 && dev < 64) done[dev] = true;
  return FC_OK;
}

}  // namespace fc

using namespace fc;

extern "C" {

const char* fc_last_error(void) { return g_err; }
int fc_version(void) { return 100; }
size_t fc_pack_floats(void) { return (size_t)kPackFloats + (size_t)tc::kPackFloatsTC + (size_t)pr::kPackFloatsP; }

int fc_pack_weights(const float* w_ih0, const float* w_hh0, const float* w_ih1, const float* w_hh1,
                    const float* w_ih2, const float* w_hh2, const float* fc_w, const float* fc_b,
                    const float* fnn_inp_w, const float* fnn_inp_b, const float* fnn_out_w, float* wpack,
                    void* stream) {
  if (!w_ih0 || !w_hh0 || !w_ih1 || !w_hh1 || !w_ih2 || !w_hh2 || !fc_w || !fc_b || !fnn_inp_w || !fnn_inp_b ||
      !fnn_out_w || !wpack)
    return fail(FC_ERR_NULL_POINTER, "fc_pack_weights: null pointer%s");
  if (!aligned16(wpack)) return fail(FC_ERR_MISALIGNED, "fc_pack_weights: wpack must be 16-byte aligned%s");
  RawWeights w;
  w.w_ih[0] = w_ih0; w.w_hh[0] = w_hh0; w.w_ih[1] = w_ih1; w.w_hh[1] = w_hh1; w.w_ih[2] = w_ih2; w.w_hh[2] = w_hh2;
  w.fc_w = fc_w; w.fc_b = fc_b; w.inp_w = fnn_inp_w; w.inp_b = fnn_inp_b; w.out_w = fnn_out_w;
  pack_weights_kernel<<<(kPackFloats + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, wpack);
  FC_CUDA(cudaGetLastError(), "pack_weights_kernel launch");
  pack_weights_tc_kernel<<<(2 * tc::kSmallOff + kSmallFloats + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, wpack);
  FC_CUDA(cudaGetLastError(), "pack_weights_tc_kernel launch");
  pack_weights_pair_kernel<<<(2 * pr::kSmallOff + kSmallFloats + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      w, wpack, (long)kPackFloats + tc::kPackFloatsTC);
  FC_CUDA(cudaGetLastError(), "pack_weights_pair_kernel launch");
  return FC_OK;
}

int fc_mpc_select_kernel(int mode) {
  if (mode < 0 || mode > 5)
    return fail(FC_ERR_BAD_SHAPE, "fc_mpc_select_kernel: mode%s must be 0 (auto), 1 (ffma), 2 (tcgen05), 3 (tcgen05 pair), 4 (replica) or 5 (pair, precise tanh)");
  g_mpc_mode = mode;
  return FC_OK;
}

size_t fc_mpc_loss_workspace_bytes(int B, int N, int with_grad) {
  if (B <= 0 || N <= 0) return 0;
  MpcPlan pl;
  if (mpc_plan(B, N, with_grad, &pl)) return 0;
  return pl.bytes;
}

// Scratch traffic of one with_grad launch, computed from the workspace layout of the kernel the plan selects: saved
// activation records (written by the forward, read once by the reverse sweep) + hidden-sequence / d-sequence scratch
// handed between layers (written and read once each).  An upper bound of the DRAM traffic (the L2 absorbs part of the
// sequence scratch); the measured figure is the ncu capture in profiles/.
size_t fc_mpc_loss_scratch_traffic_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  MpcPlan pl;
  if (mpc_plan(B, N, 1, &pl)) return 0;
  long kept = 0;                                    // kept (window, step) pairs per layer
  for (int m = 0; m < N; ++m) kept += steps_kept(m);
  size_t per_tile = 0;
  if (pl.kind == 2) {
    per_tile = (size_t)rec_base(N) * pr::kRecFloatsP * 2                                  // records: write + read
               + (size_t)N * (kLayers - 1) * kLook * pr::kSlot * 2                        // operand-format hidden sequence
               + (size_t)kept * (kLayers - 1) * (pr::kUpdWarps * pr::kMaxOwn * 32) * 2;    // d-sequence
  } else if (pl.kind == 3) {
    // 32-trajectory tiles: eleven threads of a trajectory write 5 record float4 per step, the last one 8; one hidden-sequence
    // float4 per thread (+2); d-sequence 4 (6) floats
    per_tile = (size_t)rec_base(N) * (11 * 5 + 8) * 32 * 4 * 2
               + (size_t)N * (kLayers - 1) * kLook * (12 + 2) * 32 * 4 * 2
               + (size_t)kept * (kLayers - 1) * (11 * 4 + 6) * 32 * 2;
  } else if (pl.kind == 1) {
    per_tile = (size_t)rec_base(N) * tc::kRecFloatsTC * 2;
  } else {
    per_tile = (size_t)rec_base(N) * kRecFloats * 2;
  }
  return per_tile * sizeof(float) * (size_t)pl.tiles;
}

int fc_mpc_loss(const float* X, const float* u0, const float* Z, const float* wpack, int B, int N, float alpha,
                long long B_global, int with_grad, float* cost, float* command, float* error, float* pred, float* du0,
                float* gl, void* workspace, size_t workspace_bytes, void* stream) {
  return fc_mpc_loss_noise(X, u0, Z, wpack, B, N, alpha, B_global, with_grad, cost, command, error, pred, du0, gl, workspace,
                           workspace_bytes, 0.0f, 0ull, stream);
}

static int mpc_loss_impl(const float* X, const float* u0, const float* Z, const float* wpack, int B, int N, float alpha,
                         long long B_global, int with_grad, float* cost, float* command, float* error, float* pred, float* du0,
                         float* gl, void* workspace, size_t workspace_bytes, float noise_std, unsigned long long noise_seed,
                         const float* int_w, const float* int_b, int width_dim, float* gl_wide, void* stream);

int fc_mpc_loss_noise(const float* X, const float* u0, const float* Z, const float* wpack, int B, int N, float alpha,
                      long long B_global, int with_grad, float* cost, float* command, float* error, float* pred, float* du0,
                      float* gl, void* workspace, size_t workspace_bytes, float noise_std, unsigned long long noise_seed,
                      void* stream) {
  return mpc_loss_impl(X, u0, Z, wpack, B, N, alpha, B_global, with_grad, cost, command, error, pred, du0, gl, workspace,
                       workspace_bytes, noise_std, noise_seed, nullptr, nullptr, 1, nullptr, stream);
}

size_t fc_mpc_loss_wide_workspace_bytes(int B, int N, int with_grad, int width_dim) {
  if (B <= 0 || N <= 0 || width_dim < 1) return 0;
  MpcPlan pl;
  if (mpc_plan(B, N, with_grad, &pl, width_dim)) return 0;
  return pl.bytes;
}

int fc_mpc_loss_wide(const float* X, const float* u0, const float* Z, const float* wpack, const float* fnn_int_w,
                     const float* fnn_int_b, int width_dim, int B, int N, float alpha, long long B_global, int with_grad,
                     float* cost, float* command, float* error, float* pred, float* du0, float* gl, float* gl_wide,
                     void* workspace, size_t workspace_bytes, float noise_std, unsigned long long noise_seed, void* stream) {
  if (width_dim < 1 || width_dim > 64) return fail(FC_ERR_UNSUPPORTED, "fc_mpc_loss_wide: width_dim%s=%lld out of range", "", width_dim);
  if (width_dim > 1 && (!fnn_int_w || !fnn_int_b || (with_grad && !gl_wide)))
    return fail(FC_ERR_NULL_POINTER, "fc_mpc_loss_wide: null pointer%s");
  return mpc_loss_impl(X, u0, Z, wpack, B, N, alpha, B_global, with_grad, cost, command, error, pred, du0, gl, workspace,
                       workspace_bytes, noise_std, noise_seed, fnn_int_w, fnn_int_b, width_dim, gl_wide, stream);
}

static int mpc_loss_impl(const float* X, const float* u0, const float* Z, const float* wpack, int B, int N, float alpha,
                         long long B_global, int with_grad, float* cost, float* command, float* error, float* pred, float* du0,
                         float* gl, void* workspace, size_t workspace_bytes, float noise_std, unsigned long long noise_seed,
                         const float* int_w, const float* int_b, int width_dim, float* gl_wide, void* stream) {
  if (!(noise_std >= 0.f)) return fail(FC_ERR_BAD_SHAPE, "fc_mpc_loss: noise_std%s must be >= 0");
  if (B <= 0 || N <= 0 || B_global < B) return fail(FC_ERR_BAD_SHAPE, "fc_mpc_loss: bad shape%s B=%lld N=%lld", "", B, N);
  if (N > 4096) return fail(FC_ERR_UNSUPPORTED, "fc_mpc_loss: horizon%s N=%lld too long", "", N);
  if (!X || !u0 || !Z || !wpack || !cost || !command || !error || !pred || !gl || !workspace || (with_grad && !du0))
    return fail(FC_ERR_NULL_POINTER, "fc_mpc_loss: null pointer%s");
  if (!aligned16(wpack) || !aligned16(workspace))
    return fail(FC_ERR_MISALIGNED, "fc_mpc_loss: wpack/workspace must be 16-byte aligned%s");
  MpcPlan pl;
  int rc = mpc_plan(B, N, with_grad, &pl, width_dim);
  if (rc) return rc;
  if (workspace_bytes < pl.bytes)
    return fail(FC_ERR_WORKSPACE, "fc_mpc_loss: workspace too small%s: have %lld need %lld bytes", "", (long long)workspace_bytes,
                (long long)pl.bytes);
  rc = ensure_smem_attributes();
  if (rc) return rc;
  MpcParams p;
  memset(&p, 0, sizeof(p));
  p.X = X; p.u0 = u0; p.Z = Z;
  p.wpack = pl.kind >= 2 ? wpack + kPackFloats + tc::kPackFloatsTC : (pl.kind == 1 ? wpack + kPackFloats : wpack);
  p.cost = cost; p.command = command; p.error = error; p.pred = pred; p.du0 = du0;
  p.partial = reinterpret_cast<double*>(workspace);
  p.width_dim = width_dim; p.int_w = int_w; p.int_b = int_b;
  double* after_partial = p.partial + (size_t)pl.grid * kPartialStride;
  if (width_dim > 1) { p.partial_wide = after_partial; after_partial += (size_t)pl.grid * kWidePartialStride; }
  p.work = reinterpret_cast<float*>(after_partial);
  p.work_stride = pl.work_stride;
  p.B = B; p.N = N; p.with_grad = with_grad ? 1 : 0; p.num_tiles = pl.tiles;
  p.single_tile = pl.single ? 1 : 0;
  p.alpha = alpha;
  p.grad_scale = (float)(1.0 / ((double)N * (double)B_global));
  // 1.0 = the law measured on iid data (scripts/micro/umma_test.cu); real LSTM partial sums are more coherent and
  // lose ~1.3x more (scripts/diag_trace.py scan: summed-gradient error minimal for 1.2..1.5)
  {  // power-of-two scale of the gate gradients for the fp16 split: seeds are O(1..100) / (N * B_global)
    int e = (int)floor(log2((double)N * (double)B_global));
    p.g_scale = (float)ldexp(1.0, e);
    p.g_unscale = (float)ldexp(1.0, -e);
  }
  p.acc_comp = 1.3f;
  if (const char* e = getenv("FC_TC_ACC_COMP")) p.acc_comp = (float)atof(e);   // calibration experiments only
  p.debug_timing = getenv("FC_TC_TIMING") ? 1 : 0;
  p.noise_std = noise_std; p.noise_seed = noise_seed;
  cudaStream_t st = (cudaStream_t)stream;
  if (pl.kind == 3) mpc_loss_replica_kernel<<<pl.grid, pr::kThreadsP, pr::kSmBytesP, st>>>(p);
  else if (pl.kind == 2 && pl.precise) lstm_train_pair_kernel<<<pl.grid, pr::kThreadsP, pr::kSmBytesP, st>>>(p);
  else if (pl.kind == 2) mpc_loss_pair_kernel<<<pl.grid, pr::kThreadsP, pr::kSmBytesP, st>>>(p);
  else if (pl.kind == 1) mpc_loss_tc_kernel<<<pl.grid, tc::kThreadsTC, width_dim > 1 ? tc::kSmBytesWide : tc::kSmBytesTC, st>>>(p);
  else mpc_loss_kernel<<<pl.grid, kThreads, kSmBytes, st>>>(p);
  FC_CUDA(cudaGetLastError(), "mpc_loss kernel launch");
  mpc_finalize_kernel<<<1, 1024, 0, st>>>(p.partial, pl.grid, 1.0 / (double)B_global, gl);
  FC_CUDA(cudaGetLastError(), "mpc_finalize_kernel launch");
  if (width_dim > 1 && with_grad) {
    sum_partials_kernel<<<8, 1024, 0, st>>>(p.partial_wide, pl.grid, kWidePartialStride, kWideGrads, gl_wide);
    FC_CUDA(cudaGetLastError(), "sum_partials_kernel launch");
  }
  return FC_OK;
}

// shadow roll-out: 32-trajectory tiles (replica mode) while every tile can have an SM of its own, else tile pairs
static void shadow_plan(int B, int sms, bool* replica, int* tiles, int* grid) {
  const int mode = mpc_mode();
  *replica = mode == 4 || (mode != 3 && B <= (pr::kTileP / 4) * sms);
  if (*replica) {
    *tiles = (B + pr::kTileP / 4 - 1) / (pr::kTileP / 4);
    *grid = *tiles < sms ? *tiles : sms;
  } else {
    *tiles = (B + pr::kTileP - 1) / pr::kTileP;
    const int pairs = (*tiles + pr::kTiles - 1) / pr::kTiles;
    *grid = pairs < sms ? pairs : sms;
  }
}

size_t fc_lstm_shadow_workspace_bytes(int B, int T) {
  if (B <= 0 || T <= 0) return 0;
  int sms = 0;
  if (sm_count(&sms)) return 0;
  bool replica; int tiles, grid;
  shadow_plan(B, sms, &replica, &tiles, &grid);
  return (size_t)grid * kPartialStride * sizeof(double) + (size_t)grid * pr::kTiles * pr::work_layout_p(T, 0).total * sizeof(float);
}

int fc_lstm_shadow_rollout(const float* row0, const float* u, const float* ratio, const float* wpack, int B, int T, float* y,
                           void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0 || T <= 0) return fail(FC_ERR_BAD_SHAPE, "fc_lstm_shadow_rollout: bad shape%s B=%lld T=%lld", "", B, T);
  if (T > 65536) return fail(FC_ERR_UNSUPPORTED, "fc_lstm_shadow_rollout: roll-out%s T=%lld too long", "", T);
  if (!row0 || !u || !ratio || !wpack || !y || !workspace) return fail(FC_ERR_NULL_POINTER, "fc_lstm_shadow_rollout: null pointer%s");
  if (!aligned16(wpack) || !aligned16(workspace))
    return fail(FC_ERR_MISALIGNED, "fc_lstm_shadow_rollout: wpack/workspace must be 16-byte aligned%s");
  const size_t need = fc_lstm_shadow_workspace_bytes(B, T);
  if (need == 0) return fail(FC_ERR_CUDA, "fc_lstm_shadow_rollout: no CUDA device%s");
  if (workspace_bytes < need)
    return fail(FC_ERR_WORKSPACE, "fc_lstm_shadow_rollout: workspace too small%s: have %lld need %lld bytes", "", (long long)workspace_bytes,
                (long long)need);
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  rc = ensure_smem_attributes();
  if (rc) return rc;
  bool replica; int tiles, grid;
  shadow_plan(B, sms, &replica, &tiles, &grid);
  MpcParams p;
  memset(&p, 0, sizeof(p));
  p.wpack = wpack + kPackFloats + tc::kPackFloatsTC;
  p.partial = reinterpret_cast<double*>(workspace);
  p.work = reinterpret_cast<float*>(p.partial + (size_t)grid * kPartialStride);
  p.work_stride = pr::kTiles * pr::work_layout_p(T, 0).total;
  p.B = B; p.N = T; p.with_grad = 0; p.num_tiles = tiles;
  p.acc_comp = 1.3f;
  p.g_scale = p.g_unscale = 1.0f;
  p.shadow = 1; p.sh_row0 = row0; p.sh_u = u; p.sh_y = y;
  // ratio is a HOST array of 4 floats
  for (int q = 0; q < 4; ++q) p.sh_ratio[q] = ratio[q];
  if (replica) mpc_loss_replica_kernel<<<grid, pr::kThreadsP, pr::kSmBytesP, (cudaStream_t)stream>>>(p);
  else mpc_loss_pair_kernel<<<grid, pr::kThreadsP, pr::kSmBytesP, (cudaStream_t)stream>>>(p);
  FC_CUDA(cudaGetLastError(), "mpc_loss_pair_kernel (shadow) launch");
  return FC_OK;
}

int fc_closed_loop_rk4(const float* x0, const float* ref, int n_ref, int steps_per_ref, int B, int T, float ts,
                       int substeps, const float* scale_in, const float* scale_out, const float* fnn_inp_w,
                       const float* fnn_inp_b, const float* fnn_out_w, float* meas, float* u, float* x_final,
                       void* stream) {
  return closed_loop_launch<float>(x0, ref, n_ref, steps_per_ref, B, T, ts, substeps, scale_in, scale_out, fnn_inp_w,
                                   fnn_inp_b, fnn_out_w, meas, u, x_final, stream);
}

int fc_closed_loop_rk4_f64(const double* x0, const double* ref, int n_ref, int steps_per_ref, int B, int T, double ts,
                           int substeps, const double* scale_in, const double* scale_out, const float* fnn_inp_w,
                           const float* fnn_inp_b, const float* fnn_out_w, double* meas, double* u, double* x_final,
                           void* stream) {
  return closed_loop_launch<double>(x0, ref, n_ref, steps_per_ref, B, T, ts, substeps, scale_in, scale_out, fnn_inp_w,
                                    fnn_inp_b, fnn_out_w, meas, u, x_final, stream);
}

int fc_closed_loop_rk4_noise(const float* x0, const float* ref, int n_ref, int steps_per_ref, int B, int T, float ts,
                             int substeps, const float* scale_in, const float* scale_out, const float* fnn_inp_w,
                             const float* fnn_inp_b, const float* fnn_out_w, float* meas, float* u, float* x_final,
                             const float* process_std, const float* meas_std, unsigned long long seed, void* stream) {
  return closed_loop_launch<float>(x0, ref, n_ref, steps_per_ref, B, T, ts, substeps, scale_in, scale_out, fnn_inp_w,
                                   fnn_inp_b, fnn_out_w, meas, u, x_final, stream, process_std, meas_std, seed);
}

int fc_closed_loop_rk4_f64_noise(const double* x0, const double* ref, int n_ref, int steps_per_ref, int B, int T, double ts,
                                 int substeps, const double* scale_in, const double* scale_out, const float* fnn_inp_w,
                                 const float* fnn_inp_b, const float* fnn_out_w, double* meas, double* u, double* x_final,
                                 const float* process_std, const float* meas_std, unsigned long long seed, void* stream) {
  return closed_loop_launch<double>(x0, ref, n_ref, steps_per_ref, B, T, ts, substeps, scale_in, scale_out, fnn_inp_w,
                                    fnn_inp_b, fnn_out_w, meas, u, x_final, stream, process_std, meas_std, seed);
}

int fc_build_windows(const float* Xtab, const float* ytab, const float* Ztab, long long M, int t_traj, int lookback,
                     const long long* idx, long long B, float* X, float* y, float* Z, void* stream) {
  if (M <= 0 || t_traj <= 0 || M % t_traj != 0 || lookback <= 0 || B <= 0)
    return fail(FC_ERR_BAD_SHAPE, "fc_build_windows: bad shape%s M=%lld B=%lld", "", M, B);
  if (!Xtab || !ytab || !Ztab || !idx || !X || !y || !Z) return fail(FC_ERR_NULL_POINTER, "fc_build_windows: null pointer%s");
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  const long long total = B * (4 + (long long)lookback * 5);
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;      // grid-stride, a multiple of the SM count
  cudaStream_t st = (cudaStream_t)stream;
  const bool small = M < (1ll << 31) && total < (1ll << 32) - (1ll << 24);
  if (small && lookback == 10)
    build_windows_kernel<unsigned, 10><<<(unsigned)blocks, 256, 0, st>>>(Xtab, ytab, Ztab, (unsigned)t_traj, lookback, idx, (unsigned)B, X, y, Z);
  else if (small)
    build_windows_kernel<unsigned, 0><<<(unsigned)blocks, 256, 0, st>>>(Xtab, ytab, Ztab, (unsigned)t_traj, lookback, idx, (unsigned)B, X, y, Z);
  else
    build_windows_kernel<unsigned long long, 0><<<(unsigned)blocks, 256, 0, st>>>(Xtab, ytab, Ztab, (unsigned long long)t_traj, lookback, idx,
                                                                                 (unsigned long long)B, X, y, Z);
  FC_CUDA(cudaGetLastError(), "build_windows_kernel launch");
  return FC_OK;
}

int fc_closed_loop_rk4_ex(int f64, const void* x0, const void* ref, int n_ref, int steps_per_ref, int B, int T, double ts,
                          int substeps, const void* scale_in, const void* scale_out, const float* fnn_inp_w,
                          const float* fnn_inp_b, const float* fnn_out_w, const float* fnn_int_w, const float* fnn_int_b,
                          int width_dim, void* meas, void* u, void* x_final, const float* process_std, const float* meas_std,
                          unsigned long long seed, void* stream) {
  if (f64)
    return closed_loop_launch<double>((const double*)x0, (const double*)ref, n_ref, steps_per_ref, B, T, ts, substeps,
                                      (const double*)scale_in, (const double*)scale_out, fnn_inp_w, fnn_inp_b, fnn_out_w,
                                      (double*)meas, (double*)u, (double*)x_final, stream, process_std, meas_std, seed, fnn_int_w,
                                      fnn_int_b, width_dim);
  return closed_loop_launch<float>((const float*)x0, (const float*)ref, n_ref, steps_per_ref, B, T, (float)ts, substeps,
                                   (const float*)scale_in, (const float*)scale_out, fnn_inp_w, fnn_inp_b, fnn_out_w,
                                   (float*)meas, (float*)u, (float*)x_final, stream, process_std, meas_std, seed, fnn_int_w,
                                   fnn_int_b, width_dim);
}

int fc_fp32_peak(int iters, double* flops_host, void* stream) {
  if (!flops_host || iters <= 0) return fail(FC_ERR_NULL_POINTER, "fc_fp32_peak: bad arguments%s");
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  float* sink = nullptr;
  FC_CUDA(cudaMalloc(&sink, sizeof(float)), "cudaMalloc");
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  FC_CUDA(cudaEventCreate(&e0), "cudaEventCreate");
  FC_CUDA(cudaEventCreate(&e1), "cudaEventCreate");
  const int blocks = sms * 8;
  ffma_peak_kernel<<<blocks, 256, 0, st>>>(iters / 8 + 1, sink);   // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    FC_CUDA(cudaEventRecord(e0, st), "cudaEventRecord");
    ffma_peak_kernel<<<blocks, 256, 0, st>>>(iters, sink);
    FC_CUDA(cudaEventRecord(e1, st), "cudaEventRecord");
    FC_CUDA(cudaEventSynchronize(e1), "cudaEventSynchronize");
    float ms = 0.f;
    FC_CUDA(cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime");
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  const double flop = 2.0 * 8.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
  *flops_host = flop / (best * 1e-3);
  return FC_OK;
}

// ---------------------------------------------------------------------------------------------------
// surrogate training path + optimizer update (fc_lstm_train.cuh)
// ---------------------------------------------------------------------------------------------------
// pack = [FFMA weight images (lt::kPackFloatsL) | pad to 16 bytes | pair-kernel operand images (pr::kPackFloatsP)]
static constexpr long kTrainPackPairOff = ((long)lt::kPackFloatsL + 3) / 4 * 4;
size_t fc_lstm_train_pack_floats(void) { return (size_t)kTrainPackPairOff + (size_t)pr::kPackFloatsP; }

int fc_lstm_train_pack(const float* w_ih0, const float* w_hh0, const float* w_ih1, const float* w_hh1, const float* w_ih2,
                       const float* w_hh2, float* pack, void* stream) {
  if (!w_ih0 || !w_hh0 || !w_ih1 || !w_hh1 || !w_ih2 || !w_hh2 || !pack)
    return fail(FC_ERR_NULL_POINTER, "fc_lstm_train_pack: null pointer%s");
  if (!aligned16(pack)) return fail(FC_ERR_MISALIGNED, "fc_lstm_train_pack: pack must be 16-byte aligned%s");
  lt::LstmRaw w;
  w.w_ih[0] = w_ih0; w.w_hh[0] = w_hh0; w.w_ih[1] = w_ih1; w.w_hh[1] = w_hh1; w.w_ih[2] = w_ih2; w.w_hh[2] = w_hh2;
  lt::pack_lstm_train_kernel<<<(lt::kPackFloatsL + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, pack);
  FC_CUDA(cudaGetLastError(), "pack_lstm_train_kernel launch");
  // operand images of the tensor-core path; its fc block is not used (the training mode reads fc.weight / fc.bias directly)
  RawWeights rw;
  for (int l = 0; l < kLayers; ++l) { rw.w_ih[l] = w.w_ih[l]; rw.w_hh[l] = w.w_hh[l]; }
  rw.fc_w = rw.fc_b = rw.inp_w = rw.inp_b = rw.out_w = w_hh0;
  pack_weights_pair_kernel<<<(2 * pr::kSmallOff + kSmallFloats + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rw, pack, kTrainPackPairOff);
  FC_CUDA(cudaGetLastError(), "pack_weights_pair_kernel (training) launch");
  return FC_OK;
}

// ---- tensor-core path of the surrogate training step (fc_lstm_train_tc.cuh): pair kernel in training mode (its replica
// mode with 32-sample tiles for B <= 32 x #SMs) + dw_kernel.  The automatic choice: measured faster than the FFMA kernels at
// every batch size (whole training step B = 256: 0.48 ms against 0.63 ms, 4 096: 0.49 / 0.65, 8 192: 1.08 / 1.23, 65 536:
// 2.7 / 7.1); FC_LSTM_TRAIN=tc|ffma or fc_lstm_train_select_path force one of them.
static thread_local int g_lstm_mode = -1;
static int lstm_tc_mode() {
  if (g_lstm_mode < 0) {
    const char* e = getenv("FC_LSTM_TRAIN");
    g_lstm_mode = 0;
    if (e && !strcmp(e, "ffma")) g_lstm_mode = 1;
    if (e && !strcmp(e, "tc")) g_lstm_mode = 2;
  }
  return g_lstm_mode;
}
static bool lstm_use_tc(int B) { (void)B; return lstm_tc_mode() != 1; }

static constexpr int kFcGradBlocksPerSm = 8;
struct LstmTcPlan {
  int replica;                                  // 32-sample tiles, one per CTA (B <= 32 x #SMs), else pairs of 128-sample tiles
  int tiles, grid, chunk_tiles, dgrid;          // tiles, CTAs of the training kernel, tiles per chunk of the backward, CTAs of dw_kernel
  size_t scale, partial, work, hlast, dwp, fcp, scratch, floats;   // offsets in floats
};
static int lstm_tc_plan(int B, int save, LstmTcPlan* pl) {
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  pl->replica = B <= (pr::kTileP / 4) * sms;
  const int rows = pl->replica ? pr::kTileP / 4 : pr::kTileP;
  pl->tiles = (B + rows - 1) / rows;
  const int units = pl->replica ? pl->tiles : (pl->tiles + pr::kTiles - 1) / pr::kTiles;
  pl->grid = units < sms ? units : sms;
  const int pass_tiles = (pl->replica ? 1 : pr::kTiles) * sms;
  pl->chunk_tiles = pl->tiles < pass_tiles ? pl->tiles : pass_tiles;               // one pass of the training kernel per chunk
  const size_t work_cta = pr::kTiles * (save ? pr::work_total_train() : pr::work_layout_p(1, 0).total);
  size_t o = 0;
  pl->scale = o;   o += 4;
  pl->partial = o; o += (size_t)pl->grid * kPartialStride * 2;                      // doubles
  pl->work = o;    o += (size_t)pl->grid * work_cta;
  pl->hlast = o;   o += save ? ((size_t)B * kHid + 3) / 4 * 4 : 0;
  pl->dgrid = pl->chunk_tiles < sms ? pl->chunk_tiles : sms;                        // CTAs of the weight-gradient kernel
  pl->dwp = o;     o += save ? (size_t)pl->dgrid * lt2::kDwPartialFloats : 0;
  pl->fcp = o;     o += save ? (size_t)kFcGradBlocksPerSm * sms * 256 * 2 : 0;       // doubles
  pl->scratch = o; o += save ? (size_t)pl->chunk_tiles * (pr::kTrTileFloats / (pl->replica ? 4 : 1)) : 0;
  pl->floats = o + 4;
  return FC_OK;
}

struct LstmTrainPlan {
  int tiles, pairs, grid, grid_f, wide;        // 40-sample tiles, pairs of them, CTAs (reverse / forward), 80-sample forward
  size_t rec, hseq, dseq, partial, floats;     // offsets in floats
};
static int lstm_train_plan(int B, int save, LstmTrainPlan* pl) {
  if (B <= 0) return fail(FC_ERR_BAD_SHAPE, "fc_lstm_window: B must be positive%s (B=%lld)", "", B);
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  pl->tiles = (B + lt::kTT - 1) / lt::kTT;
  pl->pairs = (pl->tiles + 1) / 2;
  pl->grid = pl->tiles < sms ? pl->tiles : sms;
  pl->wide = pl->tiles > sms;                  // more tiles than SMs: throughput matters, use the 80-sample forward
  pl->grid_f = pl->wide ? (pl->pairs < sms ? pl->pairs : sms) : pl->grid;
  size_t o = 0;
  pl->rec = o;  o += save ? (size_t)pl->pairs * lt::kRecFloatsPair : 0;
  pl->hseq = o; o += save ? (size_t)pl->pairs * lt::kHseqFloatsPair : (pl->wide ? (size_t)pl->grid_f * lt::kHseqFloatsPair : 0);
  pl->dseq = o; o += save ? (size_t)pl->grid * lt::kDseqFloatsCta : 0;
  pl->partial = o; o += save ? (size_t)pl->grid * lt::kPartialFloats : 0;
  pl->floats = o + 4;
  return FC_OK;
}

int fc_lstm_train_select_path(int mode) {
  if (mode < 0 || mode > 2) return fail(FC_ERR_BAD_SHAPE, "fc_lstm_train_select_path: mode%s must be 0 (auto), 1 (FP32 FFMA kernels) or 2 (tensor-core path)");
  g_lstm_mode = mode;
  return FC_OK;
}

int fc_lstm_train_path_for(int B) { return lstm_use_tc(B) ? 2 : 1; }

size_t fc_lstm_window_workspace_bytes(int B, int save) {
  if (B > 0 && lstm_use_tc(B)) {
    LstmTcPlan tp;
    if (lstm_tc_plan(B, save, &tp)) return 0;
    return tp.floats * sizeof(float);
  }
  LstmTrainPlan pl;
  if (lstm_train_plan(B, save, &pl)) return 0;
  return pl.floats * sizeof(float);
}

int fc_lstm_window_fwd(const float* X, const float* pack, const float* fc_w, const float* fc_b, int B, int save, float* out,
                       void* work, size_t work_bytes, void* stream) {
  if (!X || !pack || !fc_w || !fc_b || !out) return fail(FC_ERR_NULL_POINTER, "fc_lstm_window_fwd: null pointer%s");
  if (B > 0 && lstm_use_tc(B)) {
    LstmTcPlan tp;
    int rc = lstm_tc_plan(B, save, &tp);
    if (rc) return rc;
    if (!work) return fail(FC_ERR_NULL_POINTER, "fc_lstm_window_fwd: workspace required%s");
    if (work_bytes < tp.floats * sizeof(float))
      return fail(FC_ERR_WORKSPACE, "fc_lstm_window_fwd: workspace too small%s (%lld < %lld bytes)", "", (long long)work_bytes,
                  (long long)(tp.floats * sizeof(float)));
    if (!aligned16(pack) || !aligned16(work)) return fail(FC_ERR_MISALIGNED, "fc_lstm_window_fwd: pack / work must be 16-byte aligned%s");
    rc = ensure_smem_attributes();
    if (rc) return rc;
    float* wf = (float*)work;
    MpcParams p;
    memset(&p, 0, sizeof(p));
    p.wpack = pack + kTrainPackPairOff;
    p.partial = reinterpret_cast<double*>(wf + tp.partial);
    p.work = wf + tp.work;
    p.work_stride = pr::kTiles * (save ? pr::work_total_train() : pr::work_layout_p(1, 0).total);
    p.B = B; p.N = 1; p.with_grad = 0; p.num_tiles = tp.tiles;
    p.acc_comp = 1.3f; p.g_scale = p.g_unscale = 1.0f;
    p.train = 1; p.tr_x = X; p.tr_y = out; p.tr_hlast = save ? wf + tp.hlast : nullptr; p.tr_fcw = fc_w; p.tr_fcb = fc_b;
    if (tp.replica) lstm_train_replica_kernel<<<tp.grid, pr::kThreadsP, pr::kSmBytesP, (cudaStream_t)stream>>>(p);
    else lstm_train_pair_kernel<<<tp.grid, pr::kThreadsP, pr::kSmBytesP, (cudaStream_t)stream>>>(p);
    FC_CUDA(cudaGetLastError(), "lstm_train_pair_kernel (forward) launch");
    return FC_OK;
  }
  LstmTrainPlan pl;
  int rc = lstm_train_plan(B, save, &pl);
  if (rc) return rc;
  const bool need_work = save || pl.wide;
  if (need_work && !work) return fail(FC_ERR_NULL_POINTER, "fc_lstm_window_fwd: workspace required%s");
  if (need_work && work_bytes < pl.floats * sizeof(float))
    return fail(FC_ERR_WORKSPACE, "fc_lstm_window_fwd: workspace too small%s (%lld < %lld bytes)", "", (long long)work_bytes,
                (long long)(pl.floats * sizeof(float)));
  if (!aligned16(pack) || (need_work && !aligned16(work))) return fail(FC_ERR_MISALIGNED, "fc_lstm_window_fwd: pack / work must be 16-byte aligned%s");
  rc = ensure_smem_attributes();
  if (rc) return rc;
  lt::LstmFwdParams p;
  p.X = X; p.pack = pack; p.fc_w = fc_w; p.fc_b = fc_b; p.out = out; p.B = B; p.save = save;
  p.rec = save ? (float*)work + pl.rec : nullptr;
  p.hseq = need_work ? (float*)work + pl.hseq : nullptr;
  if (pl.wide)
    lt::lstm_window_fwd80_kernel<<<pl.grid_f, lt::kThreadsL, lt::kSmemFwd80, (cudaStream_t)stream>>>(p);
  else
    lt::lstm_window_fwd_kernel<<<pl.grid, lt::kThreadsL, lt::kSmemFwd, (cudaStream_t)stream>>>(p);
  FC_CUDA(cudaGetLastError(), "lstm_window_fwd_kernel launch");
  return FC_OK;
}

int fc_lstm_window_bwd(const float* X, const float* d_out, const float* pack, const float* fc_w, int B, void* work,
                       size_t work_bytes, float* g_ih0, float* g_hh0, float* g_ih1, float* g_hh1, float* g_ih2, float* g_hh2,
                       float* g_fc_w, float* g_fc_b, void* stream) {
  if (!X || !d_out || !pack || !fc_w || !work || !g_ih0 || !g_hh0 || !g_ih1 || !g_hh1 || !g_ih2 || !g_hh2 || !g_fc_w || !g_fc_b)
    return fail(FC_ERR_NULL_POINTER, "fc_lstm_window_bwd: null pointer%s");
  if (B > 0 && lstm_use_tc(B)) {
    LstmTcPlan tp;
    int rc = lstm_tc_plan(B, 1, &tp);
    if (rc) return rc;
    if (work_bytes < tp.floats * sizeof(float))
      return fail(FC_ERR_WORKSPACE, "fc_lstm_window_bwd: workspace too small%s (%lld < %lld bytes)", "", (long long)work_bytes,
                  (long long)(tp.floats * sizeof(float)));
    if (!aligned16(pack) || !aligned16(work)) return fail(FC_ERR_MISALIGNED, "fc_lstm_window_bwd: pack / work must be 16-byte aligned%s");
    rc = ensure_smem_attributes();
    if (rc) return rc;
    int sms = 0;
    rc = sm_count(&sms);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    float* wf = (float*)work;
    FC_CUDA(cudaMemsetAsync(wf + tp.scale, 0, 4 * sizeof(float), st), "cudaMemsetAsync(scale)");
    lt2::grad_absmax_kernel<<<sms, 256, 0, st>>>(d_out, (long long)B * kOut, wf + tp.scale);
    FC_CUDA(cudaGetLastError(), "grad_absmax_kernel launch");
    lt2::grad_scale_kernel<<<1, 1, 0, st>>>(wf + tp.scale);
    FC_CUDA(cudaGetLastError(), "grad_scale_kernel launch");
    FC_CUDA(cudaMemsetAsync(wf + tp.dwp, 0, (size_t)tp.dgrid * lt2::kDwPartialFloats * sizeof(float), st), "cudaMemsetAsync(dW partials)");
    for (int t0 = 0; t0 < tp.tiles; t0 += tp.chunk_tiles) {
      const int nt = tp.tiles - t0 < tp.chunk_tiles ? tp.tiles - t0 : tp.chunk_tiles;
      const int rows = tp.replica ? pr::kTileP / 4 : pr::kTileP;
      const long long s0 = (long long)t0 * rows;
      const int nb = (int)((long long)B - s0 < (long long)nt * rows ? (long long)B - s0 : (long long)nt * rows);
      const int pairs = tp.replica ? nt : (nt + pr::kTiles - 1) / pr::kTiles;
      MpcParams p;
      memset(&p, 0, sizeof(p));
      p.wpack = pack + kTrainPackPairOff;
      p.partial = reinterpret_cast<double*>(wf + tp.partial);
      p.work = wf + tp.work;
      p.work_stride = pr::kTiles * pr::work_total_train();
      p.B = nb; p.N = 1; p.with_grad = 1; p.num_tiles = nt;
      p.acc_comp = 1.3f; p.g_scale = p.g_unscale = 1.0f;
      p.train = 2; p.tr_x = X + s0 * (kLook * kFeat); p.tr_dy = d_out + s0 * kOut; p.tr_fcw = fc_w; p.tr_fcb = fc_w;   // fc.bias is not needed
      p.tr_scale = wf + tp.scale; p.tr_ws = wf + tp.scratch; p.tr_tile_base = 0;
      if (tp.replica) lstm_train_replica_kernel<<<pairs < tp.grid ? pairs : tp.grid, pr::kThreadsP, pr::kSmBytesP, st>>>(p);
      else lstm_train_pair_kernel<<<pairs < tp.grid ? pairs : tp.grid, pr::kThreadsP, pr::kSmBytesP, st>>>(p);
      FC_CUDA(cudaGetLastError(), "lstm_train_pair_kernel (forward + reverse sweep) launch");
      lt2::DwParams d;
      d.ws = wf + tp.scratch; d.tiles = nt; d.spt = tp.replica ? 1 : lt2::kStagesPerTile; d.partial = wf + tp.dwp; d.acc_comp = 1.0f;
      lt2::dw_kernel<<<tp.dgrid, lt2::kDwThreads, lt2::kDwSmem, st>>>(d);
      FC_CUDA(cudaGetLastError(), "dw_kernel launch");
    }
    lt2::DwGradOut g;
    g.g_ih[0] = g_ih0; g.g_hh[0] = g_hh0; g.g_ih[1] = g_ih1; g.g_hh[1] = g_hh1; g.g_ih[2] = g_ih2; g.g_hh[2] = g_hh2;
    lt2::dw_reduce_kernel<<<(kGates * kFeat + 5 * kGates * kHid + 255) / 256, 256, 0, st>>>(wf + tp.dwp, tp.dgrid, wf + tp.scale, g);
    FC_CUDA(cudaGetLastError(), "dw_reduce_kernel launch");
    double* fcp = reinterpret_cast<double*>(wf + tp.fcp);
    lt2::fc_grad_partial_kernel<<<kFcGradBlocksPerSm * sms, 256, 0, st>>>(wf + tp.hlast, d_out, B, fcp);
    FC_CUDA(cudaGetLastError(), "fc_grad_partial_kernel launch");
    lt2::fc_grad_reduce_kernel<<<lt2::kFcGradN, 64, 0, st>>>(fcp, kFcGradBlocksPerSm * sms, g_fc_w, g_fc_b);
    FC_CUDA(cudaGetLastError(), "fc_grad_reduce_kernel launch");
    return FC_OK;
  }
  LstmTrainPlan pl;
  int rc = lstm_train_plan(B, 1, &pl);
  if (rc) return rc;
  if (work_bytes < pl.floats * sizeof(float))
    return fail(FC_ERR_WORKSPACE, "fc_lstm_window_bwd: workspace too small%s (%lld < %lld bytes)", "", (long long)work_bytes,
                (long long)(pl.floats * sizeof(float)));
  if (!aligned16(pack) || !aligned16(work)) return fail(FC_ERR_MISALIGNED, "fc_lstm_window_bwd: pack / work must be 16-byte aligned%s");
  rc = ensure_smem_attributes();
  if (rc) return rc;
  lt::LstmBwdParams p;
  p.X = X; p.d_out = d_out; p.pack = pack; p.fc_w = fc_w; p.B = B;
  p.rec = (const float*)work + pl.rec;
  p.hseq = (const float*)work + pl.hseq;
  p.dseq = (float*)work + pl.dseq;
  p.partial = (float*)work + pl.partial;
  lt::lstm_window_bwd_kernel<<<pl.grid, lt::kThreadsL, lt::kSmemBwd, (cudaStream_t)stream>>>(p);
  FC_CUDA(cudaGetLastError(), "lstm_window_bwd_kernel launch");
  lt::LstmGradOut g;
  g.g_ih[0] = g_ih0; g.g_hh[0] = g_hh0; g.g_ih[1] = g_ih1; g.g_hh[1] = g_hh1; g.g_ih[2] = g_ih2; g.g_hh[2] = g_hh2;
  g.g_fc_w = g_fc_w; g.g_fc_b = g_fc_b;
  lt::lstm_grad_reduce_kernel<<<(51204 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p.partial, pl.grid, g);
  FC_CUDA(cudaGetLastError(), "lstm_grad_reduce_kernel launch");
  return FC_OK;
}

int fc_adamw_step(int count, float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                  const int* numel, int step, float lr, float beta1, float beta2, float eps, float weight_decay,
                  float grad_scale, void* stream) {
  if (count <= 0 || count > 8) return fail(FC_ERR_BAD_SHAPE, "fc_adamw_step: 1..8 tensors per call%s (count=%lld)", "", count);
  if (!params || !grads || !exp_avg || !exp_avg_sq || !numel) return fail(FC_ERR_NULL_POINTER, "fc_adamw_step: null pointer%s");
  if (step < 1) return fail(FC_ERR_BAD_SHAPE, "fc_adamw_step: step counts from 1%s (step=%lld)", "", step);
  lt::AdamWParams a;
  int most = 0;
  for (int k = 0; k < count; ++k) {
    if (!params[k] || !grads[k] || !exp_avg[k] || !exp_avg_sq[k] || numel[k] <= 0)
      return fail(FC_ERR_NULL_POINTER, "fc_adamw_step: null pointer or empty tensor%s at index %lld", "", k);
    a.p[k] = params[k]; a.g[k] = grads[k]; a.m[k] = exp_avg[k]; a.v[k] = exp_avg_sq[k]; a.n[k] = numel[k];
    if (numel[k] > most) most = numel[k];
  }
  a.count = count;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.grad_scale = grad_scale;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.sqrt_bc2 = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  int blocks = (most + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  lt::adamw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
  FC_CUDA(cudaGetLastError(), "adamw_kernel launch");
  return FC_OK;
}

// ---------------------------------------------------------------------------------------------------
// controller forward / backward for the u_0 path (fc_fnn.cuh)
// ---------------------------------------------------------------------------------------------------
static int fnn_grid(long long B, int per_block, int* grid) {
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  long long blocks = (B + per_block - 1) / per_block;
  const long long cap = (long long)sms * 4 < fn::kMaxBlocksF ? (long long)sms * 4 : fn::kMaxBlocksF;
  *grid = (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
  return FC_OK;
}

int fc_fnn_forward(const float* X, const float* inp_w, const float* inp_b, const float* out_w, long long B, float* u, void* stream) {
  if (B <= 0) return fail(FC_ERR_BAD_SHAPE, "fc_fnn_forward: B must be positive%s (B=%lld)", "", B);
  if (!X || !inp_w || !inp_b || !out_w || !u) return fail(FC_ERR_NULL_POINTER, "fc_fnn_forward: null pointer%s");
  int grid = 0;
  int rc = fnn_grid(B, fn::kThreadsF, &grid);
  if (rc) return rc;
  fn::fnn_forward_kernel<<<grid, fn::kThreadsF, 0, (cudaStream_t)stream>>>(X, inp_w, inp_b, out_w, B, u);
  FC_CUDA(cudaGetLastError(), "fnn_forward_kernel launch");
  return FC_OK;
}

size_t fc_fnn_backward_workspace_bytes(void) { return (size_t)fn::kMaxBlocksF * 250 * sizeof(float); }

int fc_fnn_backward(const float* X, const float* du, const float* inp_w, const float* inp_b, const float* out_w, long long B,
                    float* g_flat, void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0) return fail(FC_ERR_BAD_SHAPE, "fc_fnn_backward: B must be positive%s (B=%lld)", "", B);
  if (!X || !du || !inp_w || !inp_b || !out_w || !g_flat || !workspace) return fail(FC_ERR_NULL_POINTER, "fc_fnn_backward: null pointer%s");
  if (workspace_bytes < fc_fnn_backward_workspace_bytes()) return fail(FC_ERR_WORKSPACE, "fc_fnn_backward: workspace too small%s");
  int grid = 0;
  int rc = fnn_grid(B, fn::kThreadsF, &grid);
  if (rc) return rc;
  fn::fnn_backward_kernel<<<grid, fn::kThreadsF, 0, (cudaStream_t)stream>>>(X, du, inp_w, inp_b, out_w, B, (float*)workspace);
  FC_CUDA(cudaGetLastError(), "fnn_backward_kernel launch");
  fn::fnn_reduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((const float*)workspace, grid, g_flat);
  FC_CUDA(cudaGetLastError(), "fnn_reduce_kernel launch");
  return FC_OK;
}

#ifdef FC_TC_TRACE
// development aid: copies the event trace of the last pair-kernel launch (3 x 32768 events, counts in n[3])
int fc_debug_trace(long long* events, int* n) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(events, fc::g_trace, sizeof(long long) * 3 * 32768);
  cudaMemcpyFromSymbol(n, fc::g_trace_n, sizeof(int) * 3);
  return 0;
}
#endif
}  // extern "C"
