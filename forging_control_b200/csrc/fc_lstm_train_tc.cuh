// Surrogate training on the tensor cores (Model_NN/Functions.py:313-340 LSTMModel.forward,
// :520-569 train_model: forward -> nn.MSELoss -> loss.backward()), large batches.
//
// Forward and the data-gradient reverse sweep are the pair kernel in training mode (fc_mpc_pair_kernel.inl,
// MpcParams::train): one window per sample, every cell step recorded, the reverse sweep seeded by d loss / d y.  That
// kernel leaves, per 128-sample tile, in operand format ([piece of 8][128 samples][8 halves], fp16 hi and lo images;
// fc_pair_layout.h kTr*): the gate gradients dG of every (layer, step), the hidden sequence of every layer and the window
// features.  This file holds what follows:
//
//   dw_kernel        dW[l][gate m][k] += sum over samples of dG[l][t][s][m] * act[l][t][s][k]   (act = layer input | h_{t-1})
//                    A contraction over the SAMPLES: the K-major images the roll-out kernels write (row = sample) are read as
//                    MN-major tcgen05 operands, no transposition (scripts/micro/umma_mn_test.cu).  Three fp16 hi/lo terms,
//                    fp32 accumulation in TMEM (two accumulator buffers), 64-sample stages streamed with cp.async.bulk
//                    through a 2-deep mbarrier pipeline; warp 8 = producer, warp 9 = MMA issuer, warps 0..7 drain the
//                    accumulator of every (tile, layer) into registers (112 columns each) and write per-CTA partial sums.
//   dw_reduce_kernel per-CTA partials -> the six nn.LSTM gradient tensors (fp64 sum over CTAs, unscale, PyTorch row order)
//   fc_grad_*        d fc.weight = dy^T h_last, d fc.bias = sum dy
//   grad_scale_kernel power-of-two scale of the gate gradients from max |dy| (fp16 range of the hi/lo split)
#pragma once

namespace fc {
namespace lt2 {

constexpr int kStageRows = 32;                       // samples per pipeline stage (a quarter of a tile)
constexpr int kStagesPerTile = pr::kTileP / kStageRows;      // of a 128-sample tile (DwParams::spt at run time)
constexpr int kDgPieces = 26, kActPieces = 14;       // k-slot pieces of 8: 208 gate gradients; 56 input + 56 recurrent slots
constexpr int kStDgBytes = kDgPieces * kStageRows * 16;     // one of hi / lo
constexpr int kStActBytes = kActPieces * kStageRows * 16;
constexpr int kStageBytes = 2 * (kStDgBytes + kStActBytes); // dG hi | dG lo | act hi | act lo = 40 960
constexpr int kStages = 5;                           // 200 KiB of operands in flight per SM (two 80 KiB stages left the HBM at 50 %)
constexpr int kDwThreads = 320;                      // warps 0..7 epilogue, 8 producer, 9 MMA issuer
constexpr int kDwSmem = kStages * kStageBytes + 1024;
constexpr int kDwRows = 256, kDwCols = 112;          // partial [3 layers][256 operand rows][112]
constexpr size_t kDwPartialFloats = (size_t)kLayers * kDwRows * kDwCols;
// The tensor core adds every K block into the fp32 accumulator with truncation: a chain of S additions loses
// (0.17 + 0.135 S) ulp on average.  The accumulator is therefore drained into registers every kSegSteps
// cell steps (S = 48: 6.7 ulp, compensated to a fraction of an ulp) instead of once per (tile, layer) (S = 240: 33 ulp).
constexpr int kSegSteps = 2;
constexpr int kSegs = kLook / kSegSteps;             // accumulator segments per (tile, layer)
// MMA accumulation steps per segment: steps x stages x 2 k-steps of 16 samples x 3 terms (48 for 128-sample tiles)

struct DwParams {
  const float* ws;        // per-tile scratch written by the pair kernel (pr::kTrTileFloats / (4 / spt) floats per tile)
  int tiles;              // tiles in ws
  int spt;                // 32-sample stages per tile: 4 (128-sample tiles of the pair kernel) or 1 (its replica mode)
  float* partial;         // [grid][3][256][112], accumulated across launches (chunks of the batch)
  float acc_comp;
};

__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr) {
  // MN-major, no swizzle: LBO = 128 B between core matrices of 8 samples (K), SBO = kStageRows * 16 B between groups of 8 (M/N)
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((128u >> 4) & 0x3FFF) << 16) |
         ((uint64_t)(((uint32_t)kStageRows * 16u >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_mn(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

// barriers: full[s] 0..kStages-1, empty[s] kStages.., dfull[b], dempty[b]
constexpr int kBarEmpty = kStages, kBarDFull = 2 * kStages, kBarDEmpty = 2 * kStages + 2, kDwBars = 2 * kStages + 4;
__global__ void __launch_bounds__(kDwThreads, 1) dw_kernel(const DwParams p) {
  extern __shared__ __align__(1024) unsigned char dsm[];
  __shared__ __align__(8) unsigned long long bars[kDwBars];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t sbase = (smem_u32(dsm) + 1023u) & ~1023u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar0 + s * 8, 1); mbar_init(bar0 + (kBarEmpty + s) * 8, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar0 + (kBarDFull + b) * 8, 1); mbar_init(bar0 + (kBarDEmpty + b) * 8, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_slot;
  const int ntile = p.tiles > (int)blockIdx.x ? (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;   // tiles of this CTA
  const int spt = p.spt, rdiv = kStagesPerTile / spt, rows = kStageRows * spt;     // stages, size divisor and samples of a tile

  if (warp == 8) {
    // ---- producer: 64-sample stages of (layer l, tile, step t, half h), pieces of 1 KiB
    uint32_t it = 0;
    for (int l = 0; l < kLayers; ++l)
      for (int i = 0; i < ntile; ++i) {
        const float* tw = p.ws + (size_t)(blockIdx.x + (size_t)i * gridDim.x) * (pr::kTrTileFloats / rdiv);
        for (int t = 0; t < kLook; ++t)
          for (int h = 0; h < spt; ++h, ++it) {
            const int s = it % kStages;
            const uint32_t full = bar0 + s * 8, empty = bar0 + (kBarEmpty + s) * 8;
            if (it >= kStages) mbar_wait(empty, ((it / kStages) - 1) & 1);
            const uint32_t st = sbase + s * kStageBytes;
            const int nin = l == 0 ? 1 : 7;                        // input pieces (layer 0: the feature piece)
            const uint32_t in_bytes = (uint32_t)nin * kStageRows * 16u, rc_bytes = 7u * kStageRows * 16u;
            if (lane == 0) mbar_expect(full, 2u * kStDgBytes + 2u * (in_bytes + rc_bytes));
            __syncwarp();
            // the pair kernel stores every image stage-major ([hi | lo][stage][piece][32 rows][16 B]): six bulk copies per stage
            const float* dgb = tw + pr::kTrDgOff / rdiv + (size_t)(l * kLook + t) * (pr::kTrDgSlot / rdiv);
            const float* inb = l == 0 ? tw + pr::kTrFeatOff / rdiv + (size_t)t * (pr::kTrFeatSlot / rdiv)
                                      : tw + pr::kTrHseqOff / rdiv + (size_t)((l - 1) * (kLook + 1) + t + 1) * (pr::kTrHseqSlot / rdiv);
            const float* rcb = tw + pr::kTrHseqOff / rdiv + (size_t)(l * (kLook + 1) + t) * (pr::kTrHseqSlot / rdiv);       // h_{t-1} (slot 0 = zeros)
            if (lane < 2) {
              bulk_g2s(st + lane * kStDgBytes, dgb + (size_t)lane * (kDgPieces * rows * 4) + (size_t)h * (kDgPieces * kStageRows * 4),
                       (uint32_t)kStDgBytes, full);
            } else if (lane < 4) {
              const int hl = lane - 2;
              bulk_g2s(st + 2 * kStDgBytes + hl * kStActBytes, inb + (size_t)hl * (nin * rows * 4) + (size_t)h * (nin * kStageRows * 4), in_bytes, full);
            } else if (lane < 6) {
              const int hl = lane - 4;
              bulk_g2s(st + 2 * kStDgBytes + hl * kStActBytes + in_bytes, rcb + (size_t)hl * (7 * rows * 4) + (size_t)h * (7 * kStageRows * 4), rc_bytes, full);
            }
          }
      }
  } else if (warp == 9) {
    // ---- MMA issuer (all lanes walk the loop, lane 0 issues)
    uint32_t it = 0, nd = 0;
    for (int l = 0; l < kLayers; ++l) {
      const int n = l == 0 ? 64 : 112;
      const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int i = 0; i < ntile; ++i) {
        for (int t = 0; t < kLook; ++t)
          for (int h = 0; h < spt; ++h, ++it) {
            if (t % kSegSteps == 0 && h == 0) {                              // new accumulator segment
              if (t > 0 || i > 0 || l > 0) ++nd;
              if (nd >= 2) mbar_wait(bar0 + (kBarDEmpty + (nd & 1)) * 8, ((nd >> 1) - 1) & 1);   // accumulator buffer drained
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const int b = nd & 1;
            const int s = it % kStages;
            mbar_wait(bar0 + s * 8, (it / kStages) & 1);                      // stage landed
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
              const uint32_t st = sbase + s * kStageBytes;
              const uint32_t g_hi = st, g_lo = st + kStDgBytes, a_hi = st + 2 * kStDgBytes, a_lo = a_hi + kStActBytes;
#pragma unroll 1
              for (int mh = 0; mh < 2; ++mh) {
                const uint32_t d = tbase + b * 224 + mh * 112;
                const uint32_t moff = mh * 16 * (kStageRows * 16);
                const bool first = t % kSegSteps == 0 && h == 0;
#pragma unroll
                for (int ks = 0; ks < kStageRows / 16; ++ks)
                  umma_mn(d, mn_desc(g_lo + moff + ks * 256), mn_desc(a_hi + ks * 256), idesc, (first && ks == 0) ? 0u : 1u);
#pragma unroll
                for (int ks = 0; ks < kStageRows / 16; ++ks) umma_mn(d, mn_desc(g_hi + moff + ks * 256), mn_desc(a_lo + ks * 256), idesc, 1u);
#pragma unroll
                for (int ks = 0; ks < kStageRows / 16; ++ks) umma_mn(d, mn_desc(g_hi + moff + ks * 256), mn_desc(a_hi + ks * 256), idesc, 1u);
              }
              umma_commit(bar0 + (kBarEmpty + s) * 8);                         // stage free once these MMAs have read it
              if (t % kSegSteps == kSegSteps - 1 && h == spt - 1) umma_commit(bar0 + (kBarDFull + b) * 8);   // segment complete
            }
            __syncwarp();
          }
      }
    }
  } else {
    // ---- epilogue warps 0..7: operand row m = 128 (warp / 4) + 32 (warp % 4) + lane, all 112 columns in registers
    const int mh = warp >> 2, q = warp & 3;
    const uint32_t lane_addr = tbase + ((uint32_t)(32 * q) << 16);
    const float comp = p.acc_comp * (0.17f + 0.135f * (float)(kSegSteps * spt * 2 * 3)) * 1.1920929e-7f;
    float* out = p.partial + (size_t)blockIdx.x * kDwPartialFloats;
    uint32_t nd = 0;
    for (int l = 0; l < kLayers; ++l) {
      float acc[kDwCols];
#pragma unroll
      for (int c = 0; c < kDwCols; ++c) acc[c] = 0.f;
      for (int i = 0; i < ntile * kSegs; ++i, ++nd) {
        const int b = nd & 1;
        mbar_wait(bar0 + (kBarDFull + b) * 8, (nd >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < kDwCols; c0 += 16) {
          uint32_t r[16];
          TmemIO<16>::ld(lane_addr + b * 224 + mh * 112 + c0, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int c = 0; c < 16; ++c) { const float v = __uint_as_float(r[c]); acc[c0 + c] += fmaf(v, comp, v); }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + (kBarDEmpty + b) * 8);
      }
      float* o = out + ((size_t)l * kDwRows + (size_t)(128 * mh + 32 * q + lane)) * kDwCols;
#pragma unroll
      for (int c = 0; c < kDwCols; c += 4) {
        float4 v = *reinterpret_cast<float4*>(o + c);
        v.x += acc[c]; v.y += acc[c + 1]; v.z += acc[c + 2]; v.w += acc[c + 3];
        *reinterpret_cast<float4*>(o + c) = v;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

struct DwGradOut {
  float* g_ih[kLayers];
  float* g_hh[kLayers];
};
// element e of the six tensors: [g_ih0 200x5 | g_hh0 200x50 | g_ih1 | g_hh1 | g_ih2 | g_hh2 (200x50 each)]
__global__ void dw_reduce_kernel(const float* __restrict__ partial, int grid, const float* __restrict__ scale, DwGradOut g) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int n0 = kGates * kFeat, nh = kGates * kHid;
  if (e >= n0 + 5 * nh) return;
  int l, hh, r, c;
  if (e < n0) { l = 0; hh = 0; r = e / kFeat; c = e % kFeat; }
  else { const int e2 = e - n0, blk = e2 / nh, w = e2 % nh; l = (blk + 1) / 2; hh = (blk + 1) & 1; r = w / kHid; c = w % kHid; }
  // PyTorch row r = gate * 50 + unit -> operand row m = unit * 4 + gate; column: input k | recurrent (l = 0: 8 + u, else 56 + u)
  const int m = (r % kHid) * 4 + r / kHid;
  const int col = hh ? (l == 0 ? 8 : 56) + c : c;
  double a = 0.0;
  for (int b = 0; b < grid; ++b) a += (double)partial[(size_t)b * kDwPartialFloats + ((size_t)l * kDwRows + m) * kDwCols + col];
  const float v = (float)(a * (double)scale[1] * (1.0 / (double)pr::kScaleA));
  (hh ? g.g_hh[l] : g.g_ih[l])[r * (hh || l > 0 ? kHid : kFeat) + c] = v;
}

// scale[0] = power of two that brings max |dy| to ~2^11 (the gate gradients, a few times the seed at most, stay far below the
// fp16 range; the conversions saturate), scale[1] = 1 / scale[0].
// Two launches: block maxima folded with an atomic max on the bit pattern (non-negative floats order like unsigned integers)
// into scale[2] (zeroed before), then one thread turns it into the scale.
__global__ void __launch_bounds__(256) grad_absmax_kernel(const float* __restrict__ dy, long long n, float* __restrict__ scale) {
  __shared__ float red[8];
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = fabsf(dy[i]);
    if (v == v && v < 3.0e38f) m = fmaxf(m, v);
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    atomicMax(reinterpret_cast<unsigned int*>(scale + 2), __float_as_uint(m));
  }
}
__global__ void grad_scale_kernel(float* __restrict__ scale) {
  const float m = scale[2];
  int e = 0;
  if (m > 0.f) { frexpf(m, &e); e = 11 - e; }                // m in [2^(e-1), 2^e): max |dy| lands in [2^10, 2^11), 32x below fp16 max
  e = e > 100 ? 100 : (e < -100 ? -100 : e);
  scale[0] = ldexpf(1.0f, e);
  scale[1] = ldexpf(1.0f, -e);
}

// d fc.weight [4][50] = dy^T h_last, d fc.bias [4] = column sums of dy: per-block partial sums (fp64), then one reduce
constexpr int kFcGradN = kOut * kHid + kOut;
__global__ void __launch_bounds__(256) fc_grad_partial_kernel(const float* __restrict__ hlast, const float* __restrict__ dy, int B, double* __restrict__ partial) {
  const int tid = threadIdx.x;
  if (tid < kFcGradN) {
    const int q = tid < kOut * kHid ? tid / kHid : tid - kOut * kHid, u = tid % kHid;
    const bool w = tid < kOut * kHid;
    // contiguous slice of the batch per block; four independent sums hide the load latency
    const long long per = ((long long)B + gridDim.x - 1) / gridDim.x;
    const long long s0 = (long long)blockIdx.x * per, s1 = s0 + per < B ? s0 + per : B;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;           // the sums cancel heavily: fp64 (204 x B products, negligible work)
    long long s = s0;
    for (; s + 3 < s1; s += 4) {
      a0 = fma((double)dy[(size_t)s * kOut + q], w ? (double)hlast[(size_t)s * kHid + u] : 1.0, a0);
      a1 = fma((double)dy[(size_t)(s + 1) * kOut + q], w ? (double)hlast[(size_t)(s + 1) * kHid + u] : 1.0, a1);
      a2 = fma((double)dy[(size_t)(s + 2) * kOut + q], w ? (double)hlast[(size_t)(s + 2) * kHid + u] : 1.0, a2);
      a3 = fma((double)dy[(size_t)(s + 3) * kOut + q], w ? (double)hlast[(size_t)(s + 3) * kHid + u] : 1.0, a3);
    }
    for (; s < s1; ++s) a0 = fma((double)dy[(size_t)s * kOut + q], w ? (double)hlast[(size_t)s * kHid + u] : 1.0, a0);
    const double a = (a0 + a1) + (a2 + a3);
    partial[(size_t)blockIdx.x * 256 + tid] = a;
  }
}
__global__ void __launch_bounds__(64) fc_grad_reduce_kernel(const double* __restrict__ partial, int grid, float* __restrict__ g_fc_w, float* __restrict__ g_fc_b) {
  const int o = blockIdx.x;                                  // one block per output element
  __shared__ double red[2];
  double a = 0.0;
  for (int b = threadIdx.x; b < grid; b += 64) a += partial[(size_t)b * 256 + o];
  for (int k = 16; k; k >>= 1) a += __shfl_xor_sync(0xffffffffu, a, k);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    a = red[0] + red[1];
    if (o < kOut * kHid) g_fc_w[o] = (float)a; else g_fc_b[o - kOut * kHid] = (float)a;
  }
}

}  // namespace lt2
}  // namespace fc
