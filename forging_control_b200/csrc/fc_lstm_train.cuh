// Surrogate training path and optimizer update (8f-2), sm_100a.
//
//   lstm_window_fwd_kernel   LSTMModel.forward (Unsupervised Learning/Model_NN/Functions.py:313-340, identical to
//                            UL/Functions.py:353-379): 3 bias-free LSTM layers from a zero state over a 10-row window,
//                            read-out fc on the last hidden state; optionally records the cell activations
//   lstm_window_bwd_kernel   what loss.backward() (Model_NN/Functions.py:560) leaves in .grad of the eight surrogate
//                            parameters for an upstream gradient d_out [B,4] (nn.MSELoss stays the caller's torch
//                            op): reverse sweep over layers and time with the WEIGHT gradients, per-CTA partial sums
//   lstm_grad_reduce_kernel  sum of the per-CTA partials (fp64, fixed order) into the eight gradient tensors
//   adamw_kernel             torch.optim.AdamW.step (Model_NN/Main.py:230, UL/Main.py:195) for up to 8 tensors, 1 launch
//
// FP32 FFMA kernels: the weight-gradient contraction reduces over TRAJECTORIES (K = batch), which the
// trajectory-per-TMEM-lane layout of the tcgen05 roll-out kernels cannot feed; a tensor-core version needs the
// transposed operand staging and is listed under next steps in DESIGN.md.
//
// Tile = 40 samples, 250 compute threads of 256: thread (ug = tid/10, tg = tid%10) owns hidden units 2ug, 2ug+1 (all four
// gates) of samples 4tg..4tg+3, so cell state, d(cell) and the recurrent d(h) stay in registers; the same thread grid
// owns 8 rows x 10 columns of the 200 x 100 weight-gradient matrix [W_ih | W_hh] of the current layer, held in
// 80 registers over the 10 time steps (rows ug + 25 r, columns tg + 10 j).  All shared-memory operands are [row][samples]
// and are read as float4.
#pragma once

namespace fc {
namespace lt {

constexpr int kTT = 40;                 // samples per tile
constexpr int kThreadsL = 256;
constexpr int kActive = 250;
constexpr int kL = 10, kH = 50, kG = 200;
constexpr int kWfFloats = 55 * kG + 100 * kG + 100 * kG;     // forward images  Wf_l[k][ug*8 + gate*2 + uu]
constexpr int kWbFloats = 3 * kG * 100;                      // backward images WB_l[row][ug][ih0 ih1 hh0 hh1]
constexpr int kPackFloatsL = kWfFloats + kWbFloats;
constexpr int kRecSlots = 10;                                // float4 per thread and cell: i,f,g,o,c x 2 units
// Workspace layout: the unit is a PAIR of 40-sample tiles (80 samples), so that the 80-sample forward kernel and the
// 40-sample kernels share it.  records [pair][30 cells][2 halves][10 slots][256 threads] float4 (thread-private,
// warp-coalesced); hidden sequences [pair][30 cells][50 units][80 samples].
constexpr int kTF = 80;                 // samples per pair / per tile of the 80-sample forward kernel
constexpr size_t kRecCell = (size_t)kRecSlots * kThreadsL * 4;              // floats per (cell, half)
constexpr size_t kRecFloatsPair = (size_t)30 * 2 * kRecCell;
constexpr size_t kHseqFloatsPair = (size_t)30 * kH * kTF;
__host__ __device__ inline size_t rec_offset(int tile40, int cell) {          // floats, without the thread's slot offset
  return (size_t)(tile40 >> 1) * kRecFloatsPair + ((size_t)cell * 2 + (tile40 & 1)) * kRecCell;
}
__host__ __device__ inline size_t hseq_offset(int tile40) {                   // row stride kTF
  return (size_t)(tile40 >> 1) * kHseqFloatsPair + (size_t)(tile40 & 1) * kTT;
}
constexpr size_t kDseqFloatsCta = (size_t)kL * 2 * kThreadsL * 4;
constexpr int kPartialFloats = 3 * kG * 100 + 200 + 4;       // per CTA: dW_l [200][100] x 3, d fc.weight, d fc.bias
constexpr int kSmemFwd = (100 * kG + kL * kH * kTT + kL * 5 * kTT) * 4;                 // 168 000 B
constexpr int kSB = 44;                 // row stride (floats) of the backward shared-memory operands: 44 mod 32 = 12 puts the rows
                                        // of threads that differ in ug or tg by one into different 16-byte bank groups
constexpr int kSmemBwd = (kG * 100 + kG * kSB + 2 * 100 * kSB + 4 * kTT) * 4;           // 151 040 B

__host__ __device__ inline int wf_offset(int l) { return l == 0 ? 0 : (l == 1 ? 55 * kG : 155 * kG); }

struct LstmRaw {
  const float* w_ih[3];
  const float* w_hh[3];
};
struct LstmGradOut {
  float* g_ih[3];
  float* g_hh[3];
  float* g_fc_w;
  float* g_fc_b;
};

__global__ void __launch_bounds__(256) pack_lstm_train_kernel(LstmRaw w, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kWfFloats) {
    const int l = i < 55 * kG ? 0 : (i < 155 * kG ? 1 : 2);
    const int kin = l == 0 ? 5 : 50;
    const int r = i - wf_offset(l);
    const int k = r / kG, col = r % kG;
    const int ug = col >> 3, gate = (col >> 1) & 3, uu = col & 1;
    const int row = gate * kH + 2 * ug + uu;
    out[i] = k < kin ? w.w_ih[l][row * kin + k] : w.w_hh[l][row * kH + (k - kin)];
  } else if (i < kPackFloatsL) {
    const int r = i - kWfFloats;
    const int l = r / (kG * 100), q = r % (kG * 100);
    const int row = q / 100, ug = (q % 100) >> 2, e = q & 3;
    const int u = 2 * ug + (e & 1);
    out[i] = e < 2 ? (l == 0 ? 0.f : w.w_ih[l][row * kH + u]) : w.w_hh[l][row * kH + u];
  }
}

__device__ __forceinline__ float sigm(float x) { return DevCtx::rcp(1.f + DevCtx::ex2(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_(float x) {
  const float big = 1.f - 2.f * DevCtx::rcp(1.f + DevCtx::ex2(2.8853900817779268f * x));
  const float x2 = x * x;
  float pl = fmaf(x2, 0.021869488536155203f, -0.053968253968253971f);
  pl = fmaf(x2, pl, 0.13333333333333333f);
  pl = fmaf(x2, pl, -0.33333333333333331f);
  pl = fmaf(x2 * x, pl, x);
  return fabsf(x) < 0.3f ? pl : big;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

struct LstmFwdParams {
  const float* X;        // [B,10,5]
  const float* pack;     // pack_lstm_train_kernel output
  const float* fc_w;     // [4,50]
  const float* fc_b;     // [4]
  float* out;            // [B,4]
  float* rec;            // [tiles] records (save != 0)
  float* hseq;           // hidden sequences (save != 0; the 80-sample kernel also needs it as per-CTA scratch otherwise)
  int B, save;
};

__global__ void __launch_bounds__(kThreadsL, 1) lstm_window_fwd_kernel(const LstmFwdParams p) {
  float* Wf = fc_dyn_smem;                       // [K][200]
  float* seq = Wf + 100 * kG;                    // [10][50][40]  hidden sequence, updated in place layer by layer
  float* xs = seq + kL * kH * kTT;               // [10][5][40]
  const int tid = threadIdx.x;
  const bool active = tid < kActive;
  const int ug = active ? tid / 10 : 0, tg = active ? tid % 10 : 0;
  const int tiles = (p.B + kTT - 1) / kTT;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int b0 = tile * kTT;
    __syncthreads();
    for (int i = tid; i < kL * 5 * kTT; i += kThreadsL) {          // xs[t][k][b] <- X[b0+b][t][k]
      const int b = i % kTT, tk = i / kTT;
      xs[i] = (b0 + b < p.B) ? __ldg(p.X + (size_t)(b0 + b) * 50 + tk) : 0.f;
    }
    for (int l = 0; l < 3; ++l) {
      const int kin = l == 0 ? 5 : 50;
      __syncthreads();                                             // previous layer's GEMM reads of Wf are done
      {
        const float* src = p.pack + wf_offset(l);
        const int n = (kin + kH) * kG;
        for (int i = tid * 4; i < n; i += kThreadsL * 4) st4(Wf + i, __ldg(reinterpret_cast<const float4*>(src + i)));
      }
      __syncthreads();
      float c[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) c[e] = 0.f;
      for (int t = 0; t < kL; ++t) {
        float acc[8][4];
#pragma unroll
        for (int g = 0; g < 8; ++g)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[g][j] = 0.f;
        const float* in = l == 0 ? xs + t * 5 * kTT : seq + t * kH * kTT;
        const float* wp = Wf + ug * 8;
        const float* ap = in + tg * 4;
#pragma unroll 5
        for (int k = 0; k < kin; ++k) {
          const float4 w0 = ld4(wp + k * kG), w1 = ld4(wp + k * kG + 4), a = ld4(ap + k * kTT);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int g = 0; g < 8; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[g][j] = fmaf(wv[g], av[j], acc[g][j]);
        }
        if (t > 0) {
          const float* hp = seq + (t - 1) * kH * kTT + tg * 4;
          const float* wr = wp + kin * kG;
#pragma unroll 5
          for (int k = 0; k < kH; ++k) {
            const float4 w0 = ld4(wr + k * kG), w1 = ld4(wr + k * kG + 4), a = ld4(hp + k * kTT);
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int g = 0; g < 8; ++g)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[g][j] = fmaf(wv[g], av[j], acc[g][j]);
          }
        }
        __syncthreads();                                           // every read of seq[t] (layer input) is done
        if (active) {
          float hv[2][4];
          float gi[2][4], gf[2][4], gg[2][4], go[2][4];
#pragma unroll
          for (int uu = 0; uu < 2; ++uu)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              gi[uu][j] = sigm(acc[0 + uu][j]);
              gf[uu][j] = sigm(acc[2 + uu][j]);
              gg[uu][j] = tanh_(acc[4 + uu][j]);
              go[uu][j] = sigm(acc[6 + uu][j]);
              const float cn = fmaf(gf[uu][j], c[uu * 4 + j], gi[uu][j] * gg[uu][j]);
              c[uu * 4 + j] = cn;
              hv[uu][j] = go[uu][j] * tanh_(cn);
            }
#pragma unroll
          for (int uu = 0; uu < 2; ++uu)
            st4(seq + (t * kH + 2 * ug + uu) * kTT + tg * 4, make_float4(hv[uu][0], hv[uu][1], hv[uu][2], hv[uu][3]));
          if (p.save) {
            float* hs = p.hseq + hseq_offset(tile) + (size_t)((l * kL + t) * kH) * kTF;
#pragma unroll
            for (int uu = 0; uu < 2; ++uu)
              st4(hs + (2 * ug + uu) * kTF + tg * 4, make_float4(hv[uu][0], hv[uu][1], hv[uu][2], hv[uu][3]));
            float* r = p.rec + rec_offset(tile, l * kL + t) + tid * 4;
#pragma unroll
            for (int uu = 0; uu < 2; ++uu) {
              __stcs(reinterpret_cast<float4*>(r + (0 + uu) * kThreadsL * 4), make_float4(gi[uu][0], gi[uu][1], gi[uu][2], gi[uu][3]));
              __stcs(reinterpret_cast<float4*>(r + (2 + uu) * kThreadsL * 4), make_float4(gf[uu][0], gf[uu][1], gf[uu][2], gf[uu][3]));
              __stcs(reinterpret_cast<float4*>(r + (4 + uu) * kThreadsL * 4), make_float4(gg[uu][0], gg[uu][1], gg[uu][2], gg[uu][3]));
              __stcs(reinterpret_cast<float4*>(r + (6 + uu) * kThreadsL * 4), make_float4(go[uu][0], go[uu][1], go[uu][2], go[uu][3]));
              __stcs(reinterpret_cast<float4*>(r + (8 + uu) * kThreadsL * 4),
                     make_float4(c[uu * 4 + 0], c[uu * 4 + 1], c[uu * 4 + 2], c[uu * 4 + 3]));
            }
          }
        }
        __syncthreads();                                           // h_t visible to the next step / the read-out
      }
    }
    if (tid < 4 * kTT) {                                           // read-out fc on h of the top layer at t = 9
      const int b = tid % kTT, o = tid / kTT;
      if (b0 + b < p.B) {
        float a = __ldg(p.fc_b + o);
        const float* h = seq + 9 * kH * kTT + b;
        for (int u = 0; u < kH; ++u) a = fmaf(__ldg(p.fc_w + o * kH + u), h[u * kTT], a);
        p.out[(size_t)(b0 + b) * 4 + o] = a;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// 80-sample forward (large batches): the same arithmetic with an 8 gate-column x 8 sample register tile (4 LDS.128 per
// 64 FFMA instead of 3 per 32).  The whole-window hidden sequence of 80 samples does not fit shared memory next to the
// weights, so the layer below is streamed: it is written to the hidden-sequence workspace anyway (or to a per-CTA
// scratch of the same shape when nothing is recorded) and comes back one step ahead with cp.async.
// Shared memory: weights [100][200] | input [2][50][80] | own h [2][50][80] | window [10][5][80] = 160 000 B.
// ---------------------------------------------------------------------------------------------------
constexpr int kSmemFwd80 = (100 * kG + 2 * kH * kTF + 2 * kH * kTF + kL * 5 * kTF) * 4;

__global__ void __launch_bounds__(kThreadsL, 1) lstm_window_fwd80_kernel(const LstmFwdParams p) {
  float* Wf = fc_dyn_smem;                       // [K][200]
  float* inb = Wf + 100 * kG;                    // [2][50][80]  h of the layer below at step t (t & 1)
  float* hb = inb + 2 * kH * kTF;                // [2][50][80]  own h at step t (t & 1)
  float* xs = hb + 2 * kH * kTF;                 // [10][5][80]
  const int tid = threadIdx.x;
  const bool active = tid < kActive;
  const int ug = active ? tid / 10 : 0, tg = active ? tid % 10 : 0;
  const int pairs = (p.B + kTF - 1) / kTF;
  for (int pair = blockIdx.x; pair < pairs; pair += gridDim.x) {
    const int b0 = pair * kTF;
    float* hs = p.hseq + (size_t)(p.save ? pair : (int)blockIdx.x) * kHseqFloatsPair;
    __syncthreads();
    for (int i = tid; i < kL * 5 * kTF; i += kThreadsL) {          // xs[t][k][b] <- X[b0+b][t][k]
      const int b = i % kTF, tk = i / kTF;
      xs[i] = (b0 + b < p.B) ? __ldg(p.X + (size_t)(b0 + b) * 50 + tk) : 0.f;
    }
    for (int l = 0; l < 3; ++l) {
      const int kin = l == 0 ? 5 : 50;
      __syncthreads();                                             // previous layer: Wf / hb reads and hs writes are done
      {
        const float* src = p.pack + wf_offset(l);
        const int n = (kin + kH) * kG;
        for (int i = tid * 4; i < n; i += kThreadsL * 4) st4(Wf + i, __ldg(reinterpret_cast<const float4*>(src + i)));
      }
      auto stage = [&](int tt) {                                   // h of the layer below at step tt -> inb[tt & 1]
        const float* src = hs + (size_t)((l - 1) * kL + tt) * kH * kTF;
        float* dst = inb + (tt & 1) * kH * kTF;
        for (int i = tid * 4; i < kH * kTF; i += kThreadsL * 4) DevCtx::cp_async16(dst + i, src + i);
        DevCtx::cp_commit();
      };
      if (l > 0) stage(0);
      DevCtx::cp_wait<0>();
      __syncthreads();
      float c[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) c[e] = 0.f;
      for (int t = 0; t < kL; ++t) {
        if (l > 0 && t + 1 < kL) stage(t + 1);
        float acc[8][8];
#pragma unroll
        for (int g = 0; g < 8; ++g)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[g][j] = 0.f;
        const float* in = l == 0 ? xs + t * 5 * kTF : inb + (t & 1) * kH * kTF;
        const float* wp = Wf + ug * 8;
        const float* ap = in + tg * 4;
#pragma unroll 2
        for (int k = 0; k < kin; ++k) {
          const float4 w0 = ld4(wp + k * kG), w1 = ld4(wp + k * kG + 4);
          const float4 a0 = ld4(ap + k * kTF), a1 = ld4(ap + k * kTF + 40);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
          for (int g = 0; g < 8; ++g)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[g][j] = fmaf(wv[g], av[j], acc[g][j]);
        }
        if (t > 0) {
          const float* hp = hb + ((t - 1) & 1) * kH * kTF + tg * 4;
          const float* wr = wp + kin * kG;
#pragma unroll 2
          for (int k = 0; k < kH; ++k) {
            const float4 w0 = ld4(wr + k * kG), w1 = ld4(wr + k * kG + 4);
            const float4 a0 = ld4(hp + k * kTF), a1 = ld4(hp + k * kTF + 40);
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int g = 0; g < 8; ++g)
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[g][j] = fmaf(wv[g], av[j], acc[g][j]);
          }
        }
        if (active) {
          float* hrow = hb + (t & 1) * kH * kTF;
          float* hsg = hs + (size_t)((l * kL + t) * kH) * kTF;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float* r = p.save ? p.rec + rec_offset(2 * pair + half, l * kL + t) + tid * 4 : nullptr;
#pragma unroll
            for (int uu = 0; uu < 2; ++uu) {
              float gi[4], gf[4], gg[4], go[4], hv[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int e = uu * 8 + half * 4 + j;
                gi[j] = sigm(acc[0 + uu][half * 4 + j]);
                gf[j] = sigm(acc[2 + uu][half * 4 + j]);
                gg[j] = tanh_(acc[4 + uu][half * 4 + j]);
                go[j] = sigm(acc[6 + uu][half * 4 + j]);
                const float cn = fmaf(gf[j], c[e], gi[j] * gg[j]);
                c[e] = cn;
                hv[j] = go[j] * tanh_(cn);
              }
              const int off = (2 * ug + uu) * kTF + half * 40 + tg * 4;
              const float4 h4 = make_float4(hv[0], hv[1], hv[2], hv[3]);
              st4(hrow + off, h4);
              st4(hsg + off, h4);                                  // the layer above (and the reverse sweep) read it back
              if (p.save) {
                const int e0 = uu * 8 + half * 4;
                __stcs(reinterpret_cast<float4*>(r + (0 + uu) * kThreadsL * 4), make_float4(gi[0], gi[1], gi[2], gi[3]));
                __stcs(reinterpret_cast<float4*>(r + (2 + uu) * kThreadsL * 4), make_float4(gf[0], gf[1], gf[2], gf[3]));
                __stcs(reinterpret_cast<float4*>(r + (4 + uu) * kThreadsL * 4), make_float4(gg[0], gg[1], gg[2], gg[3]));
                __stcs(reinterpret_cast<float4*>(r + (6 + uu) * kThreadsL * 4), make_float4(go[0], go[1], go[2], go[3]));
                __stcs(reinterpret_cast<float4*>(r + (8 + uu) * kThreadsL * 4), make_float4(c[e0], c[e0 + 1], c[e0 + 2], c[e0 + 3]));
              }
            }
          }
        }
        DevCtx::cp_wait<0>();
        __syncthreads();                                           // h_t and the next input block are visible
      }
    }
    for (int i = tid; i < 4 * kTF; i += kThreadsL) {               // read-out fc on h of the top layer at t = 9
      const int b = i % kTF, o = i / kTF;
      if (b0 + b < p.B) {
        float a = __ldg(p.fc_b + o);
        const float* h = hb + (9 & 1) * kH * kTF + b;
        for (int u = 0; u < kH; ++u) a = fmaf(__ldg(p.fc_w + o * kH + u), h[u * kTF], a);
        p.out[(size_t)(b0 + b) * 4 + o] = a;
      }
    }
  }
}

struct LstmBwdParams {
  const float* X;        // [B,10,5]
  const float* d_out;    // [B,4]
  const float* pack;
  const float* fc_w;
  const float* rec;
  const float* hseq;
  float* dseq;           // [grid] thread-private d(input sequence) slots
  float* partial;        // [grid][kPartialFloats]
  int B;
};

__global__ void __launch_bounds__(kThreadsL, 1) lstm_window_bwd_kernel(const LstmBwdParams p) {
  float* WB = fc_dyn_smem;                       // [200][25][4]
  float* dG = WB + kG * 100;                     // [200][kSB]
  float* act = dG + kG * kSB;                    // [2][100][kSB]: rows 0..49 layer input at t, 50..99 own h at t-1
  float* dout = act + 2 * 100 * kSB;             // [4][40]
  const int tid = threadIdx.x;
  const bool active = tid < kActive;
  const int ug = active ? tid / 10 : 0, tg = active ? tid % 10 : 0;
  const int tiles = (p.B + kTT - 1) / kTT;
  float* part = p.partial + (size_t)blockIdx.x * kPartialFloats;
  float* dsq = p.dseq + (size_t)blockIdx.x * kDseqFloatsCta + tid * 4;
  bool first = true;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, first = false) {
    const int b0 = tile * kTT;
    const float* hs_tile = p.hseq + hseq_offset(tile);                // rows of kTF floats, this tile's 40 samples first
    const float* rec_tile = p.rec + tid * 4;
    __syncthreads();
    if (tid < 4 * kTT) {
      const int b = tid % kTT, o = tid / kTT;
      dout[o * kTT + b] = (b0 + b < p.B) ? __ldg(p.d_out + (size_t)(b0 + b) * 4 + o) : 0.f;
    }
    __syncthreads();
    if (tid < 204) {                                               // d fc.weight [4][50], d fc.bias [4]
      float a = 0.f;
      if (tid < 200) {
        const int o = tid / kH, u = tid % kH;
        const float* h = hs_tile + (size_t)((2 * kL + 9) * kH + u) * kTF;
        for (int b = 0; b < kTT; ++b) a = fmaf(dout[o * kTT + b], __ldcg(h + b), a);
      } else {
        for (int b = 0; b < kTT; ++b) a += dout[(tid - 200) * kTT + b];
      }
      float* dst = part + 3 * kG * 100 + tid;
      *dst = first ? a : *dst + a;
    }
    for (int l = 2; l >= 0; --l) {
      __syncthreads();                                             // previous layer's reads of WB / dG / act are done
      {
        const float* src = p.pack + kWfFloats + l * kG * 100;
        for (int i = tid * 4; i < kG * 100; i += kThreadsL * 4) st4(WB + i, __ldg(reinterpret_cast<const float4*>(src + i)));
      }
      float wacc[8][10];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int j = 0; j < 10; ++j) wacc[r][j] = 0.f;
      float dc[8], dhrec[8], cnext[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) dc[e] = dhrec[e] = 0.f;
      if (active) {
        const float* r9 = rec_tile + rec_offset(tile, l * kL + 9);
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) {
          const float4 v = __ldcs(reinterpret_cast<const float4*>(r9 + (8 + uu) * kThreadsL * 4));
          cnext[uu * 4 + 0] = v.x; cnext[uu * 4 + 1] = v.y; cnext[uu * 4 + 2] = v.z; cnext[uu * 4 + 3] = v.w;
        }
      }
      // operands of the weight-gradient product for step tt -> act[tt & 1], one step ahead of their use (cp.async;
      // the buffer written is not the one the current step's products read)
      auto stage = [&](int tt) {
        float* dstb = act + (tt & 1) * 100 * kSB;
        if (l > 0) {
          const float* src = hs_tile + (size_t)((l - 1) * kL + tt) * kH * kTF;
          for (int i = tid * 4; i < kH * kTT; i += kThreadsL * 4)
            DevCtx::cp_async16(dstb + (i / kTT) * kSB + i % kTT, src + (i / kTT) * kTF + i % kTT);
        } else {
          for (int i = tid; i < kH * kTT; i += kThreadsL) {
            const int b = i % kTT, k = i / kTT;
            dstb[k * kSB + b] = (k < 5 && b0 + b < p.B) ? __ldg(p.X + (size_t)(b0 + b) * 50 + tt * 5 + k) : 0.f;
          }
        }
        if (tt > 0) {
          const float* src = hs_tile + (size_t)(l * kL + tt - 1) * kH * kTF;
          for (int i = tid * 4; i < kH * kTT; i += kThreadsL * 4)
            DevCtx::cp_async16(dstb + (kH + i / kTT) * kSB + i % kTT, src + (i / kTT) * kTF + i % kTT);
        } else {
          for (int i = tid * 4; i < kH * kTT; i += kThreadsL * 4)
            st4(dstb + (kH + i / kTT) * kSB + i % kTT, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        DevCtx::cp_commit();
      };
      stage(kL - 1);
      for (int t = kL - 1; t >= 0; --t) {
        float* ab = act + (t & 1) * 100 * kSB;
        if (active) {
          // gradient arriving from above: fc (top layer, last step) or the layer above's d(input)
          float dh[8];
          if (l == 2) {
#pragma unroll
            for (int uu = 0; uu < 2; ++uu)
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float a = 0.f;
                if (t == kL - 1) {
#pragma unroll
                  for (int o = 0; o < 4; ++o) a = fmaf(__ldg(p.fc_w + o * kH + 2 * ug + uu), dout[o * kTT + tg * 4 + j], a);
                }
                dh[uu * 4 + j] = a;
              }
          } else {
#pragma unroll
            for (int uu = 0; uu < 2; ++uu) {
              const float4 v = __ldcg(reinterpret_cast<const float4*>(dsq + (size_t)(t * 2 + uu) * kThreadsL * 4));
              dh[uu * 4 + 0] = v.x; dh[uu * 4 + 1] = v.y; dh[uu * 4 + 2] = v.z; dh[uu * 4 + 3] = v.w;
            }
          }
          const float* r = rec_tile + rec_offset(tile, l * kL + t);
          if (t > 0 && (tid & 7) == 0) {                           // records of the next step (t-1): one 128-byte line per 8 lanes
            const float* rn = r - 2 * kRecCell;
#pragma unroll
            for (int sl = 0; sl < 8; ++sl) asm volatile("prefetch.global.L2 [%0];" ::"l"(rn + sl * kThreadsL * 4));
          }
          float cprev[8];
          if (t > 0) {
            const float* rp = r - 2 * kRecCell;                        // the previous cell of the same layer and half
#pragma unroll
            for (int uu = 0; uu < 2; ++uu) {
              const float4 v = __ldcs(reinterpret_cast<const float4*>(rp + (8 + uu) * kThreadsL * 4));
              cprev[uu * 4 + 0] = v.x; cprev[uu * 4 + 1] = v.y; cprev[uu * 4 + 2] = v.z; cprev[uu * 4 + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) cprev[e] = 0.f;
          }
#pragma unroll
          for (int uu = 0; uu < 2; ++uu) {
            const float4 vi = __ldcs(reinterpret_cast<const float4*>(r + (0 + uu) * kThreadsL * 4));
            const float4 vf = __ldcs(reinterpret_cast<const float4*>(r + (2 + uu) * kThreadsL * 4));
            const float4 vg = __ldcs(reinterpret_cast<const float4*>(r + (4 + uu) * kThreadsL * 4));
            const float4 vo = __ldcs(reinterpret_cast<const float4*>(r + (6 + uu) * kThreadsL * 4));
            const float gi[4] = {vi.x, vi.y, vi.z, vi.w}, gf[4] = {vf.x, vf.y, vf.z, vf.w};
            const float gg[4] = {vg.x, vg.y, vg.z, vg.w}, go[4] = {vo.x, vo.y, vo.z, vo.w};
            float di[4], df[4], dg[4], dov[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = uu * 4 + j;
              const float dht = dh[e] + dhrec[e];
              const float tc = tanh_(cnext[e]);
              const float dct = fmaf(dht * go[j], 1.f - tc * tc, dc[e]);
              di[j] = dct * gg[j] * gi[j] * (1.f - gi[j]);
              df[j] = dct * cprev[e] * gf[j] * (1.f - gf[j]);
              dg[j] = dct * gi[j] * (1.f - gg[j] * gg[j]);
              dov[j] = dht * tc * go[j] * (1.f - go[j]);
              dc[e] = dct * gf[j];
              cnext[e] = cprev[e];
            }
            const int u = 2 * ug + uu;
            st4(dG + (0 * kH + u) * kSB + tg * 4, make_float4(di[0], di[1], di[2], di[3]));
            st4(dG + (1 * kH + u) * kSB + tg * 4, make_float4(df[0], df[1], df[2], df[3]));
            st4(dG + (2 * kH + u) * kSB + tg * 4, make_float4(dg[0], dg[1], dg[2], dg[3]));
            st4(dG + (3 * kH + u) * kSB + tg * 4, make_float4(dov[0], dov[1], dov[2], dov[3]));
          }
        }
        DevCtx::cp_wait<0>();
        __syncthreads();                                           // dG, act (and WB on the first step) are visible
        if (t > 0) stage(t - 1);
        if (active) {
          // data gradients: d(input units 2ug, 2ug+1) and d(h_{t-1} units 2ug, 2ug+1) of samples 4tg..4tg+3
          float da[4][4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) da[q][j] = 0.f;
          const float* wb = WB + ug * 4;
          const float* dp = dG + tg * 4;
#pragma unroll 8
          for (int row = 0; row < kG; ++row) {
            const float4 w = ld4(wb + row * 100), d = ld4(dp + row * kSB);
            const float wv[4] = {w.x, w.y, w.z, w.w}, dv[4] = {d.x, d.y, d.z, d.w};
            if (l > 0) {                                           // the bottom layer's inputs are data: no d(input)
#pragma unroll
              for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int j = 0; j < 4; ++j) da[q][j] = fmaf(wv[q], dv[j], da[q][j]);
            }
#pragma unroll
            for (int q = 2; q < 4; ++q)
#pragma unroll
              for (int j = 0; j < 4; ++j) da[q][j] = fmaf(wv[q], dv[j], da[q][j]);
          }
          if (l > 0) {
#pragma unroll
            for (int uu = 0; uu < 2; ++uu)
              *reinterpret_cast<float4*>(dsq + (size_t)(t * 2 + uu) * kThreadsL * 4) =
                  make_float4(da[uu][0], da[uu][1], da[uu][2], da[uu][3]);
          }
#pragma unroll
          for (int uu = 0; uu < 2; ++uu)
#pragma unroll
            for (int j = 0; j < 4; ++j) dhrec[uu * 4 + j] = da[2 + uu][j];
          // weight gradients: rows ug + 25 r8 (r8 < 8) x columns tg + 10 j (j < 10) of [dW_ih | dW_hh], reduction over
          // the 40 samples (interleaved ownership: conflict-free float4 reads with the 44-float row stride)
          const float* gp = dG + ug * kSB;
          const float* cp = ab + tg * kSB;
#pragma unroll 1
          for (int bq = 0; bq < kTT; bq += 4) {
            float4 d4[8];
#pragma unroll
            for (int r8 = 0; r8 < 8; ++r8) d4[r8] = ld4(gp + r8 * 25 * kSB + bq);
#pragma unroll
            for (int j = 0; j < 10; ++j) {
              if (l == 0 && j >= 1 && j <= 4) continue;            // bottom layer: input columns 5..49 do not exist
              const float4 a = ld4(cp + j * 10 * kSB + bq);
#pragma unroll
              for (int r8 = 0; r8 < 8; ++r8) {
                float s = wacc[r8][j];
                s = fmaf(d4[r8].x, a.x, s);
                s = fmaf(d4[r8].y, a.y, s);
                s = fmaf(d4[r8].z, a.z, s);
                s = fmaf(d4[r8].w, a.w, s);
                wacc[r8][j] = s;
              }
            }
          }
        }
        __syncthreads();                                           // dG and act may be overwritten
      }
      if (active) {
        float* dst = part + (size_t)l * kG * 100 + (size_t)ug * 100 + tg;
#pragma unroll
        for (int r8 = 0; r8 < 8; ++r8)
#pragma unroll
          for (int j = 0; j < 10; ++j) {
            float* q = dst + r8 * 2500 + j * 10;
            *q = first ? wacc[r8][j] : *q + wacc[r8][j];
          }
      }
    }
  }
}

// gradient tensors <- sum over CTAs of the partials: one thread per element (consecutive elements read consecutive
// partial entries: coalesced), fp64 accumulation in a fixed order
__global__ void __launch_bounds__(256) lstm_grad_reduce_kernel(const float* __restrict__ partial, int grid, LstmGradOut g) {
  constexpr int kOut = 200 * 5 + 5 * 200 * 50 + 200 + 4;         // 51 204
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kOut) return;
  int src;
  float* dst;
  if (i < 1000) {                                                 // weight_ih_l0 [200][5]
    src = (i / 5) * 100 + (i % 5);
    dst = g.g_ih[0] + i;
  } else if (i < 51000) {
    const int r = i - 1000, blk = r / 10000, q = r % 10000;        // hh0, ih1, hh1, ih2, hh2
    const int l = (blk + 1) >> 1;
    const bool hh = (blk & 1) == 0;
    src = l * kG * 100 + (q / 50) * 100 + (hh ? 50 : 0) + (q % 50);
    dst = (hh ? g.g_hh[l] : g.g_ih[l]) + q;
  } else if (i < 51200) {
    src = 3 * kG * 100 + (i - 51000);
    dst = g.g_fc_w + (i - 51000);
  } else {
    src = 3 * kG * 100 + 200 + (i - 51200);
    dst = g.g_fc_b + (i - 51200);
  }
  double a0 = 0.0, a1 = 0.0;
  const float* pp = partial + src;
  int b = 0;
  for (; b + 1 < grid; b += 2) {
    a0 += (double)__ldcg(pp + (size_t)b * kPartialFloats);
    a1 += (double)__ldcg(pp + (size_t)(b + 1) * kPartialFloats);
  }
  if (b < grid) a0 += (double)__ldcg(pp + (size_t)b * kPartialFloats);
  *dst = (float)(a0 + a1);
}

// ---------------------------------------------------------------------------------------------------
// torch.optim.AdamW.step for up to 8 tensors in one launch
// ---------------------------------------------------------------------------------------------------
struct AdamWParams {
  float* p[8];
  const float* g[8];
  float* m[8];
  float* v[8];
  int n[8];
  int count;
  float lr, beta1, beta2, eps, weight_decay, bc1, sqrt_bc2, grad_scale;
};

__global__ void __launch_bounds__(256) adamw_kernel(const AdamWParams a) {
  for (int k = 0; k < a.count; ++k) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n[k]; i += gridDim.x * blockDim.x) {
      const float g = a.g[k][i] * a.grad_scale;
      float p = a.p[k][i];
      p = p * (1.f - a.lr * a.weight_decay);                       // decoupled weight decay (torch: param.mul_(1 - lr*wd))
      const float m = a.m[k][i] + (g - a.m[k][i]) * (1.f - a.beta1);   // exp_avg.lerp_(grad, 1 - beta1)
      const float v = a.beta2 * a.v[k][i] + (1.f - a.beta2) * g * g;
      const float denom = sqrtf(v) / a.sqrt_bc2 + a.eps;          // sqrt(v) / sqrt(1 - beta2^t) + eps
      p = p - (a.lr / a.bc1) * (m / denom);
      a.p[k][i] = p;
      a.m[k][i] = m;
      a.v[k][i] = v;
    }
  }
}

}  // namespace lt
}  // namespace fc
