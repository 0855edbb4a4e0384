// Fused MPC-loss forward + reverse-time sweep for one persistent CTA.
//
// Restates MPCLoss.forward (Unsupervised Learning/Functions.py:1353-1472), LSTMModel.forward
// (:353-379), FNNModel.forward (:261-289) and the gradients loss.backward() (:655) leaves on the
// live controller parameters and on output_controller.  The body is written against a tiny
// execution-context interface (thread id, block barrier, cp.async, fast math) so that the very same
// source is compiled (a) by nvcc into the sm_100a kernel and (b) by g++ into the CPU thread
// emulation used by the CPU test-suite to check the index arithmetic (tests/emu).
//
// Thread mapping (see fc_layout.h): lane = tg*5+cgl; thread owns trajectories tb*5+j (tb=(warp/2)*6+tg,
// j<5) and hidden units cg*5+uu (cg=(warp%2)*5+cgl, uu<5): 25 (trajectory, unit) elements, all four gates of each.
//   forward GEMM   acc[5 traj][20 gate cols]  over K = 5|50 (+50 recurrent) with FP32 FFMA
//   backward GEMM  acc[5 traj][10 cols]       over the 200 gate gradients
// Cell state c / dc and the recurrent dh stay in registers; h and dGates are exchanged through
// shared memory; cell activations (i,f,g,o,c_prev) go to HBM in thread-private float4 slots.
#pragma once
#include "fc_layout.h"

namespace fc {

#if defined(__CUDACC__)
using F4 = float4;
#else
struct alignas(16) F4 { float x, y, z, w; };
#endif

template <class Ctx>
struct MpcTile {
  Ctx& ctx;
  const MpcParams& p;
  float* sm;
  // thread coordinates
  int tid, warp, lane, tg, cg, tb;
  bool act;            // compute lane (lane < 30)
  // workspace pointers of this CTA
  float *rows, *seq, *dseq, *grow, *rec;
  // per-thread state
  float c[kElems];     // forward: cell state; backward: d(cell state)
  float hrec[kElems];  // backward: d(h) arriving from step t+1 of the same layer

  FC_HD_CTX MpcTile(Ctx& c_, const MpcParams& p_) : ctx(c_), p(p_) {
    sm = ctx.smem();
    tid = ctx.tid();
    warp = tid >> 5;
    lane = tid & 31;
    act = lane < 30;
    int l2 = act ? lane : lane - 30;   // idle lanes shadow lanes 0,1 (loads only, never store)
    // warp = 6 trajectory blocks x 5 unit groups (4 x 2 warps cover 24 x 10): per k-step a warp reads
    // 30 activations and 100 weights.  Measured on B200 (scripts/micro/lds_bench.cu): an LDS.128 whose
    // lanes touch <= 8 distinct 16-byte chunks without bank overlap costs ~2 SM cycles, the previous
    // 3 x 10 arrangement (10 chunks at an 80-byte stride) 3.5 cycles and made the kernel smem-bound.
    tg = l2 / 5;                       // trajectory block within the warp, 0..5
    cg = (warp & 1) * 5 + (l2 - tg * 5);
    tb = (warp >> 1) * 6 + tg;
    WorkLayout wl = work_layout(p.N, p.with_grad);
    float* base = p.work + (size_t)ctx.bid() * p.work_stride;
    rows = base + wl.rows;
    seq = base + wl.seq;
    dseq = base + wl.dseq;
    grow = base + wl.grow;
    rec = base + wl.rec;
  }

  // ---------------------------------------------------------------------------------------------
  // helpers
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX static float sigmoidf_(float x) { return Ctx::rcp(1.f + Ctx::ex2(-1.4426950408889634f * x)); }
  // tanh: 1 - 2/(1+e^{2x}) has an ABSOLUTE error of ~1e-7 (cancellation against 1), which is a large
  // relative error for the small gate / cell values that dominate here; below |x| = 0.3 use the odd
  // Taylor polynomial (relative error < 2e-8) instead.  Branch-free select.
  FC_HD_CTX static float tanhf_(float x) {
    const float big = 1.f - 2.f * Ctx::rcp(1.f + Ctx::ex2(2.8853900817779268f * x));
    const float x2 = x * x;
    float pl = fmaf(x2, 0.021869488536155203f, -0.053968253968253971f);
    pl = fmaf(x2, pl, 0.13333333333333333f);
    pl = fmaf(x2, pl, -0.33333333333333331f);
    pl = fmaf(x2 * x, pl, x);
    return fabsf(x) < 0.3f ? pl : big;
  }

  // cooperative global -> shared copy of n floats (n % 4 == 0, 16-byte aligned both sides)
  FC_HD_CTX void copy_async(float* dst, const float* src, int n) {
    for (int i = tid * 4; i < n; i += kThreads * 4) Ctx::cp_async16(dst + i, src + i);
  }

  // ---------------------------------------------------------------------------------------------
  // forward GEMM: acc[j][c] += A[k][traj j] * W[k][cg*20 + c]
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void fwd_fma_rows(float (&acc)[5][20], const float* A, const float* W, int nk) {
    const float* ap = A + tb * 4;
    const float* a1p = A + 96 + tb;
    const float* wp = W + cg * 20;
#pragma unroll 2
    for (int k = 0; k < nk; ++k) {
      F4 a4 = Ctx::lds4(ap + k * kTile);
      float a5 = a1p[k * kTile];
      float w[20];
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        F4 t = Ctx::lds4(wp + k * kGates + i * 4);
        w[i * 4 + 0] = t.x; w[i * 4 + 1] = t.y; w[i * 4 + 2] = t.z; w[i * 4 + 3] = t.w;
      }
      float a[5] = {a4.x, a4.y, a4.z, a4.w, a5};
#pragma unroll
      for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int cc = 0; cc < 20; ++cc) acc[j][cc] = fmaf(a[j], w[cc], acc[j][cc]);
    }
  }

  // ---------------------------------------------------------------------------------------------
  // forward cell update for the 25 owned elements; writes h to shared (and to the sequence
  // scratch for the layer above) and the activation record to HBM
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void fwd_pointwise(float (&acc)[5][20], bool first, float* ah_out, float* seq_out, float* rec_out) {
    // element index e = uu*5 + j (unit-major) for c[], hrec[], records and dseq slots
    float rv[128];
    float* rp = rec_out ? rec_out + ((size_t)warp * 32 * 32 + lane) * 4 : nullptr;
#pragma unroll
    for (int uu = 0; uu < 5; ++uu) {
      float h[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int e = uu * 5 + j;
        float gi = sigmoidf_(acc[j][uu]);
        float gf = sigmoidf_(acc[j][5 + uu]);
        float gg = tanhf_(acc[j][10 + uu]);
        float go = sigmoidf_(acc[j][15 + uu]);
        float cp = first ? 0.f : c[e];
        float cn = fmaf(gf, cp, gi * gg);
        c[e] = cn;
        h[j] = go * tanhf_(cn);
        rv[e * 5 + 0] = gi; rv[e * 5 + 1] = gf; rv[e * 5 + 2] = gg; rv[e * 5 + 3] = go; rv[e * 5 + 4] = cp;
      }
      if (uu == 4) rv[125] = rv[126] = rv[127] = 0.f;
      if (act) {
        const int unit = cg * 5 + uu;
        F4 v = {h[0], h[1], h[2], h[3]};
        Ctx::sts4(ah_out + unit * kTile + tb * 4, v);
        ah_out[unit * kTile + 96 + tb] = h[4];
        if (seq_out) {
          Ctx::stg4(seq_out + unit * kTile + tb * 4, v);
          seq_out[unit * kTile + 96 + tb] = h[4];
        }
        if (rp) {
          // floats [0, 25*(uu+1)) are final: flush the float4 slots that are complete
          const int r_lo = (25 * uu) / 4, r_hi = uu == 4 ? 32 : (25 * (uu + 1)) / 4;
#pragma unroll
          for (int r = r_lo; r < r_hi; ++r) {
            F4 w4 = {rv[r * 4], rv[r * 4 + 1], rv[r * 4 + 2], rv[r * 4 + 3]};
            Ctx::stg4_stream(rp + (size_t)r * 32 * 4, w4);
          }
        }
      }
    }
  }

  // ---------------------------------------------------------------------------------------------
  // tile set-up: reference, commands and the ten recorded rows
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void load_tile(int tile) {
    const int b0 = tile * kTile;
    for (int i = tid; i < kTile * kLook * kFeat; i += kThreads) {
      int traj = i / (kLook * kFeat), rem = i - traj * (kLook * kFeat);
      int r = rem / kFeat, f = rem - r * kFeat;
      int b = b0 + traj;
      float v = b < p.B ? p.Z[(size_t)b * (kLook * kFeat) + rem] : 0.f;
      if (r == kLook - 1 && f == kFeat - 1) v = b < p.B ? p.u0[b] : 0.f;   // Functions.py:1396
      rows[(r * kFeat + f) * kTile + plane_of(traj)] = v;
    }
    if (tid < kTile) {
      int b = b0 + tid;
      bool ok = b < p.B;
      sm[kSmRef + tid] = ok ? p.X[(size_t)b * 3 + 2] : 0.f;                     // Functions.py:1392
      sm[kSmUcur + tid] = ok ? p.u0[b] : 0.f;
      sm[kSmUprev + tid] = ok ? p.Z[(size_t)b * (kLook * kFeat) + (kLook - 2) * kFeat + 4] : 0.f;
      sm[kSmCost + tid] = 0.f;
      sm[kSmCost + kTile + tid] = 0.f;
      sm[kSmCost + 2 * kTile + tid] = 0.f;
      if (ok) p.pred[(size_t)b * p.N] = p.u0[b];                               // Functions.py:1417-1418
    }
  }

  // ---------------------------------------------------------------------------------------------
  // after window m: read-out, cost terms, next command (one thread per trajectory)
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void fwd_glue(int tile, int m, const float* ah_last) {
    if (tid < kTile) {
      const int traj = tid, pl = plane_of(traj);
      const float* sw = sm + kSmSmall;
      float x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = sw[(kFCB - kFCW) + q];
      for (int u = 0; u < kHid; ++u) {
        float hv = ah_last[u * kTile + pl];
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = fmaf(sw[q * kHid + u], hv, x[q]);
      }
      if (p.noise_std > 0.f) {                                                 // enable_noise, :1400-1402 / :1438-1440
        float e[4];
        philox_normal4(p.noise_seed, (unsigned)(tile * kTile + traj), (unsigned)m, e);
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = fmaf(p.noise_std, e[q], x[q]);
      }
      const float ref = sm[kSmRef + traj];
      const float ucur = sm[kSmUcur + traj], uprev = sm[kSmUprev + traj];
      float du = uprev - ucur;
      float cmd = p.alpha * du * du;                                           // :1405 / :1446
      float er = (x[0] - ref) * (x[0] - ref);                                  // :1408 / :1443
      float con = fmaxf(-x[1], 0.f) + fmaxf(-x[2], 0.f) + fmaxf(x[1] - kP1Max, 0.f) + fmaxf(x[2] - kP2Max, 0.f);
      sm[kSmCost + traj] += (er + cmd) + con;                                  // :1414 / :1452
      sm[kSmCost + kTile + traj] += cmd;
      sm[kSmCost + 2 * kTile + traj] += er;
      float* rnew = rows + (size_t)(kLook + m) * kFeat * kTile;                // rho_{10+m} = [x_{m+1}, u_{m+1}]
#pragma unroll
      for (int q = 0; q < 4; ++q) rnew[q * kTile + pl] = x[q];
      float unext = 0.f;
      if (m + 1 < p.N) {                                                       // :1424-1430
        const float* iw = sw + (kINPW - kFCW);
        const float* ib = sw + (kINPB - kFCW);
        const float* ow = sw + (kOUTW - kFCW);
        float v = 0.f;
        for (int u = 0; u < kFnnHid; ++u) {
          float pre = fmaf(iw[u * 3 + 2], ref, fmaf(iw[u * 3 + 1], x[3], fmaf(iw[u * 3 + 0], x[0], ib[u])));
          v = fmaf(ow[u], fmaxf(pre, 0.f), v);
        }
        unext = fminf(fmaxf(v, -1.f), 1.f);                                    // nn.Hardtanh
        sm[kSmUprev + traj] = ucur;
        sm[kSmUcur + traj] = unext;
        int b = tile * kTile + traj;
        if (b < p.B) p.pred[(size_t)b * p.N + m + 1] = unext;                  // :1455
      }
      rnew[4 * kTile + pl] = unext;
    }
  }

  // ---------------------------------------------------------------------------------------------
  // forward of one window (three layers, ten steps)
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void fwd_window(int tile, int m) {
    float* W = sm + kSmFwdW;
    float* ain = sm + kSmFwdAin;
    float* ah = sm + kSmFwdAh;
    const int tmin = t_min_of(m);
    // entry invariant: weights of layer 0 and the input of (layer 0, t = 0) are in flight as the
    // two most recent cp.async groups [ih + input] , [hh]; see prefetch_layer().
    for (int l = 0; l < kLayers; ++l) {
      const int kin = l == 0 ? kKin0 : kKin;
      for (int t = 0; t < kLook; ++t) {
        float* ain_t = ain + (t & 1) * kHid * kTile;
        float* ah_prev = ah + ((t + 1) & 1) * kHid * kTile;
        float* ah_t = ah + (t & 1) * kHid * kTile;
        if (t == 0) { Ctx::template cp_wait<1>(); } else { Ctx::template cp_wait<0>(); }
        ctx.sync();                                   // inputs of step t (and h_{t-1}) visible
        if (t + 1 < kLook) {                          // prefetch the input rows of step t+1
          float* dst = ain + ((t + 1) & 1) * kHid * kTile;
          const float* src = l == 0 ? rows + (size_t)(m + t + 1) * kFeat * kTile : seq + (size_t)(t + 1) * kHid * kTile;
          copy_async(dst, src, kin * kTile);
          Ctx::cp_commit();
        }
        float acc[5][20];
#pragma unroll
        for (int j = 0; j < 5; ++j)
#pragma unroll
          for (int cc = 0; cc < 20; ++cc) acc[j][cc] = 0.f;
        fwd_fma_rows(acc, ain_t, W, kin);
        if (t > 0) fwd_fma_rows(acc, ah_prev, W + kin * kGates, kHid);
        if (t == kLook - 1) {
          // all GEMM reads of this layer's weights are done after this barrier: start streaming
          // the next layer's (or next window's layer-0) weights under the cell update below.
          ctx.sync();
          prefetch_layer(tile, m, l + 1);
        }
        float* rec_out = nullptr;
        if (p.with_grad && t >= tmin) rec_out = rec + (size_t)(rec_base(m) + (long)l * steps_kept(m) + (t - tmin)) * kRecFloats;
        fwd_pointwise(acc, t == 0, ah_t, l + 1 < kLayers ? seq + (size_t)t * kHid * kTile : nullptr, rec_out);
      }
    }
    Ctx::template cp_wait<1>();
    ctx.sync();                                       // h of (layer 2, t = 9) visible
    fwd_glue(tile, m, ah + ((kLook - 1) & 1) * kHid * kTile);
    ctx.sync();
  }

  // issue the weight loads of layer l of window m (l == 3 -> layer 0 of window m+1) together with
  // the input rows of its first step as two cp.async groups: [W_ih + input(t=0)], [W_hh]
  FC_HD_CTX void prefetch_layer(int tile, int m, int l) {
    (void)tile;
    if (l == kLayers) { l = 0; m += 1; }
    if (m >= p.N) {                                   // nothing follows: keep the group count uniform
      Ctx::cp_commit();
      Ctx::cp_commit();
      return;
    }
    float* W = sm + kSmFwdW;
    float* ain = sm + kSmFwdAin;
    const int kin = l == 0 ? kKin0 : kKin;
    const float* wsrc = p.wpack + wf_offset(l);
    copy_async(W, wsrc, kin * kGates);
    const float* src = l == 0 ? rows + (size_t)m * kFeat * kTile : seq;
    copy_async(ain, src, kin * kTile);                // step 0 uses buffer 0
    Ctx::cp_commit();
    copy_async(W + kin * kGates, wsrc + kin * kGates, kHid * kGates);
    Ctx::cp_commit();
  }

  // ---------------------------------------------------------------------------------------------
  // backward GEMM: acc[j][c] += dG[g][traj j] * WB[g][cg][c]
  // ---------------------------------------------------------------------------------------------
  template <int CW, int NC>
  FC_HD_CTX void bwd_gemm(float (&acc)[5][NC], const float* G, const float* W) {
    const float* ap = G + tb * 4;
    const float* a1p = G + 96 + tb;
    const float* wp = W + cg * CW;
    constexpr int stride = CW * 10;
#pragma unroll 2
    for (int g = 0; g < kGates; ++g) {
      F4 a4 = Ctx::lds4(ap + g * kTile);
      float a5 = a1p[g * kTile];
      float w[CW];
#pragma unroll
      for (int i = 0; i < (NC + 3) / 4; ++i) {
        F4 t = Ctx::lds4(wp + g * stride + i * 4);
        w[i * 4 + 0] = t.x; w[i * 4 + 1] = t.y; w[i * 4 + 2] = t.z; w[i * 4 + 3] = t.w;
      }
      float a[5] = {a4.x, a4.y, a4.z, a4.w, a5};
#pragma unroll
      for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) acc[j][cc] = fmaf(a[j], w[cc], acc[j][cc]);
    }
  }

  // gate gradients of the 25 owned elements -> shared dG[g'][traj]
  FC_HD_CTX void bwd_pointwise(int l, int t, const float* rec_in, const float* dseq_in, float* G) {
    const float* rp = rec_in + ((size_t)warp * 32 * 32 + lane) * 4;
    float rv[128];
    float gxv[4][5];
    const bool top = l == kLayers - 1;
    if (top && t == kLook - 1) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        F4 v = Ctx::lds4(sm + kSmGx + q * kTile + tb * 4);
        gxv[q][0] = v.x; gxv[q][1] = v.y; gxv[q][2] = v.z; gxv[q][3] = v.w;
        gxv[q][4] = sm[kSmGx + q * kTile + 96 + tb];
      }
    }
#pragma unroll
    for (int uu = 0; uu < 5; ++uu) {
      // record floats [25uu, 25uu+25): float4 slots (25uu)/4 .. (25uu+24)/4 (first one may be loaded already)
      const int r_lo = uu == 0 ? 0 : (25 * uu - 1) / 4 + 1, r_hi = (25 * uu + 24) / 4 + 1;
#pragma unroll
      for (int r = r_lo; r < r_hi; ++r) {
        F4 v = Ctx::ldg4_stream(rp + (size_t)r * 32 * 4);
        rv[r * 4] = v.x; rv[r * 4 + 1] = v.y; rv[r * 4 + 2] = v.z; rv[r * 4 + 3] = v.w;
      }
      float dgate[4][5];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int e = uu * 5 + j;
        float gi = rv[e * 5 + 0], gf = rv[e * 5 + 1], gg = rv[e * 5 + 2], go = rv[e * 5 + 3], cp = rv[e * 5 + 4];
        float cn = fmaf(gf, cp, gi * gg);
        float tc = tanhf_(cn);
        float dh = hrec[e];
        if (top) {
          if (t == kLook - 1) {                        // through fc (Functions.py:377)
            const float* fw = sm + kSmSmall + cg * 5 + uu;
            dh += fw[0] * gxv[0][j] + fw[kHid] * gxv[1][j] + fw[2 * kHid] * gxv[2][j] + fw[3 * kHid] * gxv[3][j];
          }
        } else {
          dh += Ctx::ldcg(dseq_in + ((size_t)warp * kElems + e) * 32 + lane);
        }
        float dout = dh * tc;
        float dct = fmaf(dh * go, 1.f - tc * tc, c[e]);
        float di = dct * gg, dg = dct * gi, df = dct * cp;
        c[e] = dct * gf;
        dgate[0][j] = di * gi * (1.f - gi);
        dgate[1][j] = df * gf * (1.f - gf);
        dgate[2][j] = dg * (1.f - gg * gg);
        dgate[3][j] = dout * go * (1.f - go);
      }
      if (act) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int g = cg * 20 + q * 5 + uu;
          F4 v = {dgate[q][0], dgate[q][1], dgate[q][2], dgate[q][3]};
          Ctx::sts4(G + g * kTile + tb * 4, v);
          G[g * kTile + 96 + tb] = dgate[q][4];
        }
      }
    }
  }

  // ---------------------------------------------------------------------------------------------
  // before the reverse sweep of window m: seed d loss / d x_{m+1}, controller backward for
  // u_{m+1}, controller weight-gradient accumulation
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void bwd_glue(int tile, int m) {
    const int k = m + 1;
    const float s = p.grad_scale;
    const bool has_u = k <= p.N - 1;
    const float* sw = sm + kSmSmall;
    const float* iw = sw + (kINPW - kFCW);
    const float* ib = sw + (kINPB - kFCW);
    const float* ow = sw + (kOUTW - kFCW);
    if (tid < kTile) {
      const int traj = tid, pl = plane_of(traj);
      const bool valid = tile * kTile + traj < p.B;
      const float* rx = rows + (size_t)(kLook + m) * kFeat * kTile;
      float x0 = Ctx::ldcg(rx + 0 * kTile + pl), x1 = Ctx::ldcg(rx + 1 * kTile + pl);
      float x2 = Ctx::ldcg(rx + 2 * kTile + pl), x3 = Ctx::ldcg(rx + 3 * kTile + pl);
      const float ref = sm[kSmRef + traj];
      float g0 = 2.f * (x0 - ref) * s;
      float g1 = s * ((x1 > kP1Max ? 1.f : 0.f) - (x1 < 0.f ? 1.f : 0.f));
      float g2 = s * ((x2 > kP2Max ? 1.f : 0.f) - (x2 < 0.f ? 1.f : 0.f));
      float g3 = 0.f;
      float dv = 0.f;
      if (has_u) {
        const float* gr = grow + (size_t)k * kFeat * kTile;
        float uk = Ctx::ldcg(rows + (size_t)((kLook - 1 + k) * kFeat + 4) * kTile + pl);
        float ukm1 = Ctx::ldcg(rows + (size_t)((kLook - 2 + k) * kFeat + 4) * kTile + pl);
        float gu = Ctx::ldcg(gr + 4 * kTile + pl) - 2.f * p.alpha * (ukm1 - uk) * s;
        if (k + 1 <= p.N - 1) {
          float ukp1 = Ctx::ldcg(rows + (size_t)((kLook + k) * kFeat + 4) * kTile + pl);
          gu += 2.f * p.alpha * (uk - ukp1) * s;
        }
        float v = 0.f;
        for (int u = 0; u < kFnnHid; ++u) {
          float pre = fmaf(iw[u * 3 + 2], ref, fmaf(iw[u * 3 + 1], x3, fmaf(iw[u * 3 + 0], x0, ib[u])));
          v = fmaf(ow[u], fmaxf(pre, 0.f), v);
        }
        dv = (valid && v > -1.f && v < 1.f) ? gu : 0.f;           // hardtanh_backward
        float d0 = 0.f, d1 = 0.f;
        for (int u = 0; u < kFnnHid; ++u) {
          float pre = fmaf(iw[u * 3 + 2], ref, fmaf(iw[u * 3 + 1], x3, fmaf(iw[u * 3 + 0], x0, ib[u])));
          float dp = pre > 0.f ? dv * ow[u] : 0.f;                  // threshold_backward
          d0 = fmaf(dp, iw[u * 3 + 0], d0);
          d1 = fmaf(dp, iw[u * 3 + 1], d1);
        }
        g0 += d0 + Ctx::ldcg(gr + 0 * kTile + pl);
        g1 += Ctx::ldcg(gr + 1 * kTile + pl);
        g2 += Ctx::ldcg(gr + 2 * kTile + pl);
        g3 += d1 + Ctx::ldcg(gr + 3 * kTile + pl);
        sm[kSmDv + traj] = dv;
        sm[kSmFin + traj] = x0;
        sm[kSmFin + kTile + traj] = x3;
      }
      if (!valid) { g0 = g1 = g2 = g3 = 0.f; }
      sm[kSmGx + 0 * kTile + pl] = g0;
      sm[kSmGx + 1 * kTile + pl] = g1;
      sm[kSmGx + 2 * kTile + pl] = g2;
      sm[kSmGx + 3 * kTile + pl] = g3;
    }
    ctx.sync();
    if (has_u && tid < 4 * kFnnHid) {                  // controller weight gradients, unit-parallel
      const int u = tid % kFnnHid, part = tid / kFnnHid;
      double a_ow = 0.0, a_b = 0.0, a_w0 = 0.0, a_w1 = 0.0, a_w2 = 0.0;   // batch sums cancel heavily: fp64
      const float w0 = iw[u * 3 + 0], w1 = iw[u * 3 + 1], w2 = iw[u * 3 + 2], bb = ib[u], owu = ow[u];
      for (int traj = part * 30; traj < part * 30 + 30; ++traj) {
        float dv = sm[kSmDv + traj];
        float x0 = sm[kSmFin + traj], x3 = sm[kSmFin + kTile + traj], ref = sm[kSmRef + traj];
        float pre = fmaf(w2, ref, fmaf(w1, x3, fmaf(w0, x0, bb)));
        a_ow += (double)dv * (double)fmaxf(pre, 0.f);
        float dp = pre > 0.f ? dv * owu : 0.f;
        a_b += dp;
        a_w0 += (double)dp * (double)x0;
        a_w1 += (double)dp * (double)x3;
        a_w2 += (double)dp * (double)ref;
      }
      double* pg = reinterpret_cast<double*>(sm + kSmPg) + part * kNumFnnGrad;
      pg[u * 3 + 0] += a_w0;
      pg[u * 3 + 1] += a_w1;
      pg[u * 3 + 2] += a_w2;
      pg[150 + u] += a_b;
      pg[200 + u] += a_ow;
    }
  }

  // ---------------------------------------------------------------------------------------------
  // reverse sweep of window m
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void bwd_window(int tile, int m) {
    float* W = sm + kSmBwdW;
    float* G = sm + kSmBwdG;
    const int tmin = t_min_of(m);
    bwd_glue(tile, m);                                // ends after a barrier-protected smem write phase
    for (int l = kLayers - 1; l >= 0; --l) {
      ctx.sync();                                     // previous GEMM done with W and G
      const int wn = kGates * kWBStride;
      copy_async(W, p.wpack + wb_offset(l), wn);
      Ctx::cp_commit();
#pragma unroll
      for (int e = 0; e < kElems; ++e) { c[e] = 0.f; hrec[e] = 0.f; }
      for (int t = kLook - 1; t >= tmin; --t) {
        const float* rec_in = rec + (size_t)(rec_base(m) + (long)l * steps_kept(m) + (t - tmin)) * kRecFloats;
        const float* dseq_in = dseq + (size_t)t * kWarps * kElems * 32;
        bwd_pointwise(l, t, rec_in, dseq_in, G);
        Ctx::template cp_wait<0>();
        ctx.sync();                                   // dG (and on the first step W) visible
        if (l > 0) {
          float acc[5][10];
#pragma unroll
          for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int cc = 0; cc < 10; ++cc) acc[j][cc] = 0.f;
          bwd_gemm<12, 10>(acc, G, W);
          if (act) {
            float* dq = dseq + (size_t)t * kWarps * kElems * 32;
#pragma unroll
            for (int e = 0; e < kElems; ++e) dq[((size_t)warp * kElems + e) * 32 + lane] = acc[e % 5][e / 5];
          }
#pragma unroll
          for (int e = 0; e < kElems; ++e) hrec[e] = acc[e % 5][5 + e / 5];
        } else {
          float acc[5][6];
#pragma unroll
          for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) acc[j][cc] = 0.f;
          bwd_gemm<12, 6>(acc, G, W);
#pragma unroll
          for (int e = 0; e < kElems; ++e) hrec[e] = acc[e % 5][e / 5];
          const int kr = m + t - (kLook - 1);         // gradient of row rho_{9+kr}
          if (act && cg < kFeat && kr >= 0) {
            float* gp = grow + (size_t)(kr * kFeat + cg) * kTile;
            F4 v = Ctx::ldg4(gp + tb * 4);
            v.x += acc[0][5]; v.y += acc[1][5]; v.z += acc[2][5]; v.w += acc[3][5];
            Ctx::stg4(gp + tb * 4, v);
            gp[96 + tb] += acc[4][5];
          }
        }
        ctx.sync();                                   // GEMM done with dG before it is rewritten
      }
    }
  }

  // ---------------------------------------------------------------------------------------------
  // per-tile epilogues
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void store_costs(int tile) {
    if (tid < kTile) {
      int b = tile * kTile + tid;
      if (b < p.B) {
        const float inv = 1.f / (float)p.N;
        float cst = sm[kSmCost + tid] * inv;                                   // :1458-1460
        p.cost[b] = cst;
        p.command[b] = sm[kSmCost + kTile + tid] * inv;
        p.error[b] = sm[kSmCost + 2 * kTile + tid] * inv;
        sm[kSmCost + tid] = cst;
      } else {
        sm[kSmCost + tid] = 0.f;
      }
    }
    ctx.sync();
    if (tid == 0) {
      double acc = 0.0;
      for (int i = 0; i < kTile; ++i) acc += (double)sm[kSmCost + i];
      *reinterpret_cast<double*>(sm + kSmRed) += acc;
    }
  }

  FC_HD_CTX void store_du0(int tile) {
    if (tid < kTile) {
      const int traj = tid, pl = plane_of(traj);
      int b = tile * kTile + traj;
      if (b < p.B) {
        const float s = p.grad_scale;
        float u0 = Ctx::ldcg(rows + (size_t)((kLook - 1) * kFeat + 4) * kTile + pl);
        float um1 = Ctx::ldcg(rows + (size_t)((kLook - 2) * kFeat + 4) * kTile + pl);
        float g = Ctx::ldcg(grow + 4 * kTile + pl) - 2.f * p.alpha * (um1 - u0) * s;
        if (p.N > 1) {
          float u1 = Ctx::ldcg(rows + (size_t)(kLook * kFeat + 4) * kTile + pl);
          g += 2.f * p.alpha * (u0 - u1) * s;
        }
        p.du0[b] = g;
      }
    }
  }

  // ---------------------------------------------------------------------------------------------
  // persistent loop over tiles
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void run() {
    copy_async(sm + kSmSmall, p.wpack + kFCW, kSmallFloats);
    Ctx::cp_commit();
    for (int i = tid; i < 4 * kNumFnnGrad; i += kThreads) reinterpret_cast<double*>(sm + kSmPg)[i] = 0.0;
    if (tid == 0) *reinterpret_cast<double*>(sm + kSmRed) = 0.0;
    Ctx::template cp_wait<0>();
    ctx.sync();
    for (int tile = ctx.bid(); tile < p.num_tiles; tile += ctx.nblk()) {
      load_tile(tile);
      ctx.sync();
      prefetch_layer(tile, -1, kLayers);              // layer 0 of window 0
      for (int m = 0; m < p.N; ++m) fwd_window(tile, m);
      Ctx::template cp_wait<0>();
      store_costs(tile);
      if (p.with_grad) {
        for (size_t i = tid; i < (size_t)p.N * kFeat * kTile; i += kThreads) grow[i] = 0.f;
        ctx.sync();
        for (int m = p.N - 1; m >= 0; --m) bwd_window(tile, m);
        ctx.sync();
        store_du0(tile);
      }
      ctx.sync();
    }
    // per-CTA partial results
    double* part = p.partial + (size_t)ctx.bid() * kPartialStride;
    const double* pgd = reinterpret_cast<const double*>(sm + kSmPg);
    for (int i = tid; i < kNumFnnGrad; i += kThreads)
      part[i] = (pgd[i] + pgd[kNumFnnGrad + i]) + (pgd[2 * kNumFnnGrad + i] + pgd[3 * kNumFnnGrad + i]);
    if (tid == 0) part[kNumFnnGrad] = *reinterpret_cast<const double*>(sm + kSmRed);
  }
};

}  // namespace fc
