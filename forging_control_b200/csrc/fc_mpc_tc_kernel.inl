// tcgen05 variant of the fused MPC-loss forward + reverse sweep (fp16 hi/lo split operands).  Same mathematics and the
// same reference citations as fc_mpc_kernel.inl (MPCLoss.forward Functions.py:1353-1472,
// LSTMModel.forward :353-379, FNNModel.forward :261-289, loss.backward() :655); what changes is where the
// 99.9 % of the FLOPs run: the per-step gate contraction [128 trajectories x K] x [K x 200 gates] is a
// real dense GEMM, so it goes to the 5th-generation tensor cores:
//   * weights: fp16 hi/lo images (x 2^11) resident in shared memory (B operand, K-major, no swizzle)
//   * activations (A operand, fp16 hi/lo of x 2^10, two K-elements per 32-bit column) and the fp32
//     accumulator live in TMEM, one trajectory per lane
//   * error-compensated split  a*b ~ a_lo*b_hi + a_hi*b_lo + a_hi*b_hi  (2 x 11 bits per operand: same error
//     as an fp32 FMA chain; kind::f16 takes K=16 per instruction, half the instructions of 3xTF32)
//   * the thread that owns TMEM lane r does the cell update of trajectory r: no cross-thread exchange
// Written against the same execution-context interface as the FFMA kernel (+ TMEM / MMA / mbarrier
// operations) so that g++ can compile it into the CPU thread emulation of tests/emu.
#pragma once
#include "fc_tc_layout.h"

namespace fc {
namespace tc {

template <class Ctx>
struct MpcTileTC {
  Ctx& ctx;
  const MpcParams& p;
  float* sm;
  int tid, warp, lane, row, quarter, nown, u_first;
  bool full;             // this quarter owns 13 units (else 12: unit slot 12 is a masked dummy)
  float *rows, *seq, *dseq, *grow, *rec, *actg;
  float c[kMaxOwn];      // forward: cell state, backward: d(cell state)
  float hrec[kMaxOwn];   // backward: d(h) from step t+1
  unsigned ph[8];        // completed phases per mbarrier
#ifdef FC_TC_TIMING
  long long tm[16], tlast;  // cycle breakdown of thread 0 (development aid, -DFC_TC_TIMING)
#endif

  FC_HD_CTX MpcTileTC(Ctx& c_, const MpcParams& p_) : ctx(c_), p(p_) {
    sm = ctx.smem();
    tid = ctx.tid();
    warp = tid >> 5;
    lane = tid & 31;
    row = 32 * (warp & 3) + lane;
    quarter = warp >> 2;
    nown = units_of(quarter);
    u_first = first_unit(quarter);
    full = nown == kMaxOwn;
    WorkLayoutTC wl = work_layout_tc(p.N, p.with_grad, p.width_dim);
    float* base = p.work + (size_t)ctx.bid() * p.work_stride;
    rows = base + wl.rows;
    seq = base + wl.seq;
    dseq = base + wl.dseq;
    grow = base + wl.grow;
    rec = base + wl.rec;
    actg = base + wl.act;
#pragma unroll
    for (int i = 0; i < 8; ++i) ph[i] = 0;
#ifdef FC_TC_TIMING
    for (int i = 0; i < 16; ++i) tm[i] = 0;
    tlast = 0;
#endif
  }
#ifdef FC_TC_TIMING
  FC_HD_CTX void lap(int k) { long long t = Ctx::clock(); tm[k] += t - tlast; tlast = t; }
#else
  FC_HD_CTX void lap(int) {}
#endif

  // Activations.  MUFU (ex2 / rcp, 16 lanes per clock and SM) bounds the cell update, so reciprocals are shared:
  // 1/a, 1/b, 1/c, 1/d come from ONE rcp of the product (9 extra multiplies on the FMA pipe instead of 3 MUFU
  // operations).  Every denominator is 1 + 2^e with e clamped to kExpMax, so the product stays below 2^121.
  static constexpr float kExpMax = 30.0f;
  static constexpr float kLog2e = 1.4426950216293335f;            // fp32(log2 e)
  FC_HD_CTX static float denom_(float e2arg) { return 1.f + Ctx::ex2(fminf(e2arg, kExpMax)); }
  FC_HD_CTX static void quad_rcp(float a, float b, float c, float d, float& ra, float& rb, float& rc, float& rd) {
    const float ab = a * b, cd = c * d;
    const float r = Ctx::rcp(ab * cd);
    const float rab = r * cd, rcd = r * ab;
    ra = rab * b; rb = rab * a; rc = rcd * d; rd = rcd * c;
  }
  // tanh(x) = 1 - 2 rd with rd = 1/(1 + e^{2x}) (absolute error ~1e-7, the size of the error its argument already
  // carries; see fc_mpc_pair_kernel.inl)
  FC_HD_CTX static float tanh_from_(float, float rd) { return fmaf(-2.f, rd, 1.f); }
  // tanh of NU values with one reciprocal per four
  template <int NU>
  FC_HD_CTX static void tanh_batch(const float* x, float* y) {
    float d[NU], r[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) d[i] = denom_(2.f * kLog2e * x[i]);
#pragma unroll
    for (int i = 0; i + 3 < NU; i += 4) quad_rcp(d[i], d[i + 1], d[i + 2], d[i + 3], r[i], r[i + 1], r[i + 2], r[i + 3]);
#pragma unroll
    for (int i = NU & ~3; i < NU; ++i) r[i] = Ctx::rcp(d[i]);
#pragma unroll
    for (int i = 0; i < NU; ++i) y[i] = tanh_from_(x[i], r[i]);
  }

  FC_HD_CTX void wait_bar(int b) { ctx.bar_wait(b, ph[b]); ph[b] += 1; }

  // hi/lo fp16 split of 2*NP pre-scaled values and store into NP consecutive A-operand columns of the own lane
  // (column c holds K-elements 2c (low half) and 2c+1 (high half))
  template <int NP>
  FC_HD_CTX void st_pairs(int col_hi, int col_lo, const float* v, float scale) {
    float hi[NP], lo[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const float x0 = fminf(fmaxf(v[2 * i] * scale, -kHalfMax), kHalfMax);
      const float x1 = fminf(fmaxf(v[2 * i + 1] * scale, -kHalfMax), kHalfMax);
      Ctx::split_h2(x0, x1, hi[i], lo[i]);
    }
    ctx.template tmem_st<NP>(col_hi, hi);
    ctx.template tmem_st<NP>(col_lo, lo);
  }
  // the 13 (12) owned unit values -> the 7 (6) columns of this quarter's block starting at block column col
  FC_HD_CTX void st_own(int col_hi, int col_lo, const float* v) {
    st_pairs<4>(col_hi, col_lo, v, kScaleA);
    st_pairs<2>(col_hi + 4, col_lo + 4, v + 8, kScaleA);
    if (full) {
      const float last[2] = {v[12], 0.f};
      st_pairs<1>(col_hi + 6, col_lo + 6, last, kScaleA);
    }
  }

  // ---------------------------------------------------------------------------------------------
  // operand images
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void request_weights(bool bwd, int l) {          // tid 0 only
    const int n = bwd ? bwd_img_halves(l) : fwd_img_halves(l);   // hi + lo images of halves = that many floats
    ctx.bulk_load(sm + kSmWTC, p.wpack + (bwd ? wb_off(l) : wf_off(l)), n, kBarWeights);
  }

  // one accumulator: 3 error-compensated terms, small ones first (tid 0 only); K = 16 * ksteps
  FC_HD_CTX void issue_mma(int d_col, int n, int a_hi, int a_lo, int ksteps, int img_halves, int bar) {
    const float* b_hi = sm + kSmWTC;
    const float* b_lo = b_hi + img_halves / 2;
    ctx.mma(d_col, n, a_lo, b_hi, n, 0, ksteps, false);
    ctx.mma(d_col, n, a_hi, b_lo, n, 0, ksteps, true);
    ctx.mma(d_col, n, a_hi, b_hi, n, 0, ksteps, true);
    ctx.commit(bar);
  }

  // ---------------------------------------------------------------------------------------------
  // tile set-up (quarter-0 thread of each trajectory)
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void load_tile(int tile) {
    const int b = tile * kTileTC + row;
    const bool ok = b < p.B;
    if (quarter == 0) {
      for (int r = 0; r < kLook; ++r)
#pragma unroll
        for (int f = 0; f < kFeat; ++f) {
          float v = ok ? p.Z[(size_t)b * (kLook * kFeat) + r * kFeat + f] : 0.f;
          if (r == kLook - 1 && f == kFeat - 1) v = ok ? p.u0[b] : 0.f;              // Functions.py:1396
          rows[(size_t)(r * kFeat + f) * kTileTC + row] = v;
        }
      sm[kSmRefTC + row] = ok ? p.X[(size_t)b * 3 + 2] : 0.f;                        // :1392
      sm[kSmUcurTC + row] = ok ? p.u0[b] : 0.f;
      sm[kSmUprevTC + row] = ok ? p.Z[(size_t)b * (kLook * kFeat) + (kLook - 2) * kFeat + 4] : 0.f;
      sm[kSmCostTC + row] = 0.f;
      sm[kSmCostTC + kTileTC + row] = 0.f;
      sm[kSmCostTC + 2 * kTileTC + row] = 0.f;
      if (ok) p.pred[(size_t)b * p.N] = p.u0[b];                                     // :1417-1418
    }
  }

  // ---------------------------------------------------------------------------------------------
  // forward: cell update of NU unit slots starting at slot j0 from 4*NU accumulator columns
  // ---------------------------------------------------------------------------------------------
  // g holds RAW accumulator values (pre-activation * kScaleA * kScaleW, before the truncation compensation);
  // ak folds unscale, compensation and -log2(e) into the exponent argument of the three sigmoid gates
  struct ActK { float k1, k2; };
  FC_HD_CTX static ActK make_actk(float unscale, float corr) {
    ActK k;
    const float khi = -kLog2e * unscale;                    // exact: unscale is a power of two
    k.k1 = fmaf(khi, corr, khi);                            // one multiplier per gate (see fc_mpc_pair_kernel.inl)
    k.k2 = -2.f * k.k1;
    return k;
  }
  template <int NU>
  FC_HD_CTX void fwd_units(int j0, const float* g, const ActK& ak, bool first, float* h, float* rp, int r0) {
    float rv[NU * 5 + 3];
    float cn[NU], th[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const float di = denom_(g[i * 4 + 0] * ak.k1);
      const float df = denom_(g[i * 4 + 1] * ak.k1);
      const float dq = denom_(g[i * 4 + 3] * ak.k1);
      const float dg = denom_(g[i * 4 + 2] * ak.k2);
      float gi, gf, go, rg;
      quad_rcp(di, df, dq, dg, gi, gf, go, rg);
      const float gg = tanh_from_(0.f, rg);
      const float cp = first ? 0.f : c[j0 + i];
      cn[i] = fmaf(gf, cp, gi * gg);
      c[j0 + i] = cn[i];
      rv[i * 5 + 0] = gi; rv[i * 5 + 1] = gf; rv[i * 5 + 2] = gg; rv[i * 5 + 3] = go; rv[i * 5 + 4] = cp;
    }
    tanh_batch<NU>(cn, th);
#pragma unroll
    for (int i = 0; i < NU; ++i) h[j0 + i] = rv[i * 5 + 3] * th[i];
    if (rp) {
#pragma unroll
      for (int i = NU * 5; i < NU * 5 + 3; ++i) rv[i] = 0.f;
#pragma unroll
      for (int r = 0; r < (NU * 5 + 3) / 4; ++r) {
        F4 v = {rv[r * 4], rv[r * 4 + 1], rv[r * 4 + 2], rv[r * 4 + 3]};
        Ctx::stg4_stream(rp + (size_t)(r0 + r) * 32 * 4, v);
      }
    }
  }

  // all unit slots of one step (the accumulator barrier has been waited for by the caller)
  FC_HD_CTX void fwd_pointwise(bool first, float corr, float* h, float* rec_out) {
    const ActK ak = make_actk(1.0f / (kScaleA * kScaleW), corr);
    float* rp = rec_out ? rec_out + ((size_t)warp * kRecF4 * 32 + lane) * 4 : nullptr;
    const int col0 = kColD + 4 * u_first;
    float g[2][16];
    ctx.template tmem_ld_nowait<16>(col0, g[0]);
#pragma unroll
    for (int gi = 0; gi < 3; ++gi) {
      ctx.tmem_ld_wait();
      // software pipeline: request the next accumulator columns before working on these
      if (gi + 1 < 3) ctx.template tmem_ld_nowait<16>(col0 + (gi + 1) * 16, g[(gi + 1) & 1]);
      else ctx.template tmem_ld_nowait<4>(col0 + 48, g[(gi + 1) & 1]);
      fwd_units<4>(gi * 4, g[gi & 1], ak, first, h, rp, gi * 5);
    }
    ctx.tmem_ld_wait();
    fwd_units<1>(12, g[1], ak, first, h, rp, 15);                // slot 12: a masked dummy for quarters 2,3
  }

  // input of (layer l, step t) for the owned columns: layer 0 -> 5 row features (quarter 0 only)
  FC_HD_CTX void load_input(int l, int m, int t, float* xin) {
    if (l == 0) {
      if (quarter == 0) {
        const float* rp = rows + (size_t)(m + t) * kFeat * kTileTC + row;
#pragma unroll
        for (int f = 0; f < kFeat; ++f) xin[f] = Ctx::ldcg(rp + f * kTileTC);
      }
    } else {
      const float* sq = seq + (size_t)t * kSlot + (size_t)warp * kMaxOwn * 32 + lane;
#pragma unroll
      for (int j = 0; j < kMaxOwn; ++j) xin[j] = Ctx::ldcg(sq + j * 32);
    }
  }
  FC_HD_CTX void store_input(int l, const float* xin) {
    if (l == 0) {
      if (quarter == 0) {
        const float v[8] = {xin[0], xin[1], xin[2], xin[3], xin[4], 0.f, 0.f, 0.f};
        st_pairs<4>(kColAhi, kColAlo, v, kScaleA);         // k = 0..7
      }
    } else {
      st_own(kColAhi + block_start(quarter) / 2, kColAlo + block_start(quarter) / 2, xin);
    }
  }

  // ---------------------------------------------------------------------------------------------
  // forward window
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void fwd_window(int tile, int m, bool more_after) {
    const int tmin = t_min_of(m);
    for (int l = 0; l < kLayers; ++l) {
      const int kf = kf_of(l);
      const int rec_col = (l == 0 ? kRec0 : kRec) / 2 + block_start(quarter) / 2;   // first recurrent column of this quarter
      float h[kMaxOwn], xin[kMaxOwn];
      // A(0): zero recurrent columns and padding, input of step 0
      {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        ctx.template tmem_st<4>(kColAhi + rec_col, z); ctx.template tmem_st<2>(kColAhi + rec_col + 4, z);
        ctx.template tmem_st<4>(kColAlo + rec_col, z); ctx.template tmem_st<2>(kColAlo + rec_col + 4, z);
        if (full) { ctx.template tmem_st<1>(kColAhi + rec_col + 6, z); ctx.template tmem_st<1>(kColAlo + rec_col + 6, z); }
        if (quarter == 3) {    // padding columns behind the recurrent block: layer 0 k [60,64), layers 1,2 k [104,112)
          if (l == 0) { ctx.template tmem_st<2>(kColAhi + 30, z); ctx.template tmem_st<2>(kColAlo + 30, z); }
          else        { ctx.template tmem_st<4>(kColAhi + 52, z); ctx.template tmem_st<4>(kColAlo + 52, z); }
        }
      }
      load_input(l, m, 0, xin);
      store_input(l, xin);
      ctx.tmem_st_wait();
      wait_bar(kBarWeights);                               // operand image of this layer landed
      ctx.tc_sync();
      lap(9);
      for (int t = 0; t < kLook; ++t) {
        // at t = 0 the recurrent columns are zero: only the k-steps that cover the input columns
        lap(0);
        const int ksteps = t == 0 ? (l == 0 ? 1 : 4) : kf / 16;
        if (tid == 0) issue_mma(kColD, kNF, kColAhi, kColAlo, ksteps, fwd_img_halves(l), kBarChunk0);
        if (t + 1 < kLook) load_input(l, m, t + 1, xin);
        float* rec_out = nullptr;
        if (p.with_grad && t >= tmin) rec_out = rec + (size_t)(rec_base(m) + (long)l * steps_kept(m) + (t - tmin)) * kRecFloatsTC;
        const float corr = Ctx::kAccTruncates ? acc_correction(ksteps, p.acc_comp) : 0.0f;
        lap(1);
        wait_bar(kBarChunk0);                              // accumulator complete; A and the image are free again
        lap(2);
        if (t == kLook - 1 && tid == 0) {                  // stream the next operand image under the cell update
          if (l + 1 < kLayers) request_weights(false, l + 1);
          else if (m + 1 < p.N) request_weights(false, 0);
          else if (p.with_grad) request_weights(true, kLayers - 1);
          else if (more_after) request_weights(false, 0);
        }
        if (t + 1 < kLook) store_input(l, xin);            // input columns of step t+1: overlap with the cell update
        fwd_pointwise(t == 0, corr, h, rec_out);
        lap(3);
        if (l + 1 < kLayers) {
          float* sq = seq + (size_t)t * kSlot + (size_t)warp * kMaxOwn * 32 + lane;
#pragma unroll
          for (int j = 0; j < kMaxOwn; ++j) sq[j * 32] = h[j];
        }
        if (t + 1 < kLook) {
          st_own(kColAhi + rec_col, kColAlo + rec_col, h);
          ctx.tmem_st_wait();
        } else if (l == kLayers - 1) {
          // read-out partial sums over the owned units (Functions.py:377)
          const float* fw = sm + kSmSmallTC;
          float xq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < kMaxOwn; ++j)
            if (j < nown) {
#pragma unroll
              for (int q = 0; q < 4; ++q) xq[q] = fmaf(fw[q * kHid + u_first + j], h[j], xq[q]);
            }
          if (quarter > 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) sm[kSmFcpTC + ((quarter - 1) * 4 + q) * kTileTC + row] = xq[q];
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) hrec[q] = xq[q];   // parked until the barrier below
          }
        }
        ctx.tc_sync();
        lap(4);
      }
    }
    lap(0);
    if (p.width_dim > 1) {
      fwd_glue_wide(tile, m);
    } else {
      if (quarter == 0) fwd_glue(tile, m);
      ctx.sync();
    }
    lap(10);
  }

  // ---------------------------------------------------------------------------------------------
  // after window m (quarter-0 thread of each trajectory): read-out, cost terms, next command
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void fwd_glue(int tile, int m) {
    const float* sw = sm + kSmSmallTC;
    float x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      x[q] = ((hrec[q] + sm[kSmFcpTC + q * kTileTC + row]) + (sm[kSmFcpTC + (4 + q) * kTileTC + row] + sm[kSmFcpTC + (8 + q) * kTileTC + row])) +
             sw[(kFCB - kFCW) + q];
    if (p.noise_std > 0.f) {                                                   // enable_noise, :1400-1402 / :1438-1440
      float e[4];
      philox_normal4(p.noise_seed, (unsigned)(tile * kTileTC + row), (unsigned)m, e);
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = fmaf(p.noise_std, e[q], x[q]);
    }
    const float ref = sm[kSmRefTC + row];
    const float ucur = sm[kSmUcurTC + row], uprev = sm[kSmUprevTC + row];
    float du = uprev - ucur;
    float cmd = p.alpha * du * du;                                             // :1405 / :1446
    float er = (x[0] - ref) * (x[0] - ref);                                    // :1408 / :1443
    float con = fmaxf(-x[1], 0.f) + fmaxf(-x[2], 0.f) + fmaxf(x[1] - kP1Max, 0.f) + fmaxf(x[2] - kP2Max, 0.f);
    sm[kSmCostTC + row] += (er + cmd) + con;                                   // :1414 / :1452
    sm[kSmCostTC + kTileTC + row] += cmd;
    sm[kSmCostTC + 2 * kTileTC + row] += er;
    float* rnew = rows + (size_t)(kLook + m) * kFeat * kTileTC + row;          // rho_{10+m} = [x_{m+1}, u_{m+1}]
#pragma unroll
    for (int q = 0; q < 4; ++q) rnew[q * kTileTC] = x[q];
    float unext = 0.f;
    if (m + 1 < p.N) {                                                         // :1424-1430
      const float* iw = sw + (kINPW - kFCW);
      const float* ib = sw + (kINPB - kFCW);
      const float* ow = sw + (kOUTW - kFCW);
      float v = 0.f;
      for (int u = 0; u < kFnnHid; ++u) {
        float pre = fmaf(iw[u * 3 + 2], ref, fmaf(iw[u * 3 + 1], x[3], fmaf(iw[u * 3 + 0], x[0], ib[u])));
        v = fmaf(ow[u], fmaxf(pre, 0.f), v);
      }
      unext = fminf(fmaxf(v, -1.f), 1.f);                                      // nn.Hardtanh
      sm[kSmUprevTC + row] = ucur;
      sm[kSmUcurTC + row] = unext;
      int b = tile * kTileTC + row;
      if (b < p.B) p.pred[(size_t)b * p.N + m + 1] = unext;                    // :1455
    }
    rnew[4 * kTileTC] = unext;
  }

  // ---------------------------------------------------------------------------------------------
  // width_dim > 1 controllers (FNNModel.forward, UL/Functions.py:261-289): the hidden layer fc_int + ReLU is applied
  // width_dim - 1 times with shared weights.  All 512 threads take part: thread (row, quarter) computes the units of
  // its quarter for its trajectory; layers are separated by CTA barriers; buffers are [50][129] in shared memory.
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX float* wide_buf(int i) const { return sm + kSmWideA + i * kActFloats; }
  FC_HD_CTX void wide_layer0(float* dst) {                   // a_0 = ReLU(fc_inp(x0, x3, ref))
    const float* sw = sm + kSmSmallTC;
    const float* iw = sw + (kINPW - kFCW);
    const float* ib = sw + (kINPB - kFCW);
    const float x0 = sm[kSmWideIn + row], x3 = sm[kSmWideIn + kTileTC + row], ref = sm[kSmWideIn + 2 * kTileTC + row];
    for (int j = 0; j < nown; ++j) {
      const int u = u_first + j;
      const float pre = fmaf(iw[u * 3 + 2], ref, fmaf(iw[u * 3 + 1], x3, fmaf(iw[u * 3 + 0], x0, ib[u])));
      dst[u * kActStride + row] = fmaxf(pre, 0.f);
    }
  }
  FC_HD_CTX void wide_hidden(const float* src, float* dst) { // a_r = ReLU(fc_int(a_{r-1}))
    const float* wi = sm + kSmWideW;
    const float* bi = wi + kFnnHid * kFnnHid;
    for (int j = 0; j < nown; ++j) {
      const int i = u_first + j;
      float pre = bi[i];
      for (int k = 0; k < kFnnHid; ++k) pre = fmaf(wi[i * kFnnHid + k], src[k * kActStride + row], pre);
      dst[i * kActStride + row] = fmaxf(pre, 0.f);
    }
  }
  FC_HD_CTX float wide_output(const float* a) const {        // fc_out (no bias)
    const float* ow = sm + kSmSmallTC + (kOUTW - kFCW);
    float v = 0.f;
    for (int u = 0; u < kFnnHid; ++u) v = fmaf(ow[u], a[u * kActStride + row], v);
    return v;
  }
  // all layers; with `keep` every a_r also goes to the global scratch actg[r][u][row].  Returns the buffer index of a_R.
  FC_HD_CTX int wide_forward(bool keep) {
    int cur = 0;
    wide_layer0(wide_buf(0));
    ctx.sync();
    if (keep)
      for (int j = 0; j < nown; ++j) actg[(size_t)(u_first + j) * kTileTC + row] = wide_buf(0)[(u_first + j) * kActStride + row];
    for (int r = 1; r < p.width_dim; ++r) {
      wide_hidden(wide_buf(cur), wide_buf(cur ^ 1));
      ctx.sync();
      cur ^= 1;
      if (keep)
        for (int j = 0; j < nown; ++j)
          actg[((size_t)r * kFnnHid + u_first + j) * kTileTC + row] = wide_buf(cur)[(u_first + j) * kActStride + row];
    }
    return cur;
  }

  FC_HD_CTX void fwd_glue_wide(int tile, int m) {
    float unext = 0.f;
    float* rnew = rows + (size_t)(kLook + m) * kFeat * kTileTC + row;
    float ucur = 0.f;
    if (quarter == 0) {                                      // read-out, cost terms: as fwd_glue
      const float* sw = sm + kSmSmallTC;
      float x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        x[q] = ((hrec[q] + sm[kSmFcpTC + q * kTileTC + row]) + (sm[kSmFcpTC + (4 + q) * kTileTC + row] + sm[kSmFcpTC + (8 + q) * kTileTC + row])) +
               sw[(kFCB - kFCW) + q];
      if (p.noise_std > 0.f) {
        float e[4];
        philox_normal4(p.noise_seed, (unsigned)(tile * kTileTC + row), (unsigned)m, e);
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = fmaf(p.noise_std, e[q], x[q]);
      }
      const float ref = sm[kSmRefTC + row];
      ucur = sm[kSmUcurTC + row];
      const float uprev = sm[kSmUprevTC + row];
      float du = uprev - ucur;
      float cmd = p.alpha * du * du;
      float er = (x[0] - ref) * (x[0] - ref);
      float con = fmaxf(-x[1], 0.f) + fmaxf(-x[2], 0.f) + fmaxf(x[1] - kP1Max, 0.f) + fmaxf(x[2] - kP2Max, 0.f);
      sm[kSmCostTC + row] += (er + cmd) + con;
      sm[kSmCostTC + kTileTC + row] += cmd;
      sm[kSmCostTC + 2 * kTileTC + row] += er;
#pragma unroll
      for (int q = 0; q < 4; ++q) rnew[q * kTileTC] = x[q];
      sm[kSmWideIn + row] = x[0];
      sm[kSmWideIn + kTileTC + row] = x[3];
      sm[kSmWideIn + 2 * kTileTC + row] = ref;
    }
    ctx.sync();
    if (m + 1 < p.N) {
      const int cur = wide_forward(false);
      if (quarter == 0) {
        unext = fminf(fmaxf(wide_output(wide_buf(cur)), -1.f), 1.f);           // nn.Hardtanh
        sm[kSmUprevTC + row] = ucur;
        sm[kSmUcurTC + row] = unext;
        int b = tile * kTileTC + row;
        if (b < p.B) p.pred[(size_t)b * p.N + m + 1] = unext;
      }
    }
    if (quarter == 0) rnew[4 * kTileTC] = unext;
    ctx.sync();
  }

  // reverse: recompute the activations, then delta_R = dv * fc_out * ReLU', delta_{r-1} = ReLU' * fc_int^T delta_r;
  // gradients of fc_int (weight-shared: summed over the repeats), fc_inp, fc_out accumulated in fp64 per CTA
  FC_HD_CTX void bwd_glue_wide(int tile, int m) {
    const int k = m + 1;
    const float s = p.grad_scale;
    const bool has_u = k <= p.N - 1;
    const float* sw = sm + kSmSmallTC;
    const float* iw = sw + (kINPW - kFCW);
    const float* ow = sw + (kOUTW - kFCW);
    const float* wi = sm + kSmWideW;
    const bool valid = tile * kTileTC + row < p.B;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f, gu = 0.f;
    if (quarter == 0) {
      const float* rx = rows + (size_t)(kLook + m) * kFeat * kTileTC + row;
      float x0 = Ctx::ldcg(rx), x1 = Ctx::ldcg(rx + kTileTC), x2 = Ctx::ldcg(rx + 2 * kTileTC), x3 = Ctx::ldcg(rx + 3 * kTileTC);
      const float ref = sm[kSmRefTC + row];
      g0 = 2.f * (x0 - ref) * s;
      g1 = s * ((x1 > kP1Max ? 1.f : 0.f) - (x1 < 0.f ? 1.f : 0.f));
      g2 = s * ((x2 > kP2Max ? 1.f : 0.f) - (x2 < 0.f ? 1.f : 0.f));
      if (has_u) {
        const float* gr = grow + (size_t)k * kFeat * kTileTC + row;
        float uk = Ctx::ldcg(rows + (size_t)((kLook - 1 + k) * kFeat + 4) * kTileTC + row);
        float ukm1 = Ctx::ldcg(rows + (size_t)((kLook - 2 + k) * kFeat + 4) * kTileTC + row);
        gu = Ctx::ldcg(gr + 4 * kTileTC) - 2.f * p.alpha * (ukm1 - uk) * s;
        if (k + 1 <= p.N - 1) {
          float ukp1 = Ctx::ldcg(rows + (size_t)((kLook + k) * kFeat + 4) * kTileTC + row);
          gu += 2.f * p.alpha * (uk - ukp1) * s;
        }
        g0 += Ctx::ldcg(gr);
        g1 += Ctx::ldcg(gr + kTileTC);
        g2 += Ctx::ldcg(gr + 2 * kTileTC);
        g3 += Ctx::ldcg(gr + 3 * kTileTC);
        sm[kSmWideIn + row] = x0;
        sm[kSmWideIn + kTileTC + row] = x3;
        sm[kSmWideIn + 2 * kTileTC + row] = ref;
      }
    }
    ctx.sync();
    if (has_u) {
      const int R = p.width_dim - 1;
      const int top = wide_forward(true);                    // a_R in wide_buf(top), every a_r in actg
      if (quarter == 0) {
        const float v = wide_output(wide_buf(top));
        sm[kSmDvTC + row] = (valid && v > -1.f && v < 1.f) ? gu : 0.f;         // hardtanh_backward
      }
      ctx.sync();
      // fc_out gradient (needs a_R) and delta_R; buffers: cur = deltas, nxt = next deltas, stage = a_{r-1}
      if (tid < 4 * kFnnHid) {
        const int u = tid % kFnnHid, part = tid / kFnnHid;
        double a_ow = 0.0;
        for (int tr = part * 32; tr < part * 32 + 32; ++tr)
          a_ow += (double)sm[kSmDvTC + tr] * (double)wide_buf(top)[u * kActStride + tr];
        reinterpret_cast<double*>(sm + kSmPgTC)[part * kNumFnnGrad + 200 + u] += a_ow;
      }
      int cur = top ^ 1;                                     // free buffer
      {
        const float dv = sm[kSmDvTC + row];
        for (int j = 0; j < nown; ++j) {
          const int i = u_first + j;
          wide_buf(cur)[i * kActStride + row] = wide_buf(top)[i * kActStride + row] > 0.f ? dv * ow[i] : 0.f;
        }
      }
      ctx.sync();
      int nxt = top;                                         // a_R no longer needed
      float* stage = wide_buf(2);
      double* acc = reinterpret_cast<double*>(sm + kSmWideAcc);
      for (int r = R; r >= 1; --r) {
        // stage a_{r-1} (global scratch -> shared, padded stride)
        for (int e = tid; e < kFnnHid * kTileTC; e += kThreadsTC) {
          const int u = e / kTileTC, tr = e - u * kTileTC;
          stage[u * kActStride + tr] = Ctx::ldcg(actg + ((size_t)(r - 1) * kFnnHid + u) * kTileTC + tr);
        }
        ctx.sync();
        // d fc_int.weight[i][j] += sum_tr delta_r[i][tr] * a_{r-1}[j][tr];  d fc_int.bias[i] += sum_tr delta_r[i][tr]
        for (int e = tid; e < kWideGrads; e += kThreadsTC) {
          double a = 0.0;
          if (e < kFnnHid * kFnnHid) {
            const int i = e / kFnnHid, j = e - i * kFnnHid;
            for (int tr = 0; tr < kTileTC; ++tr) a += (double)wide_buf(cur)[i * kActStride + tr] * (double)stage[j * kActStride + tr];
          } else {
            const int i = e - kFnnHid * kFnnHid;
            for (int tr = 0; tr < kTileTC; ++tr) a += (double)wide_buf(cur)[i * kActStride + tr];
          }
          acc[e] += a;
        }
        // delta_{r-1}[j] = ReLU'(a_{r-1}[j]) * sum_i fc_int.weight[i][j] * delta_r[i]
        for (int jj = 0; jj < nown; ++jj) {
          const int j = u_first + jj;
          float sacc = 0.f;
          for (int i = 0; i < kFnnHid; ++i) sacc = fmaf(wi[i * kFnnHid + j], wide_buf(cur)[i * kActStride + row], sacc);
          wide_buf(nxt)[j * kActStride + row] = stage[j * kActStride + row] > 0.f ? sacc : 0.f;
        }
        ctx.sync();
        const int t = cur; cur = nxt; nxt = t;
      }
      // cur = delta_0: input gradient of the controller and fc_inp gradients
      if (quarter == 0) {
        float d0 = 0.f, d1 = 0.f;
        for (int u = 0; u < kFnnHid; ++u) {
          const float dp = wide_buf(cur)[u * kActStride + row];
          d0 = fmaf(dp, iw[u * 3 + 0], d0);
          d1 = fmaf(dp, iw[u * 3 + 1], d1);
        }
        g0 += d0;
        g3 += d1;
      }
      if (tid < 4 * kFnnHid) {
        const int u = tid % kFnnHid, part = tid / kFnnHid;
        double a_b = 0.0, a_w0 = 0.0, a_w1 = 0.0, a_w2 = 0.0;
        for (int tr = part * 32; tr < part * 32 + 32; ++tr) {
          const float dp = wide_buf(cur)[u * kActStride + tr];
          a_b += dp;
          a_w0 += (double)dp * (double)sm[kSmWideIn + tr];
          a_w1 += (double)dp * (double)sm[kSmWideIn + kTileTC + tr];
          a_w2 += (double)dp * (double)sm[kSmWideIn + 2 * kTileTC + tr];
        }
        double* pg = reinterpret_cast<double*>(sm + kSmPgTC) + part * kNumFnnGrad;
        pg[u * 3 + 0] += a_w0;
        pg[u * 3 + 1] += a_w1;
        pg[u * 3 + 2] += a_w2;
        pg[150 + u] += a_b;
      }
    }
    if (quarter == 0) {
      if (!valid) { g0 = g1 = g2 = g3 = 0.f; }
      sm[kSmGxTC + row] = g0;
      sm[kSmGxTC + kTileTC + row] = g1;
      sm[kSmGxTC + 2 * kTileTC + row] = g2;
      sm[kSmGxTC + 3 * kTileTC + row] = g3;
    }
  }

  // ---------------------------------------------------------------------------------------------
  // backward cell gradient, split in two parts so that everything that does not depend on the result of
  // the running MMA happens in its shadow:
  //   bwd_factors (during MMA(t+1)): record(t) -> A = o(1-tanh^2 c), Ko = tanh(c) o(1-o), Ki = g i(1-i),
  //                                  Kf = c_prev f(1-f), Kg = i(1-g^2), Gf = f          (all MUFU work)
  //   bwd_finish  (after MMA(t+1)):  dh -> dct = dc + dh A; dG = (dct Ki, dct Kf, dct Kg, dh Ko); dc = dct Gf
  // ---------------------------------------------------------------------------------------------
  struct Factors { float A[kMaxOwn], Ko[kMaxOwn], Ki[kMaxOwn], Kf[kMaxOwn], Kg[kMaxOwn], Gf[kMaxOwn]; };

  template <int NU>
  FC_HD_CTX void factors_group(const float* rp, int r0, int j0, Factors& fa) {
    float rv[NU * 5 + 3];
#pragma unroll
    for (int r = 0; r < (NU * 5 + 3) / 4; ++r) {
      F4 v = Ctx::ldg4_stream(rp + (size_t)(r0 + r) * 32 * 4);
      rv[r * 4] = v.x; rv[r * 4 + 1] = v.y; rv[r * 4 + 2] = v.z; rv[r * 4 + 3] = v.w;
    }
    float cn[NU], th[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) cn[i] = fmaf(rv[i * 5 + 1], rv[i * 5 + 4], rv[i * 5 + 0] * rv[i * 5 + 2]);
    tanh_batch<NU>(cn, th);
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const int j = j0 + i;
      float gi = rv[i * 5 + 0], gf = rv[i * 5 + 1], gg = rv[i * 5 + 2], go = rv[i * 5 + 3], cp = rv[i * 5 + 4];
      float tch = th[i];
      fa.A[j] = go * (1.f - tch * tch);
      fa.Ko[j] = tch * go * (1.f - go);
      fa.Ki[j] = gg * gi * (1.f - gi);
      fa.Kf[j] = cp * gf * (1.f - gf);
      fa.Kg[j] = gi * (1.f - gg * gg);
      fa.Gf[j] = gf;
    }
  }
  FC_HD_CTX void prefetch_record(const float* rec_in) {
    const float* rp = rec_in + ((size_t)warp * kRecF4 * 32 + lane) * 4;
#pragma unroll
    for (int r = 0; r < kRecF4; ++r) Ctx::prefetch_l2(rp + (size_t)r * 32 * 4);
  }
  FC_HD_CTX void bwd_factors(const float* rec_in, Factors& fa) {
    const float* rp = rec_in + ((size_t)warp * kRecF4 * 32 + lane) * 4;
#pragma unroll
    for (int gi = 0; gi < 3; ++gi) factors_group<4>(rp, gi * 5, gi * 4, fa);
    factors_group<1>(rp, 15, 12, fa);
  }

  // d(h) contributions that do not come from the recurrent MMA: the layer above (thread-private scratch)
  // or, for the top layer at the last step, the read-out (Functions.py:377)
  FC_HD_CTX void bwd_extra(int l, int t, float* extra) {
    if (l == kLayers - 1) {
      if (t == kLook - 1) {
        float gxv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) gxv[q] = sm[kSmGxTC + q * kTileTC + row];
#pragma unroll
        for (int j = 0; j < kMaxOwn; ++j) {
          const int u = u_first + j < kHid ? u_first + j : kHid - 1;
          const float* fw = sm + kSmSmallTC + u;
          extra[j] = fw[0] * gxv[0] + fw[kHid] * gxv[1] + fw[2 * kHid] * gxv[2] + fw[3 * kHid] * gxv[3];
        }
      } else {
#pragma unroll
        for (int j = 0; j < kMaxOwn; ++j) extra[j] = 0.f;
      }
    } else {
      const float* dsq = dseq + (size_t)t * kSlot + (size_t)warp * kMaxOwn * 32 + lane;
#pragma unroll
      for (int j = 0; j < kMaxOwn; ++j) extra[j] = Ctx::ldcg(dsq + j * 32);
    }
  }

  template <int NU>
  FC_HD_CTX void finish_group(int j0, const Factors& fa, const float* extra, float* dg) {
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const int j = j0 + i;
      const float dh = hrec[j] + extra[j];
      const float dct = fmaf(dh, fa.A[j], c[j]);
      c[j] = dct * fa.Gf[j];
      dg[i * 4 + 0] = dct * fa.Ki[j];
      dg[i * 4 + 1] = dct * fa.Kf[j];
      dg[i * 4 + 2] = dct * fa.Kg[j];
      dg[i * 4 + 3] = dh * fa.Ko[j];
    }
  }
  FC_HD_CTX void bwd_finish(const Factors& fa, const float* extra) {
    const int col0 = 2 * u_first;                          // dG k-index = unit*4+gate, two per column
#pragma unroll
    for (int gi = 0; gi < 3; ++gi) {
      float dg[16];
      finish_group<4>(gi * 4, fa, extra, dg);
      st_pairs<8>(kColGhi + col0 + gi * 8, kColGlo + col0 + gi * 8, dg, p.g_scale);
    }
    float dg[4];
    finish_group<1>(12, fa, extra, dg);                    // slot 12: computed by everybody, stored by the owners
    if (full) st_pairs<2>(kColGhi + col0 + 24, kColGlo + col0 + 24, dg, p.g_scale);
  }

  // ---------------------------------------------------------------------------------------------
  // before the reverse sweep of window m (quarter-0 thread per trajectory + 200 accumulation threads)
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void bwd_glue(int tile, int m) {
    const int k = m + 1;
    const float s = p.grad_scale;
    const bool has_u = k <= p.N - 1;
    const float* sw = sm + kSmSmallTC;
    const float* iw = sw + (kINPW - kFCW);
    const float* ib = sw + (kINPB - kFCW);
    const float* ow = sw + (kOUTW - kFCW);
    if (quarter == 0) {
      const bool valid = tile * kTileTC + row < p.B;
      const float* rx = rows + (size_t)(kLook + m) * kFeat * kTileTC + row;
      float x0 = Ctx::ldcg(rx), x1 = Ctx::ldcg(rx + kTileTC), x2 = Ctx::ldcg(rx + 2 * kTileTC), x3 = Ctx::ldcg(rx + 3 * kTileTC);
      const float ref = sm[kSmRefTC + row];
      float g0 = 2.f * (x0 - ref) * s;
      float g1 = s * ((x1 > kP1Max ? 1.f : 0.f) - (x1 < 0.f ? 1.f : 0.f));
      float g2 = s * ((x2 > kP2Max ? 1.f : 0.f) - (x2 < 0.f ? 1.f : 0.f));
      float g3 = 0.f;
      float dv = 0.f;
      if (has_u) {
        const float* gr = grow + (size_t)k * kFeat * kTileTC + row;
        float uk = Ctx::ldcg(rows + (size_t)((kLook - 1 + k) * kFeat + 4) * kTileTC + row);
        float ukm1 = Ctx::ldcg(rows + (size_t)((kLook - 2 + k) * kFeat + 4) * kTileTC + row);
        float gu = Ctx::ldcg(gr + 4 * kTileTC) - 2.f * p.alpha * (ukm1 - uk) * s;
        if (k + 1 <= p.N - 1) {
          float ukp1 = Ctx::ldcg(rows + (size_t)((kLook + k) * kFeat + 4) * kTileTC + row);
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
        g0 += d0 + Ctx::ldcg(gr);
        g1 += Ctx::ldcg(gr + kTileTC);
        g2 += Ctx::ldcg(gr + 2 * kTileTC);
        g3 += d1 + Ctx::ldcg(gr + 3 * kTileTC);
        sm[kSmDvTC + row] = dv;
        sm[kSmFinTC + row] = x0;
        sm[kSmFinTC + kTileTC + row] = x3;
      }
      if (!valid) { g0 = g1 = g2 = g3 = 0.f; }
      sm[kSmGxTC + row] = g0;
      sm[kSmGxTC + kTileTC + row] = g1;
      sm[kSmGxTC + 2 * kTileTC + row] = g2;
      sm[kSmGxTC + 3 * kTileTC + row] = g3;
    }
    ctx.sync();
    if (has_u && tid < 4 * kFnnHid) {                  // controller weight gradients, unit-parallel
      const int u = tid % kFnnHid, part = tid / kFnnHid;
      double a_ow = 0.0, a_b = 0.0, a_w0 = 0.0, a_w1 = 0.0, a_w2 = 0.0;   // batch sums cancel heavily: fp64
      const float w0 = iw[u * 3 + 0], w1 = iw[u * 3 + 1], w2 = iw[u * 3 + 2], bb = ib[u], owu = ow[u];
      for (int tr = part * 32; tr < part * 32 + 32; ++tr) {
        float dv = sm[kSmDvTC + tr];
        float x0 = sm[kSmFinTC + tr], x3 = sm[kSmFinTC + kTileTC + tr], ref = sm[kSmRefTC + tr];
        float pre = fmaf(w2, ref, fmaf(w1, x3, fmaf(w0, x0, bb)));
        a_ow += (double)dv * (double)fmaxf(pre, 0.f);
        float dp = pre > 0.f ? dv * owu : 0.f;
        a_b += dp;
        a_w0 += (double)dp * (double)x0;
        a_w1 += (double)dp * (double)x3;
        a_w2 += (double)dp * (double)ref;
      }
      double* pg = reinterpret_cast<double*>(sm + kSmPgTC) + part * kNumFnnGrad;
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
  FC_HD_CTX void bwd_window(int tile, int m, bool more_after) {
    const int tmin = t_min_of(m);
    const float corr_b = Ctx::kAccTruncates ? acc_correction(kKB / 16, p.acc_comp) : 0.0f;
    const float unscale_b = p.g_unscale / kScaleW;         // exact power of two
    lap(0);
    if (p.width_dim > 1) bwd_glue_wide(tile, m);
    else bwd_glue(tile, m);
    ctx.sync();
    lap(11);
    for (int l = kLayers - 1; l >= 0; --l) {
      const int nb = nb_of(l);
#pragma unroll
      for (int j = 0; j < kMaxOwn; ++j) { c[j] = 0.f; hrec[j] = 0.f; }
      if (quarter == 3) {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        ctx.template tmem_st<4>(kColGhi + 100, z);
        ctx.template tmem_st<4>(kColGlo + 100, z);
      }
      Factors fa;
      float extra[kMaxOwn];
      const float* rec_l = rec + (size_t)(rec_base(m) + (long)l * steps_kept(m)) * kRecFloatsTC;
      bwd_factors(rec_l + (size_t)(kLook - 1 - tmin) * kRecFloatsTC, fa);
      bwd_extra(l, kLook - 1, extra);
      wait_bar(kBarWeights);
      lap(12);
      for (int t = kLook - 1; t >= tmin; --t) {
        lap(0);
        bwd_finish(fa, extra);
        ctx.tmem_st_wait();
        ctx.tc_sync();
        lap(5);
        if (tid == 0) issue_mma(kColD, nb, kColGhi, kColGlo, kKB / 16, bwd_img_halves(l), kBarChunk0);
        // in the shadow of the MMA: record and upstream gradient of the next step
        if (t > tmin) {
          // HBM -> L2 for the step after next (or the first step of the next layer / window)
          if (t - 2 >= tmin) prefetch_record(rec_l + (size_t)(t - 2 - tmin) * kRecFloatsTC);
          else if (l > 0) prefetch_record(rec + (size_t)(rec_base(m) + (long)(l - 1) * steps_kept(m) + (kLook - 1 - tmin)) * kRecFloatsTC);
          else if (m > 0) prefetch_record(rec + (size_t)(rec_base(m - 1) + (long)(kLayers - 1) * steps_kept(m - 1) + (kLook - 1 - t_min_of(m - 1))) * kRecFloatsTC);
          bwd_factors(rec_l + (size_t)(t - 1 - tmin) * kRecFloatsTC, fa);
          bwd_extra(l, t - 1, extra);
        }
        lap(6);
        wait_bar(kBarChunk0);
        lap(7);
        if (t == tmin && tid == 0) {                   // all MMAs that read this image are complete
          if (l > 0) request_weights(true, l - 1);
          else if (m > 0) request_weights(true, kLayers - 1);
          else if (more_after) request_weights(false, 0);
        }
        if (l > 0) {
          float d[32];
          ctx.template tmem_ld_nowait<16>(kColD + 26 * quarter, d);
          ctx.template tmem_ld_nowait<8>(kColD + 26 * quarter + 16, d + 16);
          ctx.template tmem_ld_nowait<2>(kColD + 26 * quarter + 24, d + 24);
          ctx.tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 2 * kMaxOwn; ++j) d[j] *= unscale_b;
          float* dq = dseq + (size_t)t * kSlot + (size_t)warp * kMaxOwn * 32 + lane;
#pragma unroll
          for (int j = 0; j < kMaxOwn; ++j) {
            dq[j * 32] = fmaf(d[j], corr_b, d[j]);                     // d(input unit) -> the layer below, same thread
            hrec[j] = fmaf(d[kMaxOwn + j], corr_b, d[kMaxOwn + j]);
          }
        } else {
          float d[16], df[8];
          ctx.template tmem_ld_nowait<8>(kColD + 13 * quarter, d);
          ctx.template tmem_ld_nowait<4>(kColD + 13 * quarter + 8, d + 8);
          ctx.template tmem_ld_nowait<1>(kColD + 13 * quarter + 12, d + 12);
          if (quarter == 0) ctx.template tmem_ld_nowait<8>(kColD + 52, df);
          ctx.tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < kMaxOwn; ++j) d[j] *= unscale_b;
#pragma unroll
          for (int f = 0; f < kFeat; ++f) df[f] *= unscale_b;
#pragma unroll
          for (int j = 0; j < kMaxOwn; ++j) hrec[j] = fmaf(d[j], corr_b, d[j]);
          const int kr = m + t - (kLook - 1);          // gradient of row rho_{9+kr}
          if (quarter == 0 && kr >= 0) {
            float* gp = grow + (size_t)kr * kFeat * kTileTC + row;
#pragma unroll
            for (int f = 0; f < kFeat; ++f) gp[f * kTileTC] = Ctx::ldcg(gp + f * kTileTC) + fmaf(df[f], corr_b, df[f]);
          }
        }
        lap(8);
      }
    }
  }

  // ---------------------------------------------------------------------------------------------
  // per-tile epilogues
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void store_costs(int tile) {
    if (quarter == 0) {
      int b = tile * kTileTC + row;
      if (b < p.B) {
        const float inv = 1.f / (float)p.N;
        float cst = sm[kSmCostTC + row] * inv;                                 // :1458-1460
        p.cost[b] = cst;
        p.command[b] = sm[kSmCostTC + kTileTC + row] * inv;
        p.error[b] = sm[kSmCostTC + 2 * kTileTC + row] * inv;
        sm[kSmCostTC + row] = cst;
      } else {
        sm[kSmCostTC + row] = 0.f;
      }
    }
    ctx.sync();
    if (tid == 0) {
      double acc = 0.0;
      for (int i = 0; i < kTileTC; ++i) acc += (double)sm[kSmCostTC + i];
      *reinterpret_cast<double*>(sm + kSmRedTC) += acc;
    }
  }

  FC_HD_CTX void store_du0(int tile) {
    if (quarter == 0) {
      int b = tile * kTileTC + row;
      if (b < p.B) {
        const float s = p.grad_scale;
        float u0 = Ctx::ldcg(rows + (size_t)((kLook - 1) * kFeat + 4) * kTileTC + row);
        float um1 = Ctx::ldcg(rows + (size_t)((kLook - 2) * kFeat + 4) * kTileTC + row);
        float g = Ctx::ldcg(grow + 4 * kTileTC + row) - 2.f * p.alpha * (um1 - u0) * s;
        if (p.N > 1) {
          float u1 = Ctx::ldcg(rows + (size_t)(kLook * kFeat + 4) * kTileTC + row);
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
    ctx.tc_setup(sm + kSmBarTC);
#ifdef FC_TC_TIMING
    tlast = Ctx::clock();
#endif
    for (int i = tid; i < kSmallFloats; i += kThreadsTC) sm[kSmSmallTC + i] = p.wpack[kSmallOff + i];
    for (int i = tid; i < 4 * kNumFnnGrad; i += kThreadsTC) reinterpret_cast<double*>(sm + kSmPgTC)[i] = 0.0;
    if (p.width_dim > 1) {
      for (int i = tid; i < kFnnHid * kFnnHid; i += kThreadsTC) sm[kSmWideW + i] = p.int_w[i];
      for (int i = tid; i < kFnnHid; i += kThreadsTC) sm[kSmWideW + kFnnHid * kFnnHid + i] = p.int_b[i];
      for (int i = tid; i < kWideGrads; i += kThreadsTC) reinterpret_cast<double*>(sm + kSmWideAcc)[i] = 0.0;
    }
    if (tid == 0) *reinterpret_cast<double*>(sm + kSmRedTC) = 0.0;
    ctx.sync();
    if (tid == 0 && ctx.bid() < p.num_tiles) request_weights(false, 0);
    for (int tile = ctx.bid(); tile < p.num_tiles; tile += ctx.nblk()) {
      const bool more = tile + ctx.nblk() < p.num_tiles;
      load_tile(tile);
      ctx.sync();
      for (int m = 0; m < p.N; ++m) fwd_window(tile, m, more);
      store_costs(tile);
      if (p.with_grad) {
        if (quarter == 0) {
          for (int k = 0; k < p.N; ++k)
#pragma unroll
            for (int f = 0; f < kFeat; ++f) grow[(size_t)(k * kFeat + f) * kTileTC + row] = 0.f;
        }
        ctx.sync();
        for (int m = p.N - 1; m >= 0; --m) bwd_window(tile, m, more);
        ctx.sync();
        store_du0(tile);
      }
      ctx.sync();
    }
    double* part = p.partial + (size_t)ctx.bid() * kPartialStride;
    const double* pgd = reinterpret_cast<const double*>(sm + kSmPgTC);
    for (int i = tid; i < kNumFnnGrad; i += kThreadsTC)
      part[i] = (pgd[i] + pgd[kNumFnnGrad + i]) + (pgd[2 * kNumFnnGrad + i] + pgd[3 * kNumFnnGrad + i]);
    if (tid == 0) part[kNumFnnGrad] = *reinterpret_cast<const double*>(sm + kSmRedTC);
    if (p.width_dim > 1 && p.with_grad) {
      double* pw = p.partial_wide + (size_t)ctx.bid() * kWidePartialStride;
      for (int i = tid; i < kWideGrads; i += kThreadsTC) pw[i] = reinterpret_cast<const double*>(sm + kSmWideAcc)[i];
    }
    ctx.sync();
#ifdef FC_TC_TIMING
    if (p.debug_timing && tid == 0 && ctx.bid() == 0) Ctx::report(tm);
#endif
    ctx.tc_teardown();
  }
};

}  // namespace tc
}  // namespace fc
