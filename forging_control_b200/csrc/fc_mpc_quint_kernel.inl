// Five-group ("quint") two-tile tcgen05 variant of the fused MPC-loss forward + reverse sweep.  Same mathematics and
// reference citations as fc_mpc_kernel.inl (MPCLoss.forward Functions.py:1353-1472, LSTMModel.forward :353-379,
// FNNModel.forward :261-289, loss.backward() :655), the same fp16 hi/lo split tensor-core contraction as
// fc_mpc_tc_kernel.inl and the same schedule as fc_mpc_pair_kernel.inl: a CTA owns TWO 128-trajectory tiles and
// alternates between them cell step by cell step, so that the gate contraction of one tile (tcgen05.mma, asynchronous)
// runs under the cell update (MUFU / FMA pipes) of the other.
//
// What changes against the pair kernel (see fc_quint_layout.h for the layouts):
//   * 24 warps: the 50 hidden units are split over FIVE groups of exactly 10 (20 cell-update warps, 5 per SM
//     sub-partition, perfectly balanced) instead of three thirds of 16/16/18 units (12 warps): more warps to cover the
//     hand-shake / TMEM / record-load latencies that left 45 % of the issue slots empty in the round-1 kernel;
//   * the cell state (forward) / d(cell state) (reverse) of BOTH tiles stays in registers (2 x 10 per thread): the
//     park-and-swap through spare TMEM columns at every item is gone.  The per-item code is instantiated per tile
//     (template <int X>) so that the register arrays are addressed statically;
//   * all per-trajectory scalar work (roll-out rows, layer-0 features, read-out, cost terms, controller, row-feature
//     gradients) belongs to warps 0..3 (one per TMEM quadrant); the cell-update warps do nothing else.  Warp 0 is also
//     the MMA issuer: its scalar work sits between "accumulator complete" and "operands ready", where it idles anyway.
#pragma once
#include "fc_quint_layout.h"

namespace fc {
namespace q5 {

template <class Ctx>
struct MpcQuint {
  Ctx& ctx;
  const MpcParams& p;
  float* sm;
  int tid, warp, lane, row, grp, uw;
  bool issuer;           // warp 0: issues every MMA, requests the weight images (and is the scalar warp of quadrant 0)
  bool scalar;           // warps 0..3: per-trajectory scalar work of the rows of their TMEM quadrant
  bool update;           // warps 4..23: cell update of the units [10 grp, 10 grp + 10) of their rows
  float* wbase;          // workspace of tile 0 of this CTA; tile 1 follows at +tstride
  size_t tstride;
  int ntl;               // live tiles in this pass (1 or 2)
  int tile0;             // global index of tile 0 of this pass
  // forward: cell state, reverse sweep: d(cell state); tile 0 / tile 1, both resident in registers
  float cA[kOwn], cB[kOwn];
  unsigned ph;           // phase parities of the five mbarriers, one bit each (bit = barrier id)
#ifdef FC_TC_TIMING
  long long tm[24], tlast;  // cycle breakdown (development aid, -DFC_TC_TIMING)
  FC_HD_CTX void lap(int k) { long long t = Ctx::clock(); tm[k] += t - tlast; tlast = t; }
#else
  FC_HD_CTX void lap(int) {}
#endif

  FC_HD_CTX MpcQuint(Ctx& c_, const MpcParams& p_) : ctx(c_), p(p_) {
    sm = ctx.smem();
    tid = ctx.tid();
    warp = tid >> 5;
    lane = tid & 31;
    row = 32 * (warp & 3) + lane;
    issuer = warp == 0;
    update = warp >= 4 && warp < 4 + kUpdWarpsQ;
    scalar = warp < 4;
    uw = update ? warp - 4 : 0;                              // index among the cell-update warps
    grp = uw >> 2;
    tstride = work_layout_q(p.N, p.with_grad).total;
    wbase = p.work + (size_t)ctx.bid() * p.work_stride;
    ph = 0;
    ntl = 0; tile0 = 0;
#pragma unroll
    for (int j = 0; j < kOwn; ++j) { cA[j] = 0.f; cB[j] = 0.f; }
#ifdef FC_TC_TIMING
    for (int i = 0; i < 24; ++i) tm[i] = 0;
    tlast = 0;
#endif
  }

  // workspace pointers of tile X
  FC_HD_CTX float* w_rows(int X) const { return wbase + X * tstride; }
  FC_HD_CTX float* w_cost(int X) const { return w_rows(X) + (size_t)(p.N + kLook) * kFeat * kTileQ; }
  FC_HD_CTX float* w_seq(int X) const { return w_cost(X) + 3 * kTileQ; }
  FC_HD_CTX float* w_dseq(int X) const { return w_seq(X) + (size_t)kLook * kSeqSlot; }
  FC_HD_CTX float* w_grow(int X) const { return w_dseq(X) + (size_t)kLook * kDseqSlot; }
  FC_HD_CTX float* w_rec(int X) const { return w_rows(X) + work_layout_q(p.N, p.with_grad).rec; }

  // ---------------------------------------------------------------------------------------------
  // activations (see fc_mpc_tc_kernel.inl): one reciprocal per four denominators, polynomial tanh near 0
  // ---------------------------------------------------------------------------------------------
  static constexpr float kExpMax = 30.0f;
  static constexpr float kLog2e = 1.4426950216293335f;            // fp32(log2 e)
  static constexpr float kLog2eLo = 1.92596e-8f;                  // log2 e - fp32(log2 e)
  FC_HD_CTX static float denom_(float e2arg) { return 1.f + Ctx::ex2(fminf(e2arg, kExpMax)); }
  FC_HD_CTX static void quad_rcp(float a, float b, float c, float d, float& ra, float& rb, float& rc, float& rd) {
    const float ab = a * b, cd = c * d;
    const float r = Ctx::rcp(ab * cd);
    const float rab = r * cd, rcd = r * ab;
    ra = rab * b; rb = rab * a; rc = rcd * d; rd = rcd * c;
  }
  FC_HD_CTX static float tanh_from_(float x, float rd) {
    // below |x| = 0.2: odd Taylor polynomial up to x^7 (next term 62/2835 x^9: relative 6e-8 at 0.2); above: 1 - 2 rd
    const float big = fmaf(-2.f, rd, 1.f);
    const float x2 = x * x;
    float pl = fmaf(x2, -0.053968253968253971f, 0.13333333333333333f);
    pl = fmaf(x2, pl, -0.33333333333333331f);
    pl = fmaf(x2 * x, pl, x);
    return fabsf(x) < 0.2f ? pl : big;
  }
  template <int NU>
  FC_HD_CTX static void tanh_batch(const float* x, float* y) {
    float d[NU], r[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) d[i] = denom_(2.f * kLog2e * x[i]);
#pragma unroll
    for (int i = 0; i + 3 < NU; i += 4) quad_rcp(d[i], d[i + 1], d[i + 2], d[i + 3], r[i], r[i + 1], r[i + 2], r[i + 3]);
    if ((NU & 3) == 2) {
      const float r2 = Ctx::rcp(d[NU - 2] * d[NU - 1]);
      r[NU - 2] = r2 * d[NU - 1]; r[NU - 1] = r2 * d[NU - 2];
    } else {
#pragma unroll
      for (int i = NU & ~3; i < NU; ++i) r[i] = Ctx::rcp(d[i]);
    }
#pragma unroll
    for (int i = 0; i < NU; ++i) y[i] = tanh_from_(x[i], r[i]);
  }
  struct ActK { float khi, klo, us, corr; };
  FC_HD_CTX static ActK make_actk(float unscale, float corr) {
    ActK k;
    k.khi = -kLog2e * unscale;                              // exact: unscale is a power of two
    k.klo = fmaf(k.khi, corr, -kLog2eLo * unscale);
    k.us = unscale; k.corr = corr;
    return k;
  }

  // ---------------------------------------------------------------------------------------------
  // hand-shakes
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void wait_full(int X) {
    if (scalar && !issuer) ctx.bar_wait_relaxed(kBarFull + X, ph >> (kBarFull + X));     // off the critical path
    else ctx.bar_wait(kBarFull + X, ph >> (kBarFull + X));
    ph ^= 1u << (kBarFull + X);
  }
  // all operand writes / accumulator reads of this thread for tile X are done
  FC_HD_CTX void arrive_ready(int X, bool smem_operand = true) {
    if (smem_operand) ctx.operand_fence();
    else ctx.tmem_fence();
    ctx.warp_sync();
    if (lane == 0) ctx.bar_arrive(kBarReady + X);            // one arrival per warp
  }
  FC_HD_CTX void wait_ready(int X) {                                                             // issuer warp
    ctx.bar_wait(kBarReady + X, ph >> (kBarReady + X));
    ph ^= 1u << (kBarReady + X);
  }
  FC_HD_CTX void wait_weights() { ctx.bar_wait(kBarWeightsQ, ph >> kBarWeightsQ); ph ^= 1u << kBarWeightsQ; }                    // issuer warp
  FC_HD_CTX void request_weights(bool bwd, int l) {                                              // tid 0 only
    const int n = bwd ? bwd_img_halves(l) : fwd_img_halves(l);   // hi + lo images of halves = that many floats
    ctx.bulk_load(sm + kSmWQ, p.wpack + (bwd ? wb_off(l) : wf_off(l)), n, kBarWeightsQ);
  }

  // ---------------------------------------------------------------------------------------------
  // shared-memory operand images: K-major, no swizzle: half (row r, k) at (k/8)*1024 + r*8 + k%8
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX float* op_ptr(int halves_off, int k) const {       // k multiple of 2
    return sm + kSmOpQ + ((halves_off + (k >> 3) * (kTileQ * 8) + row * 8 + (k & 7)) >> 1);
  }
  // the 10 owned units' values (scaled here) as fp16 hi/lo: one piece of 8 halves (units 0..7) + one pair (units 8, 9)
  struct Own16 { F4 hi8, lo8; float hi2, lo2; };
  // -> slots [8 grp, 8 grp + 8) and (40 + 2 grp, + 1) of the block that starts at k = kbase (multiple of 8)
  FC_HD_CTX void st_own(int img_hi, int img_lo, int kbase, const Own16& o) {
    Ctx::sts4(op_ptr(img_hi, kbase + 8 * grp), o.hi8);
    Ctx::sts4(op_ptr(img_lo, kbase + 8 * grp), o.lo8);
    Ctx::sts1(op_ptr(img_hi, kbase + 40 + 2 * grp), o.hi2);
    Ctx::sts1(op_ptr(img_lo, kbase + 40 + 2 * grp), o.lo2);
  }
  FC_HD_CTX void st_own_zero(int img_hi, int img_lo, int kbase) {
    const F4 z = {0.f, 0.f, 0.f, 0.f};
    Ctx::sts4(op_ptr(img_hi, kbase + 8 * grp), z);
    Ctx::sts4(op_ptr(img_lo, kbase + 8 * grp), z);
    Ctx::sts1(op_ptr(img_hi, kbase + 40 + 2 * grp), 0.f);
    Ctx::sts1(op_ptr(img_lo, kbase + 40 + 2 * grp), 0.f);
  }
  // zero padding slots 50..55 of a block (the last group's threads, one row each); disjoint from every slot a copy writes
  FC_HD_CTX void zero_pad_slots(int img_hi, int img_lo, int kbase) {
    Ctx::sts1(op_ptr(img_hi, kbase + 50), 0.f);
    Ctx::sts2(op_ptr(img_hi, kbase + 52), 0.f, 0.f);
    Ctx::sts1(op_ptr(img_lo, kbase + 50), 0.f);
    Ctx::sts2(op_ptr(img_lo, kbase + 52), 0.f, 0.f);
  }
  // pieces <-> the hidden-sequence scratch of the layer below: [t][warp][hi8 x32 | lo8 x32 | (hi2,lo2) x32], thread-private
  FC_HD_CTX float* seq_ptr(int X, int t) const { return w_seq(X) + (size_t)t * kSeqSlot + (size_t)uw * kSeqWarp; }
  FC_HD_CTX void stg_own(int X, int t, const Own16& o) {
    float* sq = seq_ptr(X, t);
    Ctx::stg4(sq + lane * 4, o.hi8);
    Ctx::stg4(sq + 128 + lane * 4, o.lo8);
    Ctx::stg2(sq + 256 + lane * 2, o.hi2, o.lo2);
  }
  // asynchronous copy scratch -> input block of the operand images (cp.async, no registers); cp_wait before the arrive
  FC_HD_CTX void copy_input(int X, int t) {
    const int img_hi = op_fwd_halves(X), img_lo = img_hi + kOpLoHalves;
    const float* sq = seq_ptr(X, t);
    Ctx::cp_async16(op_ptr(img_hi, 8 * grp), sq + lane * 4);
    Ctx::cp_async16(op_ptr(img_lo, 8 * grp), sq + 128 + lane * 4);
    Ctx::cp_async4(op_ptr(img_hi, 40 + 2 * grp), sq + 256 + lane * 2);
    Ctx::cp_async4(op_ptr(img_lo, 40 + 2 * grp), sq + 256 + lane * 2 + 1);
    Ctx::cp_commit();
  }

  // hi/lo fp16 split of 2*NP values (already in the scaled domain) into NP consecutive TMEM operand columns of the own lane
  template <int NP>
  FC_HD_CTX void st_pairs(int col_hi, int col_lo, const float* v) {
    float hi[NP], lo[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) Ctx::split_h2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);   // saturating conversion
    ctx.template tmem_st<NP>(col_hi, hi);
    ctx.template tmem_st<NP>(col_lo, lo);
  }
  // the same into 2*NP halves of the shared-memory dG image (tile 1), k0 multiple of 8, NP multiple of 4
  template <int NP>
  FC_HD_CTX void st_pairs_smem(int k0, const float* v) {
#pragma unroll
    for (int ch = 0; ch < NP / 4; ++ch) {
      float hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) Ctx::split_h2(v[ch * 8 + 2 * i], v[ch * 8 + 2 * i + 1], hi[i], lo[i]);
      Ctx::sts4(op_ptr(0, k0 + ch * 8), F4{hi[0], hi[1], hi[2], hi[3]});
      Ctx::sts4(op_ptr(kOpGLoHalves, k0 + ch * 8), F4{lo[0], lo[1], lo[2], lo[3]});
    }
  }

  // ---------------------------------------------------------------------------------------------
  // MMA issue (lane 0 of warp 0): 3 error-compensated terms, small ones first
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void issue_fwd(int X, int l, int ksteps) {
    const float* b_hi = sm + kSmWQ;
    const float* b_lo = b_hi + fwd_img_halves(l) / 2;
    const float* a_hi = sm + kSmOpQ + op_fwd_halves(X) / 2;
    const float* a_lo = a_hi + kOpLoHalves / 2;
    const int d = col_d_fwd(X);
    ctx.mma_ss(d, kNF, a_lo, b_hi, kNF, ksteps, false);
    ctx.mma_ss(d, kNF, a_hi, b_lo, kNF, ksteps, true);
    ctx.mma_ss(d, kNF, a_hi, b_hi, kNF, ksteps, true);
    ctx.commit(kBarFull + X);
  }
  FC_HD_CTX void issue_bwd(int X, int l) {
    const int nb = nb_of(l);
    const float* b_hi = sm + kSmWQ;
    const float* b_lo = b_hi + bwd_img_halves(l) / 2;
    const int d = col_d_bwd(X);
    if (X == 0) {
      ctx.mma(d, nb, kColGlo, b_hi, nb, 0, kKB / 16, false);
      ctx.mma(d, nb, kColGhi, b_lo, nb, 0, kKB / 16, true);
      ctx.mma(d, nb, kColGhi, b_hi, nb, 0, kKB / 16, true);
    } else {
      const float* a_hi = sm + kSmOpQ;
      const float* a_lo = a_hi + kOpGLoHalves / 2;
      ctx.mma_ss(d, nb, a_lo, b_hi, nb, kKB / 16, false);
      ctx.mma_ss(d, nb, a_hi, b_lo, nb, kKB / 16, true);
      ctx.mma_ss(d, nb, a_hi, b_hi, nb, kKB / 16, true);
    }
    ctx.commit(kBarFull + X);
  }

  // ---------------------------------------------------------------------------------------------
  // tile set-up (scalar-work thread of each trajectory)
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void load_tile(int X) {
    const int b = (tile0 + X) * kTileQ + row;
    const bool ok = b < p.B;
    float* rows = w_rows(X);
    for (int r = 0; r < kLook; ++r)
#pragma unroll
      for (int f = 0; f < kFeat; ++f) {
        float v = ok ? p.Z[(size_t)b * (kLook * kFeat) + r * kFeat + f] : 0.f;
        if (r == kLook - 1 && f == kFeat - 1) v = ok ? p.u0[b] : 0.f;              // Functions.py:1396
        rows[(size_t)(r * kFeat + f) * kTileQ + row] = v;
      }
    sm[kSmRefQ + X * kTileQ + row] = ok ? p.X[(size_t)b * 3 + 2] : 0.f;            // :1392
    float* cg = w_cost(X);
    cg[row] = 0.f; cg[kTileQ + row] = 0.f; cg[2 * kTileQ + row] = 0.f;
    if (ok) p.pred[(size_t)b * p.N] = p.u0[b];                                     // :1417-1418
  }

  // ---------------------------------------------------------------------------------------------
  // forward: cell update of one PAIR of units (chunk k of the thread's 10) from 8 raw accumulator columns.
  // Register economy is what this kernel lives on (80 registers, the cell state of both tiles resident): nothing is
  // buffered across pairs -- the gate activations of a unit leave as one float4 as soon as they exist, the previous
  // cell state of the pair as one float2 before it is overwritten, the new hidden values as one packed fp16 hi/lo word
  // each.  Returns h of the two units in h0, h1.
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void fwd_pair(const float* g, const ActK& ak, float* c2, float* rg4, float* rc2, float& h0, float& h1) {
    if (rc2) Ctx::stg2_stream(rc2, c2[0], c2[1]);
    float go[2], cn[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float xi = g[i * 4 + 0], xf = g[i * 4 + 1], xo = g[i * 4 + 3];
      float xg = g[i * 4 + 2] * ak.us;
      xg = fmaf(xg, ak.corr, xg);
      const float di = denom_(fmaf(xi, ak.khi, xi * ak.klo));
      const float df = denom_(fmaf(xf, ak.khi, xf * ak.klo));
      const float dq = denom_(fmaf(xo, ak.khi, xo * ak.klo));
      const float dg = denom_(2.f * kLog2e * xg);
      float gi, gf, rg;
      quad_rcp(di, df, dq, dg, gi, gf, go[i], rg);
      const float gg = tanh_from_(xg, rg);
      if (rg4) Ctx::stg4_stream(rg4 + i * 128, F4{gi, gf, gg, go[i]});
      cn[i] = fmaf(gf, c2[i], gi * gg);
      c2[i] = cn[i];
    }
    const float d0 = denom_(2.f * kLog2e * cn[0]), d1 = denom_(2.f * kLog2e * cn[1]);
    const float r = Ctx::rcp(d0 * d1);
    h0 = go[0] * tanh_from_(cn[0], r * d1);
    h1 = go[1] * tanh_from_(cn[1], r * d0);
  }

  // layer 0: the 5 row features of step t (scalar-work thread of the row): k = 0..7, 3 zero
  FC_HD_CTX void load_features(int X, int m, int t, float* xin) {
    const float* rp = w_rows(X) + (size_t)(m + t) * kFeat * kTileQ + row;
#pragma unroll
    for (int f = 0; f < kFeat; ++f) xin[f] = Ctx::ldcg(rp + f * kTileQ);
  }
  FC_HD_CTX void store_features(int X, const float* xin) {
    const int img_hi = op_fwd_halves(X), img_lo = img_hi + kOpLoHalves;
    float hi[4], lo[4];
    Ctx::split_h2(xin[0] * kScaleA, xin[1] * kScaleA, hi[0], lo[0]);
    Ctx::split_h2(xin[2] * kScaleA, xin[3] * kScaleA, hi[1], lo[1]);
    Ctx::split_h2(xin[4] * kScaleA, 0.f, hi[2], lo[2]);
    Ctx::sts4(op_ptr(img_hi, 0), F4{hi[0], hi[1], hi[2], 0.f});
    Ctx::sts4(op_ptr(img_lo, 0), F4{lo[0], lo[1], lo[2], 0.f});
  }

  // ---------------------------------------------------------------------------------------------
  // forward window: both tiles, interleaved cell step by cell step
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void fwd_prologue(int X, int l, int m) {
    // operand of step 0: zero recurrent block, input of step 0 (the previous MMA on this tile has been waited for)
    const int img_hi = op_fwd_halves(X), img_lo = img_hi + kOpLoHalves;
    lap(18);
    if (update) {
      const int krec = l == 0 ? kRec0 : kRec;
      st_own_zero(img_hi, img_lo, krec);
      if (grp == kGroups - 1) {                              // zero padding of both blocks (slots 50..55)
        zero_pad_slots(img_hi, img_lo, krec);
        if (l > 0) zero_pad_slots(img_hi, img_lo, 0);
      }
      if (l > 0) {
        copy_input(X, 0);
        Ctx::template cp_wait<0>();
      }
      arrive_ready(X);
    } else {
      if (l == 0) {
        float xin[kFeat];
        load_features(X, m, 0, xin);
        store_features(X, xin);
      }
      if (!issuer) arrive_ready(X);
      else if (l == 0) { ctx.operand_fence(); ctx.warp_sync(); }   // all 32 feature rows of the issuer warp visible to the tensor core
    }
    if (issuer) {                                            // all lanes follow, lane 0 issues
      wait_ready(X);
      if (X == 0) wait_weights();                            // operand image of this layer landed
      if (lane == 0) issue_fwd(X, l, l == 0 ? 1 : 4);        // recurrent part is zero: only the k-steps over the input
      ctx.warp_sync();
    }
    lap(14);
  }

  // issuer warp (warp 0, all 32 lanes stay together; lane 0 issues): one step of the issue loop
  FC_HD_CTX void fwd_item_issuer(int X, int l, int m, int t, bool more_after) {
    float xin[kFeat];
    lap(23);
    if (l == 0 && t + 1 < kLook) load_features(X, m, t + 1, xin);     // quadrant 0's layer-0 features of the next step
    wait_full(X);                                            // MMA(X, l, t) complete: operand free
    lap(1);
    if (t == kLook - 1 && X == ntl - 1 && lane == 0) {       // stream the next weight image under the cell updates
      if (l + 1 < kLayers) request_weights(false, l + 1);
      else if (m + 1 < p.N) request_weights(false, 0);
      else if (p.with_grad) request_weights(true, kLayers - 1);
      else if (more_after) request_weights(false, 0);
    }
    if (t + 1 < kLook) {
      if (l == 0) { store_features(X, xin); ctx.operand_fence(); ctx.warp_sync(); }
      wait_ready(X);
      lap(4);
      if (lane == 0) issue_fwd(X, l, kf_of(l) / 16);
      ctx.warp_sync();
      lap(5);
    }
  }

  // warps 1..3: layer-0 features of their rows, in step with the cell-update warps
  FC_HD_CTX void fwd_item_scalar(int X, int l, int m, int t) {
    float xin[kFeat];
    if (l == 0 && t + 1 < kLook) load_features(X, m, t + 1, xin);
    wait_full(X);
    if (t + 1 < kLook) {
      if (l == 0) store_features(X, xin);
      arrive_ready(X);
    }
  }

  template <int X>
  FC_HD_CTX void fwd_item(int l, int m, int t) {
    float (&c)[kOwn] = X ? cB : cA;
    const int tmin = t_min_of(m);
    lap(16);
    float* rp = nullptr;                                     // this thread's slots of the activation record of (m, l, t)
    if (p.with_grad && t >= tmin)
      rp = w_rec(X) + (size_t)(rec_base(m) + (long)l * steps_kept(m) + (t - tmin)) * kRecFloatsQ + (size_t)uw * kRecWarp;
    const int ksteps = t == 0 ? (l == 0 ? 1 : 4) : kf_of(l) / 16;
    const float corr = Ctx::kAccTruncates ? acc_correction(ksteps, p.acc_comp) : 0.0f;
    const ActK ak = make_actk(1.0f / (kScaleA * kScaleW), corr);
    const int img_hi = op_fwd_halves(X), img_lo = img_hi + kOpLoHalves;
    const bool read_out_step = l == kLayers - 1 && t + 1 == kLook;
    const int col0 = col_d_fwd(X) + 4 * kOwn * grp;
    lap(17);
    wait_full(X);                                            // accumulator complete; operand free
    lap(1);
    if (l > 0 && t + 1 < kLook) copy_input(X, t + 1);        // input block of step t+1: lands during the cell update
    if (t == 0) {
#pragma unroll
      for (int j = 0; j < kOwn; ++j) c[j] = 0.f;             // zero state (LSTMModel.initialize_hidden_states, :331-351)
    }
    float hi[kOwn / 2], lo[kOwn / 2];                        // packed fp16 hi / lo words of the five unit pairs
    float xq[4] = {0.f, 0.f, 0.f, 0.f};
    float g[2][8];
    ctx.template tmem_ld_nowait<8>(col0, g[0]);
#pragma unroll
    for (int k = 0; k < kOwn / 2; ++k) {
      ctx.tmem_ld_wait();
      if (k + 1 < kOwn / 2) ctx.template tmem_ld_nowait<8>(col0 + 8 * (k + 1), g[(k + 1) & 1]);   // next pair requested first
      float h0, h1;
      fwd_pair(g[k & 1], ak, c + 2 * k, rp ? rp + (size_t)(2 * k) * 128 + lane * 4 : nullptr,
               rp ? rp + kRecCp + k * 64 + lane * 2 : nullptr, h0, h1);
      if (read_out_step) {                                   // read-out partial sums over the owned units (Functions.py:377)
        const float* fw = sm + kSmSmallQ + kOwn * grp + 2 * k;
#pragma unroll
        for (int q = 0; q < 4; ++q) xq[q] = fmaf(fw[q * kHid + 1], h1, fmaf(fw[q * kHid], h0, xq[q]));
      }
      Ctx::split_h2(h0 * kScaleA, h1 * kScaleA, hi[k], lo[k]);
    }
    if (l + 1 < kLayers || t + 1 < kLook) {
      Own16 o;
      o.hi8 = F4{hi[0], hi[1], hi[2], hi[3]};
      o.lo8 = F4{lo[0], lo[1], lo[2], lo[3]};
      o.hi2 = hi[4]; o.lo2 = lo[4];
      if (l + 1 < kLayers) stg_own(X, t, o);                 // input of the layer above, already in operand format
      if (t + 1 < kLook) st_own(img_hi, img_lo, l == 0 ? kRec0 : kRec, o);
    }
    if (t + 1 < kLook) {
      if (l > 0) Ctx::template cp_wait<0>();
      lap(2);
      arrive_ready(X);
      lap(3);
    } else if (read_out_step) {
      // handed to the scalar-work thread of the row through spare TMEM columns of the own lane
      ctx.template tmem_st<4>(kColFcp + 4 * kGroups * X + 4 * grp, xq);
      ctx.tmem_st_wait();
    }
    if (t + 1 == kLook) lap(13);
  }

  FC_HD_CTX void fwd_window(int m, bool more_after) {
    for (int l = 0; l < kLayers; ++l) {
      for (int X = 0; X < ntl; ++X) fwd_prologue(X, l, m);
      for (int t = 0; t < kLook; ++t) {
        if (issuer) {
          for (int X = 0; X < ntl; ++X) fwd_item_issuer(X, l, m, t, more_after);
        } else if (update) {
          fwd_item<0>(l, m, t);
          if (ntl > 1) fwd_item<1>(l, m, t);
        } else {
          for (int X = 0; X < ntl; ++X) fwd_item_scalar(X, l, m, t);
        }
      }
    }
    lap(19);
    ctx.tc_sync();                                           // read-out partial sums visible
    if (scalar)
      for (int X = 0; X < ntl; ++X) {
        if (p.shadow) shadow_glue(X, m);
        else fwd_glue(X, m);
      }
    lap(15);
  }

  // read-out of window m: sum of the five groups' partial sums + bias (scalar-work thread of the row)
  FC_HD_CTX void read_out(int X, float* x) {
    const float* sw = sm + kSmSmallQ;
    float fp[4 * kGroups];
    ctx.template tmem_ld_nowait<16>(kColFcp + 4 * kGroups * X, fp);
    ctx.template tmem_ld_nowait<4>(kColFcp + 4 * kGroups * X + 16, fp + 16);
    ctx.tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) x[q] = ((((fp[q] + fp[4 + q]) + fp[8 + q]) + fp[12 + q]) + fp[16 + q]) + sw[(kFCB - kFCW) + q];
  }

  // ---------------------------------------------------------------------------------------------
  // LSTM shadow roll-out (Functions.py:969-1011, 1196-1231): the window starts as ten copies of the first row; after
  // window m the surrogate output is logged and [output * scale_out / scale_in, u_{m+1}] becomes the newest row
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void load_tile_shadow(int X) {
    const int b = (tile0 + X) * kTileQ + row;
    const bool ok = b < p.B;
    float* rows = w_rows(X);
#pragma unroll
    for (int f = 0; f < kFeat; ++f) {
      const float v = ok ? p.sh_row0[(size_t)b * kFeat + f] : 0.f;
      for (int r = 0; r < kLook; ++r) rows[(size_t)(r * kFeat + f) * kTileQ + row] = v;
    }
  }
  FC_HD_CTX void shadow_glue(int X, int m) {
    float x[4];
    read_out(X, x);
    const int b = (tile0 + X) * kTileQ + row;
    const bool ok = b < p.B;
    float* rnew = w_rows(X) + (size_t)(kLook + m) * kFeat * kTileQ + row;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (ok) p.sh_y[((size_t)b * p.N + m) * 4 + q] = x[q];
      rnew[q * kTileQ] = x[q] * p.sh_ratio[q];
    }
    rnew[4 * kTileQ] = (ok && m + 1 < p.N) ? p.sh_u[(size_t)b * p.N + m + 1] : 0.f;
  }

  // ---------------------------------------------------------------------------------------------
  // after window m (scalar-work thread of each trajectory): read-out, cost terms, next command
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void fwd_glue(int X, int m) {
    const float* sw = sm + kSmSmallQ;
    float x[4];
    read_out(X, x);
    if (p.noise_std > 0.f) {                                                   // enable_noise, :1400-1402 / :1438-1440
      float e[4];
      philox_normal4(p.noise_seed, (unsigned)((tile0 + X) * kTileQ + row), (unsigned)m, e);
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = fmaf(p.noise_std, e[q], x[q]);
    }
    float* rows = w_rows(X);
    const float ref = sm[kSmRefQ + X * kTileQ + row];
    const float ucur = Ctx::ldcg(rows + (size_t)((kLook - 1 + m) * kFeat + 4) * kTileQ + row);
    const float uprev = Ctx::ldcg(rows + (size_t)((kLook - 2 + m) * kFeat + 4) * kTileQ + row);
    float du = uprev - ucur;
    float cmd = p.alpha * du * du;                                             // :1405 / :1446
    float er = (x[0] - ref) * (x[0] - ref);                                    // :1408 / :1443
    float con = fmaxf(-x[1], 0.f) + fmaxf(-x[2], 0.f) + fmaxf(x[1] - kP1Max, 0.f) + fmaxf(x[2] - kP2Max, 0.f);
    float* cg = w_cost(X);
    cg[row] = Ctx::ldcg(cg + row) + ((er + cmd) + con);                        // :1414 / :1452
    cg[kTileQ + row] = Ctx::ldcg(cg + kTileQ + row) + cmd;
    cg[2 * kTileQ + row] = Ctx::ldcg(cg + 2 * kTileQ + row) + er;
    float* rnew = rows + (size_t)(kLook + m) * kFeat * kTileQ + row;           // rho_{10+m} = [x_{m+1}, u_{m+1}]
#pragma unroll
    for (int q = 0; q < 4; ++q) rnew[q * kTileQ] = x[q];
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
      int b = (tile0 + X) * kTileQ + row;
      if (b < p.B) p.pred[(size_t)b * p.N + m + 1] = unext;                    // :1455
    }
    rnew[4 * kTileQ] = unext;
  }

  // ---------------------------------------------------------------------------------------------
  // before the reverse sweep of window m (scalar-work thread per trajectory, then 200 accumulation threads)
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void bwd_glue_tile(int X, int m) {
    const int k = m + 1;
    const float s = p.grad_scale;
    const bool has_u = k <= p.N - 1;
    const float* sw = sm + kSmSmallQ;
    const float* iw = sw + (kINPW - kFCW);
    const float* ib = sw + (kINPB - kFCW);
    const float* ow = sw + (kOUTW - kFCW);
    const bool valid = (tile0 + X) * kTileQ + row < p.B;
    const float* rows = w_rows(X);
    const float* rx = rows + (size_t)(kLook + m) * kFeat * kTileQ + row;
    float x0 = Ctx::ldcg(rx), x1 = Ctx::ldcg(rx + kTileQ), x2 = Ctx::ldcg(rx + 2 * kTileQ), x3 = Ctx::ldcg(rx + 3 * kTileQ);
    const float ref = sm[kSmRefQ + X * kTileQ + row];
    float g0 = 2.f * (x0 - ref) * s;
    float g1 = s * ((x1 > kP1Max ? 1.f : 0.f) - (x1 < 0.f ? 1.f : 0.f));
    float g2 = s * ((x2 > kP2Max ? 1.f : 0.f) - (x2 < 0.f ? 1.f : 0.f));
    float g3 = 0.f;
    float dv = 0.f;
    if (has_u) {
      const float* gr = w_grow(X) + (size_t)k * kFeat * kTileQ + row;
      float uk = Ctx::ldcg(rows + (size_t)((kLook - 1 + k) * kFeat + 4) * kTileQ + row);
      float ukm1 = Ctx::ldcg(rows + (size_t)((kLook - 2 + k) * kFeat + 4) * kTileQ + row);
      float gu = Ctx::ldcg(gr + 4 * kTileQ) - 2.f * p.alpha * (ukm1 - uk) * s;
      if (k + 1 <= p.N - 1) {
        float ukp1 = Ctx::ldcg(rows + (size_t)((kLook + k) * kFeat + 4) * kTileQ + row);
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
      g1 += Ctx::ldcg(gr + kTileQ);
      g2 += Ctx::ldcg(gr + 2 * kTileQ);
      g3 += d1 + Ctx::ldcg(gr + 3 * kTileQ);
      sm[kSmDvQ + X * kTileQ + row] = dv;
      sm[kSmFinQ + (X * 2) * kTileQ + row] = x0;
      sm[kSmFinQ + (X * 2 + 1) * kTileQ + row] = x3;
    }
    if (!valid) { g0 = g1 = g2 = g3 = 0.f; }
    sm[kSmGxQ + (X * 4 + 0) * kTileQ + row] = g0;
    sm[kSmGxQ + (X * 4 + 1) * kTileQ + row] = g1;
    sm[kSmGxQ + (X * 4 + 2) * kTileQ + row] = g2;
    sm[kSmGxQ + (X * 4 + 3) * kTileQ + row] = g3;
  }
  FC_HD_CTX void bwd_glue(int m) {
    const bool has_u = m + 1 <= p.N - 1;
    const float* sw = sm + kSmSmallQ;
    const float* iw = sw + (kINPW - kFCW);
    const float* ib = sw + (kINPB - kFCW);
    const float* ow = sw + (kOUTW - kFCW);
    if (scalar)
      for (int X = 0; X < ntl; ++X) bwd_glue_tile(X, m);
    ctx.sync();
    // controller weight gradients, unit-parallel: 200 threads of the cell-update warps (warps 0..3 just worked)
    const int gt = tid - 128;
    if (has_u && gt >= 0 && gt < 4 * kFnnHid) {
      const int u = gt % kFnnHid, part = gt / kFnnHid;
      double a_ow = 0.0, a_b = 0.0, a_w0 = 0.0, a_w1 = 0.0, a_w2 = 0.0;   // batch sums cancel heavily: fp64
      const float w0 = iw[u * 3 + 0], w1 = iw[u * 3 + 1], w2 = iw[u * 3 + 2], bb = ib[u], owu = ow[u];
      for (int X = 0; X < ntl; ++X)
        for (int tr = part * 32; tr < part * 32 + 32; ++tr) {
          float dv = sm[kSmDvQ + X * kTileQ + tr];
          float x0 = sm[kSmFinQ + (X * 2) * kTileQ + tr], x3 = sm[kSmFinQ + (X * 2 + 1) * kTileQ + tr];
          float ref = sm[kSmRefQ + X * kTileQ + tr];
          float pre = fmaf(w2, ref, fmaf(w1, x3, fmaf(w0, x0, bb)));
          a_ow += (double)dv * (double)fmaxf(pre, 0.f);
          float dp = pre > 0.f ? dv * owu : 0.f;
          a_b += dp;
          a_w0 += (double)dp * (double)x0;
          a_w1 += (double)dp * (double)x3;
          a_w2 += (double)dp * (double)ref;
        }
      double* pg = reinterpret_cast<double*>(sm + kSmPgQ) + part * kNumFnnGrad;
      pg[u * 3 + 0] += a_w0;
      pg[u * 3 + 1] += a_w1;
      pg[u * 3 + 2] += a_w2;
      pg[150 + u] += a_b;
      pg[200 + u] += a_ow;
    }
  }

  // ---------------------------------------------------------------------------------------------
  // backward cell gradient of NU unit slots: record(t) -> factors, dh -> d(cell), d(gates)
  //   A = o(1-tanh^2 c), Ko = tanh(c) o(1-o), Ki = g i(1-i), Kf = c_prev f(1-f), Kg = i(1-g^2)
  //   dct = dc + dh A; dG = (dct Ki, dct Kf, dct Kg, dh Ko); dc = dct f
  // ---------------------------------------------------------------------------------------------
  // one pair of units: gate activations (two float4) + previous cell states (one float2) of the record, d(h) of the
  // two units -> d(cell) in c2, the eight gate gradients in dg
  FC_HD_CTX static void bwd_pair(const F4& r0, const F4& r1, float cp0, float cp1, float dh0, float dh1, float* c2, float* dg) {
    const float cn0 = fmaf(r0.y, cp0, r0.x * r0.z), cn1 = fmaf(r1.y, cp1, r1.x * r1.z);
    const float d0 = denom_(2.f * kLog2e * cn0), d1 = denom_(2.f * kLog2e * cn1);
    const float r = Ctx::rcp(d0 * d1);
    const float t0 = tanh_from_(cn0, r * d1), t1 = tanh_from_(cn1, r * d0);
    // s(1-s) and 1-t^2 as single fused operations: fmaf(-s, s, s), fmaf(-t, t, 1)
    {
      const float gi = r0.x, gf = r0.y, gg = r0.z, go = r0.w;
      const float A = go * fmaf(-t0, t0, 1.f);
      const float dct = fmaf(dh0, A, c2[0]);
      c2[0] = dct * gf;
      dg[0] = dct * (gg * fmaf(-gi, gi, gi));
      dg[1] = dct * (cp0 * fmaf(-gf, gf, gf));
      dg[2] = dct * (gi * fmaf(-gg, gg, 1.f));
      dg[3] = dh0 * (t0 * fmaf(-go, go, go));
    }
    {
      const float gi = r1.x, gf = r1.y, gg = r1.z, go = r1.w;
      const float A = go * fmaf(-t1, t1, 1.f);
      const float dct = fmaf(dh1, A, c2[1]);
      c2[1] = dct * gf;
      dg[4] = dct * (gg * fmaf(-gi, gi, gi));
      dg[5] = dct * (cp1 * fmaf(-gf, gf, gf));
      dg[6] = dct * (gi * fmaf(-gg, gg, 1.f));
      dg[7] = dh1 * (t1 * fmaf(-go, go, go));
    }
  }

  FC_HD_CTX void prefetch_record(const float* rec_in) {
    // the warp's record slots are contiguous (6400 B): one bulk prefetch from lane 0
    if (lane == 0) Ctx::prefetch_l2_bulk(rec_in + (size_t)uw * kRecWarp, (unsigned)kRecWarp * 4u);
  }

  // d(h) of the unit pair k at step t that does not come from the recurrent MMA: the layer above (thread-private
  // scratch) or, for the top layer at the last step, the read-out (Functions.py:377)
  FC_HD_CTX void bwd_extra(int X, int l, int t, int k, float& e0, float& e1) {
    if (l == kLayers - 1) {
      e0 = e1 = 0.f;
      if (t == kLook - 1) {
        const float* fw = sm + kSmSmallQ + kOwn * grp + 2 * k;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float gx = sm[kSmGxQ + (X * 4 + q) * kTileQ + row] * p.g_scale;               // into the scaled domain
          e0 = fmaf(fw[q * kHid], gx, e0);
          e1 = fmaf(fw[q * kHid + 1], gx, e1);
        }
      }
    } else {
      const float* dsq = w_dseq(X) + (size_t)t * kDseqSlot + (size_t)uw * kOwn * 32 + lane + (size_t)(2 * k) * 32;
      e0 = Ctx::ldcg(dsq);
      e1 = Ctx::ldcg(dsq + 32);
    }
  }

  // result of MMA(X, l, t) for the unit pair k: d(input) of step t -> the layer below, returns d(h_prev)
  FC_HD_CTX void bwd_collect_pair(int X, int l, int t, int k, float& dh0, float& dh1) {
    const float corr_b = Ctx::kAccTruncates ? acc_correction(kKB / 16, p.acc_comp) : 0.0f;
    // the whole reverse sweep of a window runs in the scaled domain (gradients x g_scale, an exact power of two):
    // the gate gradients then need no multiplication before their fp16 split; only the weight scale is removed here
    const float unscale_b = 1.0f / kScaleW;
    const int dcol = col_d_bwd(X);
    if (l > 0) {
      float d[4];
      ctx.template tmem_ld<4>(dcol + 2 * kOwn * grp + 4 * k, d);
#pragma unroll
      for (int j = 0; j < 4; ++j) { d[j] *= unscale_b; d[j] = fmaf(d[j], corr_b, d[j]); }
      float* dq = w_dseq(X) + (size_t)t * kDseqSlot + (size_t)uw * kOwn * 32 + lane + (size_t)(2 * k) * 32;
      dq[0] = d[0];                                         // d(input unit) -> the layer below, same thread
      dq[32] = d[1];
      dh0 = d[2]; dh1 = d[3];
    } else {
      float d[2];
      ctx.template tmem_ld<2>(dcol + kOwn * grp + 2 * k, d);
      d[0] *= unscale_b; d[1] *= unscale_b;
      dh0 = fmaf(d[0], corr_b, d[0]); dh1 = fmaf(d[1], corr_b, d[1]);
    }
  }
  // layer 0: gradient of the row features of step t (scalar-work thread of the row)
  FC_HD_CTX void bwd_collect_features(int X, int m, int t) {
    const float corr_b = Ctx::kAccTruncates ? acc_correction(kKB / 16, p.acc_comp) : 0.0f;
    const float unscale_b = p.g_unscale / kScaleW;
    const int kr = m + t - (kLook - 1);                    // gradient of row rho_{9+kr}
    if (kr >= 0) {
      float df[8];
      ctx.template tmem_ld<8>(col_d_bwd(X) + 56, df);
      float* gp = w_grow(X) + (size_t)kr * kFeat * kTileQ + row;
#pragma unroll
      for (int f = 0; f < kFeat; ++f) {
        const float v = df[f] * unscale_b;
        Ctx::red_add(gp + f * kTileQ, fmaf(v, corr_b, v));   // same thread, one add per step: deterministic order
      }
    }
  }

  // issuer warp: one step of the issue loop
  FC_HD_CTX void bwd_item_issuer(int X, int l, int m, int t) {
    if (t < kLook - 1) {
      lap(23);
      wait_full(X);                                        // MMA(X, l, t+1) complete
      lap(6);
      if (l == 0) { bwd_collect_features(X, m, t + 1); ctx.warp_sync(); }   // quadrant 0's row-feature gradients, before the accumulator is reused
    }
    wait_ready(X);
    if (X == 0 && t == kLook - 1) wait_weights();          // backward image of this layer landed
    lap(9);
    if (lane == 0) issue_bwd(X, l);
    ctx.warp_sync();
    lap(10);
  }

  // warps 1..3: row-feature gradients of layer 0, otherwise only the phase counts of the hand-shakes
  FC_HD_CTX void bwd_item_scalar(int X, int l, int m, int t) {
    if (t < kLook - 1) {
      wait_full(X);
      if (l == 0) bwd_collect_features(X, m, t + 1);
    }
    arrive_ready(X, false);                                  // nothing written to shared memory
  }

  template <int X>
  FC_HD_CTX void bwd_item(int l, int m, int t) {
    float (&c)[kOwn] = X ? cB : cA;
    const int tmin = t_min_of(m);
    lap(20);
    const float* rec_l = w_rec(X) + (size_t)(rec_base(m) + (long)l * steps_kept(m)) * kRecFloatsQ;
    // HBM -> L2 for the next step of this tile (or the first step of the next layer / window)
    if (t - 1 >= tmin) prefetch_record(rec_l + (size_t)(t - 1 - tmin) * kRecFloatsQ);
    else if (l > 0) prefetch_record(w_rec(X) + (size_t)(rec_base(m) + (long)(l - 1) * steps_kept(m) + (kLook - 1 - tmin)) * kRecFloatsQ);
    else if (m > 0) prefetch_record(w_rec(X) + (size_t)(rec_base(m - 1) + (long)(kLayers - 1) * steps_kept(m - 1) + (kLook - 1 - t_min_of(m - 1))) * kRecFloatsQ);
    const float* rp = rec_l + (size_t)(t - tmin) * kRecFloatsQ + (size_t)uw * kRecWarp;
    const float* rg = rp + lane * 4;                       // gate activations: unit j at + j*128
    const float* rc = rp + kRecCp + lane * 2;              // previous cell states: pair k at + k*64
    // software pipeline over the five unit pairs: the record of pair k+1 and its d(h) from the layer above are
    // requested before pair k is worked on; pair 0 is in flight during the wait for the MMA
    F4 ra[2], rb[2];
    float cpa[2], cpb[2], ex[2][2];
    ra[0] = Ctx::ldg4_stream(rg); rb[0] = Ctx::ldg4_stream(rg + 128);
    Ctx::ldg2_stream(rc, cpa[0], cpb[0]);
    bwd_extra(X, l, t, 0, ex[0][0], ex[0][1]);
    const bool top = t == kLook - 1;
    if (top) {
#pragma unroll
      for (int j = 0; j < kOwn; ++j) c[j] = 0.f;
    } else {
      lap(12);
      wait_full(X);                                        // MMA(X, l, t+1) complete
      lap(6);
    }
    const int kg = 4 * kOwn * grp;                         // first gate-gradient index of the group (multiple of 8)
#pragma unroll
    for (int k = 0; k < kOwn / 2; ++k) {
      const int cur = k & 1, nxt = cur ^ 1;
      if (k + 1 < kOwn / 2) {
        ra[nxt] = Ctx::ldg4_stream(rg + (size_t)(2 * k + 2) * 128);
        rb[nxt] = Ctx::ldg4_stream(rg + (size_t)(2 * k + 3) * 128);
        Ctx::ldg2_stream(rc + (k + 1) * 64, cpa[nxt], cpb[nxt]);
        bwd_extra(X, l, t, k + 1, ex[nxt][0], ex[nxt][1]);
      }
      float dh0 = ex[cur][0], dh1 = ex[cur][1];
      if (!top) {
        float m0, m1;
        bwd_collect_pair(X, l, t + 1, k, m0, m1);
        dh0 += m0; dh1 += m1;
      }
      float dg[8], hi[4], lo[4];
      bwd_pair(ra[cur], rb[cur], cpa[cur], cpb[cur], dh0, dh1, c + 2 * k, dg);
#pragma unroll
      for (int i = 0; i < 4; ++i) Ctx::split_h2(dg[2 * i], dg[2 * i + 1], hi[i], lo[i]);       // saturating conversion
      if (X == 0) {
        ctx.template tmem_st<4>(kColGhi + (kg >> 1) + 4 * k, hi);
        ctx.template tmem_st<4>(kColGlo + (kg >> 1) + 4 * k, lo);
      } else {
        Ctx::sts4(op_ptr(0, kg + 8 * k), F4{hi[0], hi[1], hi[2], hi[3]});
        Ctx::sts4(op_ptr(kOpGLoHalves, kg + 8 * k), F4{lo[0], lo[1], lo[2], lo[3]});
      }
    }
    if (X == 0) ctx.tmem_st_wait();
    lap(7);
    arrive_ready(X, X != 0);                                 // tile 0: the operand went to TMEM, no shared-memory writes
    lap(8);
  }

  // after the last step of a layer: collect the result of MMA(X, l, tmin)
  FC_HD_CTX void bwd_tail(int X, int l, int m, bool more_after) {
    const int tmin = t_min_of(m);
    lap(22);
    wait_full(X);
    lap(6);
    if (X == ntl - 1 && tid == 0) {                        // all MMAs that read this image are complete (issuer lane)
      if (l > 0) request_weights(true, l - 1);
      else if (m > 0) request_weights(true, kLayers - 1);
      else if (more_after) request_weights(false, 0);
    }
    if (update) {
      if (l > 0) {
#pragma unroll
        for (int k = 0; k < kOwn / 2; ++k) {
          float m0, m1;
          bwd_collect_pair(X, l, tmin, k, m0, m1);         // d(h) before the first kept step is not needed
        }
      }
    } else {
      if (l == 0) bwd_collect_features(X, m, tmin);
    }
    lap(14);
  }

  FC_HD_CTX void bwd_window(int m, bool more_after) {
    const int tmin = t_min_of(m);
    lap(21);
    bwd_glue(m);
    ctx.sync();
    lap(15);
    for (int l = kLayers - 1; l >= 0; --l) {
      for (int t = kLook - 1; t >= tmin; --t) {
        if (issuer) {
          for (int X = 0; X < ntl; ++X) bwd_item_issuer(X, l, m, t);
        } else if (update) {
          bwd_item<0>(l, m, t);
          if (ntl > 1) bwd_item<1>(l, m, t);
        } else {
          for (int X = 0; X < ntl; ++X) bwd_item_scalar(X, l, m, t);
        }
      }
      for (int X = 0; X < ntl; ++X) bwd_tail(X, l, m, more_after);
    }
    ctx.tc_sync();                                         // all accumulator reads of this window done
  }

  // ---------------------------------------------------------------------------------------------
  // per-pass epilogues (scalar-work threads)
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void store_costs() {
    if (scalar) {
      for (int X = 0; X < kTilesQ; ++X) {
        float mine = 0.f;
        int b = (tile0 + X) * kTileQ + row;
        if (X < ntl && b < p.B) {
          const float inv = 1.f / (float)p.N;
          const float* cg = w_cost(X);
          mine = Ctx::ldcg(cg + row) * inv;                                    // :1458-1460
          p.cost[b] = mine;
          p.command[b] = Ctx::ldcg(cg + kTileQ + row) * inv;
          p.error[b] = Ctx::ldcg(cg + 2 * kTileQ + row) * inv;
        }
        sm[kSmGxQ + X * kTileQ + row] = mine;              // gx area is free between the sweeps
      }
    }
    ctx.sync();
    if (tid == 0) {
      double acc = 0.0;
      for (int i = 0; i < kTilesQ * kTileQ; ++i) acc += (double)sm[kSmGxQ + i];
      *reinterpret_cast<double*>(sm + kSmRedQ) += acc;
    }
    ctx.sync();
  }

  FC_HD_CTX void store_du0() {
    if (scalar)
      for (int X = 0; X < ntl; ++X) {
        int b = (tile0 + X) * kTileQ + row;
        if (b < p.B) {
          const float s = p.grad_scale;
          const float* rows = w_rows(X);
          float u0 = Ctx::ldcg(rows + (size_t)((kLook - 1) * kFeat + 4) * kTileQ + row);
          float um1 = Ctx::ldcg(rows + (size_t)((kLook - 2) * kFeat + 4) * kTileQ + row);
          float g = Ctx::ldcg(w_grow(X) + 4 * kTileQ + row) - 2.f * p.alpha * (um1 - u0) * s;
          if (p.N > 1) {
            float u1 = Ctx::ldcg(rows + (size_t)(kLook * kFeat + 4) * kTileQ + row);
            g += 2.f * p.alpha * (u0 - u1) * s;
          }
          p.du0[b] = g;
        }
      }
  }

  // zero padding of the dG operands (k = 200..207): TMEM columns 100..103 of tile 0, the last 16-byte piece of tile 1
  FC_HD_CTX void zero_dg_padding() {
    if (scalar) {
      float z[4] = {0.f, 0.f, 0.f, 0.f};
      ctx.template tmem_st<4>(kColGhi + 100, z);
      ctx.template tmem_st<4>(kColGlo + 100, z);
      ctx.tmem_st_wait();
      const F4 z4 = {0.f, 0.f, 0.f, 0.f};
      Ctx::sts4(op_ptr(0, kGates), z4);
      Ctx::sts4(op_ptr(kOpGLoHalves, kGates), z4);
    }
  }

  // ---------------------------------------------------------------------------------------------
  // persistent loop over tile pairs
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void run() {
    ctx.tc_setup(sm + kSmBarQ);
#ifdef FC_TC_TIMING
    tlast = Ctx::clock();
#endif
    if (tid == 0) {
      ctx.bar_init(kBarReady, kUpdWarpsQ + kScalarWarpsQ);   // cell-update warps + scalar warps
      ctx.bar_init(kBarReady + 1, kUpdWarpsQ + kScalarWarpsQ);
    }
    for (int i = tid; i < kSmallFloats; i += kThreadsQ) sm[kSmSmallQ + i] = p.wpack[kSmallOff + i];
    for (int i = tid; i < 4 * kNumFnnGrad; i += kThreadsQ) reinterpret_cast<double*>(sm + kSmPgQ)[i] = 0.0;
    if (tid == 0) *reinterpret_cast<double*>(sm + kSmRedQ) = 0.0;
    ctx.bar_init_fence();
    ctx.sync();
    const int npairs = (p.num_tiles + kTilesQ - 1) / kTilesQ;
    if (tid == 0 && ctx.bid() < npairs) request_weights(false, 0);
    for (int pp = ctx.bid(); pp < npairs; pp += ctx.nblk()) {
      const bool more = pp + ctx.nblk() < npairs;
      tile0 = pp * kTilesQ;
      ntl = p.num_tiles - tile0 < kTilesQ ? p.num_tiles - tile0 : kTilesQ;
      if (scalar)
        for (int X = 0; X < ntl; ++X) {
          if (p.shadow) load_tile_shadow(X);
          else load_tile(X);
        }
      ctx.sync();
      for (int m = 0; m < p.N; ++m) fwd_window(m, more);
      if (!p.shadow) store_costs();
      if (p.with_grad) {
        if (scalar)
          for (int X = 0; X < ntl; ++X) {
            float* grow = w_grow(X);
            for (int k = 0; k < p.N; ++k)
#pragma unroll
              for (int f = 0; f < kFeat; ++f) grow[(size_t)(k * kFeat + f) * kTileQ + row] = 0.f;
          }
        zero_dg_padding();
        ctx.sync();
        for (int m = p.N - 1; m >= 0; --m) bwd_window(m, more);
        store_du0();
      }
      ctx.tc_sync();
    }
    double* part = p.partial + (size_t)ctx.bid() * kPartialStride;
    const double* pgd = reinterpret_cast<const double*>(sm + kSmPgQ);
    for (int i = tid; i < kNumFnnGrad; i += kThreadsQ)
      part[i] = (pgd[i] + pgd[kNumFnnGrad + i]) + (pgd[2 * kNumFnnGrad + i] + pgd[3 * kNumFnnGrad + i]);
    if (tid == 0) part[kNumFnnGrad] = *reinterpret_cast<const double*>(sm + kSmRedQ);
    ctx.sync();
#ifdef FC_TC_TIMING
    if (p.debug_timing && (tid == 0 || tid == 160 || tid == 736) && ctx.bid() == 0) Ctx::report_pair(tid, tm);
#endif
    ctx.tc_teardown();
  }
};

}  // namespace q5
}  // namespace fc
