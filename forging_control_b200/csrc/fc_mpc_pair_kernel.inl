// Two-tile ("pair") tcgen05 variant of the fused MPC-loss forward + reverse sweep.  Same mathematics and reference
// citations as fc_mpc_kernel.inl (MPCLoss.forward Functions.py:1353-1472, LSTMModel.forward :353-379,
// FNNModel.forward :261-289, loss.backward() :655) and the same fp16 hi/lo split tensor-core contraction as
// fc_mpc_tc_kernel.inl.  What changes is the schedule: a CTA owns TWO 128-trajectory tiles and alternates between
// them, LSTM cell step by cell step.  The gate contraction of one tile (tcgen05.mma, asynchronous) runs while all
// 16 warps do the cell update (MUFU / FMA pipes) of the other tile, instead of the two phases taking turns:
//
//   thread 0 :  ... issue MMA(tile 0, step t+1) | cell update tile 1, step t | issue MMA(tile 1, t+1) | ...
//   tensor   :  ...        MMA(tile 0, t+1) running ............ | MMA(tile 1, t+1) running ...........
//   16 warps :  ... cell update (tile 1, t) .................... | cell update (tile 0, t+1) ..........
//
// Hand-shakes are per tile and by mbarrier, never CTA-wide: full[tile] (tcgen05.commit: accumulator complete, operand
// free) and ready[tile] (512 arrivals: accumulator consumed, next operand written).  See fc_pair_layout.h for where
// the operands live (TMEM / shared memory budget) and DESIGN.md section 2.3.
// Warp 0 is the MMA issuer and does nothing else (tcgen05.mma blocks its issuer for about the duration of the
// chain; all 32 lanes walk the issue loop together, lane 0 issues).  Warps 4..15 do the cell updates: third th = w/4 - 1 owns the hidden units [16 th, 16 th + 16) (18 for
// th = 2).  The per-trajectory scalar work of a row (roll-out rows, layer-0 features, read-out, cost terms,
// controller) belongs to the otherwise idle warps 1..3 for their TMEM quadrants and to warp 4 for quadrant 0.
#pragma once
#include "fc_pair_layout.h"

namespace fc {
namespace pr {

// R = 1: two 128-trajectory tiles per CTA (the schedule described above).
// R = 4 ("replica" mode, small and mid-size batches): ONE tile of 32 trajectories per CTA, replicated into the four
// 32-row groups of the M = 128 operand images, so that every TMEM quadrant holds the gate pre-activations of the same 32
// trajectories and the twelve cell-update warps split the 50 hidden units of a trajectory twelve ways (4 units per thread,
// 6 for the last one) instead of three ways: the dependent chain of a step shrinks from ~5 000 to ~1 200 cycles, which is
// what a latency-bound batch (B = 15 .. 4 736) needs.  Both operands come from shared memory in both directions (SS-mode
// MMA); the per-trajectory scalar work runs on warp 1; read-out partial sums are exchanged through shared memory.
// PT = true (surrogate training): tanh keeps its odd polynomial below |x| = 0.2.  With freshly initialised surrogate weights
// the cell values are small and the weight gradients are differences of nearly equal terms, so the RELATIVE accuracy of
// small tanh values shows (fc.weight gradient 2.7e-5 of the reference without it, 3e-7 with it); the roll-out kernels work
// on a trained surrogate and do not need it (see tanh_from_).
template <class Ctx, int R = 1, bool PT = false>
struct MpcPair {
  static_assert(R == 1 || R == 4, "replica factor");
  static constexpr int kRowsT = kTileP / R;          // distinct trajectories per tile
  static constexpr int kOwn = R == 1 ? kMaxOwn : 6;  // register slots for the units a thread owns
  Ctx& ctx;
  const MpcParams& p;
  float* sm;
  int tid, warp, lane, row, traj, quad, th, nown, u_first, uw;
  bool service;          // warps 0..3: thread 0 issues the MMAs, the rest only joins the CTA-wide barriers
  bool owner;            // third 0 of the cell-update warps: also collects the row-feature gradients of its row
  bool scalar;           // does the per-trajectory scalar work of its row: warps 1..3, and warp 4 for the rows of warp 0
  bool last;             // third 2: 18 units (else 16)
  float* wbase;          // workspace of tile 0 of this CTA; tile 1 follows at +tstride
  size_t tstride;
  int ntl;               // live tiles in this pass (1 or 2)
  int tile0;             // global index of tile 0 of this pass
  // forward: cell state, backward: d(cell state) of the tile of the CURRENT item; the state of the other tile is
  // parked in spare TMEM columns of the own lane and exchanged at the start of every item (items alternate strictly;
  // keeping both in registers was measured 12 % slower: spills)
  float c[kOwn];
  unsigned phF0, phF1, phR0, phR1, phW;
#if defined(FC_TC_TRACE)
#ifndef FC_TRACE_W1
#define FC_TRACE_W1 5
#define FC_TRACE_W2 13
#endif
  // event trace (development aid, -DFC_TC_TRACE): lane 0 of warps 0 (issuer), 5 (third 0) and 13 (third 2) of CTA 0 log
  // (clock << 8 | event) at every lap point; read back with fc_debug_trace
  int trace_slot, trace_n;
  FC_HD_CTX void lap(int k) {
    if (trace_slot >= 0 && trace_n < Ctx::kTraceMax) Ctx::trace_put(trace_slot, trace_n++, (Ctx::clock() << 8) | (long long)k);
  }
#elif defined(FC_TC_TIMING)
  long long tm[24], tlast;  // cycle breakdown (development aid, -DFC_TC_TIMING): threads 0 and 160 of CTA 0 report
  FC_HD_CTX void lap(int k) { long long t = Ctx::clock(); tm[k] += t - tlast; tlast = t; }
#else
  FC_HD_CTX void lap(int) {}
#endif

  FC_HD_CTX MpcPair(Ctx& c_, const MpcParams& p_) : ctx(c_), p(p_) {
    sm = ctx.smem();
    tid = ctx.tid();
    warp = tid >> 5;
    lane = tid & 31;
    row = 32 * (warp & 3) + lane;
    quad = warp & 3;
    traj = R == 1 ? row : lane;
    service = warp < 4;
    th = service ? 0 : (warp >> 2) - 1;
    if (R == 1) {
      nown = units_of(th);
      u_first = first_unit(th);
      last = th == 2;
      owner = !service && th == 0;
      scalar = (service && warp != 0) || warp == 4;
    } else {
      // units [16 th + 4 quad, + 4); the last thread of a trajectory also takes units 48, 49 (44..49)
      u_first = 16 * th + 4 * quad;
      last = !service && th == 2 && quad == 3;
      nown = last ? 6 : 4;
      owner = !service && th == 0 && quad == 0;
      scalar = warp == 1;
    }
    uw = service ? 0 : warp - 4;                             // index among the cell-update warps
    tstride = p.train == 2 ? work_total_train() : work_layout_p(p.N, p.with_grad).total;
    gsc = p.g_scale;
    wbase = p.work + (size_t)ctx.bid() * p.work_stride;
    phF0 = phF1 = phR0 = phR1 = phW = 0;
    ntl = 0; tile0 = 0;
#ifdef FC_TC_TRACE
    trace_n = 0;
    trace_slot = (ctx.bid() == 0 && lane == 0) ? (warp == 0 ? 0 : warp == FC_TRACE_W1 ? 1 : warp == FC_TRACE_W2 ? 2 : -1) : -1;
#endif
#ifdef FC_TC_TIMING
    for (int i = 0; i < 24; ++i) tm[i] = 0;
    tlast = 0;
#endif
  }

  // window bookkeeping: the roll-out keeps steps t_min_of(m)..9 of window m; surrogate training has one window and keeps all
  FC_HD_CTX int tmin_of(int m) const { return p.train ? 0 : t_min_of(m); }
  FC_HD_CTX int kept_of(int m) const { return p.train ? kLook : steps_kept(m); }
  FC_HD_CTX long recb_of(int m) const { return p.train ? 0 : rec_base(m); }
  float gsc;             // scale of the gate gradients (roll-out: kernel parameter; training: device scalar)
  // training: scratch of tile X for the weight-gradient kernel
  // (a tile of kRowsT samples holds kRowsT / 32 stages of every image: all sizes scale with 1 / R)
  FC_HD_CTX float* tr_tile(int X) const { return p.tr_ws + (size_t)(p.tr_tile_base + tile0 + X) * (kTrTileFloats / R); }
  FC_HD_CTX float* tr_hseq(int X, int l, int slot) const { return tr_tile(X) + kTrHseqOff / R + (size_t)(l * (kLook + 1) + slot) * (kTrHseqSlot / R); }
  // R == 4: the thread's 8 bytes (4 units) of the hi / lo image of the hidden state of layer l at step t (slot t + 1)
  FC_HD_CTX float* seq_half_r(int X, int l, int t, int hl) const {
    return tr_hseq(X, l, t + 1) + (size_t)hl * (7 * kRowsT * 4) + (size_t)((u_first >> 3) * 32 + lane) * 4 + ((u_first & 7) >> 1);
  }
  FC_HD_CTX float* seq_tail_r(int X, int l, int t, int hl) const {      // piece 6: units 48..55
    return tr_hseq(X, l, t + 1) + (size_t)hl * (7 * kRowsT * 4) + (size_t)(6 * 32 + lane) * 4;
  }

  // workspace pointers of tile X
  FC_HD_CTX float* w_rows(int X) const { return wbase + X * tstride; }
  FC_HD_CTX float* w_cost(int X) const { return w_rows(X) + (size_t)(p.N + kLook) * kFeat * kTileP; }
  FC_HD_CTX float* w_seq(int X) const { return w_cost(X) + 3 * kTileP; }
  FC_HD_CTX float* w_dseq(int X) const { return w_seq(X) + (size_t)kLook * kSlot; }
  FC_HD_CTX float* w_grow(int X) const { return w_dseq(X) + (size_t)kLook * kSlot; }
  FC_HD_CTX float* w_rec(int X) const { return w_rows(X) + work_layout_p(p.N, p.with_grad).rec; }

  // ---------------------------------------------------------------------------------------------
  // activations (see fc_mpc_tc_kernel.inl): one reciprocal per four denominators, polynomial tanh near 0
  // ---------------------------------------------------------------------------------------------
  static constexpr float kExpMax = 30.0f;
  static constexpr float kLog2e = 1.4426950216293335f;            // fp32(log2 e)
#ifdef FC_ABL_NO_MUFU         // timing ablation only: no MUFU.EX2 / MUFU.RCP
  FC_HD_CTX static float denom_(float e2arg) { return 1.f + fminf(e2arg, kExpMax) * fminf(e2arg, kExpMax); }
#else
  FC_HD_CTX static float denom_(float e2arg) { return 1.f + Ctx::ex2(fminf(e2arg, kExpMax)); }
#endif
  // four reciprocals from TWO MUFU operations (one per pair of denominators): 8 issue slots.  One reciprocal of the product
  // of all four (9 multiplies, the one-tile kernel's form) costs 10 issue slots; with the cell update bound by its
  // instruction count the pair form measured +0.6 % sustained / +1.1 % one pass (same-box A/B) at a quarter more XU load.
  FC_HD_CTX static void quad_rcp(float a, float b, float c, float d, float& ra, float& rb, float& rc, float& rd) {
    const float r1 = Ctx::rcp(a * b), r2 = Ctx::rcp(c * d);
    ra = r1 * b; rb = r1 * a; rc = r2 * d; rd = r2 * c;
  }
  // tanh(x) = 1 - 2 rd with rd = 1/(1 + e^{2x}): absolute error ~1e-7.  The gate pre-activations and cell states it is
  // applied to carry an absolute error of that size already (fp16 hi/lo product, fp32 cell recurrence), so the odd
  // polynomial rounds 1 and 2 used below |x| = 0.2 for RELATIVE accuracy bought nothing measurable: summed gradients
  // 2e-7..5e-7 of the fp64 oracle without it against 1.4e-7..2.6e-7 with it (profiles/r02b_activation_variants.md),
  // for 12 of 62 instructions per cell unit
  FC_HD_CTX static float tanh_from_(float x, float rd) {
    const float big = fmaf(-2.f, rd, 1.f);
    if constexpr (!PT) return big;
    // below |x| = 0.2: odd Taylor polynomial up to x^7 (next term 62/2835 x^9: relative 6e-8 at 0.2)
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
    // the cell state of a 10-step window that starts from zero is bounded by 10 (|i g| < 1, f < 1): 2 log2(e) |c| < 28.9,
    // no clamp needed for the product of four denominators to stay finite
    for (int i = 0; i < NU; ++i) d[i] = 1.f + Ctx::ex2(2.f * kLog2e * x[i]);
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
  // unscale, the accumulator compensation (1 + corr) and -log2(e) folded into ONE multiplier of the raw accumulator
  // value (rounded once: 6e-8 relative on the exponent argument, i.e. <= 2e-8 on a sigmoid)
  struct ActK { float k1, k2, kx; };
  FC_HD_CTX static ActK make_actk(float unscale, float corr) {
    ActK k;
    const float khi = -kLog2e * unscale;                    // exact: unscale is a power of two
    k.k1 = fmaf(khi, corr, khi);                            // sigmoid gates: exponent argument = raw * k1
    k.k2 = -2.f * k.k1;                                     // tanh gate: 2 log2(e) x
    k.kx = fmaf(unscale, corr, unscale);                    // PT: the pre-activation itself
    return k;
  }

  // ---------------------------------------------------------------------------------------------
  // hand-shakes
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void wait_full(int X) {
    if (R == 1 && service && warp != 0) ctx.bar_wait_relaxed(kBarFull + X, X ? phF1 : phF0);  // warps 1..3: off the critical path
    else ctx.bar_wait(kBarFull + X, X ? phF1 : phF0);
    if (X) phF1 += 1; else phF0 += 1;
  }
  FC_HD_CTX void swap_cells() {                            // cell-update warps only
#ifdef FC_ABL_NO_SWAP        // timing ablation only
    return;
#endif
    if constexpr (R == 1) if (ntl > 1) {
      float o[kMaxOwn];
      const int col = kColPark + kMaxOwn * th;
      ctx.template tmem_ld_nowait<16>(col, o);
      ctx.template tmem_ld_nowait<2>(col + 16, o + 16);
      ctx.tmem_ld_wait();
      ctx.template tmem_st<16>(col, c);
      ctx.template tmem_st<2>(col + 16, c + 16);
#pragma unroll
      for (int j = 0; j < kMaxOwn; ++j) c[j] = o[j];
      ctx.tmem_st_wait();
    }
  }
  // all operand writes / accumulator reads of this thread for tile X are done
  FC_HD_CTX void arrive_ready(int X, bool smem_operand = true) {
    if (smem_operand) ctx.operand_fence();
    else ctx.tmem_fence();
    ctx.warp_sync();
    if (lane == 0) ctx.bar_arrive(kBarReady + X);            // one arrival per warp
  }
  FC_HD_CTX void wait_ready(int X) {                                                             // issuer warp
    ctx.bar_wait(kBarReady + X, X ? phR1 : phR0);
    if (X) phR1 += 1; else phR0 += 1;
  }
  FC_HD_CTX void wait_weights() { ctx.bar_wait(kBarWeightsP, phW); phW += 1; }                    // issuer warp
  FC_HD_CTX void request_weights(bool bwd, int l) {                                              // tid 0 only
    const int n = bwd ? bwd_img_halves(l) : fwd_img_halves(l);   // hi + lo images of halves = that many floats
    ctx.bulk_load(sm + kSmWP, p.wpack + (bwd ? wb_off(l) : wf_off(l)), n, kBarWeightsP);
  }

  // ---------------------------------------------------------------------------------------------
  // shared-memory operand images: K-major, no swizzle: half (row r, k) at (k/8)*1024 + r*8 + k%8
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX float* op_ptr(int halves_off, int k) const {       // k multiple of 2
    return sm + kSmOpP + ((halves_off + (k >> 3) * (kTileP * 8) + row * 8 + (k & 7)) >> 1);
  }
  // R == 4: copy g (0..3) of this lane's trajectory, i.e. image row 32 g + lane
  FC_HD_CTX float* op_ptr_rep(int halves_off, int k, int g) const {
    return sm + kSmOpP + ((halves_off + (k >> 3) * (kTileP * 8) + (32 * g + lane) * 8 + (k & 7)) >> 1);
  }
  // R == 4: fp16 hi/lo words of the 4 (6) owned units: w = {hi01, hi23, lo01, lo23, hi45, lo45}
  FC_HD_CTX void split_units_r(const float* v, float scale, float* w) const {
    Ctx::split_h2(v[0] * scale, v[1] * scale, w[0], w[2]);
    Ctx::split_h2(v[2] * scale, v[3] * scale, w[1], w[3]);
    if (last) Ctx::split_h2(v[kOwn - 2] * scale, v[kOwn - 1] * scale, w[4], w[5]);
    else { w[4] = 0.f; w[5] = 0.f; }
  }
  // ... into halves [kbase + u_first, ...) of the hi and lo images, all four copies: 8 bytes (half a piece) per image,
  // the last thread also the whole piece of units 48..55 (48, 49 and the zero padding)
  FC_HD_CTX void st_pieces_r(int img_hi, int img_lo, int kbase, const float* w) {
    const int k0 = kbase + u_first;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      Ctx::sts2(op_ptr_rep(img_hi, k0, g), w[0], w[1]);
      Ctx::sts2(op_ptr_rep(img_lo, k0, g), w[2], w[3]);
      if (last) {
        Ctx::sts4(op_ptr_rep(img_hi, k0 + 4, g), F4{w[4], 0.f, 0.f, 0.f});
        Ctx::sts4(op_ptr_rep(img_lo, k0 + 4, g), F4{w[5], 0.f, 0.f, 0.f});
      }
    }
  }
  // ... and into the thread-private slots of the hidden-sequence scratch: slot 0 = {hi01, hi23, lo01, lo23}, the last
  // thread: slot 1 = {hi45, 0, 0, 0}, slot 2 = {lo45, 0, 0, 0}
  FC_HD_CTX void stg_pieces_r(int X, int l, int t, const float* w) {
    if (p.train == 2) {                                      // operand-format hidden sequence kept for the weight-gradient kernel
      Ctx::stg2(seq_half_r(X, l, t, 0), w[0], w[1]);
      Ctx::stg2(seq_half_r(X, l, t, 1), w[2], w[3]);
      if (last) { Ctx::stg4(seq_tail_r(X, l, t, 0), F4{w[4], 0.f, 0.f, 0.f}); Ctx::stg4(seq_tail_r(X, l, t, 1), F4{w[5], 0.f, 0.f, 0.f}); }
      return;
    }
    float* sq = seq_ptr(X, t);
    Ctx::stg4(sq, F4{w[0], w[1], w[2], w[3]});
    if (last) { Ctx::stg4(sq + 128, F4{w[4], 0.f, 0.f, 0.f}); Ctx::stg4(sq + 256, F4{w[5], 0.f, 0.f, 0.f}); }
  }
  // the owned units' values v[0..nown) (scaled here) as fp16 hi/lo pieces of 8 halves (third 2: the third piece
  // holds units 48,49 and the zero padding up to slot 56)
  FC_HD_CTX void split_units(const float* v, float scale, F4* hi4, F4* lo4) const {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = ch * 8 + 2 * i;
        const float x0 = j < kMaxOwn ? v[j < kMaxOwn ? j : 0] * scale : 0.f;
        const float x1 = j + 1 < kMaxOwn ? v[j + 1 < kMaxOwn ? j + 1 : 0] * scale : 0.f;
        Ctx::split_h2(x0, x1, hi[i], lo[i]);
      }
      hi4[ch] = F4{hi[0], hi[1], hi[2], hi[3]};
      lo4[ch] = F4{lo[0], lo[1], lo[2], lo[3]};
    }
  }
  // pieces -> halves [kbase + u_first, ...) of the hi and lo operand images (kbase multiple of 8)
  FC_HD_CTX void st_pieces(int img_hi, int img_lo, int kbase, const F4* hi4, const F4* lo4) {
    const int k0 = kbase + u_first;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      if (ch < 2 || last) { Ctx::sts4(op_ptr(img_hi, k0 + ch * 8), hi4[ch]); Ctx::sts4(op_ptr(img_lo, k0 + ch * 8), lo4[ch]); }
  }
  // pieces <-> the hidden-sequence scratch of the layer below: [t][warp][piece*2 + hi/lo][lane] float4, thread-private
  FC_HD_CTX float* seq_ptr(int X, int t) const { return w_seq(X) + (size_t)t * kSlot + ((size_t)uw * 6 * 32 + lane) * 4; }
  // piece ch (hi / lo) of the hidden state of layer l at step t: thread-private slot of the per-CTA scratch, or (training
  // with the reverse sweep) its place in the operand-format hidden sequence kept for the weight-gradient kernel
  FC_HD_CTX float* seq_piece(int X, int l, int t, int ch, int hl) const {
    if (p.train == 2)   // [hi | lo][32-sample stage = TMEM quadrant][7 pieces][32 rows][16 B]: a stage of the weight-gradient kernel is contiguous
      return tr_hseq(X, l, t + 1) + (size_t)hl * (7 * kRowsT * 4) + (size_t)(quad * (7 * 32) + (2 * th + ch) * 32 + lane) * 4;
    return seq_ptr(X, t) + (ch * 2 + hl) * 128;
  }
  FC_HD_CTX void stg_pieces(int X, int l, int t, const F4* hi4, const F4* lo4) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      if (ch < 2 || last) { Ctx::stg4(seq_piece(X, l, t, ch, 0), hi4[ch]); Ctx::stg4(seq_piece(X, l, t, ch, 1), lo4[ch]); }
  }
  // asynchronous copy scratch -> input block of the operand images (cp.async, no registers); cp_wait before the arrive
  FC_HD_CTX void copy_input(int X, int l, int t) {            // input of layer l at step t = h of layer l - 1
    const int img_hi = op_fwd_halves(X), img_lo = img_hi + kOpLoHalves;
    const float* sq = seq_ptr(X, t);
    if constexpr (R == 4) {
      const bool kept = p.train == 2;
      const float* s_hi = kept ? seq_half_r(X, l - 1, t, 0) : sq;
      const float* s_lo = kept ? seq_half_r(X, l - 1, t, 1) : sq + 2;
      const float* t_hi = kept ? seq_tail_r(X, l - 1, t, 0) : sq + 128;
      const float* t_lo = kept ? seq_tail_r(X, l - 1, t, 1) : sq + 256;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        Ctx::cp_async8(op_ptr_rep(img_hi, u_first, g), s_hi);
        Ctx::cp_async8(op_ptr_rep(img_lo, u_first, g), s_lo);
        if (last) {
          Ctx::cp_async16(op_ptr_rep(img_hi, u_first + 4, g), t_hi);
          Ctx::cp_async16(op_ptr_rep(img_lo, u_first + 4, g), t_lo);
        }
      }
      Ctx::cp_commit();
      return;
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      if (ch < 2 || last) {
        Ctx::cp_async16(op_ptr(img_hi, u_first + ch * 8), seq_piece(X, l - 1, t, ch, 0));
        Ctx::cp_async16(op_ptr(img_lo, u_first + ch * 8), seq_piece(X, l - 1, t, ch, 1));
      }
    Ctx::cp_commit();
  }
  FC_HD_CTX void st_units_zero(int img_hi, int img_lo, int kbase) {
    const int k0 = kbase + u_first;
    const F4 z = {0.f, 0.f, 0.f, 0.f};
    if constexpr (R == 4) {
      const float w[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      st_pieces_r(img_hi, img_lo, kbase, w);
      return;
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      if (ch < 2 || last) { Ctx::sts4(op_ptr(img_hi, k0 + ch * 8), z); Ctx::sts4(op_ptr(img_lo, k0 + ch * 8), z); }
  }

  // hi/lo fp16 split of 2*NP values (already in the scaled domain) into NP consecutive TMEM operand columns of the own lane
  // gp != nullptr (surrogate training): the pieces also go to the global dG image [26 pieces][128][16 B] (hi, then lo) of
  // this (tile, layer, step); k0 = first k-slot (multiple of 8)
  FC_HD_CTX void stg_dg_piece(float* gp, int k, F4 hi, F4 lo) const {
    float* q = gp + (size_t)((R == 1 ? quad : 0) * (26 * 32) + (k >> 3) * 32 + lane) * 4;     // [stage][26 pieces][32 rows][16 B]
    Ctx::stg4(q, hi);
    Ctx::stg4(q + 26 * kRowsT * 4, lo);
  }
  template <int NP>
  FC_HD_CTX void st_pairs(int col_hi, int col_lo, const float* v, float* gp = nullptr, int k0 = 0) {
    float hi[NP], lo[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      Ctx::split_h2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);   // saturating conversion
    }
    ctx.template tmem_st<NP>(col_hi, hi);
    ctx.template tmem_st<NP>(col_lo, lo);
    if (gp) {
#pragma unroll
      for (int ch = 0; ch < NP / 4; ++ch)
        stg_dg_piece(gp, k0 + ch * 8, F4{hi[ch * 4], hi[ch * 4 + 1], hi[ch * 4 + 2], hi[ch * 4 + 3]},
                     F4{lo[ch * 4], lo[ch * 4 + 1], lo[ch * 4 + 2], lo[ch * 4 + 3]});
    }
  }
  // the same into 2*NP halves of the shared-memory dG image (tile 1), k0 multiple of 8, NP multiple of 4
  template <int NP>
  FC_HD_CTX void st_pairs_smem(int k0, const float* v, float* gp = nullptr) {
#pragma unroll
    for (int ch = 0; ch < NP / 4; ++ch) {
      float hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        Ctx::split_h2(v[ch * 8 + 2 * i], v[ch * 8 + 2 * i + 1], hi[i], lo[i]);
      }
      if (gp) stg_dg_piece(gp, k0 + ch * 8, F4{hi[0], hi[1], hi[2], hi[3]}, F4{lo[0], lo[1], lo[2], lo[3]});
      if constexpr (R == 1) {
        Ctx::sts4(op_ptr(0, k0 + ch * 8), F4{hi[0], hi[1], hi[2], hi[3]});
        Ctx::sts4(op_ptr(kOpGLoHalves, k0 + ch * 8), F4{lo[0], lo[1], lo[2], lo[3]});
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          Ctx::sts4(op_ptr_rep(0, k0 + ch * 8, g), F4{hi[0], hi[1], hi[2], hi[3]});
          Ctx::sts4(op_ptr_rep(kOpGLoHalves, k0 + ch * 8, g), F4{lo[0], lo[1], lo[2], lo[3]});
        }
      }
    }
  }

  // ---------------------------------------------------------------------------------------------
  // MMA issue (tid 0 only): 3 error-compensated terms, small ones first
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void issue_fwd(int X, int l, int ksteps) {
    const float* b_hi = sm + kSmWP;
    const float* b_lo = b_hi + fwd_img_halves(l) / 2;
    const float* a_hi = sm + kSmOpP + op_fwd_halves(X) / 2;
    const float* a_lo = a_hi + kOpLoHalves / 2;
    const int d = col_d_fwd(X);
#ifdef FC_ABL_ONE_TERM       // timing ablation only (results lose the error compensation)
    ctx.mma_ss(d, kNF, a_hi, b_hi, kNF, ksteps, false);
#else
    ctx.mma_ss(d, kNF, a_lo, b_hi, kNF, ksteps, false);
    ctx.mma_ss(d, kNF, a_hi, b_lo, kNF, ksteps, true);
    ctx.mma_ss(d, kNF, a_hi, b_hi, kNF, ksteps, true);
#endif
    ctx.commit(kBarFull + X);
  }
  FC_HD_CTX void issue_bwd(int X, int l) {
    const int nb = nb_of(l);
    const float* b_hi = sm + kSmWP;
    const float* b_lo = b_hi + bwd_img_halves(l) / 2;
    const int d = col_d_bwd(X);
    if (R == 1 && X == 0) {
#ifdef FC_ABL_ONE_TERM
      ctx.mma(d, nb, kColGhi, b_hi, nb, 0, kKB / 16, false);
#else
      ctx.mma(d, nb, kColGlo, b_hi, nb, 0, kKB / 16, false);
      ctx.mma(d, nb, kColGhi, b_lo, nb, 0, kKB / 16, true);
      ctx.mma(d, nb, kColGhi, b_hi, nb, 0, kKB / 16, true);
#endif
    } else {
      const float* a_hi = sm + kSmOpP;
      const float* a_lo = a_hi + kOpGLoHalves / 2;
#ifdef FC_ABL_ONE_TERM
      ctx.mma_ss(d, nb, a_hi, b_hi, nb, kKB / 16, false);
#else
      ctx.mma_ss(d, nb, a_lo, b_hi, nb, kKB / 16, false);
      ctx.mma_ss(d, nb, a_hi, b_lo, nb, kKB / 16, true);
      ctx.mma_ss(d, nb, a_hi, b_hi, nb, kKB / 16, true);
#endif
    }
    ctx.commit(kBarFull + X);
  }

  // ---------------------------------------------------------------------------------------------
  // tile set-up (scalar-work thread of each trajectory)
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void load_tile(int X) {
    const int b = (tile0 + X) * kRowsT + traj;
    const bool ok = b < p.B;
    float* rows = w_rows(X);
    const int row = traj;                                    // per-trajectory arrays are indexed by the trajectory
    if (p.train) {                                           // surrogate training: the window is the sample itself
      for (int r = 0; r < kLook; ++r)
#pragma unroll
        for (int f = 0; f < kFeat; ++f)
          rows[(size_t)(r * kFeat + f) * kTileP + row] = ok ? p.tr_x[(size_t)b * (kLook * kFeat) + r * kFeat + f] : 0.f;
      return;
    }
    for (int r = 0; r < kLook; ++r)
#pragma unroll
      for (int f = 0; f < kFeat; ++f) {
        float v = ok ? p.Z[(size_t)b * (kLook * kFeat) + r * kFeat + f] : 0.f;
        if (r == kLook - 1 && f == kFeat - 1) v = ok ? p.u0[b] : 0.f;              // Functions.py:1396
        rows[(size_t)(r * kFeat + f) * kTileP + row] = v;
      }
    sm[kSmRefP + X * kTileP + row] = ok ? p.X[(size_t)b * 3 + 2] : 0.f;            // :1392
    float* cg = w_cost(X);
    cg[row] = 0.f; cg[kTileP + row] = 0.f; cg[2 * kTileP + row] = 0.f;
    if (ok) p.pred[(size_t)b * p.N] = p.u0[b];                                     // :1417-1418
  }

  // ---------------------------------------------------------------------------------------------
  // forward: cell update of NU unit slots starting at slot j0 from 4*NU raw accumulator columns
  // ---------------------------------------------------------------------------------------------
  template <int NU>
  FC_HD_CTX void fwd_units(int j0, const float* g, const ActK& ak, bool first, float* h, float* rp, int r0) {
    constexpr int NR = (NU * 5 + 3) / 4 * 4;
    float rv[NR];
    float cn[NU], tch[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const float di = denom_(g[i * 4 + 0] * ak.k1);
      const float df = denom_(g[i * 4 + 1] * ak.k1);
      const float dq = denom_(g[i * 4 + 3] * ak.k1);
      const float dg = denom_(g[i * 4 + 2] * ak.k2);
      float gi, gf, go, rg;
      quad_rcp(di, df, dq, dg, gi, gf, go, rg);
      const float gg = tanh_from_(PT ? g[i * 4 + 2] * ak.kx : 0.f, rg);
      const float cp = first ? 0.f : c[j0 + i];
      cn[i] = fmaf(gf, cp, gi * gg);
      c[j0 + i] = cn[i];
      rv[i * 5 + 0] = gi; rv[i * 5 + 1] = gf; rv[i * 5 + 2] = gg; rv[i * 5 + 3] = go; rv[i * 5 + 4] = cp;
    }
    tanh_batch<NU>(cn, tch);
#pragma unroll
    for (int i = 0; i < NU; ++i) h[j0 + i] = rv[i * 5 + 3] * tch[i];
    if (rp) {
#pragma unroll
      for (int i = NU * 5; i < NR; ++i) rv[i] = 0.f;
#pragma unroll
      for (int r = 0; r < NR / 4; ++r) {
        F4 v = {rv[r * 4], rv[r * 4 + 1], rv[r * 4 + 2], rv[r * 4 + 3]};
        Ctx::stg4_stream(rp + (size_t)(r0 + r) * 32 * 4, v);
      }
    }
  }

  // all unit slots of one step (the accumulator barrier has been waited for by the caller)
  FC_HD_CTX void fwd_pointwise(int X, bool first, float corr, float* h, float* rec_out) {
    const ActK ak = make_actk(1.0f / (kScaleA * kScaleW), corr);
    float* rp = rec_out ? rec_out + ((size_t)uw * kRecF4 * 32 + lane) * 4 : nullptr;
    const int col0 = col_d_fwd(X) + 4 * u_first;
    if constexpr (R == 4) {
      float g4[24];
      ctx.template tmem_ld_nowait<16>(col0, g4);
      if (last) ctx.template tmem_ld_nowait<8>(col0 + 16, g4 + 16);
      ctx.tmem_ld_wait();
      fwd_units<4>(0, g4, ak, first, h, rp, 0);
      if (last) fwd_units<2>(4, g4 + 16, ak, first, h, rp, 5);   // units 48, 49: record float4 5..7
      return;
    }
    float g[2][16];
    ctx.template tmem_ld_nowait<16>(col0, g[0]);
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
      ctx.tmem_ld_wait();
      // software pipeline: request the next accumulator columns before working on these
      if (gi + 1 < 4) ctx.template tmem_ld_nowait<16>(col0 + (gi + 1) * 16, g[(gi + 1) & 1]);
      else if (last) ctx.template tmem_ld_nowait<8>(col0 + 64, g[(gi + 1) & 1]);
      fwd_units<4>(gi * 4, g[gi & 1], ak, first, h, rp, gi * 5);
    }
    if (last) {
      ctx.tmem_ld_wait();
      fwd_units<2>(16, g[0], ak, first, h, rp, 20);        // units 48, 49: record float4 20..22
    }
  }

  // layer 0: the 5 row features of step t (scalar-work thread of the row): k = 0..7, 3 zero
  FC_HD_CTX void load_features(int X, int m, int t, float* xin) {
    const float* rp = w_rows(X) + (size_t)(m + t) * kFeat * kTileP + traj;
#pragma unroll
    for (int f = 0; f < kFeat; ++f) xin[f] = Ctx::ldcg(rp + f * kTileP);
  }
  FC_HD_CTX void store_features(int X, const float* xin, int t) {
    const int img_hi = op_fwd_halves(X), img_lo = img_hi + kOpLoHalves;
    float hi[4], lo[4];
    Ctx::split_h2(xin[0] * kScaleA, xin[1] * kScaleA, hi[0], lo[0]);
    Ctx::split_h2(xin[2] * kScaleA, xin[3] * kScaleA, hi[1], lo[1]);
    Ctx::split_h2(xin[4] * kScaleA, 0.f, hi[2], lo[2]);
    if (p.train == 2) {                                      // layer-0 input of step t, kept for the weight-gradient kernel
      float* fp = tr_tile(X) + kTrFeatOff / R + (size_t)t * (kTrFeatSlot / R) + (size_t)traj * 4;
      Ctx::stg4(fp, F4{hi[0], hi[1], hi[2], 0.f});
      Ctx::stg4(fp + kRowsT * 4, F4{lo[0], lo[1], lo[2], 0.f});
    }
    if constexpr (R == 1) {
      Ctx::sts4(op_ptr(img_hi, 0), F4{hi[0], hi[1], hi[2], 0.f});
      Ctx::sts4(op_ptr(img_lo, 0), F4{lo[0], lo[1], lo[2], 0.f});
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        Ctx::sts4(op_ptr_rep(img_hi, 0, g), F4{hi[0], hi[1], hi[2], 0.f});
        Ctx::sts4(op_ptr_rep(img_lo, 0, g), F4{lo[0], lo[1], lo[2], 0.f});
      }
    }
  }

  // ---------------------------------------------------------------------------------------------
  // forward window: both tiles, interleaved cell step by cell step
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void fwd_prologue(int X, int l, int m) {
    // operand of step 0: zero recurrent block, input of step 0 (the previous MMA on this tile has been waited for)
    const int img_hi = op_fwd_halves(X), img_lo = img_hi + kOpLoHalves;
    lap(18);
    if (!service) {
      st_units_zero(img_hi, img_lo, l == 0 ? kRec0 : kRec);
      if (p.train == 2) {                                     // slot 0 of the kept hidden sequence = zeros (h before step 0)
        const F4 z = {0.f, 0.f, 0.f, 0.f};
        if constexpr (R == 1) {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch)
            if (ch < 2 || last) { Ctx::stg4(seq_piece(X, l, -1, ch, 0), z); Ctx::stg4(seq_piece(X, l, -1, ch, 1), z); }
        } else {
          const float w0[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          stg_pieces_r(X, l, -1, w0);
        }
      }
      if (l > 0) {
        copy_input(X, l, 0);
        Ctx::template cp_wait<0>();
      } else if (scalar) {
        float xin[kFeat];
        load_features(X, m, 0, xin);
        store_features(X, xin, 0);
      }
      arrive_ready(X);
    } else if (scalar) {
      if (l == 0) {
        float xin[kFeat];
        load_features(X, m, 0, xin);
        store_features(X, xin, 0);
      }
      arrive_ready(X);
    }
    if (warp == 0) {                                         // issuer warp: all lanes follow, lane 0 issues
      wait_ready(X);
      if (X == 0) wait_weights();                            // operand image of this layer landed
      if (lane == 0) issue_fwd(X, l, l == 0 ? 1 : 4);        // recurrent part is zero: only the k-steps over the input
      ctx.warp_sync();
    }
    lap(14);
  }

  // issuer warp (warp 0, all 32 lanes stay together; lane 0 issues): one step of the issue loop
  FC_HD_CTX void fwd_item_issuer(int X, int l, int m, int t, bool more_after) {
    lap(23);
    wait_full(X);                                            // MMA(X, l, t) complete (keeps the phase count; no stall)
    lap(1);
    if (t == kLook - 1 && X == ntl - 1 && lane == 0) {       // stream the next weight image under the cell updates
      if (l + 1 < kLayers) request_weights(false, l + 1);
      else if (m + 1 < p.N) request_weights(false, 0);
      else if (p.with_grad) request_weights(true, kLayers - 1);
      else if (more_after) request_weights(false, 0);
    }
    if (t + 1 < kLook) {
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
      if (l == 0) store_features(X, xin, t + 1);
      arrive_ready(X);
    }
  }

  FC_HD_CTX void fwd_item(int X, int l, int m, int t) {
    const int tmin = tmin_of(m);
    float h[kOwn], xin[kFeat];
    lap(16);
    swap_cells();
    lap(11);
    if (l == 0 && scalar && t + 1 < kLook) load_features(X, m, t + 1, xin);
    float* rec_out = nullptr;
    if (p.with_grad && t >= tmin)
#ifdef FC_ABL_NO_REC_TRAFFIC  // timing ablation only: every record lands in the same L2-resident slot
      rec_out = w_rec(X);
#else
      rec_out = w_rec(X) + (size_t)(recb_of(m) + (long)l * kept_of(m) + (t - tmin)) * kRecFloatsP;
#endif
    const int ksteps = t == 0 ? (l == 0 ? 1 : 4) : kf_of(l) / 16;
    const float corr = Ctx::kAccTruncates ? acc_correction(ksteps, p.acc_comp) : 0.0f;
    const int img_hi = op_fwd_halves(X), img_lo = img_hi + kOpLoHalves;
    lap(17);
    wait_full(X);                                            // accumulator complete; operand free
    lap(1);
    if (l > 0 && t + 1 < kLook) copy_input(X, l, t + 1);     // input block of step t+1: lands during the cell update
    if (l == 0 && scalar && t + 1 < kLook) store_features(X, xin, t + 1);
    fwd_pointwise(X, t == 0, corr, h, rec_out);
    lap(24);
    if (l + 1 < kLayers || t + 1 < kLook || p.train == 2) {
      if constexpr (R == 1) {
        F4 hi4[3], lo4[3];
        split_units(h, kScaleA, hi4, lo4);
        if (l + 1 < kLayers || p.train == 2) stg_pieces(X, l, t, hi4, lo4);   // input of the layer above, already in operand format
        if (t + 1 < kLook) st_pieces(img_hi, img_lo, l == 0 ? kRec0 : kRec, hi4, lo4);
      } else {
        float w[6];
        split_units_r(h, kScaleA, w);
        if (l + 1 < kLayers || p.train == 2) stg_pieces_r(X, l, t, w);
        if (t + 1 < kLook) st_pieces_r(img_hi, img_lo, l == 0 ? kRec0 : kRec, w);
      }
    }
    lap(25);
    if (t + 1 < kLook) {
      if (l > 0) Ctx::template cp_wait<0>();
      lap(2);
      arrive_ready(X);
      lap(3);
    } else if (l == kLayers - 1) {
      // read-out partial sums over the owned units (Functions.py:377), handed to the scalar-work thread of the row
      // through spare TMEM columns of the own lane
      const float* fw = sm + kSmSmallP;
      float xq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < kOwn; ++j)
        if (j < nown) {
#pragma unroll
          for (int q = 0; q < 4; ++q) xq[q] = fmaf(fw[q * kHid + u_first + j], h[j], xq[q]);
        }
      if (p.train == 1 && p.tr_hlast) {                       // top-layer hidden state of the last step: input of the fc gradients
        const int b = (tile0 + X) * kRowsT + traj;
        if (b < p.B)
#pragma unroll
          for (int j = 0; j < kOwn; ++j)
            if (j < nown) p.tr_hlast[(size_t)b * kHid + u_first + j] = h[j];
      }
      if constexpr (R == 1) {
        ctx.template tmem_st<4>(kColFcp + 12 * X + 4 * th, xq);
        ctx.tmem_st_wait();
      } else {
        // twelve partial sums per trajectory, exchanged through the free tail of the operand region
#pragma unroll
        for (int q = 0; q < 4; ++q) sm[kSmExP + (uw * 4 + q) * 32 + lane] = xq[q];
      }
    }
    if (t + 1 == kLook) lap(13);
  }

  FC_HD_CTX void fwd_window(int m, bool more_after) {
    for (int l = 0; l < kLayers; ++l) {
      for (int X = 0; X < ntl; ++X) fwd_prologue(X, l, m);
      for (int t = 0; t < kLook; ++t)
        for (int X = 0; X < ntl; ++X) {
          if (warp == 0) fwd_item_issuer(X, l, m, t, more_after);
          else if (!service) fwd_item(X, l, m, t);
          else if (scalar) fwd_item_scalar(X, l, m, t);
        }
    }
    lap(19);
    ctx.tc_sync();                                           // read-out partial sums visible
    if (scalar)
      for (int X = 0; X < ntl; ++X) {
        if (p.train) train_glue(X);
        else if (p.shadow) shadow_glue(X, m);
        else fwd_glue(X, m);
      }
    lap(15);
  }

  // ---------------------------------------------------------------------------------------------
  // LSTM shadow roll-out (Functions.py:969-1011, 1196-1231): the window starts as ten copies of the first row; after
  // window m the surrogate output is logged and [output * scale_out / scale_in, u_{m+1}] becomes the newest row
  // ---------------------------------------------------------------------------------------------
  // surrogate training: y = fc(h) of the window (Model_NN/Functions.py:338-340), and the seed of its reverse sweep
  FC_HD_CTX void train_glue(int X) {
    const float* sw = sm + kSmSmallP;
    float fp[12];
    readout_partials(X, fp);
    const int b = (tile0 + X) * kRowsT + traj;
    if (p.tr_y && b < p.B)
#pragma unroll
      for (int q = 0; q < 4; ++q) p.tr_y[(size_t)b * 4 + q] = ((fp[q] + fp[4 + q]) + fp[8 + q]) + sw[(kFCB - kFCW) + q];
  }
  FC_HD_CTX void train_seed(int X) {
    const int b = (tile0 + X) * kRowsT + traj;
#pragma unroll
    for (int q = 0; q < 4; ++q) sm[kSmGxP + (X * 4 + q) * kTileP + traj] = b < p.B ? p.tr_dy[(size_t)b * 4 + q] : 0.f;
  }
  FC_HD_CTX void load_tile_shadow(int X) {
    const int row = traj;                                    // per-trajectory arrays are indexed by the trajectory
    const int b = (tile0 + X) * kRowsT + traj;
    const bool ok = b < p.B;
    float* rows = w_rows(X);
#pragma unroll
    for (int f = 0; f < kFeat; ++f) {
      const float v = ok ? p.sh_row0[(size_t)b * kFeat + f] : 0.f;
      for (int r = 0; r < kLook; ++r) rows[(size_t)(r * kFeat + f) * kTileP + row] = v;
    }
  }
  // read-out partial sums of the trajectory as three values per output (R == 4: the twelve partial sums of the
  // exchange buffer folded four at a time, fixed order)
  FC_HD_CTX void readout_partials(int X, float* fp) {
    if constexpr (R == 1) {
      ctx.template tmem_ld_nowait<8>(kColFcp + 12 * X, fp);
      ctx.template tmem_ld_nowait<4>(kColFcp + 12 * X + 8, fp + 8);
      ctx.tmem_ld_wait();
    } else {
#pragma unroll
      for (int t3 = 0; t3 < 3; ++t3)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float* e = sm + kSmExP + ((t3 * 4) * 4 + q) * 32 + lane;
          fp[t3 * 4 + q] = ((e[0] + e[4 * 32]) + e[8 * 32]) + e[12 * 32];
        }
    }
  }
  FC_HD_CTX void shadow_glue(int X, int m) {
    const int row = traj;                                    // per-trajectory arrays are indexed by the trajectory
    const float* sw = sm + kSmSmallP;
    float fp[12];
    readout_partials(X, fp);
    const int b = (tile0 + X) * kRowsT + traj;
    const bool ok = b < p.B;
    float* rnew = w_rows(X) + (size_t)(kLook + m) * kFeat * kTileP + row;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float x = ((fp[q] + fp[4 + q]) + fp[8 + q]) + sw[(kFCB - kFCW) + q];
      if (ok) p.sh_y[((size_t)b * p.N + m) * 4 + q] = x;
      rnew[q * kTileP] = x * p.sh_ratio[q];
    }
    rnew[4 * kTileP] = (ok && m + 1 < p.N) ? p.sh_u[(size_t)b * p.N + m + 1] : 0.f;
  }

  // ---------------------------------------------------------------------------------------------
  // after window m (scalar-work thread of each trajectory): read-out, cost terms, next command
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void fwd_glue(int X, int m) {
    const int row = traj;                                    // per-trajectory arrays are indexed by the trajectory
    const float* sw = sm + kSmSmallP;
    float fp[12];
    readout_partials(X, fp);
    float x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) x[q] = ((fp[q] + fp[4 + q]) + fp[8 + q]) + sw[(kFCB - kFCW) + q];
    if (p.noise_std > 0.f) {                                                   // enable_noise, :1400-1402 / :1438-1440
      float e[4];
      philox_normal4(p.noise_seed, (unsigned)((tile0 + X) * kRowsT + traj), (unsigned)m, e);
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = fmaf(p.noise_std, e[q], x[q]);
    }
    float* rows = w_rows(X);
    const float ref = sm[kSmRefP + X * kTileP + row];
    const float ucur = Ctx::ldcg(rows + (size_t)((kLook - 1 + m) * kFeat + 4) * kTileP + row);
    const float uprev = Ctx::ldcg(rows + (size_t)((kLook - 2 + m) * kFeat + 4) * kTileP + row);
    float du = uprev - ucur;
    float cmd = p.alpha * du * du;                                             // :1405 / :1446
    float er = (x[0] - ref) * (x[0] - ref);                                    // :1408 / :1443
    float con = fmaxf(-x[1], 0.f) + fmaxf(-x[2], 0.f) + fmaxf(x[1] - kP1Max, 0.f) + fmaxf(x[2] - kP2Max, 0.f);
    float* cg = w_cost(X);
    cg[row] = Ctx::ldcg(cg + row) + ((er + cmd) + con);                        // :1414 / :1452
    cg[kTileP + row] = Ctx::ldcg(cg + kTileP + row) + cmd;
    cg[2 * kTileP + row] = Ctx::ldcg(cg + 2 * kTileP + row) + er;
    float* rnew = rows + (size_t)(kLook + m) * kFeat * kTileP + row;           // rho_{10+m} = [x_{m+1}, u_{m+1}]
#pragma unroll
    for (int q = 0; q < 4; ++q) rnew[q * kTileP] = x[q];
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
      int b = (tile0 + X) * kRowsT + traj;
      if (b < p.B) p.pred[(size_t)b * p.N + m + 1] = unext;                    // :1455
    }
    rnew[4 * kTileP] = unext;
  }

  // ---------------------------------------------------------------------------------------------
  // before the reverse sweep of window m (scalar-work thread per trajectory, then 200 accumulation threads)
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void bwd_glue_tile(int X, int m) {
    const int row = traj;                                    // per-trajectory arrays are indexed by the trajectory
    const int k = m + 1;
    const float s = p.grad_scale;
    const bool has_u = k <= p.N - 1;
    const float* sw = sm + kSmSmallP;
    const float* iw = sw + (kINPW - kFCW);
    const float* ib = sw + (kINPB - kFCW);
    const float* ow = sw + (kOUTW - kFCW);
    const bool valid = (tile0 + X) * kRowsT + traj < p.B;
    const float* rows = w_rows(X);
    const float* rx = rows + (size_t)(kLook + m) * kFeat * kTileP + row;
    float x0 = Ctx::ldcg(rx), x1 = Ctx::ldcg(rx + kTileP), x2 = Ctx::ldcg(rx + 2 * kTileP), x3 = Ctx::ldcg(rx + 3 * kTileP);
    const float ref = sm[kSmRefP + X * kTileP + row];
    float g0 = 2.f * (x0 - ref) * s;
    float g1 = s * ((x1 > kP1Max ? 1.f : 0.f) - (x1 < 0.f ? 1.f : 0.f));
    float g2 = s * ((x2 > kP2Max ? 1.f : 0.f) - (x2 < 0.f ? 1.f : 0.f));
    float g3 = 0.f;
    float dv = 0.f;
    if (has_u) {
      const float* gr = w_grow(X) + (size_t)k * kFeat * kTileP + row;
      float uk = Ctx::ldcg(rows + (size_t)((kLook - 1 + k) * kFeat + 4) * kTileP + row);
      float ukm1 = Ctx::ldcg(rows + (size_t)((kLook - 2 + k) * kFeat + 4) * kTileP + row);
      float gu = Ctx::ldcg(gr + 4 * kTileP) - 2.f * p.alpha * (ukm1 - uk) * s;
      if (k + 1 <= p.N - 1) {
        float ukp1 = Ctx::ldcg(rows + (size_t)((kLook + k) * kFeat + 4) * kTileP + row);
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
      g1 += Ctx::ldcg(gr + kTileP);
      g2 += Ctx::ldcg(gr + 2 * kTileP);
      g3 += d1 + Ctx::ldcg(gr + 3 * kTileP);
      sm[kSmDvP + X * kTileP + row] = dv;
      sm[kSmFinP + (X * 2) * kTileP + row] = x0;
      sm[kSmFinP + (X * 2 + 1) * kTileP + row] = x3;
    }
    if (!valid) { g0 = g1 = g2 = g3 = 0.f; }
    sm[kSmGxP + (X * 4 + 0) * kTileP + row] = g0;
    sm[kSmGxP + (X * 4 + 1) * kTileP + row] = g1;
    sm[kSmGxP + (X * 4 + 2) * kTileP + row] = g2;
    sm[kSmGxP + (X * 4 + 3) * kTileP + row] = g3;
  }
  FC_HD_CTX void bwd_glue(int m) {
    const bool has_u = m + 1 <= p.N - 1;
    const float* sw = sm + kSmSmallP;
    const float* iw = sw + (kINPW - kFCW);
    const float* ib = sw + (kINPB - kFCW);
    const float* ow = sw + (kOUTW - kFCW);
    if (scalar)
      for (int X = 0; X < ntl; ++X) bwd_glue_tile(X, m);
    ctx.sync();
    // controller weight gradients, unit-parallel: warps 0..6 (the idle warps first)
    if (has_u && tid < 4 * kFnnHid) {
      const int u = tid % kFnnHid, part = tid / kFnnHid;
      double a_ow = 0.0, a_b = 0.0, a_w0 = 0.0, a_w1 = 0.0, a_w2 = 0.0;   // batch sums cancel heavily: fp64
      const float w0 = iw[u * 3 + 0], w1 = iw[u * 3 + 1], w2 = iw[u * 3 + 2], bb = ib[u], owu = ow[u];
      for (int X = 0; X < ntl; ++X)
        for (int tr = part * 32; tr < part * 32 + 32 && tr < kRowsT; ++tr) {
          float dv = sm[kSmDvP + X * kTileP + tr];
          float x0 = sm[kSmFinP + (X * 2) * kTileP + tr], x3 = sm[kSmFinP + (X * 2 + 1) * kTileP + tr];
          float ref = sm[kSmRefP + X * kTileP + tr];
          float pre = fmaf(w2, ref, fmaf(w1, x3, fmaf(w0, x0, bb)));
          a_ow += (double)dv * (double)fmaxf(pre, 0.f);
          float dp = pre > 0.f ? dv * owu : 0.f;
          a_b += dp;
          a_w0 += (double)dp * (double)x0;
          a_w1 += (double)dp * (double)x3;
          a_w2 += (double)dp * (double)ref;
        }
      double* pg = reinterpret_cast<double*>(sm + kSmPgP) + part * kNumFnnGrad;
      pg[u * 3 + 0] += a_w0;
      pg[u * 3 + 1] += a_w1;
      pg[u * 3 + 2] += a_w2;
      pg[150 + u] += a_b;
      pg[200 + u] += a_ow;
    }
  }

  // ---------------------------------------------------------------------------------------------
  // backward cell gradient of NU unit slots: record(t) -> factors, dh -> d(cell), d(gates)
  //   (saving tanh(c) in the record instead of recomputing it was measured 12 % slower: +20 % record traffic)
  //   A = o(1-tanh^2 c), Ko = tanh(c) o(1-o), Ki = g i(1-i), Kf = c_prev f(1-f), Kg = i(1-g^2)
  //   dct = dc + dh A; dG = (dct Ki, dct Kf, dct Kg, dh Ko); dc = dct f
  // ---------------------------------------------------------------------------------------------
  template <int NU>
  FC_HD_CTX static void rec_load(const float* rp, int r0, float* rv) {
    constexpr int NR = (NU * 5 + 3) / 4 * 4;
#pragma unroll
    for (int r = 0; r < NR / 4; ++r) {
      F4 v = Ctx::ldg4_stream(rp + (size_t)(r0 + r) * 32 * 4);
      rv[r * 4] = v.x; rv[r * 4 + 1] = v.y; rv[r * 4 + 2] = v.z; rv[r * 4 + 3] = v.w;
    }
  }
  template <int NU>
  FC_HD_CTX void bwd_units(int j0, const float* rv, const float* dh, float* dg) {
    float cn[NU], tch[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) cn[i] = fmaf(rv[i * 5 + 1], rv[i * 5 + 4], rv[i * 5 + 0] * rv[i * 5 + 2]);
    tanh_batch<NU>(cn, tch);
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const int j = j0 + i;
      const float gi = rv[i * 5 + 0], gf = rv[i * 5 + 1], gg = rv[i * 5 + 2], go = rv[i * 5 + 3], cp = rv[i * 5 + 4];
      // s(1-s) and 1-t^2 as single fused operations: fmaf(-s, s, s), fmaf(-t, t, 1)
      const float A = go * fmaf(-tch[i], tch[i], 1.f);
      const float dct = fmaf(dh[j], A, c[j]);
      c[j] = dct * gf;
      dg[i * 4 + 0] = dct * (gg * fmaf(-gi, gi, gi));
      dg[i * 4 + 1] = dct * (cp * fmaf(-gf, gf, gf));
      dg[i * 4 + 2] = dct * (gi * fmaf(-gg, gg, 1.f));
      dg[i * 4 + 3] = dh[j] * (tch[i] * fmaf(-go, go, go));
    }
  }

  FC_HD_CTX void prefetch_record(const float* rec_in) {
    // the warp's record slots are contiguous (nf4 x 512 B): one bulk prefetch from lane 0 instead of nf4 per-lane ones
    const int nf4 = R == 1 ? (last ? 23 : 20) : (last ? 8 : 5);
    if (lane == 0) Ctx::prefetch_l2_bulk(rec_in + (size_t)uw * kRecF4 * 32 * 4, (unsigned)nf4 * 512u);
  }

  // d(h) of step t that does not come from the recurrent MMA: the layer above (thread-private scratch) or,
  // for the top layer at the last step, the read-out (Functions.py:377)
  FC_HD_CTX void bwd_extra(int X, int l, int t, float* extra) {
    if (l == kLayers - 1) {
      if (t == kLook - 1) {
        float gxv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) gxv[q] = sm[kSmGxP + (X * 4 + q) * kTileP + traj] * gsc;   // into the scaled domain
#pragma unroll
        for (int j = 0; j < kOwn; ++j) {
          const int u = u_first + j < kHid ? u_first + j : kHid - 1;
          const float* fw = sm + kSmSmallP + u;
          extra[j] = fw[0] * gxv[0] + fw[kHid] * gxv[1] + fw[2 * kHid] * gxv[2] + fw[3 * kHid] * gxv[3];
        }
      } else {
#pragma unroll
        for (int j = 0; j < kOwn; ++j) extra[j] = 0.f;
      }
    } else {
      const float* dsq = w_dseq(X) + (size_t)t * kSlot + (size_t)uw * kMaxOwn * 32 + lane;
#pragma unroll
      for (int j = 0; j < kOwn; ++j) extra[j] = j < nown ? Ctx::ldcg(dsq + j * 32) : 0.f;
    }
  }

  // result of MMA(X, l, t): d(input) of step t -> the layer below, d(h_prev) -> dh   (cell-update warps)
  FC_HD_CTX void bwd_collect(int X, int l, int t, float* dh) {
    const float corr_b = Ctx::kAccTruncates ? acc_correction(kKB / 16, p.acc_comp) : 0.0f;
    // the whole reverse sweep of a window runs in the scaled domain (gradients x g_scale, an exact power of two):
    // the gate gradients then need no multiplication before their fp16 split; only the weight scale is removed here
    const float unscale_b = 1.0f / kScaleW;
    const int dcol = col_d_bwd(X);
    if constexpr (R == 4) {
      // third th owns columns [36 th, 36 th + 36) (layer 0: [18 th, 18 th + 18)): 18 slots d(input unit), 18 slots
      // d(h_prev unit); this thread's units are slots 4 quad .. 4 quad + 3 (+ 2 for the last thread)
      float di[kOwn], dr[kOwn];
      const int cr = l > 0 ? dcol + 36 * th + 18 + 4 * quad : dcol + 18 * th + 4 * quad;
      if (l > 0) {
        ctx.template tmem_ld_nowait<4>(dcol + 36 * th + 4 * quad, di);
        if (last) ctx.template tmem_ld_nowait<2>(dcol + 36 * th + 4 * quad + 4, di + 4);
      }
      ctx.template tmem_ld_nowait<4>(cr, dr);
      if (last) ctx.template tmem_ld_nowait<2>(cr + 4, dr + 4);
      ctx.tmem_ld_wait();
      float* dq = w_dseq(X) + (size_t)t * kSlot + (size_t)uw * kMaxOwn * 32 + lane;
#pragma unroll
      for (int j = 0; j < kOwn; ++j)
        if (j < nown) {
          if (l > 0) { const float v = di[j] * unscale_b; dq[j * 32] = fmaf(v, corr_b, v); }
          const float v = dr[j] * unscale_b;
          dh[j] = fmaf(v, corr_b, v);
        } else {
          dh[j] = 0.f;
        }
      return;
    }
    if (l > 0) {
      float d[36];
      ctx.template tmem_ld_nowait<32>(dcol + 36 * th, d);
      ctx.template tmem_ld_nowait<4>(dcol + 36 * th + 32, d + 32);
      ctx.tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 2 * kMaxOwn; ++j) { d[j] *= unscale_b; d[j] = fmaf(d[j], corr_b, d[j]); }
      float* dq = w_dseq(X) + (size_t)t * kSlot + (size_t)uw * kMaxOwn * 32 + lane;
#pragma unroll
      for (int j = 0; j < kMaxOwn; ++j) {
        if (j < nown) dq[j * 32] = d[j];                   // d(input unit) -> the layer below, same thread
        dh[j] = d[kMaxOwn + j];
      }
    } else {
      float d[18];
      ctx.template tmem_ld_nowait<16>(dcol + 18 * th, d);
      ctx.template tmem_ld_nowait<2>(dcol + 18 * th + 16, d + 16);
      ctx.tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < kMaxOwn; ++j) { d[j] *= unscale_b; dh[j] = fmaf(d[j], corr_b, d[j]); }
    }
  }
  // layer 0: gradient of the row features of step t (owner thread of the row)
  FC_HD_CTX void bwd_collect_features(int X, int m, int t) {
    const int row = traj;                                    // per-trajectory arrays are indexed by the trajectory
    const float corr_b = Ctx::kAccTruncates ? acc_correction(kKB / 16, p.acc_comp) : 0.0f;
    const float unscale_b = p.g_unscale / kScaleW;
    const int kr = m + t - (kLook - 1);                    // gradient of row rho_{9+kr}
    if (kr >= 0) {
      float df[8];
      ctx.template tmem_ld<8>(col_d_bwd(X) + 56, df);
      float* gp = w_grow(X) + (size_t)kr * kFeat * kTileP + row;
#pragma unroll
      for (int f = 0; f < kFeat; ++f) {
        const float v = df[f] * unscale_b;
        Ctx::red_add(gp + f * kTileP, fmaf(v, corr_b, v));   // same thread, one add per step: deterministic order
      }
    }
  }

  // issuer warp: one step of the issue loop
  FC_HD_CTX void bwd_item_issuer(int X, int l, int t) {
    if (t < kLook - 1) {
      lap(23);
      wait_full(X);                                        // MMA(X, l, t+1) complete (keeps the phase count)
      lap(6);
    }
    wait_ready(X);
    if (X == 0 && t == kLook - 1) wait_weights();          // backward image of this layer landed
    lap(9);
    if (lane == 0) issue_bwd(X, l);
    ctx.warp_sync();
    lap(10);
  }

  // warps 1..3: stay in step (phase counts of the hand-shakes)
  FC_HD_CTX void bwd_item_scalar(int X, int t) {
    if (t < kLook - 1) wait_full(X);
    arrive_ready(X, false);                                  // nothing written
  }

  FC_HD_CTX void bwd_item(int X, int l, int m, int t) {
    const int tmin = tmin_of(m);
    float dh[kOwn];
    lap(20);
    swap_cells();
    lap(11);
    const float* rec_l = w_rec(X) + (size_t)(recb_of(m) + (long)l * kept_of(m)) * kRecFloatsP;
    // HBM -> L2 for the next step of this tile (or the first step of the next layer / window)
    if (t - 1 >= tmin) prefetch_record(rec_l + (size_t)(t - 1 - tmin) * kRecFloatsP);
    else if (l > 0) prefetch_record(w_rec(X) + (size_t)(recb_of(m) + (long)(l - 1) * kept_of(m) + (kLook - 1 - tmin)) * kRecFloatsP);
    else if (m > 0) prefetch_record(w_rec(X) + (size_t)(rec_base(m - 1) + (long)(kLayers - 1) * steps_kept(m - 1) + (kLook - 1 - t_min_of(m - 1))) * kRecFloatsP);
#ifdef FC_ABL_NO_REC_TRAFFIC
    const float* rp = w_rec(X) + ((size_t)uw * kRecF4 * 32 + lane) * 4;
#else
    const float* rp = rec_l + (size_t)(t - tmin) * kRecFloatsP + ((size_t)uw * kRecF4 * 32 + lane) * 4;
#endif
    float* dg_out = p.train == 2 ? tr_tile(X) + kTrDgOff / R + (size_t)(l * kLook + t) * (kTrDgSlot / R) : nullptr;
    float rv[2][20];
    rec_load<4>(rp, 0, rv[0]);                             // first record group: in flight during the wait
    if constexpr (R == 4) { if (last) rec_load<2>(rp, 5, rv[1]); }
    {
      float extra[kOwn];
      bwd_extra(X, l, t, extra);
      if (t == kLook - 1) {
#pragma unroll
        for (int j = 0; j < kOwn; ++j) { dh[j] = extra[j]; c[j] = 0.f; }
      } else {
        lap(12);
        wait_full(X);                                      // MMA(X, l, t+1) complete
        lap(6);
        bwd_collect(X, l, t + 1, dh);
        if (l == 0 && owner && !p.train) bwd_collect_features(X, m, t + 1);
#pragma unroll
        for (int j = 0; j < kOwn; ++j) dh[j] += extra[j];
      }
    }
    if constexpr (R == 4) {
      float dg[24];
      bwd_units<4>(0, rv[0], dh, dg);
      st_pairs_smem<8>(4 * u_first, dg, dg_out);
      if (last) {
        bwd_units<2>(4, rv[1], dh, dg + 16);
        st_pairs_smem<4>(4 * u_first + 16, dg + 16, dg_out);
      }
      lap(7);
      arrive_ready(X, true);
      lap(8);
      return;
    }
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
      // software pipeline: request the next record group before working on this one
      if (gi + 1 < 4) rec_load<4>(rp, (gi + 1) * 5, rv[(gi + 1) & 1]);
      else if (last) rec_load<2>(rp, 20, rv[(gi + 1) & 1]);
      float dg[16];
      bwd_units<4>(gi * 4, rv[gi & 1], dh, dg);
      if (X == 0) st_pairs<8>(kColGhi + 2 * u_first + gi * 8, kColGlo + 2 * u_first + gi * 8, dg, dg_out, 4 * u_first + gi * 16);
      else st_pairs_smem<8>(4 * u_first + gi * 16, dg, dg_out);
    }
    if (last) {
      float dg[8];
      bwd_units<2>(16, rv[0], dh, dg);
      if (X == 0) st_pairs<4>(kColGhi + 2 * u_first + 32, kColGlo + 2 * u_first + 32, dg, dg_out, 4 * u_first + 64);
      else st_pairs_smem<4>(4 * u_first + 64, dg, dg_out);
    }
    if (X == 0) ctx.tmem_st_wait();
    lap(7);
    arrive_ready(X, X != 0);                                 // tile 0: the operand went to TMEM, no shared-memory writes
    lap(8);
  }

  // after the last step of a layer: collect the result of MMA(X, l, tmin)
  FC_HD_CTX void bwd_tail(int X, int l, int m, bool more_after) {
    const int tmin = tmin_of(m);
    if (service && warp != 0 && !scalar) return;
    lap(22);
    wait_full(X);
    lap(6);
    if (X == ntl - 1 && tid == 0) {                        // all MMAs that read this image are complete (issuer lane)
      if (l > 0) request_weights(true, l - 1);
      else if (m > 0) request_weights(true, kLayers - 1);
      else if (more_after) request_weights(false, 0);
    }
    if (!service) {
      if (l > 0) {
        float dh[kOwn];
        bwd_collect(X, l, tmin, dh);                       // d(h) before the first kept step is not needed
      } else if (owner && !p.train) {
        bwd_collect_features(X, m, tmin);
      }
    }
    lap(14);
  }

  FC_HD_CTX void bwd_window(int m, bool more_after) {
    const int tmin = tmin_of(m);
    lap(21);
    if (p.train) {
      if (scalar)
        for (int X = 0; X < ntl; ++X) train_seed(X);
    } else {
      bwd_glue(m);
    }
    ctx.sync();
    lap(15);
    for (int l = kLayers - 1; l >= 0; --l) {
      for (int t = kLook - 1; t >= tmin; --t)
        for (int X = 0; X < ntl; ++X) {
          if (warp == 0) bwd_item_issuer(X, l, t);
          else if (!service) bwd_item(X, l, m, t);
          else if (scalar) bwd_item_scalar(X, t);
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
      for (int X = 0; X < kTiles; ++X) {
        float mine = 0.f;
        int b = (tile0 + X) * kRowsT + traj;
        if (X < ntl && b < p.B) {
          const float inv = 1.f / (float)p.N;
          const float* cg = w_cost(X);
          mine = Ctx::ldcg(cg + traj) * inv;                                   // :1458-1460
          p.cost[b] = mine;
          p.command[b] = Ctx::ldcg(cg + kTileP + traj) * inv;
          p.error[b] = Ctx::ldcg(cg + 2 * kTileP + traj) * inv;
        }
        sm[kSmGxP + X * kTileP + traj] = mine;             // gx area is free between the sweeps
        if constexpr (R == 4) {                            // the rows that do not exist contribute nothing to the sum below
#pragma unroll
          for (int g = 1; g < 4; ++g) sm[kSmGxP + X * kTileP + 32 * g + lane] = 0.f;
        }
      }
    }
    ctx.sync();
    if (tid == 0) {
      double acc = 0.0;
      for (int i = 0; i < kTiles * kTileP; ++i) acc += (double)sm[kSmGxP + i];
      *reinterpret_cast<double*>(sm + kSmRedP) += acc;
    }
    ctx.sync();
  }

  FC_HD_CTX void store_du0() {
    const int row = traj;                                    // per-trajectory arrays are indexed by the trajectory
    if (scalar)
      for (int X = 0; X < ntl; ++X) {
        int b = (tile0 + X) * kRowsT + traj;
        if (b < p.B) {
          const float s = p.grad_scale;
          const float* rows = w_rows(X);
          float u0 = Ctx::ldcg(rows + (size_t)((kLook - 1) * kFeat + 4) * kTileP + row);
          float um1 = Ctx::ldcg(rows + (size_t)((kLook - 2) * kFeat + 4) * kTileP + row);
          float g = Ctx::ldcg(w_grow(X) + 4 * kTileP + row) - 2.f * p.alpha * (um1 - u0) * s;
          if (p.N > 1) {
            float u1 = Ctx::ldcg(rows + (size_t)(kLook * kFeat + 4) * kTileP + row);
            g += 2.f * p.alpha * (u0 - u1) * s;
          }
          p.du0[b] = g;
        }
      }
  }

  // zero padding of the dG operands (k = 200..207): TMEM columns 100..103 of tile 0, the last 16-byte piece of tile 1
  FC_HD_CTX void zero_dg_padding() {
    if (scalar) {
      const F4 z4 = {0.f, 0.f, 0.f, 0.f};
      if constexpr (R == 1) {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        ctx.template tmem_st<4>(kColGhi + 100, z);
        ctx.template tmem_st<4>(kColGlo + 100, z);
        ctx.tmem_st_wait();
        Ctx::sts4(op_ptr(0, kGates), z4);
        Ctx::sts4(op_ptr(kOpGLoHalves, kGates), z4);
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          Ctx::sts4(op_ptr_rep(0, kGates, g), z4);
          Ctx::sts4(op_ptr_rep(kOpGLoHalves, kGates, g), z4);
        }
      }
    }
  }

  // ---------------------------------------------------------------------------------------------
  // persistent loop over tile pairs
  // ---------------------------------------------------------------------------------------------
  FC_HD_CTX void run() {
    ctx.tc_setup(sm + kSmBarP);
#ifdef FC_TC_TIMING
    tlast = Ctx::clock();
#endif
    if (tid == 0) {
      ctx.bar_init(kBarReady, kUpdWarps + (R == 1 ? 3 : 1));   // cell-update warps + warps 1..3 (R == 4: + warp 1)
      ctx.bar_init(kBarReady + 1, kUpdWarps + (R == 1 ? 3 : 1));
    }
    for (int i = tid; i < kSmallFloats; i += kThreadsP)
      sm[kSmSmallP + i] = !p.train ? p.wpack[kSmallOff + i]
                                   : (i < kOut * kHid ? p.tr_fcw[i] : (i < kOut * kHid + kOut ? p.tr_fcb[i - kOut * kHid] : 0.f));
    if (p.train == 2) gsc = Ctx::ldcg(p.tr_scale);
    for (int i = tid; i < 4 * kNumFnnGrad; i += kThreadsP) reinterpret_cast<double*>(sm + kSmPgP)[i] = 0.0;
    if (tid == 0) *reinterpret_cast<double*>(sm + kSmRedP) = 0.0;
    for (int i = tid; i < kSmPgP - kSmRefP; i += kThreadsP) sm[kSmRefP + i] = 0.f;   // per-row arrays (R == 4 uses 32 rows of 128)
    ctx.bar_init_fence();
    ctx.sync();
    const bool one = R != 1 || p.single_tile;                // one tile per CTA work item
    const int npairs = one ? p.num_tiles : (p.num_tiles + kTiles - 1) / kTiles;      // CTA work items
    if (tid == 0 && ctx.bid() < npairs) request_weights(false, 0);
    for (int pp = ctx.bid(); pp < npairs; pp += ctx.nblk()) {
      const bool more = pp + ctx.nblk() < npairs;
      tile0 = one ? pp : pp * kTiles;
      ntl = one ? 1 : (p.num_tiles - tile0 < kTiles ? p.num_tiles - tile0 : kTiles);
      if (scalar)
        for (int X = 0; X < ntl; ++X) {
          if (p.shadow) load_tile_shadow(X);
          else load_tile(X);
        }
      ctx.sync();
      for (int m = 0; m < p.N; ++m) fwd_window(m, more);
      if (!p.shadow && !p.train) store_costs();
      if (p.with_grad) {
        if (scalar && !p.train)
          for (int X = 0; X < ntl; ++X) {
            float* grow = w_grow(X);
            for (int k = 0; k < p.N; ++k)
#pragma unroll
              for (int f = 0; f < kFeat; ++f) grow[(size_t)(k * kFeat + f) * kTileP + traj] = 0.f;
          }
        zero_dg_padding();
        ctx.sync();
        for (int m = p.N - 1; m >= 0; --m) bwd_window(m, more);
        if (!p.train) store_du0();
      }
      ctx.tc_sync();
    }
    double* part = p.partial + (size_t)ctx.bid() * kPartialStride;
    const double* pgd = reinterpret_cast<const double*>(sm + kSmPgP);
    for (int i = tid; i < kNumFnnGrad; i += kThreadsP)
      part[i] = (pgd[i] + pgd[kNumFnnGrad + i]) + (pgd[2 * kNumFnnGrad + i] + pgd[3 * kNumFnnGrad + i]);
    if (tid == 0) part[kNumFnnGrad] = *reinterpret_cast<const double*>(sm + kSmRedP);
    ctx.sync();
#ifdef FC_TC_TRACE
    if (trace_slot >= 0) Ctx::trace_count(trace_slot, trace_n);
#endif
#ifdef FC_TC_TIMING
    if (p.debug_timing && (tid == 0 || tid == 160 || tid == 288 || tid == 416) && ctx.bid() == 0) Ctx::report_pair(tid, tm);
#endif
    ctx.tc_teardown();
  }
};

}  // namespace pr
}  // namespace fc
