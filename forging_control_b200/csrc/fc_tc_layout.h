// Layouts of the tcgen05 (3xTF32) variant of the fused MPC-loss kernel.  Shared by the device code,
// the weight packer and the CPU thread-emulation harness.  See DESIGN.md section 2.2.
//
// CTA tile = 128 trajectories = the 128 TMEM lanes.  512 threads: thread (warp w, lane i) owns TMEM
// lane / trajectory row r = 32*(w%4)+i (the only lanes warp w may touch with tcgen05.ld/st) and the
// hidden units of quarter q = w/4:  q=0 -> units 0..12, q=1 -> 13..25, q=2 -> 26..37, q=3 -> 38..49
// (13,13,12,12 units; every thread runs 13 unit slots, the 13th is masked for q >= 2).
//
// forward  phase: D[128 x 208] (TMEM fp32) = A[128 x K] (TMEM, tf32 hi/lo) * WF^T (smem [208 x K], tf32 hi/lo)
//                 gate column n = unit*4 + gate (i,f,g,o); columns 200..207 are zero padding
//                 A column k: layers 1,2: [0,50) input units | [50,100) recurrent units | 4 zero
//                             layer 0   : [0,5) row features | 3 zero | [8,58) recurrent units | 6 zero
// backward phase: D[128 x Nb] = dG[128 x 200] (TMEM hi/lo, column g = unit*4+gate) * WB^T (smem [Nb x 200])
//                 layers 1,2 (Nb=112): quarter q owns columns [26q, 26q+26): 13 slots d(input unit) then 13 slots
//                                      d(h_prev unit) of its units (unused slots have zero weights); 8 zero
//                 layer 0    (Nb=64) : quarter q owns [13q, 13q+13) d(h_prev unit); [52,57) d(row feature)
//                                      (read by quarter 0); 7 zero
// smem operand images are the canonical K-major / no-swizzle UMMA layout: [k/4][row][4 floats]
// (core matrix = 8 rows x 16 bytes contiguous; LBO = rows*16 B between K chunks, SBO = 128 B).
#pragma once
#include "fc_layout.h"

namespace fc {
namespace tc {

constexpr int kTileTC = 128;
constexpr int kNF = 208;                 // forward gate columns incl. padding
constexpr int kKF0 = 64, kKF = 104;      // forward K per layer (multiple of 8)
constexpr int kRec0 = 8, kRec = 50;      // first recurrent A column (layer 0 / layers 1,2)
constexpr int kKB = 200;                 // backward K
constexpr int kNB0 = 64, kNB = 112;      // backward output columns per layer
constexpr int kWarpsTC = 16;
constexpr int kThreadsTC = kWarpsTC * 32;
constexpr int kMaxOwn = 13;              // unit slots per thread
FC_HD int units_of(int q) { return q < 2 ? 13 : 12; }
FC_HD int first_unit(int q) { return q < 2 ? 13 * q : 26 + 12 * (q - 2); }
FC_HD int unit_quarter(int u) { return u < 13 ? 0 : (u < 26 ? 1 : (u < 38 ? 2 : 3)); }

FC_HD int kf_of(int l) { return l == 0 ? kKF0 : kKF; }
FC_HD int nb_of(int l) { return l == 0 ? kNB0 : kNB; }
FC_HD int fwd_img_floats(int l) { return kNF * kf_of(l); }          // one of hi / lo
FC_HD int bwd_img_floats(int l) { return nb_of(l) * kKB; }

// packed weight buffer (floats): per layer [hi image | lo image] forward, then backward, then the
// small fc/fnn block of the FFMA layout (kSmallFloats)
FC_HD int wf_off(int l) { return l == 0 ? 0 : 2 * fwd_img_floats(0) + (l - 1) * 2 * fwd_img_floats(1); }
constexpr int kFwdTotal = 2 * kNF * kKF0 + 4 * kNF * kKF;              // 113152
FC_HD int wb_off(int l) { return kFwdTotal + (l == 0 ? 0 : 2 * bwd_img_floats(0) + (l - 1) * 2 * bwd_img_floats(1)); }
constexpr int kBwdTotal = 2 * kNB0 * kKB + 4 * kNB * kKB;              // 115200
constexpr int kSmallOff = kFwdTotal + kBwdTotal;                       // 228352
constexpr int kPackFloatsTC = kSmallOff + kSmallFloats;                // 228808

// PyTorch gate row of gate column / dG column c = unit*4+gate
FC_HD int gate_row(int c) { return (c & 3) * kHid + (c >> 2); }

// value of element `idx` of the packed buffer BEFORE the hi/lo split (lo images repeat the hi index)
FC_HD float tc_packed_value(const RawWeights& w, int idx) {
  if (idx >= kSmallOff) return packed_value(w, kFCW + (idx - kSmallOff));
  if (idx < kFwdTotal) {
    int l = idx < wf_off(1) ? 0 : (idx < wf_off(2) ? 1 : 2);
    int r = (idx - wf_off(l)) % fwd_img_floats(l);
    int kc = r / (kNF * 4), rem = r - kc * (kNF * 4);
    int n = rem / 4, k = kc * 4 + (rem & 3);
    if (n >= kGates) return 0.f;
    int row = gate_row(n);
    if (l == 0) {
      if (k < kFeat) return w.w_ih[0][row * kFeat + k];
      if (k >= kRec0 && k < kRec0 + kHid) return w.w_hh[0][row * kHid + (k - kRec0)];
      return 0.f;
    }
    if (k < kHid) return w.w_ih[l][row * kHid + k];
    if (k < 2 * kHid) return w.w_hh[l][row * kHid + (k - kHid)];
    return 0.f;
  }
  int l = idx < wb_off(1) ? 0 : (idx < wb_off(2) ? 1 : 2);
  int nb = nb_of(l);
  int r = (idx - wb_off(l)) % bwd_img_floats(l);
  int kc = r / (nb * 4), rem = r - kc * (nb * 4);
  int n = rem / 4, g = kc * 4 + (rem & 3);
  int row = gate_row(g);
  if (l == 0) {
    if (n < 52) {
      int q = n / 13, sl = n - q * 13;
      return sl < units_of(q) ? w.w_hh[0][row * kHid + first_unit(q) + sl] : 0.f;
    }
    if (n < 57) return w.w_ih[0][row * kFeat + (n - 52)];
    return 0.f;
  }
  if (n < 104) {
    int q = n / 26, r2 = n - q * 26, sl = r2 % 13;
    if (sl >= units_of(q)) return 0.f;
    int u = first_unit(q) + sl;
    return r2 < 13 ? w.w_ih[l][row * kHid + u] : w.w_hh[l][row * kHid + u];
  }
  return 0.f;
}
FC_HD bool tc_is_lo(int idx) {
  if (idx >= kSmallOff) return false;
  if (idx < kFwdTotal) {
    int l = idx < wf_off(1) ? 0 : (idx < wf_off(2) ? 1 : 2);
    return (idx - wf_off(l)) >= fwd_img_floats(l);
  }
  int l = idx < wb_off(1) ? 0 : (idx < wb_off(2) ? 1 : 2);
  return (idx - wb_off(l)) >= bwd_img_floats(l);
}

// TMEM columns
constexpr int kColD = 0;
constexpr int kColAhi = 256, kColAlo = 384;          // forward A operand
constexpr int kColGhi = 112, kColGlo = 312;          // backward A operand (dG), 200 columns each

// One accumulator group per step: every tcgen05.mma costs ~100 cycles whatever its N (the A operand is
// fetched from TMEM per instruction), so column chunking to overlap cell update and MMA does not pay
// (measured: 39 MMAs N=208 4340 cycles, 78 MMAs N=96/112 7490, 156 MMAs N=48/64 15400).

// per-CTA global workspace (floats); every slot is private to one thread
//   rows [(N+10)][5][128], seq [10][16][13][32], dseq [10][16][13][32], grow [N][5][128],
//   rec [nrec][16][17][32] float4   (13 units x (i,f,g,o,c_prev) = 65 floats -> 17 float4)
constexpr int kSlot = kWarpsTC * kMaxOwn * 32;               // 6656
constexpr int kRecF4 = 17;
constexpr int kRecFloatsTC = kWarpsTC * kRecF4 * 32 * 4;    // 34816
struct WorkLayoutTC {
  size_t rows, seq, dseq, grow, rec, total;
};
FC_HD WorkLayoutTC work_layout_tc(int N, int with_grad) {
  WorkLayoutTC w;
  w.rows = 0;
  w.seq = w.rows + (size_t)(N + kLook) * kFeat * kTileTC;
  w.dseq = w.seq + (size_t)kLook * kSlot;
  w.grow = w.dseq + (with_grad ? (size_t)kLook * kSlot : 0);
  w.rec = w.grow + (with_grad ? (size_t)N * kFeat * kTileTC : 0);
  w.rec = (w.rec + 31) / 32 * 32;
  w.total = w.rec + (with_grad ? (size_t)rec_base(N) * kRecFloatsTC : 0);
  w.total = (w.total + 31) / 32 * 32;
  return w;
}

// shared memory (floats)
constexpr int kSmSmallTC = 0;                               // fc + fnn weights (456)
constexpr int kSmRefTC = kSmSmallTC + kSmallFloats;         // [128]
constexpr int kSmUcurTC = kSmRefTC + kTileTC;
constexpr int kSmUprevTC = kSmUcurTC + kTileTC;
constexpr int kSmCostTC = kSmUprevTC + kTileTC;             // [3][128]
constexpr int kSmGxTC = kSmCostTC + 3 * kTileTC;            // [4][128]
constexpr int kSmDvTC = kSmGxTC + 4 * kTileTC;              // [128]
constexpr int kSmFinTC = kSmDvTC + kTileTC;                 // [2][128]
constexpr int kSmFcpTC = kSmFinTC + 2 * kTileTC;            // [3][4][128] read-out partial sums of quarters 1..3
constexpr int kSmPgTC = kSmFcpTC + 12 * kTileTC;            // double [4][250]
constexpr int kSmRedTC = kSmPgTC + 8 * kNumFnnGrad;         // double [4]
static_assert(kSmPgTC % 2 == 0, "double alignment");
constexpr int kSmBarTC = ((kSmRedTC + 8 + 3) / 4) * 4;      // 8 mbarriers (64-bit) + tmem base
constexpr int kSmWTC = ((kSmBarTC + 24 + 255) / 256) * 256; // operand images, 1 KiB aligned
constexpr int kSmWFloats = 2 * kNB * kKB;                   // 44800 >= 2*208*104 = 43264
constexpr int kSmFloatsTC = kSmWTC + kSmWFloats;
constexpr size_t kSmBytesTC = (size_t)kSmFloatsTC * sizeof(float);
static_assert(kSmBytesTC <= 227 * 1024, "shared memory budget exceeded (tcgen05 variant)");
static_assert(2 * kNF * kKF <= kSmWFloats, "forward image does not fit");

// Expected-value compensation of the tensor-core accumulator.  tcgen05.mma adds every K=8 block into the
// fp32 accumulator with TRUNCATION toward zero, so a chain of S accumulation steps of the hi*hi term
// shrinks |D| by (0.17 + 0.135*S) * 2^-23 * |D| on average, independent of the data distribution
// (measured on B200 with scripts/micro/umma_test.cu: S=8 -> 1.25 ulp, S=13 -> 1.95, S=25 -> 3.55; a
// software model of a truncating accumulator gives the same law).  Left alone this systematic shrink does
// not average out over the batch and costs a factor ~4 in gradient accuracy; multiplying the accumulator
// by (1 + beta) -- as fmaf(x, beta, x), beta is a fraction of an ulp -- removes the mean and halves the rms
// error (to that of an fp32 FMA chain).
// returns beta; apply as fmaf(x, beta, x) so that fractions of an ulp are honoured in expectation
FC_HD float acc_correction(int steps, float scale) { return scale * (0.17f + 0.135f * (float)steps) * 1.1920929e-7f; }

// mbarrier ids
constexpr int kBarChunk0 = 0;      // 0..3: forward chunk accumulators ready / backward: 0,1
constexpr int kBarWeights = 4;     // bulk copy of an operand image landed

}  // namespace tc
}  // namespace fc
