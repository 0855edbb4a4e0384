// Layouts of the tcgen05 variant of the fused MPC-loss kernel (fp16 hi/lo split operands, fp32 accumulate).
// Shared by the device code, the weight packer and the CPU thread-emulation harness.  See DESIGN.md section 2.2.
//
// CTA tile = 128 trajectories = the 128 TMEM lanes.  512 threads: thread (warp w, lane i) owns TMEM
// lane / trajectory row r = 32*(w%4)+i (the only lanes warp w may touch with tcgen05.ld/st) and the
// hidden units of quarter q = w/4:  q=0 -> units 0..12, q=1 -> 13..25, q=2 -> 26..37, q=3 -> 38..49
// (13,13,12,12 units; every thread runs 13 unit slots, the 13th is masked for q >= 2).
//
// Operands are fp16 hi/lo pairs of power-of-two pre-scaled values (x*S = hi + lo, 2 x 11 bits): kind::f16 MMAs
// take K=16 per instruction, half the instruction count of 3xTF32 for the same accuracy (measured,
// profiles/r01_umma_fp16x2_feasibility.txt).  Two fp16 K-elements share one 32-bit TMEM column.
//
// K ordering: hidden units are laid out in per-quarter blocks of even size so that every thread owns whole
// TMEM columns: block slot kb(u) = BS[q] + (u - first_unit(q)), BS = {0,14,28,40}, 52 slots (slot 13 of
// quarters 0,1 is zero padding).
// forward  phase: D[128 x 208] (TMEM fp32) = A[128 x K] (TMEM fp16 hi/lo) * WF^T (smem [208 x K] fp16 hi/lo)
//                 gate column n = unit*4 + gate (i,f,g,o); columns 200..207 are zero padding
//                 A k-index: layers 1,2 (K=112): [0,52) input slots | [52,104) recurrent slots | 8 zero
//                            layer 0   (K=64) : [0,5) row features | 3 zero | [8,60) recurrent slots | 4 zero
// backward phase: D[128 x Nb] = dG[128 x 208] (TMEM fp16 hi/lo, k = unit*4+gate, 8 zero) * WB^T (smem [Nb x 208])
//                 layers 1,2 (Nb=112): quarter q owns columns [26q, 26q+26): 13 slots d(input unit) then 13 slots
//                                      d(h_prev unit) of its units (unused slots have zero weights); 8 zero
//                 layer 0    (Nb=64) : quarter q owns [13q, 13q+13) d(h_prev unit); [52,57) d(row feature)
//                                      (read by quarter 0); 7 zero
// smem operand images are the canonical K-major / no-swizzle UMMA layout for 16-bit types: [k/8][row][8 halves]
// (core matrix = 8 rows x 16 bytes contiguous; LBO = rows*16 B between K chunks, SBO = 128 B).
#pragma once
#include "fc_layout.h"

namespace fc {
namespace tc {

constexpr int kTileTC = 128;
constexpr int kNF = 208;                 // forward gate columns incl. padding
constexpr int kKF0 = 64, kKF = 112;      // forward K per layer (multiple of 16)
constexpr int kSlots = 52;               // unit slots (per-quarter blocks 14,14,12,12)
constexpr int kRec0 = 8, kRec = 52;      // first recurrent k-index (layer 0 / layers 1,2)
constexpr int kKB = 208;                 // backward K (200 gate gradients + 8 zero)
constexpr int kNB0 = 64, kNB = 112;      // backward output columns per layer
constexpr int kWarpsTC = 16;
constexpr int kThreadsTC = kWarpsTC * 32;
constexpr int kMaxOwn = 13;              // unit slots per thread
// power-of-two operand scales (exact): activations / row features x 2^10 (|x| < 58), weights x 2^11 (|w| < 29);
// gate gradients x 2^g_exp chosen per launch from N * B_global
constexpr float kScaleA = 1024.0f, kScaleW = 2048.0f;
constexpr float kHalfMax = 60000.0f;     // saturation bound applied before the fp16 conversion

FC_HD int units_of(int q) { return q < 2 ? 13 : 12; }
FC_HD int first_unit(int q) { return q < 2 ? 13 * q : 26 + 12 * (q - 2); }
FC_HD int block_start(int q) { return q == 0 ? 0 : (q == 1 ? 14 : (q == 2 ? 28 : 40)); }
// unit of block slot kb (or -1 for padding slots)
FC_HD int slot_unit(int kb) {
  int q = kb < 14 ? 0 : (kb < 28 ? 1 : (kb < 40 ? 2 : 3));
  int j = kb - block_start(q);
  return j < units_of(q) ? first_unit(q) + j : -1;
}

FC_HD int kf_of(int l) { return l == 0 ? kKF0 : kKF; }
FC_HD int nb_of(int l) { return l == 0 ? kNB0 : kNB; }
// operand images are fp16: sizes in HALVES for one of hi / lo (multiples of 8); hi + lo = that many floats
FC_HD int fwd_img_halves(int l) { return kNF * kf_of(l); }
FC_HD int bwd_img_halves(int l) { return nb_of(l) * kKB; }

// packed weight buffer (offsets in floats = 2 halves): per layer [hi image | lo image] forward, then backward,
// then the small fc/fnn block of the FFMA layout (kSmallFloats, fp32)
FC_HD int wf_off(int l) { return l == 0 ? 0 : fwd_img_halves(0) + (l - 1) * fwd_img_halves(1); }
constexpr int kFwdTotal = kNF * kKF0 + 2 * kNF * kKF;                  // 59904
FC_HD int wb_off(int l) { return kFwdTotal + (l == 0 ? 0 : bwd_img_halves(0) + (l - 1) * bwd_img_halves(1)); }
constexpr int kBwdTotal = kNB0 * kKB + 2 * kNB * kKB;                  // 59904
constexpr int kSmallOff = kFwdTotal + kBwdTotal;                       // 119808
constexpr int kPackFloatsTC = kSmallOff + kSmallFloats;                // 120264

// PyTorch gate row of gate column / dG index c = unit*4+gate
FC_HD int gate_row(int c) { return (c & 3) * kHid + (c >> 2); }

// UNSCALED weight behind half-element h of the forward image of layer l: h = (k/8)*(208*8) + n*8 + k%8
FC_HD float fwd_weight(const RawWeights& w, int l, int h) {
  int kc = h / (kNF * 8), rem = h - kc * (kNF * 8);
  int n = rem / 8, k = kc * 8 + (rem & 7);
  if (n >= kGates) return 0.f;
  int row = gate_row(n);
  if (l == 0) {
    if (k < kFeat) return w.w_ih[0][row * kFeat + k];
    if (k >= kRec0 && k < kRec0 + kSlots) {
      int u = slot_unit(k - kRec0);
      return u < 0 ? 0.f : w.w_hh[0][row * kHid + u];
    }
    return 0.f;
  }
  if (k < kSlots) {
    int u = slot_unit(k);
    return u < 0 ? 0.f : w.w_ih[l][row * kHid + u];
  }
  if (k < 2 * kSlots) {
    int u = slot_unit(k - kSlots);
    return u < 0 ? 0.f : w.w_hh[l][row * kHid + u];
  }
  return 0.f;
}
// backward image of layer l: h = (g/8)*(Nb*8) + n*8 + g%8, g = gate-gradient index unit*4+gate (>= 200: zero)
FC_HD float bwd_weight(const RawWeights& w, int l, int h) {
  const int nb = nb_of(l);
  int kc = h / (nb * 8), rem = h - kc * (nb * 8);
  int n = rem / 8, g = kc * 8 + (rem & 7);
  if (g >= kGates) return 0.f;
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
// decode a half index of the packed buffer: which image, layer, hi/lo and element
struct TcSlot { int kind; int l; int lo; int h; };   // kind 0 = forward, 1 = backward
FC_HD TcSlot tc_decode_half(long hidx) {             // hidx counts halves from the start of the TC pack buffer
  TcSlot s;
  const long f2 = 2L * kFwdTotal;
  if (hidx < f2) {
    s.kind = 0;
    s.l = hidx < 2L * wf_off(1) ? 0 : (hidx < 2L * wf_off(2) ? 1 : 2);
    long r = hidx - 2L * wf_off(s.l);
    s.lo = r >= fwd_img_halves(s.l) ? 1 : 0;
    s.h = (int)(r - (s.lo ? fwd_img_halves(s.l) : 0));
  } else {
    s.kind = 1;
    s.l = hidx < 2L * wb_off(1) ? 0 : (hidx < 2L * wb_off(2) ? 1 : 2);
    long r = hidx - 2L * wb_off(s.l);
    s.lo = r >= bwd_img_halves(s.l) ? 1 : 0;
    s.h = (int)(r - (s.lo ? bwd_img_halves(s.l) : 0));
  }
  return s;
}

// TMEM columns
constexpr int kColD = 0;
constexpr int kColAhi = 256, kColAlo = 320;          // forward A operand: K/2 = 56 columns each (2 fp16 per column)
constexpr int kColGhi = 112, kColGlo = 216;          // backward A operand (dG): 104 columns each

// One accumulator group per step: every tcgen05.mma costs ~100 cycles whatever its N (the A operand is
// fetched from TMEM per instruction), so column chunking to overlap cell update and MMA does not pay
// (measured with tf32: 39 MMAs N=208 4340 cycles, 78 MMAs N=96/112 7490, 156 MMAs N=48/64 15400).

// per-CTA global workspace (floats); every slot is private to one thread
//   rows [(N+10)][5][128], seq [10][16][13][32], dseq [10][16][13][32], grow [N][5][128],
//   rec [nrec][16][17][32] float4   (13 units x (i,f,g,o,c_prev) = 65 floats -> 17 float4)
constexpr int kSlot = kWarpsTC * kMaxOwn * 32;               // 6656
constexpr int kRecF4 = 17;
constexpr int kRecFloatsTC = kWarpsTC * kRecF4 * 32 * 4;    // 34816
struct WorkLayoutTC {
  size_t rows, seq, dseq, grow, rec, total, act;
};
constexpr int kActStride = kTileTC + 1;                      // padded row stride of the [50][128] controller buffers
constexpr int kActFloats = kFnnHid * kActStride;             // 6450
FC_HD WorkLayoutTC work_layout_tc(int N, int with_grad, int width_dim = 1) {
  WorkLayoutTC w;
  w.rows = 0;
  w.seq = w.rows + (size_t)(N + kLook) * kFeat * kTileTC;
  w.dseq = w.seq + (size_t)kLook * kSlot;
  w.grow = w.dseq + (with_grad ? (size_t)kLook * kSlot : 0);
  w.rec = w.grow + (with_grad ? (size_t)N * kFeat * kTileTC : 0);
  w.rec = (w.rec + 31) / 32 * 32;
  w.total = w.rec + (with_grad ? (size_t)rec_base(N) * kRecFloatsTC : 0);
  w.total = (w.total + 31) / 32 * 32;
  w.act = w.total;                                            // width_dim > 1: hidden activations a_0..a_R, [R+1][50][128]
  if (width_dim > 1 && with_grad) w.total += (size_t)width_dim * kFnnHid * kTileTC;
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
constexpr int kSmWFloats = kNB * kKB;                       // hi + lo fp16 images: 2*112*208 halves = 23296 floats
constexpr int kSmFloatsTC = kSmWTC + kSmWFloats;
constexpr size_t kSmBytesTC = (size_t)kSmFloatsTC * sizeof(float);
static_assert(kSmBytesTC <= 227 * 1024, "shared memory budget exceeded (tcgen05 variant)");
// width_dim > 1 controllers: extra shared memory behind the operand image (only launches with width_dim > 1 ask for it)
constexpr int kSmWideW = kSmFloatsTC;                         // fc_int.weight [50][50] | fc_int.bias [50] (+2 pad)
constexpr int kSmWideIn = kSmWideW + kFnnHid * kFnnHid + kFnnHid + 2;   // [3][128] controller inputs x0, x3, ref
constexpr int kSmWideA = kSmWideIn + 3 * kTileTC;             // three [50][129] buffers (activations / deltas / staging)
constexpr int kSmWideAcc = ((kSmWideA + 3 * kActFloats + 1) / 2) * 2;   // double [2550] gradient accumulators
constexpr int kSmFloatsWide = kSmWideAcc + 2 * kWideGrads;
constexpr size_t kSmBytesWide = (size_t)kSmFloatsWide * sizeof(float);
static_assert(kSmBytesWide <= 227 * 1024, "shared memory budget exceeded (tcgen05 variant, wide controller)");
static_assert(kNF * kKF <= kSmWFloats, "forward image does not fit");

// Expected-value compensation of the tensor-core accumulator.  tcgen05.mma adds every K-block into the fp32
// accumulator with TRUNCATION toward zero, so a chain of S accumulation steps of the hi*hi term shrinks |D| by
// (0.17 + 0.135*S) * 2^-23 * |D| on average, independent of the data distribution (measured on B200 with
// scripts/micro/umma_test.cu, kind::tf32: S=8 -> 1.25 ulp, S=13 -> 1.95, S=25 -> 3.55; kind::f16
// (umma_f16_test.cu): S=7 -> 1.17, S=13 -> 2.0; a software model of a truncating accumulator gives the same law).
// Left alone this systematic shrink does not average out over the batch and costs a factor ~4 in gradient
// accuracy; multiplying the accumulator by (1 + beta) -- as fmaf(x, beta, x), beta is a fraction of an ulp --
// removes the mean and halves the rms error (to that of an fp32 FMA chain).
// returns beta; apply as fmaf(x, beta, x) so that fractions of an ulp are honoured in expectation
FC_HD float acc_correction(int steps, float scale) { return scale * (0.17f + 0.135f * (float)steps) * 1.1920929e-7f; }

// mbarrier ids
constexpr int kBarChunk0 = 0;      // accumulator of the current step complete
constexpr int kBarWeights = 4;     // bulk copy of an operand image landed

}  // namespace tc
}  // namespace fc
