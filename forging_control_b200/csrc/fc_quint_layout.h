// Layouts of the five-group ("quint") two-tile tcgen05 variant of the fused MPC-loss kernel.  See DESIGN.md section 2.3b.
//
// Same arithmetic and the same two-tiles-per-CTA schedule as fc_pair_layout.h (fp16 hi/lo split operands, fp32
// accumulate in TMEM, one trajectory per TMEM lane, the MMA of one tile under the cell update of the other), but the
// hidden units are split over FIVE groups of exactly 10 units instead of three of 16/16/18:
//   * 20 cell-update warps (5 per SM sub-partition instead of 3) cover each other's latencies (mbarrier hand-shakes,
//     TMEM round trips, activation-record loads) -- the round-1 kernel left 45 % of the issue slots empty;
//   * the groups are perfectly balanced (the 18-unit third was 12 % slower than the 16-unit ones, everybody waited);
//   * 10 units per thread are few enough to keep the cell state of BOTH tiles in registers (2 x 10), so the
//     park-and-swap of the cell state through spare TMEM columns at every item disappears.
//
// 768 threads = 24 warps (6 per SM sub-partition: 80 registers per thread; a 25th warp would put 7 warps on one
// sub-partition and cap everybody at 72); thread (warp w, lane i) works on TMEM lane / trajectory row r = 32*(w%4)+i
// of BOTH tiles.
//   warps 0..3    per-trajectory scalar work (roll-out rows, layer-0 features, read-out, cost terms, controller, row-feature
//                 gradients) of the rows of their TMEM quadrant; warp 0 is also the MMA issuer (all 32 lanes walk the
//                 issue loop, lane 0 issues): it idles two thirds of the time waiting for the cell updates, which is
//                 where its scalar work goes
//   warps 4..23   cell update: group g = (w-4)/4 owns the hidden units [10 g, 10 g + 10)
//
// Operand K slots.  The K-major operand images are written in 16-byte pieces (8 halves); 10 units per group do not
// align to pieces, so the unit -> slot map is permuted: the first 8 units of group g fill piece g (slots 8g..8g+7), the
// last 2 units of the groups share pieces 5 and 6 (slot 40 + 2g + j).  Every thread writes one 16-byte piece and one
// 4-byte pair per image.  Slots 50..55 are zero.  The weight images are packed with the same permutation.
//
// forward  : D[128 x 208] = A[128 x K] * WF^T (smem [208 x K]); gate column n = unit*4 + gate (i,f,g,o), 8 zero
//            A k-index: layers 1,2 (K=112): [0,56) slot of input unit | [56,112) slot of recurrent unit
//                       layer 0   (K=64) : [0,5) row features | 3 zero | [8,64) slot of recurrent unit
// backward : D[128 x Nb] = dG[128 x 208] (k = unit*4+gate, natural order: group g = pieces 5g..5g+4; 8 zero) * WB^T
//            layers 1,2 (Nb=112): group g owns columns [20 g, 20 g + 20), interleaved by pairs of units so that one
//                                 4-column TMEM load serves a pair: column 20 g + 4 k + q = d(input unit 10g+2k+q) for q < 2,
//                                 d(h_prev unit 10g+2k+q-2) for q >= 2
//            layer 0    (Nb=64) : group g owns [10 g, 10 g + 10) d(h_prev unit); [56,61) d(row feature) (scalar warps)
#pragma once
#include "fc_layout.h"

namespace fc {
namespace q5 {

constexpr int kTileQ = 128;
constexpr int kTilesQ = 2;               // tiles in flight per CTA
constexpr int kGroups = 5;
constexpr int kOwn = 10;                 // hidden units per group / thread
constexpr int kNF = 208;
constexpr int kKF0 = 64, kKF = 112;
constexpr int kRec0 = 8, kRec = 56;      // first recurrent k-index (layer 0 / layers 1,2)
constexpr int kKB = 208;
constexpr int kNB0 = 64, kNB = 112;
constexpr int kWarpsQ = 24;
constexpr int kThreadsQ = kWarpsQ * 32;  // 768
constexpr int kUpdWarpsQ = 20;           // cell-update warps (4..23)
constexpr int kScalarWarpsQ = 3;         // warps 1, 2, 3 arrive on the ready barriers (warp 0 is the one that waits)
constexpr float kScaleA = 1024.0f, kScaleW = 2048.0f;   // exact power-of-two operand scales (see fc_tc_layout.h)

FC_HD int slot_of(int u) { const int g = u / kOwn, j = u - g * kOwn; return j < 8 ? 8 * g + j : 40 + 2 * g + (j - 8); }
FC_HD int unit_of_slot(int s) {           // -1: zero padding
  if (s < 40) return kOwn * (s >> 3) + (s & 7);
  if (s < 50) return kOwn * ((s - 40) >> 1) + 8 + ((s - 40) & 1);
  return -1;
}

FC_HD int kf_of(int l) { return l == 0 ? kKF0 : kKF; }
FC_HD int nb_of(int l) { return l == 0 ? kNB0 : kNB; }
FC_HD int fwd_img_halves(int l) { return kNF * kf_of(l); }     // one of hi / lo
FC_HD int bwd_img_halves(int l) { return nb_of(l) * kKB; }

// packed weight buffer (offsets in floats = 2 halves): per layer [hi image | lo image] forward, then backward,
// then the small fc/fnn block (kSmallFloats, fp32)
FC_HD int wf_off(int l) { return l == 0 ? 0 : fwd_img_halves(0) + (l - 1) * fwd_img_halves(1); }
constexpr int kFwdTotal = kNF * kKF0 + 2 * kNF * kKF;
FC_HD int wb_off(int l) { return kFwdTotal + (l == 0 ? 0 : bwd_img_halves(0) + (l - 1) * bwd_img_halves(1)); }
constexpr int kBwdTotal = kNB0 * kKB + 2 * kNB * kKB;
constexpr int kSmallOff = kFwdTotal + kBwdTotal;
constexpr int kPackFloatsQ = kSmallOff + kSmallFloats;

FC_HD int gate_row(int c) { return (c & 3) * kHid + (c >> 2); }   // PyTorch gate row of column unit*4+gate

// UNSCALED weight behind half-element h of the forward image of layer l: h = (k/8)*(208*8) + n*8 + k%8
FC_HD float fwd_weight(const RawWeights& w, int l, int h) {
  const int kc = h / (kNF * 8), rem = h - kc * (kNF * 8);
  const int n = rem / 8, k = kc * 8 + (rem & 7);
  if (n >= kGates) return 0.f;
  const int row = gate_row(n);
  if (l == 0) {
    if (k < kFeat) return w.w_ih[0][row * kFeat + k];
    if (k >= kRec0) { const int u = unit_of_slot(k - kRec0); return u >= 0 ? w.w_hh[0][row * kHid + u] : 0.f; }
    return 0.f;
  }
  if (k < kRec) { const int u = unit_of_slot(k); return u >= 0 ? w.w_ih[l][row * kHid + u] : 0.f; }
  const int u = unit_of_slot(k - kRec);
  return u >= 0 ? w.w_hh[l][row * kHid + u] : 0.f;
}
// backward image of layer l: h = (g/8)*(Nb*8) + n*8 + g%8, g = gate-gradient index unit*4+gate (>= 200: zero)
FC_HD float bwd_weight(const RawWeights& w, int l, int h) {
  const int nb = nb_of(l);
  const int kc = h / (nb * 8), rem = h - kc * (nb * 8);
  const int n = rem / 8, g = kc * 8 + (rem & 7);
  if (g >= kGates) return 0.f;
  const int row = gate_row(g);
  if (l == 0) {
    if (n < kHid) return w.w_hh[0][row * kHid + n];
    if (n >= 56 && n < 56 + kFeat) return w.w_ih[0][row * kFeat + (n - 56)];
    return 0.f;
  }
  if (n >= 2 * kHid) return 0.f;
  const int grp = n / (2 * kOwn), r = n - grp * 2 * kOwn, k = r >> 2, q = r & 3;
  const int u = grp * kOwn + 2 * k + (q & 1);
  return q < 2 ? w.w_ih[l][row * kHid + u] : w.w_hh[l][row * kHid + u];
}
struct QSlot { int kind; int l; int lo; int h; };   // kind 0 = forward, 1 = backward
FC_HD QSlot decode_half(long hidx) {                // hidx counts halves from the start of the quint pack buffer
  QSlot s;
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
FC_HD int col_d_fwd(int tile) { return kNF * tile; }          // 0, 208
constexpr int kColFcp = 416;                                  // + 20*tile + 4*group: read-out partial sums
FC_HD int col_d_bwd(int tile) { return kNB * tile; }          // 0, 112
constexpr int kColGhi = 224, kColGlo = 328;                   // dG operand of tile 0 (104 columns each)

// per-TILE global workspace (floats); every slot is private to one thread
//   rows [(N+10)][5][128], cost [3][128],
//   seq  [10][20 warps][hi8: 32 x float4 | lo8: 32 x float4 | pairs: 32 x float2]   (fp16 pieces in operand format)
//   dseq [10][20 warps][10][32], grow [N][5][128],
//   rec  [nrec][20 warps][ 10 x 32 float4 (i,f,g,o of unit j) | 5 x 32 float2 (c_prev of the unit pair k) ]
constexpr int kSeqWarp = 32 * 4 + 32 * 4 + 32 * 2;           // 320 floats per warp and step
constexpr int kSeqSlot = kUpdWarpsQ * kSeqWarp;              // 6400 floats per step
constexpr int kDseqSlot = kUpdWarpsQ * kOwn * 32;            // 6400
constexpr int kRecWarp = kOwn * 32 * 4 + (kOwn / 2) * 32 * 2;   // 1600 floats = 6400 B per warp, contiguous
constexpr int kRecCp = kOwn * 32 * 4;                         // offset of the c_prev pairs inside a warp's block
constexpr int kRecFloatsQ = kUpdWarpsQ * kRecWarp;            // 32000
struct WorkLayoutQ {
  size_t rows, cost, seq, dseq, grow, rec, total;
};
FC_HD WorkLayoutQ work_layout_q(int N, int with_grad) {
  WorkLayoutQ w;
  w.rows = 0;
  w.cost = w.rows + (size_t)(N + kLook) * kFeat * kTileQ;
  w.seq = w.cost + 3 * kTileQ;
  w.dseq = w.seq + (size_t)kLook * kSeqSlot;
  w.grow = w.dseq + (with_grad ? (size_t)kLook * kDseqSlot : 0);
  w.rec = w.grow + (with_grad ? (size_t)N * kFeat * kTileQ : 0);
  w.rec = (w.rec + 31) / 32 * 32;
  w.total = w.rec + (with_grad ? (size_t)rec_base(N) * kRecFloatsQ : 0);
  w.total = (w.total + 31) / 32 * 32;
  return w;
}

// shared memory (floats)
constexpr int kSmSmallQ = 0;                                 // fc + fnn weights (456)
constexpr int kSmRefQ = kSmSmallQ + kSmallFloats;            // [2][128]
constexpr int kSmGxQ = kSmRefQ + kTilesQ * kTileQ;           // [2][4][128]
constexpr int kSmDvQ = kSmGxQ + kTilesQ * 4 * kTileQ;        // [2][128]
constexpr int kSmFinQ = kSmDvQ + kTilesQ * kTileQ;           // [2][2][128]
constexpr int kSmPgQ = kSmFinQ + kTilesQ * 2 * kTileQ;       // double [4][250]
static_assert(kSmPgQ % 2 == 0, "double alignment");
constexpr int kSmRedQ = kSmPgQ + 8 * kNumFnnGrad;            // double
constexpr int kSmBarQ = ((kSmRedQ + 2 + 3) / 4) * 4;         // 8 mbarriers (64-bit) + tmem base
constexpr int kSmWQ = ((kSmBarQ + 24 + 255) / 256) * 256;    // weight image, 1 KiB aligned
constexpr int kSmWFloatsQ = kNB * kKB;                       // hi + lo fp16 images = 23296 floats
constexpr int kSmOpQ = kSmWQ + kSmWFloatsQ;                  // operand region
constexpr int kOpTileFloats = kTileQ * kKF;                  // A operand of one tile: hi + lo = 2*128*112 halves
constexpr int kSmOpFloats = kTilesQ * kOpTileFloats;         // 28672 floats = 114 688 B
constexpr int kSmFloatsQ = kSmOpQ + kSmOpFloats;
constexpr size_t kSmBytesQ = (size_t)kSmFloatsQ * sizeof(float);
static_assert(kSmBytesQ <= 227 * 1024, "shared memory budget exceeded (quint variant)");
static_assert(kTileQ * kKB <= kSmOpFloats, "dG operand of tile 1 does not fit the operand region");
// halves offsets inside the operand region
constexpr int kOpLoHalves = kTileQ * kKF;                    // forward: lo image of a tile follows its hi image
FC_HD int op_fwd_halves(int tile) { return tile * 2 * kTileQ * kKF; }
constexpr int kOpGLoHalves = kTileQ * kKB;                   // backward: lo image of dG (tile 1) follows the hi image

FC_HD float acc_correction(int steps, float scale) { return scale * (0.17f + 0.135f * (float)steps) * 1.1920929e-7f; }

// mbarrier ids
constexpr int kBarFull = 0;        // + tile: accumulator of the tile complete (tcgen05.commit)
constexpr int kBarReady = 2;       // + tile: operand of the tile written (one arrival per warp)
constexpr int kBarWeightsQ = 4;    // bulk copy of a weight image landed

}  // namespace q5
}  // namespace fc
