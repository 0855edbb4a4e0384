// Layouts of the two-tile ("pair") tcgen05 variant of the fused MPC-loss kernel.  See DESIGN.md section 2.3.
//
// Same arithmetic as fc_tc_layout.h (fp16 hi/lo split operands, fp32 accumulate in TMEM, one trajectory per
// TMEM lane) but every CTA works on TWO 128-trajectory tiles at once and alternates between them: while the
// tensor core runs the gate contraction of one tile, the 16 warps do the cell update of the other, so that
// neither the tensor pipe nor the MUFU/FMA pipes wait for each other.  What makes two tiles fit:
//   TMEM (512 columns)   forward : accumulators D[tile] = 208 columns each (416) + 24 read-out exchange columns
//                        backward: D[tile] = 112 columns each (224) + the dG operand of tile 0 (2 x 104)
//   shared memory        one weight image (93 184 B) + an operand region (114 688 B):
//                        forward : A operand (activations) of both tiles, hi and lo    (SS-mode MMA)
//                        backward: dG operand of tile 1, hi and lo                      (tile 0: TS-mode)
//
// 512 threads: thread (warp w, lane i) works on TMEM lane / trajectory row r = 32*(w%4)+i of BOTH tiles.
//   warps 0..3  (service): thread 0 issues every MMA (tcgen05.mma blocks its issuer for about the duration of the
//               chain, so the issuer must not carry cell-update work); all 128 threads do the per-trajectory scalar
//               work of their row (roll-out rows, layer-0 features, read-out, cost, controller)
//   warps 4..15 (cell update): third th = w/4 - 1 owns the hidden units [16 th, 16 th + 16) (th = 2: [32, 50), 18 units).
//               The ranges start at multiples of 8 so that every thread writes whole 16-byte pieces of the K-major
//               operand images ([k/8][row][8 halves], core matrix = 8 rows x 16 bytes).
//
// forward  : D[128 x 208] = A[128 x K] * WF^T (smem [208 x K]); gate column n = unit*4 + gate (i,f,g,o), 8 zero
//            A k-index: layers 1,2 (K=112): [0,56) input unit k | [56,112) recurrent unit k-56   (units >= 50: zero)
//                       layer 0   (K=64) : [0,5) row features | 3 zero | [8,64) recurrent unit k-8
// backward : D[128 x Nb] = dG[128 x 208] (k = unit*4+gate, 8 zero) * WB^T (smem [Nb x 208])
//            layers 1,2 (Nb=112): third th owns columns [36 th, 36 th + 36): 18 slots d(input unit), 18 slots d(h_prev unit)
//            layer 0    (Nb=64) : third th owns [18 th, 18 th + 18) d(h_prev unit); [56,61) d(row feature) (service warps)
#pragma once
#include "fc_layout.h"

namespace fc {
namespace pr {

constexpr int kTileP = 128;
constexpr int kTiles = 2;                // tiles in flight per CTA
constexpr int kNF = 208;
constexpr int kKF0 = 64, kKF = 112;
constexpr int kRec0 = 8, kRec = 56;      // first recurrent k-index (layer 0 / layers 1,2)
constexpr int kKB = 208;
constexpr int kNB0 = 64, kNB = 112;
constexpr int kWarpsP = 16;
constexpr int kThreadsP = kWarpsP * 32;
constexpr int kMaxOwn = 18;
constexpr int kUpdWarps = 12;             // cell-update warps (4..15)
constexpr float kScaleA = 1024.0f, kScaleW = 2048.0f;   // exact power-of-two operand scales (see fc_tc_layout.h)
constexpr float kHalfMax = 60000.0f;

FC_HD int units_of(int th) { return th < 2 ? 16 : 18; }
FC_HD int first_unit(int th) { return 16 * th; }

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
constexpr int kPackFloatsP = kSmallOff + kSmallFloats;

FC_HD int gate_row(int c) { return (c & 3) * kHid + (c >> 2); }   // PyTorch gate row of column unit*4+gate

// UNSCALED weight behind half-element h of the forward image of layer l: h = (k/8)*(208*8) + n*8 + k%8
FC_HD float fwd_weight(const RawWeights& w, int l, int h) {
  int kc = h / (kNF * 8), rem = h - kc * (kNF * 8);
  int n = rem / 8, k = kc * 8 + (rem & 7);
  if (n >= kGates) return 0.f;
  int row = gate_row(n);
  if (l == 0) {
    if (k < kFeat) return w.w_ih[0][row * kFeat + k];
    if (k >= kRec0 && k < kRec0 + kHid) return w.w_hh[0][row * kHid + (k - kRec0)];
    return 0.f;
  }
  if (k < kHid) return w.w_ih[l][row * kHid + k];
  if (k >= kRec && k < kRec + kHid) return w.w_hh[l][row * kHid + (k - kRec)];
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
    if (n < 54) {
      int th = n / 18, sl = n - th * 18;
      return sl < units_of(th) ? w.w_hh[0][row * kHid + first_unit(th) + sl] : 0.f;
    }
    if (n >= 56 && n < 56 + kFeat) return w.w_ih[0][row * kFeat + (n - 56)];
    return 0.f;
  }
  if (n >= 108) return 0.f;
  int th = n / 36, r2 = n - th * 36, sl = r2 % 18;
  if (sl >= units_of(th)) return 0.f;
  int u = first_unit(th) + sl;
  return r2 < 18 ? w.w_ih[l][row * kHid + u] : w.w_hh[l][row * kHid + u];
}
struct PrSlot { int kind; int l; int lo; int h; };   // kind 0 = forward, 1 = backward
FC_HD PrSlot decode_half(long hidx) {                // hidx counts halves from the start of the pair pack buffer
  PrSlot s;
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
constexpr int kColFcp = 416;                                  // + 12*tile + 4*third: read-out partial sums
constexpr int kColPark = 440;                                 // + 18*third: cell state of the tile that is not being updated
FC_HD int col_d_bwd(int tile) { return kNB * tile; }          // 0, 112
constexpr int kColGhi = 224, kColGlo = 328;                   // dG operand of tile 0 (104 columns each)

// per-TILE global workspace (floats); every slot is private to one thread
//   rows [(N+10)][5][128], cost [3][128], seq [10][12][6][32] float4 (fp16 hi/lo pieces in operand format),
//   dseq [10][12][18][32], grow [N][5][128],
//   rec [nrec][12][23][32] float4   (<= 18 units x (i,f,g,o,c_prev) = 90 floats -> 23 float4)
constexpr int kSlot = kUpdWarps * 6 * 32 * 4;                // 9216 floats per step (seq; dseq uses the first 6912)
constexpr int kRecF4 = 23;
constexpr int kRecFloatsP = kUpdWarps * kRecF4 * 32 * 4;     // 35328
struct WorkLayoutP {
  size_t rows, cost, seq, dseq, grow, rec, total;
};
FC_HD WorkLayoutP work_layout_p(int N, int with_grad) {
  WorkLayoutP w;
  w.rows = 0;
  w.cost = w.rows + (size_t)(N + kLook) * kFeat * kTileP;
  w.seq = w.cost + 3 * kTileP;
  w.dseq = w.seq + (size_t)kLook * kSlot;
  w.grow = w.dseq + (with_grad ? (size_t)kLook * kSlot : 0);
  w.rec = w.grow + (with_grad ? (size_t)N * kFeat * kTileP : 0);
  w.rec = (w.rec + 31) / 32 * 32;
  w.total = w.rec + (with_grad ? (size_t)rec_base(N) * kRecFloatsP : 0);
  w.total = (w.total + 31) / 32 * 32;
  return w;
}

// ---- surrogate training (MpcParams::train): per-TILE scratch handed to the weight-gradient kernel (fc_lstm_train_tc.cuh),
// everything in operand format = [piece of 8 k-slots][128 samples][8 halves], hi image then lo image:
//   hseq [3 layers][11 slots][2][4 stages][7 pieces][32][16 B]   slot 0 = zeros (h before the first step), slot t + 1 = h_t
//   feat [10 steps][2][128][16 B]                                the 5 window features of step t (k-slots 0..7 of layer 0)
//   dG   [3 layers][10 steps][2][4 stages][26 pieces][32][16 B]  gate gradients x g_scale, k = unit * 4 + gate
// (stage = 32 samples = one TMEM quadrant: what one pipeline stage of the weight-gradient kernel loads is contiguous)
constexpr int kTrHseqSlot = 2 * 7 * kTileP * 4;              // floats per (layer, slot)
constexpr int kTrFeatSlot = 2 * kTileP * 4;
constexpr int kTrDgSlot = 2 * 26 * kTileP * 4;
constexpr size_t kTrHseqOff = 0;
constexpr size_t kTrFeatOff = kTrHseqOff + (size_t)kLayers * (kLook + 1) * kTrHseqSlot;
constexpr size_t kTrDgOff = kTrFeatOff + (size_t)kLook * kTrFeatSlot;
constexpr size_t kTrTileFloats = kTrDgOff + (size_t)kLayers * kLook * kTrDgSlot;
// per-tile workspace of the kernel itself in training mode: the roll-out layout with all 30 cell records of ONE window
FC_HD size_t work_total_train() {
  WorkLayoutP w = work_layout_p(1, 1);
  return (w.rec + (size_t)kLayers * kLook * kRecFloatsP + 31) / 32 * 32;
}

// shared memory (floats)
constexpr int kSmSmallP = 0;                                 // fc + fnn weights (456)
constexpr int kSmRefP = kSmSmallP + kSmallFloats;            // [2][128]
constexpr int kSmGxP = kSmRefP + kTiles * kTileP;            // [2][4][128]
constexpr int kSmDvP = kSmGxP + kTiles * 4 * kTileP;         // [2][128]
constexpr int kSmFinP = kSmDvP + kTiles * kTileP;            // [2][2][128]
constexpr int kSmPgP = kSmFinP + kTiles * 2 * kTileP;        // double [4][250]
static_assert(kSmPgP % 2 == 0, "double alignment");
constexpr int kSmRedP = kSmPgP + 8 * kNumFnnGrad;            // double
constexpr int kSmBarP = ((kSmRedP + 2 + 3) / 4) * 4;         // 8 mbarriers (64-bit) + tmem base
constexpr int kSmWP = ((kSmBarP + 24 + 255) / 256) * 256;    // weight image, 1 KiB aligned
constexpr int kSmWFloatsP = kNB * kKB;                       // hi + lo fp16 images = 23296 floats
constexpr int kSmOpP = kSmWP + kSmWFloatsP;                  // operand region
constexpr int kOpTileFloats = kTileP * kKF;                  // A operand of one tile: hi + lo = 2*128*112 halves
constexpr int kSmOpFloats = kTiles * kOpTileFloats;          // 28672 floats = 114 688 B
constexpr int kSmFloatsP = kSmOpP + kSmOpFloats;
constexpr size_t kSmBytesP = (size_t)kSmFloatsP * sizeof(float);
static_assert(kSmBytesP <= 227 * 1024, "shared memory budget exceeded (pair variant)");
static_assert(kTileP * kKB <= kSmOpFloats, "dG operand of tile 1 does not fit the operand region");
// halves offsets inside the operand region
constexpr int kOpLoHalves = kTileP * kKF;                    // forward: lo image of a tile follows its hi image
FC_HD int op_fwd_halves(int tile) { return tile * 2 * kTileP * kKF; }
constexpr int kOpGLoHalves = kTileP * kKB;                   // backward: lo image of dG (tile 1) follows the hi image
// replica mode (R = 4, one tile per CTA): read-out partial sums [12 update warps][4 outputs][32 trajectories] in the tail of the
// operand region that neither the forward images of tile 0 nor the dG images reach
constexpr int kSmExP = kSmOpP + kTileP * kKB;
static_assert(kTileP * kKB + kUpdWarps * 4 * 32 <= kSmOpFloats, "exchange buffer of the replica mode does not fit");
static_assert(kOpTileFloats <= kTileP * kKB, "exchange buffer overlaps the forward images of tile 0");

FC_HD float acc_correction(int steps, float scale) { return scale * (0.17f + 0.135f * (float)steps) * 1.1920929e-7f; }

// mbarrier ids
constexpr int kBarFull = 0;        // + tile: accumulator of the tile complete (tcgen05.commit)
constexpr int kBarReady = 2;       // + tile: operand of the tile written (one arrival per warp)
constexpr int kBarWeightsP = 4;    // bulk copy of a weight image landed

}  // namespace pr
}  // namespace fc
