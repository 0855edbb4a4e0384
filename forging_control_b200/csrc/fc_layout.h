// Layout constants shared by the sm_100a kernels, the weight packer and the CPU thread-emulation
// harness used by the CPU test-suite (tests/emu).  No CUDA-only constructs in this file.
//
// Hot path: MPCLoss.forward + loss.backward() of the reference
// (Unsupervised Learning/Functions.py:1353-1472, :655) for LSTMModel(5,50,4,3) (:295-379) and
// FNNModel(3,50,1,width_dim=1) (:215-289).  See DESIGN.md for the derivation of every layout.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FC_HD __host__ __device__ __forceinline__
#else
#define FC_HD inline
#endif

namespace fc {

// ---- problem constants (hard-coded in the reference: look-back 10 at Functions.py:1434) --------
constexpr int kLook = 10;     // window rows
constexpr int kFeat = 5;      // row features: y_dot, p1, p2, z, u
constexpr int kHid = 50;      // LSTM hidden size
constexpr int kGates = 200;   // 4 * kHid, PyTorch order i|f|g|o
constexpr int kLayers = 3;
constexpr int kOut = 4;       // surrogate read-out: y_dot, p1, p2, z
constexpr int kFnnHid = 50;
constexpr float kP1Max = 2.122366f;   // Functions.py:1411
constexpr float kP2Max = 1.036233f;

// ---- CTA tiling ---------------------------------------------------------------------------------
// 256 threads = 8 warps.  In every warp lanes 0..29 are compute lanes: lane = tg*5 + cgl with
// tg in [0,6) (trajectory block of 5) and cgl in [0,5) (hidden-unit group of 5 = 20 gate columns);
// warp w covers trajectory blocks (w/2)*6.. and unit groups (w%2)*5..  A CTA tile is 24 trajectory
// blocks * 5 = 120 trajectories x 10 unit groups.
constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kTile = 120;        // trajectories per CTA tile
constexpr int kTB = 24;           // trajectory blocks of 5 per tile
constexpr int kElems = 25;        // (trajectory, unit) elements owned by a thread: 5 x 5

// "plane" layout of a 120-vector over trajectories: trajectories j=0..3 of block tb live at
// tb*4+j (one 16-byte chunk per block), trajectory j=4 lives at 96+tb.
FC_HD int plane_of(int traj) {
  int tb = traj / 5, j = traj - tb * 5;
  return j < 4 ? tb * 4 + j : 96 + tb;
}

// ---- packed weight buffer (floats) ----------------------------------------------------------------
// forward:  WF_l[k][cg][q][uu]      k = input feature (ih rows first, then hh), col = cg*20+q*5+uu
// backward: WB_l[g'][cg][c]         g' = cg_g*20+q*5+uu (same permuted gate order),
//           layers 1,2: c<5 -> d(input unit cg*5+c), 5<=c<10 -> d(h_prev unit cg*5+c-5), 10,11 pad
//           layer 0:    c<5 -> d(h_prev unit cg*5+c), c==5 -> d(row feature cg) for cg<5, 6..11 pad
constexpr int kKin0 = kFeat, kKin = kHid;
constexpr int kWF0 = 0;
constexpr int kWF1 = kWF0 + (kKin0 + kHid) * kGates;   // 11000
constexpr int kWF2 = kWF1 + (kKin + kHid) * kGates;    // 31000
constexpr int kWB0 = kWF2 + (kKin + kHid) * kGates;    // 51000
constexpr int kWB0Stride = 120;                         // 10 cg * 12 (same layout as layers 1,2)
constexpr int kWBStride = 120;                          // 10 cg * 12
constexpr int kWB1 = kWB0 + kGates * kWB0Stride;        // 75000
constexpr int kWB2 = kWB1 + kGates * kWBStride;         // 99000
constexpr int kFCW = kWB2 + kGates * kWBStride;         // 123000  fc.weight [4][50]
constexpr int kFCB = kFCW + kOut * kHid;                // 123200  fc.bias [4]
constexpr int kINPW = kFCB + kOut;                      // 123204  fc_inp.weight [50][3]
constexpr int kINPB = kINPW + kFnnHid * 3;              // 123354  fc_inp.bias [50]
constexpr int kOUTW = kINPB + kFnnHid;                  // 123404  fc_out.weight [50]
constexpr int kPackFloats = ((kOUTW + kFnnHid + 3) / 4) * 4;   // 123456
constexpr int kSmallFloats = kPackFloats - kFCW;        // 456: fc + fnn block copied to smem

FC_HD int wf_offset(int l) { return l == 0 ? kWF0 : (l == 1 ? kWF1 : kWF2); }
FC_HD int wb_offset(int l) { return l == 0 ? kWB0 : (l == 1 ? kWB1 : kWB2); }

// raw (PyTorch state_dict) tensors
struct RawWeights {
  const float* w_ih[kLayers];   // [200][5] , [200][50], [200][50]
  const float* w_hh[kLayers];   // [200][50]
  const float* fc_w;            // [4][50]
  const float* fc_b;            // [4]
  const float* inp_w;           // [50][3]
  const float* inp_b;           // [50]
  const float* out_w;           // [1][50]
};

// value of packed element `idx` (used by the device pack kernel and by the emulation harness)
FC_HD float packed_value(const RawWeights& w, int idx) {
  if (idx < kWB0) {                                     // forward blocks
    int l = idx < kWF1 ? 0 : (idx < kWF2 ? 1 : 2);
    int r = idx - wf_offset(l);
    int k = r / kGates, col = r - k * kGates;
    int cg = col / 20, q = (col % 20) / 5, uu = col % 5;
    int row = q * kHid + cg * 5 + uu;
    int kin = l == 0 ? kKin0 : kKin;
    return k < kin ? w.w_ih[l][row * kin + k] : w.w_hh[l][row * kHid + (k - kin)];
  }
  if (idx < kFCW) {                                     // backward blocks
    int l = idx < kWB1 ? 0 : (idx < kWB2 ? 1 : 2);
    int r = idx - wb_offset(l);
    int stride = kWBStride, cw = 12;
    int g = r / stride, rem = r - g * stride;
    int cgo = rem / cw, c = rem - cgo * cw;
    int cgg = g / 20, q = (g % 20) / 5, uu = g % 5;
    int row = q * kHid + cgg * 5 + uu;
    if (l == 0) {
      if (c < 5) return w.w_hh[0][row * kHid + cgo * 5 + c];
      if (c == 5 && cgo < kFeat) return w.w_ih[0][row * kFeat + cgo];
      return 0.f;
    }
    if (c < 5) return w.w_ih[l][row * kHid + cgo * 5 + c];
    if (c < 10) return w.w_hh[l][row * kHid + cgo * 5 + (c - 5)];
    return 0.f;
  }
  if (idx < kFCB) return w.fc_w[idx - kFCW];
  if (idx < kINPW) return w.fc_b[idx - kFCB];
  if (idx < kINPB) return w.inp_w[idx - kINPW];
  if (idx < kOUTW) return w.inp_b[idx - kINPB];
  if (idx < kOUTW + kFnnHid) return w.out_w[idx - kOUTW];
  return 0.f;
}

// ---- per-CTA global workspace (floats) ----------------------------------------------------------
// rows  [(N+10)][5][120]  roll-out rows rho_0..rho_{N+9} (plane layout); rho_{9+k} = [x_k, u_k]
// seq   [10][50][120]     forward: hidden sequence of the layer below (plane layout)
// dseq  [10][8][25][32]   backward: d(hidden sequence), thread-private slots
// grow  [N][5][120]       gradient of rows rho_9..rho_{9+N-1}
// rec   [nrec][8][32][32][4]  saved cell activations (i,f,g,o,c_prev), thread-private float4 slots
constexpr int kRecFloats = kWarps * 32 * 32 * 4;       // 32768 floats = 128 KiB per (window,layer,step)
FC_HD int steps_kept(int m) { return m + 1 < kLook ? m + 1 : kLook; }          // 10 - t_min(m)
FC_HD int t_min_of(int m) { return kLook - steps_kept(m); }
FC_HD long rec_base(int m) {       // number of records before window m
  long c = m <= kLook ? (long)m * (m + 1) / 2 : 55 + (long)kLook * (m - kLook);
  return kLayers * c;
}
struct WorkLayout {
  size_t rows, seq, dseq, grow, rec, total;   // offsets in floats
};
FC_HD WorkLayout work_layout(int N, int with_grad) {
  WorkLayout w;
  w.rows = 0;
  w.seq = w.rows + (size_t)(N + kLook) * kFeat * kTile;
  w.dseq = w.seq + (size_t)kLook * kHid * kTile;
  w.grow = w.dseq + (with_grad ? (size_t)kLook * kWarps * kElems * 32 : 0);
  w.rec = w.grow + (with_grad ? (size_t)N * kFeat * kTile : 0);
  w.rec = (w.rec + 31) / 32 * 32;                      // 128-byte alignment of the record area
  w.total = w.rec + (with_grad ? (size_t)rec_base(N) * kRecFloats : 0);
  w.total = (w.total + 31) / 32 * 32;
  return w;
}

// per-CTA partial results (double precision: the batch reductions are cancellation-heavy): controller
// gradients then loss sum
constexpr int kNumFnnGrad = 250;            // inp_w[150] | inp_b[50] | out_w[50]
constexpr int kWideGrads = kFnnHid * kFnnHid + kFnnHid;   // 2550: fc_int.weight | fc_int.bias
constexpr int kWidePartialStride = 2560;
constexpr int kPartialStride = 256;         // [0..249] grads, [250] sum of per-trajectory cost

// ---- shared memory carve-up (floats) ---------------------------------------------------------------
// small block (always resident)
constexpr int kSmSmall = 0;                               // fc + fnn weights (kSmallFloats = 456)
constexpr int kSmRef = kSmSmall + kSmallFloats;           // [120] reference (scaled)
constexpr int kSmUcur = kSmRef + kTile;                   // [120]
constexpr int kSmUprev = kSmUcur + kTile;                 // [120]
constexpr int kSmCost = kSmUprev + kTile;                 // [3][120] cost, command, error accumulators
constexpr int kSmGx = kSmCost + 3 * kTile;                // [4][120] d loss / d x_{m+1} (plane layout)
constexpr int kSmDv = kSmGx + 4 * kTile;                  // [120] d loss / d v (pre-saturation command)
constexpr int kSmFin = kSmDv + kTile;                     // [2][120] controller inputs x[0], x[3]
constexpr int kSmPg = kSmFin + 2 * kTile;                 // double [4][250] controller gradient partials (8-byte aligned)
constexpr int kSmRed = kSmPg + 8 * kNumFnnGrad;           // double [4] scratch (loss sum)
static_assert(kSmPg % 2 == 0, "double alignment");
constexpr int kSmBig = ((kSmRed + 8 + 3) / 4) * 4;        // start of the phase-dependent area
// forward: W [100][200] | ain [2][50][120] | ah [2][50][120]
constexpr int kSmFwdW = kSmBig;
constexpr int kSmFwdAin = kSmFwdW + 100 * kGates;
constexpr int kSmFwdAh = kSmFwdAin + 2 * kHid * kTile;
constexpr int kSmFwdEnd = kSmFwdAh + 2 * kHid * kTile;
// backward: W [200][120] | dG [200][120]
constexpr int kSmBwdW = kSmBig;
constexpr int kSmBwdG = kSmBwdW + kGates * kWBStride;
constexpr int kSmBwdEnd = kSmBwdG + kGates * kTile;
constexpr int kSmFloats = kSmFwdEnd > kSmBwdEnd ? kSmFwdEnd : kSmBwdEnd;
constexpr size_t kSmBytes = (size_t)kSmFloats * sizeof(float);
static_assert(kSmBytes <= 227 * 1024, "shared memory budget of one sm_100a CTA exceeded");

struct MpcParams {
  const float* X;        // [B][3]  controller input (scaled); reference = column 2
  const float* u0;       // [B]     controller output for X (output_controller.squeeze())
  const float* Z;        // [B][10][5] look-back window (scaled): y_dot,p1,p2,z,u
  const float* wpack;    // packed weights, kPackFloats
  float* cost;           // [B]
  float* command;        // [B]
  float* error;          // [B]
  float* pred;           // [B][N]
  float* du0;            // [B] d loss / d u0 (with_grad)
  double* partial;       // [grid][kPartialStride]
  float* work;           // [grid][work_stride]
  size_t work_stride;    // floats
  int B;
  int N;
  int with_grad;
  int num_tiles;
  float alpha;
  float grad_scale;      // 1 / (N * B_global)
  float acc_comp;        // scale of the tensor-core accumulator compensation (1 = calibrated value)
  int debug_timing;      // CTA 0 prints a cycle breakdown (development aid)
  float g_scale;         // tcgen05 kernel: power-of-two scale of the gate gradients before the fp16 split
  float g_unscale;       // 1 / g_scale
  // width_dim > 1 controllers (FNNModel.forward, UL/Functions.py:261-289: width_dim-1 repeats of the weight-shared
  // fc_int + ReLU); one-tile tcgen05 kernel only
  int width_dim;            // 1 = no hidden repeat
  const float* int_w;       // fc_int.weight [50][50]
  const float* int_b;       // fc_int.bias [50]
  double* partial_wide;     // [grid][kWidePartialStride]: d loss / d fc_int.weight [2500] | fc_int.bias [50]
  // enable_noise (UL/Functions.py:1400-1402, :1438-1440): x += noise_std * N(0,1) after every surrogate call;
  // counter-based generator (Philox4x32-10, key = seed, counter = (trajectory, window)), see philox_normal4
  float noise_std;
  unsigned long long noise_seed;
  // LSTM shadow roll-out (pair kernel, forward only; Functions.py:969-1011, 1196-1231): N windows, the command of
  // every step is an input, no cost / controller / reverse sweep
  int shadow;
  const float* sh_row0;  // [B][5]  first window row (scaled), repeated 10 times
  const float* sh_u;     // [B][N]  scaled commands; sh_u[b][m+1] closes the row appended after window m
  float* sh_y;           // [B][N][4] surrogate outputs (scaled)
  float sh_ratio[4];     // scale_out / scale_in: output -> next input row
  // surrogate training on the pair kernel (Model_NN/Functions.py:313-340, :520-569; fc_lstm_train_tc.cuh): one window per
  // sample, N = 1, every cell step recorded.  train = 1: forward only (y); train = 2: forward with records + reverse sweep
  // seeded by d loss / d y, writing the gate gradients and the layer inputs in operand format for the weight-gradient kernel
  // pair kernel with ONE tile per CTA (mid-size batches, #tiles <= #SMs: every tile gets an SM of its own and the dedicated
  // issuer warp of the pair kernel is worth ~6 % over the one-tile kernel)
  int single_tile;
  int train;
  const float* tr_x;     // [B][10][5] windows
  float* tr_y;           // [B][4] outputs (train = 1)
  float* tr_hlast;       // [B][50] top-layer hidden state of the last step (train = 1, may be null)
  const float* tr_dy;    // [B][4] upstream gradient (train = 2)
  const float* tr_fcw;   // fc.weight [4][50] (raw state_dict tensor)
  const float* tr_fcb;   // fc.bias [4]
  const float* tr_scale; // device: {scale, 1 / scale} of the gate gradients (power of two)
  float* tr_ws;          // per-tile scratch for the weight-gradient kernel, see lt2::TileScratch
  int tr_tile_base;      // index of this launch's first tile in tr_ws
};

// Four standard normals for (trajectory b, window m): Philox4x32-10 + Box-Muller.  Counter-based, so every kernel
// (and the oracle's restatement, oracle/mpc_loss_oracle.py::philox_normal4) draws the same noise for the same seed.
FC_HD void philox_normal4(unsigned long long seed, unsigned b, unsigned m, float* out) {
  unsigned c0 = b, c1 = m, c2 = 0u, c3 = 0u;
  unsigned k0 = (unsigned)(seed & 0xffffffffull), k1 = (unsigned)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const float s = 2.3283064365386963e-10f;                      // 2^-32
  const float u0 = ((float)c0 + 0.5f) * s, u1 = ((float)c1 + 0.5f) * s, u2 = ((float)c2 + 0.5f) * s, u3 = ((float)c3 + 0.5f) * s;
  const float r0 = sqrtf(-2.0f * logf(u0 < 1e-30f ? 1e-30f : (u0 > 0.99999994f ? 0.99999994f : u0)));
  const float r1 = sqrtf(-2.0f * logf(u2 < 1e-30f ? 1e-30f : (u2 > 0.99999994f ? 0.99999994f : u2)));
  const float t0 = 6.2831853071795865f * u1, t1 = 6.2831853071795865f * u3;
  out[0] = r0 * cosf(t0); out[1] = r0 * sinf(t0); out[2] = r1 * cosf(t1); out[3] = r1 * sinf(t1);
}

}  // namespace fc
