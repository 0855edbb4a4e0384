// TEST HARNESS ONLY -- never linked into the product library.
//
// CPU thread-emulation of one CUDA thread block for the fused MPC-loss kernel body
// (forging_control_b200/csrc/fc_mpc_kernel.inl).  Every CUDA thread becomes an OS thread, the block
// barrier becomes a std::barrier, cp.async becomes an immediate copy and the MUFU approximations
// become exp2f / 1/x.  It exists so that the CPU test-suite (no GPU in the build container) can
// check the index arithmetic, the tiling and the reverse-sweep dataflow of the kernel source
// against the oracle.  It is NOT a fallback: the product library has no CPU path.
#include <barrier>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define FC_HD_CTX inline
#include "../../forging_control_b200/csrc/fc_mpc_kernel.inl"
#include "../../forging_control_b200/csrc/fc_mpc_tc_kernel.inl"
#include "../../forging_control_b200/csrc/fc_mpc_pair_kernel.inl"
#include <atomic>
#include <cstdint>

namespace {

struct EmuBlock {
  std::barrier<> bar;
  std::vector<float> smem;
  int bid, nblk;
  EmuBlock(int b, int n, int nthreads = fc::kThreads) : bar(nthreads), smem(fc::kSmFloats, 0.f), bid(b), nblk(n) {}
};

struct EmuCtx {
  EmuBlock* blk;
  int t;
  int tid() const { return t; }
  int bid() const { return blk->bid; }
  int nblk() const { return blk->nblk; }
  float* smem() const { return blk->smem.data(); }
  void sync() const { blk->bar.arrive_and_wait(); }
  static float ex2(float x) { return exp2f(x); }
  static float rcp(float x) { return 1.0f / x; }
  static fc::F4 lds4(const float* p) { fc::F4 v; std::memcpy(&v, p, 16); return v; }
  static void sts4(float* p, fc::F4 v) { std::memcpy(p, &v, 16); }
  static fc::F4 ldg4(const float* p) { return lds4(p); }
  static fc::F4 ldg4_stream(const float* p) { return lds4(p); }
  static void stg4(float* p, fc::F4 v) { sts4(p, v); }
  static void stg4_stream(float* p, fc::F4 v) { sts4(p, v); }
  static void stg2_stream(float* p, float a, float b) { p[0] = a; p[1] = b; }
  static void ldg2_stream(const float* p, float& a, float& b) { a = p[0]; b = p[1]; }
  static float ldcg(const float* p) { return *p; }
  static void cp_async16(float* d, const float* s) { std::memcpy(d, s, 16); }
  static void cp_async4(float* d, const float* s) { std::memcpy(d, s, 4); }
  static void cp_async8(float* d, const float* s) { std::memcpy(d, s, 8); }
  static void sts1(float* p, float v) { *p = v; }
  static void stg2(float* p, float a, float b) { p[0] = a; p[1] = b; }
  static void cp_commit() {}
  template <int N> static void cp_wait() {}
};


// ---- tcgen05 variant: TMEM = [128 lanes][512 columns] array, MMA executed synchronously by the issuing
// thread with tf32-truncated operands, mbarriers = counters --------------------------------------------
struct EmuBlockTC : EmuBlock {
  std::vector<float> tmem;
  std::atomic<unsigned> bars[8];      // arrivals
  std::atomic<int> tx[8];             // pending bytes of split bulk copies (bulk_expect / bulk_copy)
  unsigned counts[8];                 // arrivals per phase
  std::vector<std::unique_ptr<std::barrier<>>> wbar;   // one barrier per warp (__syncwarp)
  EmuBlockTC(int b, int n, int smem_floats = fc::tc::kSmFloatsTC, int nthreads = fc::tc::kThreadsTC)
      : EmuBlock(b, n, nthreads), tmem(128 * 512, 0.f) {
    smem.assign(smem_floats, 0.f);
    for (auto& x : bars) x.store(0);
    for (auto& x : tx) x.store(0);
    for (auto& x : counts) x = 1;
    for (int w = 0; w < nthreads / 32; ++w) wbar.emplace_back(new std::barrier<>(32));
  }
};


struct EmuCtxTC : EmuCtx {
  EmuBlockTC* tb;
  EmuCtxTC(EmuBlockTC* b, int t_) : EmuCtx{b, t_}, tb(b) {}
  int row() const { return 32 * ((t >> 5) & 3) + (t & 31); }
  void tc_setup(float*) { sync(); }
  void tc_teardown() { sync(); }
  void tc_sync() const { sync(); }
  template <int N> void tmem_ld(int col, float* v) const { for (int i = 0; i < N; ++i) v[i] = tb->tmem[row() * 512 + col + i]; }
  template <int N> void tmem_ld_nowait(int col, float* v) const { tmem_ld<N>(col, v); }
  void tmem_ld_wait() const {}
  template <int N> void tmem_st(int col, const float* v) const { for (int i = 0; i < N; ++i) tb->tmem[row() * 512 + col + i] = v[i]; }
  void tmem_st_wait() const {}
  static uint16_t h_bits(float x) { x = x > 65504.f ? 65504.f : (x < -65504.f ? -65504.f : x); _Float16 h = (_Float16)x; uint16_t u; std::memcpy(&u, &h, 2); return u; }
  static float h_val(uint16_t u) { _Float16 h; std::memcpy(&h, &u, 2); return (float)h; }
  static void split_h2(float x0, float x1, float& hi, float& lo) {
    uint16_t h0 = h_bits(x0), h1 = h_bits(x1);
    uint16_t l0 = h_bits(x0 - h_val(h0)), l1 = h_bits(x1 - h_val(h1));
    uint32_t ph = (uint32_t)h0 | ((uint32_t)h1 << 16), pl = (uint32_t)l0 | ((uint32_t)l1 << 16);
    std::memcpy(&hi, &ph, 4); std::memcpy(&lo, &pl, 4);
  }
  // the emulation accumulates exactly (round to nearest): it checks index arithmetic and dataflow, not the
  // truncating accumulator of the hardware, so the kernel's accumulator compensation is switched off
  static constexpr bool kAccTruncates = false;
  static long long clock() { return 0; }
  static void prefetch_l2(const float*) {}
  static void prefetch_l2_bulk(const float*, unsigned) {}
  static void report(const long long*) {}
  void mma(int d_col, int n, int a_col, const float* b_img, int n_img, int row0, int ksteps, bool accumulate) const {
    const int K = ksteps * 16;
    const uint16_t* bh = reinterpret_cast<const uint16_t*>(b_img);
    for (int m = 0; m < 128; ++m) {
      const uint32_t* a = reinterpret_cast<const uint32_t*>(&tb->tmem[m * 512 + a_col]);
      for (int j = 0; j < n; ++j) {
        double s = accumulate ? (double)tb->tmem[m * 512 + d_col + j] : 0.0;
        for (int k = 0; k < K; ++k) {
          const uint16_t ab = (uint16_t)((a[k >> 1] >> ((k & 1) * 16)) & 0xffffu);
          s += (double)h_val(ab) * (double)h_val(bh[(k / 8) * (n_img * 8) + (row0 + j) * 8 + (k & 7)]);
        }
        tb->tmem[m * 512 + d_col + j] = (float)s;
      }
    }
  }
  void commit(int bar) const { tb->bars[bar].fetch_add(1); }
  void bar_wait(int bar, unsigned phase) const {
    // mbarrier.try_wait.parity semantics: succeeds once the phase with the given parity has completed, i.e. as soon as
    // the parity of the number of completed phases differs from it (a waiter is never more than one phase behind)
    while (((tb->bars[bar].load() / tb->counts[bar]) & 1u) == (phase & 1u)) std::this_thread::yield();
  }
  void bar_wait_relaxed(int bar, unsigned phase) const { bar_wait(bar, phase); }
  // pair kernel
  void bar_init(int bar, int count) const { tb->counts[bar] = (unsigned)count; tb->bars[bar].store(0); }
  void bar_init_fence() const {}
  void warp_sync() const { tb->wbar[t >> 5]->arrive_and_wait(); }
  static void report_pair(int, const long long*) {}
  void bar_arrive(int bar) const { tb->bars[bar].fetch_add(1); }
  void operand_fence() const {}
  void tmem_fence() const {}
  static void sts2(float* p, float a, float b) { p[0] = a; p[1] = b; }
  static void red_add(float* p, float v) { *p += v; }
  void mma_ss(int d_col, int n, const float* a_img, const float* b_img, int n_img, int ksteps, bool accumulate) const {
    const int K = ksteps * 16;
    const uint16_t* ah = reinterpret_cast<const uint16_t*>(a_img);
    const uint16_t* bh = reinterpret_cast<const uint16_t*>(b_img);
    for (int m = 0; m < 128; ++m)
      for (int j = 0; j < n; ++j) {
        double s = accumulate ? (double)tb->tmem[m * 512 + d_col + j] : 0.0;
        for (int k = 0; k < K; ++k)
          s += (double)h_val(ah[(k / 8) * (128 * 8) + m * 8 + (k & 7)]) * (double)h_val(bh[(k / 8) * (n_img * 8) + j * 8 + (k & 7)]);
        tb->tmem[m * 512 + d_col + j] = (float)s;
      }
  }
  // transaction count of the mbarrier: the arrival is published when the announced bytes have landed
  void bulk_expect(int bar, int bytes) const { tb->tx[bar].fetch_add(bytes); }
  void bulk_copy(float* dst, const float* src, int bytes, int bar) const {
    std::memcpy(dst, src, (size_t)bytes);
    if (tb->tx[bar].fetch_sub(bytes) == bytes) tb->bars[bar].fetch_add(1);
  }
  void bulk_load(float* dst, const float* src, int nfloats, int bar) const {
    std::memcpy(dst, src, (size_t)nfloats * 4);
    tb->bars[bar].fetch_add(1);
  }
};

}  // namespace

static float g_noise_std = 0.f;
static unsigned long long g_noise_seed = 0ull;

extern "C" {

// enable_noise for the following emulated launches (0 = off)
void fc_emu_set_noise(float std, unsigned long long seed) { g_noise_std = std; g_noise_seed = seed; }

int fc_emu_pack_floats() { return fc::kPackFloats; }

void fc_emu_pack_weights(const float* w_ih0, const float* w_hh0, const float* w_ih1, const float* w_hh1,
                         const float* w_ih2, const float* w_hh2, const float* fc_w, const float* fc_b,
                         const float* inp_w, const float* inp_b, const float* out_w, float* out) {
  fc::RawWeights w;
  w.w_ih[0] = w_ih0; w.w_hh[0] = w_hh0; w.w_ih[1] = w_ih1; w.w_hh[1] = w_hh1; w.w_ih[2] = w_ih2; w.w_hh[2] = w_hh2;
  w.fc_w = fc_w; w.fc_b = fc_b; w.inp_w = inp_w; w.inp_b = inp_b; w.out_w = out_w;
  for (int i = 0; i < fc::kPackFloats; ++i) out[i] = fc::packed_value(w, i);
}

// runs `grid` emulated CTAs one after the other; returns 0
int fc_emu_mpc_loss(const float* X, const float* u0, const float* Z, const float* wpack, int B, int N, float alpha,
                    long long B_global, int with_grad, int grid, float* cost, float* command, float* error,
                    float* pred, float* du0, float* gl /*[256]*/) {
  fc::MpcParams p;
  std::memset(&p, 0, sizeof(p));
  p.X = X; p.u0 = u0; p.Z = Z; p.wpack = wpack;
  p.cost = cost; p.command = command; p.error = error; p.pred = pred; p.du0 = du0;
  p.B = B; p.N = N; p.with_grad = with_grad; p.alpha = alpha;
  p.grad_scale = 1.0f / ((float)N * (float)B_global);
  p.noise_std = g_noise_std; p.noise_seed = g_noise_seed;
  p.acc_comp = 1.0f;
  { int e = (int)std::floor(std::log2((double)N * (double)B_global)); p.g_scale = (float)std::ldexp(1.0, e); p.g_unscale = (float)std::ldexp(1.0, -e); }
  p.num_tiles = (B + fc::kTile - 1) / fc::kTile;
  if (grid > p.num_tiles) grid = p.num_tiles;
  fc::WorkLayout wl = fc::work_layout(N, with_grad);
  p.work_stride = wl.total;
  std::vector<float> work((size_t)grid * wl.total, 0.f);
  std::vector<double> partial((size_t)grid * fc::kPartialStride, 0.0);
  p.work = work.data();
  p.partial = partial.data();
  for (int b = 0; b < grid; ++b) {
    EmuBlock blk(b, grid);
    std::vector<std::thread> th;
    th.reserve(fc::kThreads);
    for (int t = 0; t < fc::kThreads; ++t)
      th.emplace_back([&blk, &p, t]() {
        EmuCtx ctx{&blk, t};
        fc::MpcTile<EmuCtx> k(ctx, p);
        k.run();
      });
    for (auto& x : th) x.join();
  }
  for (int i = 0; i < 256; ++i) gl[i] = 0.f;
  for (int i = 0; i <= fc::kNumFnnGrad; ++i) {
    double a = 0.0;
    for (int b = 0; b < grid; ++b) a += partial[(size_t)b * fc::kPartialStride + i];
    gl[i] = (float)(i == fc::kNumFnnGrad ? a / (double)B_global : a);
  }
  return 0;
}


int fc_emu_pack_floats_tc() { return fc::tc::kPackFloatsTC; }

void fc_emu_pack_weights_tc(const float* w_ih0, const float* w_hh0, const float* w_ih1, const float* w_hh1,
                            const float* w_ih2, const float* w_hh2, const float* fc_w, const float* fc_b,
                            const float* inp_w, const float* inp_b, const float* out_w, float* out) {
  fc::RawWeights w;
  w.w_ih[0] = w_ih0; w.w_hh[0] = w_hh0; w.w_ih[1] = w_ih1; w.w_hh[1] = w_hh1; w.w_ih[2] = w_ih2; w.w_hh[2] = w_hh2;
  w.fc_w = fc_w; w.fc_b = fc_b; w.inp_w = inp_w; w.inp_b = inp_b; w.out_w = out_w;
  uint16_t* oh = reinterpret_cast<uint16_t*>(out);
  const long n_halves = 2L * fc::tc::kSmallOff;
  for (long i = 0; i < n_halves; ++i) {
    const fc::tc::TcSlot s = fc::tc::tc_decode_half(i);
    const float v = (s.kind == 0 ? fc::tc::fwd_weight(w, s.l, s.h) : fc::tc::bwd_weight(w, s.l, s.h)) * fc::tc::kScaleW;
    const uint16_t hi = EmuCtxTC::h_bits(v);
    oh[i] = s.lo ? EmuCtxTC::h_bits(v - EmuCtxTC::h_val(hi)) : hi;
  }
  for (int j = 0; j < fc::kSmallFloats; ++j) out[fc::tc::kSmallOff + j] = fc::packed_value(w, fc::kFCW + j);
}

int fc_emu_mpc_loss_tc(const float* X, const float* u0, const float* Z, const float* wpack, int B, int N, float alpha,
                       long long B_global, int with_grad, int grid, float* cost, float* command, float* error,
                       float* pred, float* du0, float* gl /*[256]*/) {
  fc::MpcParams p;
  std::memset(&p, 0, sizeof(p));
  p.X = X; p.u0 = u0; p.Z = Z; p.wpack = wpack;
  p.cost = cost; p.command = command; p.error = error; p.pred = pred; p.du0 = du0;
  p.B = B; p.N = N; p.with_grad = with_grad; p.alpha = alpha;
  p.grad_scale = 1.0f / ((float)N * (float)B_global);
  p.noise_std = g_noise_std; p.noise_seed = g_noise_seed;
  p.acc_comp = 1.0f;
  { int e = (int)std::floor(std::log2((double)N * (double)B_global)); p.g_scale = (float)std::ldexp(1.0, e); p.g_unscale = (float)std::ldexp(1.0, -e); }
  p.num_tiles = (B + fc::tc::kTileTC - 1) / fc::tc::kTileTC;
  if (grid > p.num_tiles) grid = p.num_tiles;
  fc::tc::WorkLayoutTC wl = fc::tc::work_layout_tc(N, with_grad);
  p.work_stride = wl.total;
  std::vector<float> work((size_t)grid * wl.total, 0.f);
  std::vector<double> partial((size_t)grid * fc::kPartialStride, 0.0);
  p.work = work.data();
  p.partial = partial.data();
  for (int b = 0; b < grid; ++b) {
    EmuBlockTC blk(b, grid);
    std::vector<std::thread> th;
    th.reserve(fc::tc::kThreadsTC);
    for (int t = 0; t < fc::tc::kThreadsTC; ++t)
      th.emplace_back([&blk, &p, t]() {
        EmuCtxTC ctx(&blk, t);
        fc::tc::MpcTileTC<EmuCtxTC> k(ctx, p);
        k.run();
      });
    for (auto& x : th) x.join();
  }
  for (int i = 0; i < 256; ++i) gl[i] = 0.f;
  for (int i = 0; i <= fc::kNumFnnGrad; ++i) {
    double a = 0.0;
    for (int b = 0; b < grid; ++b) a += partial[(size_t)b * fc::kPartialStride + i];
    gl[i] = (float)(i == fc::kNumFnnGrad ? a / (double)B_global : a);
  }
  return 0;
}


int fc_emu_pack_floats_pair() { return fc::pr::kPackFloatsP; }

void fc_emu_pack_weights_pair(const float* w_ih0, const float* w_hh0, const float* w_ih1, const float* w_hh1,
                              const float* w_ih2, const float* w_hh2, const float* fc_w, const float* fc_b,
                              const float* inp_w, const float* inp_b, const float* out_w, float* out) {
  fc::RawWeights w;
  w.w_ih[0] = w_ih0; w.w_hh[0] = w_hh0; w.w_ih[1] = w_ih1; w.w_hh[1] = w_hh1; w.w_ih[2] = w_ih2; w.w_hh[2] = w_hh2;
  w.fc_w = fc_w; w.fc_b = fc_b; w.inp_w = inp_w; w.inp_b = inp_b; w.out_w = out_w;
  uint16_t* oh = reinterpret_cast<uint16_t*>(out);
  const long n_halves = 2L * fc::pr::kSmallOff;
  for (long i = 0; i < n_halves; ++i) {
    const fc::pr::PrSlot s = fc::pr::decode_half(i);
    const float v = (s.kind == 0 ? fc::pr::fwd_weight(w, s.l, s.h) : fc::pr::bwd_weight(w, s.l, s.h)) * fc::pr::kScaleW;
    const uint16_t hi = EmuCtxTC::h_bits(v);
    oh[i] = s.lo ? EmuCtxTC::h_bits(v - EmuCtxTC::h_val(hi)) : hi;
  }
  for (int j = 0; j < fc::kSmallFloats; ++j) out[fc::pr::kSmallOff + j] = fc::packed_value(w, fc::kFCW + j);
}

}  // extern "C"
template <int R>
static int emu_mpc_loss_pair(const float* X, const float* u0, const float* Z, const float* wpack, int B, int N, float alpha,
                         long long B_global, int with_grad, int grid, float* cost, float* command, float* error,
                         float* pred, float* du0, float* gl /*[256]*/) {
  fc::MpcParams p;
  std::memset(&p, 0, sizeof(p));
  p.X = X; p.u0 = u0; p.Z = Z; p.wpack = wpack;
  p.cost = cost; p.command = command; p.error = error; p.pred = pred; p.du0 = du0;
  p.B = B; p.N = N; p.with_grad = with_grad; p.alpha = alpha;
  p.grad_scale = 1.0f / ((float)N * (float)B_global);
  p.noise_std = g_noise_std; p.noise_seed = g_noise_seed;
  p.acc_comp = 1.0f;
  { int e = (int)std::floor(std::log2((double)N * (double)B_global)); p.g_scale = (float)std::ldexp(1.0, e); p.g_unscale = (float)std::ldexp(1.0, -e); }
  p.num_tiles = (B + fc::pr::kTileP / R - 1) / (fc::pr::kTileP / R);
  const int npairs = R == 1 ? (p.num_tiles + fc::pr::kTiles - 1) / fc::pr::kTiles : p.num_tiles;
  if (grid > npairs) grid = npairs;
  fc::pr::WorkLayoutP wl = fc::pr::work_layout_p(N, with_grad);
  p.work_stride = fc::pr::kTiles * wl.total;
  std::vector<float> work((size_t)grid * p.work_stride, 0.f);
  std::vector<double> partial((size_t)grid * fc::kPartialStride, 0.0);
  p.work = work.data();
  p.partial = partial.data();
  for (int b = 0; b < grid; ++b) {
    EmuBlockTC blk(b, grid, fc::pr::kSmFloatsP);
    std::vector<std::thread> th;
    th.reserve(fc::pr::kThreadsP);
    for (int t = 0; t < fc::pr::kThreadsP; ++t)
      th.emplace_back([&blk, &p, t]() {
        EmuCtxTC ctx(&blk, t);
        fc::pr::MpcPair<EmuCtxTC, R> k(ctx, p);
        k.run();
      });
    for (auto& x : th) x.join();
  }
  for (int i = 0; i < 256; ++i) gl[i] = 0.f;
  for (int i = 0; i <= fc::kNumFnnGrad; ++i) {
    double a = 0.0;
    for (int b = 0; b < grid; ++b) a += partial[(size_t)b * fc::kPartialStride + i];
    gl[i] = (float)(i == fc::kNumFnnGrad ? a / (double)B_global : a);
  }
  return 0;
}

extern "C" {
int fc_emu_mpc_loss_pair(const float* X, const float* u0, const float* Z, const float* wpack, int B, int N, float alpha,
                         long long B_global, int with_grad, int grid, float* cost, float* command, float* error,
                         float* pred, float* du0, float* gl /*[256]*/) {
  return emu_mpc_loss_pair<1>(X, u0, Z, wpack, B, N, alpha, B_global, with_grad, grid, cost, command, error, pred, du0, gl);
}
// replica mode of the pair-kernel source (one 32-trajectory tile per CTA); same packed weights as the pair kernel
int fc_emu_pack_floats_replica() { return fc::pr::kPackFloatsP; }
void fc_emu_pack_weights_replica(const float* w_ih0, const float* w_hh0, const float* w_ih1, const float* w_hh1,
                                 const float* w_ih2, const float* w_hh2, const float* fc_w, const float* fc_b,
                                 const float* inp_w, const float* inp_b, const float* out_w, float* out) {
  fc_emu_pack_weights_pair(w_ih0, w_hh0, w_ih1, w_hh1, w_ih2, w_hh2, fc_w, fc_b, inp_w, inp_b, out_w, out);
}
int fc_emu_mpc_loss_replica(const float* X, const float* u0, const float* Z, const float* wpack, int B, int N, float alpha,
                            long long B_global, int with_grad, int grid, float* cost, float* command, float* error,
                            float* pred, float* du0, float* gl /*[256]*/) {
  return emu_mpc_loss_pair<4>(X, u0, Z, wpack, B, N, alpha, B_global, with_grad, grid, cost, command, error, pred, du0, gl);
}
}  // extern "C"

template <int R>
static int emu_lstm_shadow_pair(const float* row0, const float* u, const float* ratio, const float* wpack, int B, int T, int grid,
                            float* y /*[B][T][4]*/) {
  fc::MpcParams p;
  std::memset(&p, 0, sizeof(p));
  p.wpack = wpack;
  p.B = B; p.N = T; p.with_grad = 0;
  p.acc_comp = 1.0f; p.g_scale = p.g_unscale = 1.0f;
  p.shadow = 1; p.sh_row0 = row0; p.sh_u = u; p.sh_y = y;
  for (int q = 0; q < 4; ++q) p.sh_ratio[q] = ratio[q];
  p.num_tiles = (B + fc::pr::kTileP / R - 1) / (fc::pr::kTileP / R);
  const int npairs = R == 1 ? (p.num_tiles + fc::pr::kTiles - 1) / fc::pr::kTiles : p.num_tiles;
  if (grid > npairs) grid = npairs;
  p.work_stride = fc::pr::kTiles * fc::pr::work_layout_p(T, 0).total;
  std::vector<float> work((size_t)grid * p.work_stride, 0.f);
  std::vector<double> partial((size_t)grid * fc::kPartialStride, 0.0);
  p.work = work.data();
  p.partial = partial.data();
  for (int b = 0; b < grid; ++b) {
    EmuBlockTC blk(b, grid, fc::pr::kSmFloatsP);
    std::vector<std::thread> th;
    th.reserve(fc::pr::kThreadsP);
    for (int t = 0; t < fc::pr::kThreadsP; ++t)
      th.emplace_back([&blk, &p, t]() {
        EmuCtxTC ctx(&blk, t);
        fc::pr::MpcPair<EmuCtxTC, R> k(ctx, p);
        k.run();
      });
    for (auto& x : th) x.join();
  }
  return 0;
}

extern "C" {
int fc_emu_lstm_shadow_pair(const float* row0, const float* u, const float* ratio, const float* wpack, int B, int T, int grid,
                            float* y /*[B][T][4]*/) {
  return emu_lstm_shadow_pair<1>(row0, u, ratio, wpack, B, T, grid, y);
}
int fc_emu_lstm_shadow_replica(const float* row0, const float* u, const float* ratio, const float* wpack, int B, int T, int grid,
                               float* y /*[B][T][4]*/) {
  return emu_lstm_shadow_pair<4>(row0, u, ratio, wpack, B, T, grid, y);
}

// surrogate training on the pair-kernel source (MpcParams::train): mode 1 = forward (y, h_last), mode 2 = forward with
// records + reverse sweep seeded by dy, which fills the per-tile scratch (gate gradients / hidden sequences / features in
// operand format) that the tcgen05 weight-gradient kernel consumes; ws must hold tiles * fc_emu_train_tile_floats() floats
long fc_emu_train_tile_floats() { return (long)fc::pr::kTrTileFloats; }
}  // extern "C"
template <int R>
static int emu_lstm_train(int mode, const float* X, const float* dy, const float* wpack, const float* fc_w, const float* fc_b, int B,
                      float g_scale, int grid, float* y /*[B][4]*/, float* hlast /*[B][50]*/, float* ws) {
  fc::MpcParams p;
  std::memset(&p, 0, sizeof(p));
  p.wpack = wpack;
  p.B = B; p.N = 1; p.with_grad = mode == 2 ? 1 : 0;
  p.acc_comp = 1.0f; p.g_scale = p.g_unscale = 1.0f;
  p.train = mode; p.tr_x = X; p.tr_y = mode == 1 ? y : nullptr; p.tr_hlast = mode == 1 ? hlast : nullptr; p.tr_dy = dy;
  p.tr_fcw = fc_w; p.tr_fcb = fc_b;
  float scale[2] = {g_scale, 1.0f / g_scale};
  p.tr_scale = scale; p.tr_ws = ws; p.tr_tile_base = 0;
  p.num_tiles = (B + fc::pr::kTileP / R - 1) / (fc::pr::kTileP / R);
  const int npairs = R == 1 ? (p.num_tiles + fc::pr::kTiles - 1) / fc::pr::kTiles : p.num_tiles;
  if (grid > npairs) grid = npairs;
  p.work_stride = fc::pr::kTiles * (mode == 2 ? fc::pr::work_total_train() : fc::pr::work_layout_p(1, 0).total);
  std::vector<float> work((size_t)grid * p.work_stride, 0.f);
  std::vector<double> partial((size_t)grid * fc::kPartialStride, 0.0);
  p.work = work.data();
  p.partial = partial.data();
  for (int b = 0; b < grid; ++b) {
    EmuBlockTC blk(b, grid, fc::pr::kSmFloatsP);
    std::vector<std::thread> th;
    th.reserve(fc::pr::kThreadsP);
    for (int t = 0; t < fc::pr::kThreadsP; ++t)
      th.emplace_back([&blk, &p, t]() {
        EmuCtxTC ctx(&blk, t);
        fc::pr::MpcPair<EmuCtxTC, R, true> k(ctx, p);
        k.run();
      });
    for (auto& x : th) x.join();
  }
  return 0;
}

extern "C" {
// replica = 0: pairs of 128-sample tiles; replica = 1: 32-sample tiles (tile scratch = fc_emu_train_tile_floats() / 4)
int fc_emu_lstm_train(int mode, const float* X, const float* dy, const float* wpack, const float* fc_w, const float* fc_b, int B,
                      float g_scale, int grid, float* y /*[B][4]*/, float* hlast /*[B][50]*/, float* ws, int replica) {
  return replica ? emu_lstm_train<4>(mode, X, dy, wpack, fc_w, fc_b, B, g_scale, grid, y, hlast, ws)
                 : emu_lstm_train<1>(mode, X, dy, wpack, fc_w, fc_b, B, g_scale, grid, y, hlast, ws);
}

// one-tile tcgen05 kernel with a wide controller (width_dim > 1): gl_wide [2560] = d fc_int.weight | d fc_int.bias
int fc_emu_mpc_loss_tc_wide(const float* X, const float* u0, const float* Z, const float* wpack, const float* int_w,
                            const float* int_b, int width_dim, int B, int N, float alpha, long long B_global, int with_grad,
                            int grid, float* cost, float* command, float* error, float* pred, float* du0, float* gl /*[256]*/,
                            float* gl_wide /*[2560]*/) {
  fc::MpcParams p;
  std::memset(&p, 0, sizeof(p));
  p.X = X; p.u0 = u0; p.Z = Z; p.wpack = wpack;
  p.cost = cost; p.command = command; p.error = error; p.pred = pred; p.du0 = du0;
  p.B = B; p.N = N; p.with_grad = with_grad; p.alpha = alpha;
  p.grad_scale = 1.0f / ((float)N * (float)B_global);
  p.noise_std = g_noise_std; p.noise_seed = g_noise_seed;
  p.acc_comp = 1.0f;
  { int e = (int)std::floor(std::log2((double)N * (double)B_global)); p.g_scale = (float)std::ldexp(1.0, e); p.g_unscale = (float)std::ldexp(1.0, -e); }
  p.width_dim = width_dim; p.int_w = int_w; p.int_b = int_b;
  p.num_tiles = (B + fc::tc::kTileTC - 1) / fc::tc::kTileTC;
  if (grid > p.num_tiles) grid = p.num_tiles;
  fc::tc::WorkLayoutTC wl = fc::tc::work_layout_tc(N, with_grad, width_dim);
  p.work_stride = wl.total;
  std::vector<float> work((size_t)grid * wl.total, 0.f);
  std::vector<double> partial((size_t)grid * fc::kPartialStride, 0.0), partial_wide((size_t)grid * fc::kWidePartialStride, 0.0);
  p.work = work.data();
  p.partial = partial.data();
  p.partial_wide = partial_wide.data();
  for (int b = 0; b < grid; ++b) {
    EmuBlockTC blk(b, grid, fc::tc::kSmFloatsWide);
    std::vector<std::thread> th;
    th.reserve(fc::tc::kThreadsTC);
    for (int t = 0; t < fc::tc::kThreadsTC; ++t)
      th.emplace_back([&blk, &p, t]() {
        EmuCtxTC ctx(&blk, t);
        fc::tc::MpcTileTC<EmuCtxTC> k(ctx, p);
        k.run();
      });
    for (auto& x : th) x.join();
  }
  for (int i = 0; i < 256; ++i) gl[i] = 0.f;
  for (int i = 0; i <= fc::kNumFnnGrad; ++i) {
    double a = 0.0;
    for (int b = 0; b < grid; ++b) a += partial[(size_t)b * fc::kPartialStride + i];
    gl[i] = (float)(i == fc::kNumFnnGrad ? a / (double)B_global : a);
  }
  for (int i = 0; i < fc::kWideGrads; ++i) {
    double a = 0.0;
    for (int b = 0; b < grid; ++b) a += partial_wide[(size_t)b * fc::kWidePartialStride + i];
    gl_wide[i] = (float)a;
  }
  return 0;
}
}  // extern "C"
