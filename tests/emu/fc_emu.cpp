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
#include <thread>
#include <vector>

#define FC_HD_CTX inline
#include "../../forging_control_b200/csrc/fc_mpc_kernel.inl"

namespace {

struct EmuBlock {
  std::barrier<> bar;
  std::vector<float> smem;
  int bid, nblk;
  EmuBlock(int b, int n) : bar(fc::kThreads), smem(fc::kSmFloats, 0.f), bid(b), nblk(n) {}
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
  static float ldcg(const float* p) { return *p; }
  static void cp_async16(float* d, const float* s) { std::memcpy(d, s, 16); }
  static void cp_commit() {}
  template <int N> static void cp_wait() {}
};

}  // namespace

extern "C" {

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
  p.num_tiles = (B + fc::kTile - 1) / fc::kTile;
  if (grid > p.num_tiles) grid = p.num_tiles;
  fc::WorkLayout wl = fc::work_layout(N, with_grad);
  p.work_stride = wl.total;
  std::vector<float> work((size_t)grid * wl.total, 0.f);
  std::vector<float> partial((size_t)grid * fc::kPartialStride, 0.f);
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
  for (int b = 0; b < grid; ++b) {
    for (int i = 0; i < fc::kNumFnnGrad; ++i) gl[i] += partial[(size_t)b * fc::kPartialStride + i];
    gl[fc::kNumFnnGrad] += partial[(size_t)b * fc::kPartialStride + fc::kNumFnnGrad];
  }
  gl[fc::kNumFnnGrad] /= (float)B_global;
  return 0;
}

}  // extern "C"
