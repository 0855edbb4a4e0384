// Controller FNNModel.forward (Unsupervised Learning/Functions.py:261-289, width_dim = 1: Hardtanh(fc_out(ReLU(fc_inp(x)))))
// and its backward for the u_0 path of a training step (`output = model(X)`, Functions.py:643, differentiated by
// loss.backward() at :655): two launches instead of ~12 ATen kernels that stream [B,50] intermediates through HBM.
//   fnn_forward_kernel    one thread per sample, weights (250 floats) broadcast from shared memory; nothing is saved
//   fnn_backward_kernel   recomputes the hidden layer, masks (hardtanh_backward: -1 < v < 1, relu: h > 0), and reduces the 250
//                         weight gradients with warp-shuffle butterflies: lane (j & 31) keeps the five sums of hidden unit j
//   fnn_reduce_kernel     per-block partials -> flat gradient [fc_inp.weight 150 | fc_inp.bias 50 | fc_out.weight 50] (fp64)
// HBM-bound by construction: 12 B read + 4 B written per sample (forward), 16 B read per sample (backward).
#pragma once

namespace fc {
namespace fn {

constexpr int kHidF = 50;
constexpr int kThreadsF = 256;
constexpr int kMaxBlocksF = 148 * 4;

__device__ __forceinline__ void load_weights(float* sw, const float* inp_w, const float* inp_b, const float* out_w) {
  for (int i = threadIdx.x; i < 250; i += blockDim.x)
    sw[i] = i < 150 ? __ldg(inp_w + i) : (i < 200 ? __ldg(inp_b + i - 150) : __ldg(out_w + i - 200));
  __syncthreads();
}

__global__ void __launch_bounds__(kThreadsF) fnn_forward_kernel(const float* __restrict__ X, const float* inp_w, const float* inp_b,
                                                               const float* out_w, long long B, float* __restrict__ u) {
  __shared__ float sw[250];
  load_weights(sw, inp_w, inp_b, out_w);
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
    const float x0 = __ldg(X + b * 3), x1 = __ldg(X + b * 3 + 1), x2 = __ldg(X + b * 3 + 2);
    float v = 0.f;
#pragma unroll 10
    for (int j = 0; j < kHidF; ++j) {
      // torch.nn.Linear accumulates k = 0, 1, 2 then adds the bias (addmm): same order here
      float z = fmaf(sw[j * 3 + 2], x2, fmaf(sw[j * 3 + 1], x1, sw[j * 3] * x0)) + sw[150 + j];
      v = fmaf(sw[200 + j], fmaxf(z, 0.f), v);
    }
    u[b] = fminf(fmaxf(v, -1.f), 1.f);
  }
}

__global__ void __launch_bounds__(kThreadsF) fnn_backward_kernel(const float* __restrict__ X, const float* __restrict__ du,
                                                                const float* inp_w, const float* inp_b, const float* out_w,
                                                                long long B, float* __restrict__ partial /*[grid][250]*/) {
  __shared__ float sw[250];
  __shared__ float red[kThreadsF / 32][250];
  load_weights(sw, inp_w, inp_b, out_w);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[2][5];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int q = 0; q < 5; ++q) acc[s][q] = 0.f;
  const long long chunks = (B + 31) / 32;
  for (long long ch = (long long)blockIdx.x * (kThreadsF / 32) + warp; ch < chunks; ch += (long long)gridDim.x * (kThreadsF / 32)) {
    const long long b = ch * 32 + lane;
    const bool valid = b < B;
    const float x0 = valid ? __ldg(X + b * 3) : 0.f, x1 = valid ? __ldg(X + b * 3 + 1) : 0.f, x2 = valid ? __ldg(X + b * 3 + 2) : 0.f;
    float v = 0.f;
#pragma unroll 10
    for (int j = 0; j < kHidF; ++j) {
      const float z = fmaf(sw[j * 3 + 2], x2, fmaf(sw[j * 3 + 1], x1, sw[j * 3] * x0)) + sw[150 + j];
      v = fmaf(sw[200 + j], fmaxf(z, 0.f), v);
    }
    const float dv = (valid && v > -1.f && v < 1.f) ? __ldg(du + b) : 0.f;        // hardtanh_backward
#pragma unroll 2
    for (int j = 0; j < kHidF; ++j) {
      const float z = fmaf(sw[j * 3 + 2], x2, fmaf(sw[j * 3 + 1], x1, sw[j * 3] * x0)) + sw[150 + j];
      const float h = fmaxf(z, 0.f);
      const float dh = h > 0.f ? dv * sw[200 + j] : 0.f;                          // threshold_backward of ReLU
      float r[5] = {dh * x0, dh * x1, dh * x2, dh, dv * h};
#pragma unroll
      for (int q = 0; q < 5; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r[q] += __shfl_xor_sync(0xffffffffu, r[q], o);
      }
      if (lane == (j & 31)) {
#pragma unroll
        for (int q = 0; q < 5; ++q) acc[j >> 5][q] += r[q];
      }
    }
  }
  // lane l holds unit l (slot 0) and unit 32 + l (slot 1, l < 18): flat layout [inp_w 150 | inp_b 50 | out_w 50]
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int j = s * 32 + lane;
    if (j < kHidF) {
      red[warp][j * 3] = acc[s][0]; red[warp][j * 3 + 1] = acc[s][1]; red[warp][j * 3 + 2] = acc[s][2];
      red[warp][150 + j] = acc[s][3]; red[warp][200 + j] = acc[s][4];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 250; i += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kThreadsF / 32; ++w) a += red[w][i];
    partial[(size_t)blockIdx.x * 250 + i] = a;
  }
}

__global__ void __launch_bounds__(256) fnn_reduce_kernel(const float* __restrict__ partial, int grid, float* __restrict__ g) {
  const int i = threadIdx.x;
  if (i >= 250) return;
  double a = 0.0;
  for (int b = 0; b < grid; ++b) a += (double)partial[(size_t)b * 250 + i];
  g[i] = (float)a;
}

}  // namespace fn
}  // namespace fc
