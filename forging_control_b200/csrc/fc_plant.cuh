// 5-state hydraulic forging press (Unsupervised Learning/template_model.py:20-156) and the
// fixed-step RK4 of FeasibilityRecovery.Ruge_Kuta (Unsupervised Learning/Functions.py:1759-1775),
// one trajectory per thread, state in registers.  R = float or double.
#pragma once
#include <math.h>

namespace fc {

template <typename R> struct Mth;
// float: the transcendental functions of the right-hand side go straight to the MUFU unit (lg2 / ex2 / rsqrt / rcp,
// 1-2 ulp each) instead of the IEEE library routines (powf alone is ~80 instructions, and the RHS is evaluated 16 times
// per 1 ms step).  The exponents of the forging-force law are small (|M2|, |M3|, A < 0.4), so pow = ex2(y * lg2(x)) keeps
// a relative error of ~1e-7; measured against the fp64 oracle the one-step error stays below 1e-6 of the state scale
// (tests/test_gpu_closed_loop.py), two orders below the 1e-4 bar of the path.
template <> struct Mth<float> {
  static __device__ __forceinline__ float lg2_(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  static __device__ __forceinline__ float ex2_(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  static __device__ __forceinline__ float rcp_(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  static __device__ __forceinline__ float sqrt_(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  static __device__ __forceinline__ float log_(float x) { return 0.69314718055994531f * lg2_(x); }
  static __device__ __forceinline__ float exp_(float x) { return ex2_(1.4426950408889634f * x); }
  static __device__ __forceinline__ float pow_(float x, float y) { return ex2_(y * lg2_(x)); }
  static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
  static __device__ __forceinline__ float div_(float a, float b) { return a * rcp_(b); }
};
template <> struct Mth<double> {
  static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
  static __device__ __forceinline__ double log_(double x) { return log(x); }
  static __device__ __forceinline__ double exp_(double x) { return exp(x); }
  static __device__ __forceinline__ double pow_(double x, double y) { return pow(x, y); }
  static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
  static __device__ __forceinline__ double div_(double a, double b) { return a / b; }
};

// constants, template_model.py:20-62, 88-92 (evaluated in double, rounded once to R)
namespace pc {
constexpr double kPi = 3.14159265358979323846;
constexpr double M = 90000.0, Bv = 25000.0, FT = 200000.0, D1 = 0.6, D2 = 0.5;
constexpr double A1 = kPi * D1 * D1 / 4, A2 = kPi * D2 * D2 / 4, G = 9.81;
constexpr double KB = 22e9, V1_0 = 0.3, V2_0 = 0.1, KL_1 = 8e-13, KL_2 = 14e-14;
constexpr double CD = 0.63, RHO = 858.0, D = 0.006, PS = 32e6, PT = 101325.0;
constexpr double MU = 0.3, K = 1.115, W0 = 0.2, H0 = 0.5, B0 = 0.1;
constexpr double A = 0.14 + 0.36 * (B0 / W0) - 0.054 * (B0 / W0) * (B0 / W0);
constexpr double T1 = 0.005;
constexpr double M0 = 1200e6, M1 = -0.0025, M2 = -0.0587, M3 = 0.1165, M4 = -0.0065, TEMP = 900.0;
constexpr double EXP_M1T = 0.10539922456186433;   // exp(M1*TEMP) = exp(-2.25)
constexpr double FLOOR_EPS = 1e-6;
}  // namespace pc

template <typename R>
__device__ __forceinline__ R smooth_floor(R p) {                 // template_model.py:107-112
  return R(0.5) * (p + Mth<R>::sqrt_(p * p + R(pc::FLOOR_EPS)));
}
// float: for |p| > 8 Pa the 1e-6 under the root is below half an ulp of p^2, the root IS |p| in fp32 and the floor is
// max(p, 0) exactly; the square root is only evaluated in the +-8 Pa band around zero (pressures live at 1e5..3e7 Pa)
template <>
__device__ __forceinline__ float smooth_floor<float>(float p) {
  if (fabsf(p) > 8.0f) return fmaxf(p, 0.0f);
  return 0.5f * (p + Mth<float>::sqrt_(p * p + float(pc::FLOOR_EPS)));
}

template <typename R>
__device__ __forceinline__ R valve_flow(R kv, R dp) {            // template_model.py:120-125
  R q = kv * Mth<R>::sqrt_(R(2.0 / pc::RHO) * Mth<R>::abs_(dp));
  return dp > R(0) ? q : (dp < R(0) ? -q : R(0));
}

template <typename R>
__device__ __forceinline__ void press_rhs(const R (&x)[5], R u, R (&dx)[5]) {
  const R y = x[0], v = x[1], p1 = x[2], p2 = x[3], z = x[4];
  R Fd = R(0);
  if (sizeof(R) == 4 && y > R(0) && v >= R(0)) {
    // float: template_model.py:74-99 with the powers, the logarithm and the exponential folded onto ONE base-2
    // logarithm of H0/h1 and ONE exponential (10 MUFU operations instead of 16; algebraically identical):
    //   w1 = W0 ratio^A, ratio W0 / w1 = ratio^(1-A), e = ln ratio, Fd = Kd Ad M0 e^(M1 T) 2^(M2 lg2 e + M3 lg2 e_dot + lg2(e) M4 / e)
    const float fy = (float)y, fv = (float)v;
    const float rh = Mth<float>::rcp_(float(pc::H0) - fy);
    const float L = Mth<float>::lg2_(float(pc::H0) * rh);                     // lg2(ratio)
    const float w1 = float(pc::W0) * Mth<float>::ex2_(float(pc::A) * L);
    const float b1 = float(pc::B0) * (1.0f + 0.67f * (Mth<float>::ex2_(float(1.0 - pc::A) * L) - 1.0f));
    const float Kd = float(pc::K) * (1.0f + float(pc::MU) * b1 * Mth<float>::rcp_(2.0f * fy) + fy * Mth<float>::rcp_(4.0f * b1));
    const float e = 0.69314718055994531f * L;
    const float e_dot = fv * rh;
    const float ex = float(pc::M2) * Mth<float>::lg2_(e) + float(pc::M3) * Mth<float>::lg2_(e_dot) +
                     float(1.4426950408889634 * pc::M4) * Mth<float>::rcp_(e);
    Fd = (R)(Kd * (w1 * b1) * float(pc::M0 * pc::EXP_M1T) * Mth<float>::ex2_(ex));
  } else if (y > R(0) && v >= R(0)) {                             // template_model.py:74-99
    const R h1 = R(pc::H0) - y;
    const R ratio = Mth<R>::div_(R(pc::H0), h1);
    const R w1 = R(pc::W0) * Mth<R>::pow_(ratio, R(pc::A));
    const R b1 = R(pc::B0) * (R(1) + R(0.67) * (Mth<R>::div_(ratio * R(pc::W0), w1) - R(1)));
    const R Kd = R(pc::K) * (R(1) + Mth<R>::div_(R(pc::MU) * b1, R(2) * y) + Mth<R>::div_(y, R(4) * b1));
    const R Ad = w1 * b1;
    const R e = Mth<R>::log_(ratio);
    const R e_dot = Mth<R>::div_(v, h1);
    Fd = Kd * Ad * R(pc::M0 * pc::EXP_M1T) * Mth<R>::pow_(e, R(pc::M2)) * Mth<R>::pow_(e_dot, R(pc::M3)) *
         Mth<R>::exp_(Mth<R>::div_(R(pc::M4), e));
  }
  const R p1e = smooth_floor(p1), p2e = smooth_floor(p2);
  const R kv = R(pc::kPi * pc::D * pc::CD) * z;
  R qPB, qAT;
  if (z >= R(0)) {                                                // template_model.py:128-129
    qPB = valve_flow(kv, R(pc::PS) - p1e);
    qAT = valve_flow(kv, p2e - R(pc::PT));
  } else {
    qPB = valve_flow(kv, p1e - R(pc::PT));
    qAT = valve_flow(kv, R(pc::PS) - p2e);
  }
  const R V1 = R(pc::V1_0 / 2) + R(pc::A1) * y;
  const R V2 = R(pc::V2_0 / 2) - R(pc::A2) * y;
  const R Ft = Mth<R>::abs_(v) <= R(0.5) ? R(pc::FT / 0.5) * v : R(pc::FT);      // template_model.py:142
  dx[0] = v;
  dx[1] = (R(3 * pc::kPi * pc::D1 * pc::D1 / 4) * p1e - R(pc::kPi * pc::D2 * pc::D2 / 2) * p2e - R(pc::Bv) * v - Ft - Fd) * R(1.0 / pc::M) + R(pc::G);
  dx[2] = Mth<R>::div_(R(pc::KB), V1) * (qPB * R(1.0 / 3.0) - R(pc::A1) * v - R(pc::KL_1) * p1e);
  dx[3] = Mth<R>::div_(R(pc::KB), V2) * (-qAT * R(0.5) + R(pc::A2) * v - R(pc::KL_2) * p2e);
  dx[4] = (u - z) * R(1.0 / pc::T1);
}

// NOISE: the process noise w of do-mpc's model.set_rhs(..., process_noise=True) (template_model.py:145-149) is an
// ADDITIVE TERM OF THE RIGHT-HAND SIDE, dx/dt = f(x, u) + w, held constant over one simulator step.
template <typename R, bool NOISE = false>
__device__ __forceinline__ void rk4_substep(R (&x)[5], R u, R h, const R (&w)[5]) {   // Functions.py:1767-1775
  R k1[5], k2[5], k3[5], k4[5], xt[5];
  const R h2 = h * R(0.5), h6 = h * R(1.0 / 6.0);
  press_rhs(x, u, k1);
#pragma unroll
  for (int i = 0; i < 5; ++i) { if (NOISE) k1[i] += w[i]; xt[i] = x[i] + h2 * k1[i]; }
  press_rhs(xt, u, k2);
#pragma unroll
  for (int i = 0; i < 5; ++i) { if (NOISE) k2[i] += w[i]; xt[i] = x[i] + h2 * k2[i]; }
  press_rhs(xt, u, k3);
#pragma unroll
  for (int i = 0; i < 5; ++i) { if (NOISE) k3[i] += w[i]; xt[i] = x[i] + h * k3[i]; }
  press_rhs(xt, u, k4);
#pragma unroll
  for (int i = 0; i < 5; ++i) { if (NOISE) k4[i] += w[i]; x[i] = x[i] + h6 * (k1[i] + R(2) * k2[i] + R(2) * k3[i] + k4[i]); }
}

// process / measurement noise of do-mpc's Simulator.make_step(u0, v0, w0) as driven by NeuralNetwork.loop
// (UL/Functions.py:1176-1183): every state is declared with process_noise=True (template_model.py:145-149), i.e.
// x_next = integrate(dx/dt = f(x, u) + w0 over t_step) with ONE draw w0 ~ N(0, process_std) per make_step, held
// constant over the step (a rate: UL/Main.py:88-96 uses 0.5 m/s on y, 5e7 Pa/s on the pressures);
// y = measurement(x_next) + v0, v0 ~ N(0, meas_std) per state; the controller reads the noisy measurement.
// Normals: philox_normal4 with counter (trajectory, 3*step + k).  std all zero = off.
// hidden-layer repeats of the controller (FNNModel.forward, UL/Functions.py:261-289): width_dim - 1 applications of the
// weight-shared fc_int + ReLU
struct ClosedLoopWide {
  const float* int_w;   // [50][50]
  const float* int_b;   // [50]
  int width_dim;        // <= 1: none
};

struct ClosedLoopNoise {
  float process_std[5];
  float meas_std[5];
  unsigned long long seed;
  int on;
};

// closed loop: scaler -> FNN (float32) -> saturation -> inverse scaler -> RK4 plant step
// WIDE / NOISE are compile-time switches: the plain variant (reference configuration: width_dim = 1, zero noise)
// keeps its register and shared-memory footprint (the generic one measured 35 % slower).
template <typename R, bool WIDE, bool NOISE>
__global__ void __launch_bounds__(128) closed_loop_kernel(
    const R* __restrict__ x0, const R* __restrict__ ref, int n_ref, int steps_per_ref, int B, int T, R ts,
    int substeps, const R* __restrict__ scale_in, const R* __restrict__ scale_out,
    const float* __restrict__ inp_w, const float* __restrict__ inp_b, const float* __restrict__ out_w,
    R* __restrict__ meas, R* __restrict__ ucmd, R* __restrict__ x_final, ClosedLoopNoise nz, ClosedLoopWide wd) {
  __shared__ float s_w[50 * 3], s_b[50], s_o[50];
  __shared__ float s_iw[WIDE ? 50 * 50 : 1], s_ib[WIDE ? 50 : 1];
  for (int i = threadIdx.x; i < 150; i += blockDim.x) s_w[i] = inp_w[i];
  for (int i = threadIdx.x; i < 50; i += blockDim.x) { s_b[i] = inp_b[i]; s_o[i] = out_w[i]; }
  if (WIDE) {
    for (int i = threadIdx.x; i < 2500; i += blockDim.x) s_iw[i] = wd.int_w[i];
    for (int i = threadIdx.x; i < 50; i += blockDim.x) s_ib[i] = wd.int_b[i];
  }
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const R si0 = scale_in[0], si1 = scale_in[1], so = scale_out[0];
  R x[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) x[i] = x0[(size_t)b * 5 + i];
  if (meas) {
#pragma unroll
    for (int i = 0; i < 5; ++i) meas[(size_t)i * B + b] = x[i];           // Functions.py:1134-1138
  }
  const R h = ts / R(substeps);
  R ym1 = x[1], ym4 = x[4];                                                  // measured y_dot, z seen by the controller
  for (int t = 0; t < T; ++t) {
    int ir = t / steps_per_ref;
    ir = ir < n_ref ? ir : n_ref - 1;
    const R r = ref[(size_t)ir * B + b];
    // NN_make_step, Functions.py:1596-1604: MaxAbs scale, reference scaled by the y_dot scaler
    const float f0 = (float)(ym1 / si0), f1 = (float)(ym4 / si1), f2 = (float)(r / si0);
    float vv = 0.f;
    if (!WIDE) {
#pragma unroll 10
      for (int k = 0; k < 50; ++k) {
        float pre = fmaf(s_w[k * 3 + 2], f2, fmaf(s_w[k * 3 + 1], f1, fmaf(s_w[k * 3], f0, s_b[k])));
        vv = fmaf(s_o[k], fmaxf(pre, 0.f), vv);
      }
    } else {
      float a[50], a2[50];
      for (int k = 0; k < 50; ++k)
        a[k] = fmaxf(fmaf(s_w[k * 3 + 2], f2, fmaf(s_w[k * 3 + 1], f1, fmaf(s_w[k * 3], f0, s_b[k]))), 0.f);
      for (int r = 1; r < wd.width_dim; ++r) {
        for (int i = 0; i < 50; ++i) {
          float pre = s_ib[i];
          for (int k = 0; k < 50; ++k) pre = fmaf(s_iw[i * 50 + k], a[k], pre);
          a2[i] = fmaxf(pre, 0.f);
        }
        for (int i = 0; i < 50; ++i) a[i] = a2[i];
      }
      for (int k = 0; k < 50; ++k) vv = fmaf(s_o[k], a[k], vv);
    }
    const float us = fminf(fmaxf(vv, -1.f), 1.f);                          // nn.Hardtanh
    const R u = (R)us * so;
    if (ucmd) ucmd[(size_t)t * B + b] = u;
    float e[12];
    R w[5] = {R(0), R(0), R(0), R(0), R(0)};
    if (NOISE) {
      philox_normal4(nz.seed, (unsigned)b, 3u * (unsigned)t, e);
      philox_normal4(nz.seed, (unsigned)b, 3u * (unsigned)t + 1u, e + 4);
      philox_normal4(nz.seed, (unsigned)b, 3u * (unsigned)t + 2u, e + 8);
#pragma unroll
      for (int i = 0; i < 5; ++i) w[i] = (R)(nz.process_std[i] * e[i]);      // one draw per step, a term of the RHS
    }
    for (int s = 0; s < substeps; ++s) rk4_substep<R, NOISE>(x, u, h, w);
    R y[5] = {x[0], x[1], smooth_floor(x[2]), smooth_floor(x[3]), x[4]};     // template_model.py:154-155
    if (NOISE) {
#pragma unroll
      for (int i = 0; i < 5; ++i) y[i] += (R)(nz.meas_std[i] * e[5 + i]);
    }
    ym1 = y[1]; ym4 = y[4];
    if (meas) {
      R* mp = meas + (size_t)(t + 1) * 5 * B + b;
#pragma unroll
      for (int i = 0; i < 5; ++i) mp[(size_t)i * B] = y[i];
    }
  }
  if (x_final) {
#pragma unroll
    for (int i = 0; i < 5; ++i) x_final[(size_t)b * 5 + i] = x[i];
  }
}

}  // namespace fc
