/* forging_b200 -- C ABI of the B200-native hot path of marcowus/forging-control.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; each entry point below names the
 * reference interface (file:line under "Unsupervised Learning/") whose arithmetic it replaces.  The
 * Python host side (forging_control_b200/Functions.py) binds these with ctypes, see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host or its comment says HOST (the few
 *     5-/4-float parameter vectors: noise standard deviations, scaler ratios); the library borrows it
 *     for the duration of the call and never frees or retains caller memory;
 *   - all calls are enqueue-only on `stream` (a cudaStream_t passed as void*), no hidden
 *     synchronisation, re-entrant per device;
 *   - every function returns 0 on success or a negative fc_status; fc_last_error() returns a
 *     thread-local message for the last failure;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef FORGING_B200_H_
#define FORGING_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum fc_status {
  FC_OK = 0,
  FC_ERR_BAD_SHAPE = -1,     /* B <= 0, N <= 0, T <= 0 ...                                  */
  FC_ERR_UNSUPPORTED = -2,   /* dims other than LSTM(5,50,4,3) / FNN(3,50,1,width 1), ...   */
  FC_ERR_NULL_POINTER = -3,
  FC_ERR_MISALIGNED = -4,    /* pointer not 16-byte aligned where required                  */
  FC_ERR_WORKSPACE = -5,     /* workspace too small                                         */
  FC_ERR_CUDA = -6           /* CUDA runtime error, text in fc_last_error()                 */
} fc_status;

const char* fc_last_error(void);
int fc_version(void);

/* ---- weights ---------------------------------------------------------------------------------
 * Packs the state_dict tensors of LSTMModel(5,50,4,3) (Functions.py:295-379: lstm.weight_ih_l{k}
 * [200,in], lstm.weight_hh_l{k} [200,50], fc.weight [4,50], fc.bias [4]; gate rows i|f|g|o) and of
 * FNNModel(3,50,1,width_dim=1) (Functions.py:215-289: fc_inp.weight [50,3], fc_inp.bias [50],
 * fc_out.weight [1,50]) into the tiled layouts the kernels read (fc_pack_floats() floats).        */
size_t fc_pack_floats(void);
int fc_pack_weights(const float* w_ih0, const float* w_hh0, const float* w_ih1, const float* w_hh1,
                    const float* w_ih2, const float* w_hh2, const float* fc_w, const float* fc_b,
                    const float* fnn_inp_w, const float* fnn_inp_b, const float* fnn_out_w,
                    float* wpack, void* stream);

/* ---- fused MPC loss: replaces MPCLoss.forward (Functions.py:1353-1472) and the autograd sweep
 * triggered by loss.backward() (Functions.py:655) ------------------------------------------------
 *   X [B,3] controller input (reference = column 2), u0 [B] = output_controller.squeeze(),
 *   Z [B,10,5] look-back window; N = prediction horizon; alpha = command-rate weight.
 *   B_global: batch size the mean is taken over (== B on one GPU; the global batch when the batch
 *   is sharded over ranks, so that summing gl over ranks gives the single-GPU result).
 *   Outputs: cost/command/error [B], pred [B,N] (loss_features of the reference),
 *   gl [256]: gl[0..149] d loss/d fc_inp.weight, gl[150..199] d/d fc_inp.bias, gl[200..249]
 *   d/d fc_out.weight (contributions of the commands u_1..u_{N-1}; the u_0 path is returned as
 *   du0 [B] = d loss / d output_controller), gl[250] = sum_b cost[b] / B_global (the loss).
 *   with_grad = 0 computes the forward only (du0 may be NULL, gl[0..249] are zero).              */
size_t fc_mpc_loss_workspace_bytes(int B, int N, int with_grad);
/* Kernel behind fc_mpc_loss: 0 = automatic (B <= 32 x #SMs: the replica mode of the pair kernel, 32-trajectory tiles;
 * up to #SMs tiles of 128 trajectories: the one-tile tcgen05 kernel; beyond: the two-tile tcgen05 pair kernel; measured
 * fastest), 1 = always the FP32 FFMA kernel, 2 = always the one-tile tcgen05 kernel, 3 = always the pair kernel,
 * 4 = always the replica mode, 5 = the pair kernel with the small-argument tanh polynomial (relative instead of absolute
 * tanh accuracy: for a surrogate with tiny cell values such as freshly initialised weights; 7 % slower).  Also settable
 * with the environment variable FC_MPC_KERNEL=ffma|tc|pair|replica|pair-precise before the first call.  All kernels meet
 * the same parity bar on the reference's shipped weights.                                              */
int fc_mpc_select_kernel(int mode);

/* Scratch traffic (bytes) of one with_grad launch for (B, N), computed from the workspace layout of the kernel the
 * current selection picks: activation records saved by the forward roll-out and read once by the reverse sweep
 * (what autograd keeps alive between MPCLoss.forward, Functions.py:1353-1472, and loss.backward(), :655) plus the
 * hidden-sequence scratch handed between LSTM layers.  Used by bench.py for `roofline.traffic` ("computed"); an
 * upper bound of the DRAM traffic, the measured value is the ncu capture under profiles/.  0 on error.           */
size_t fc_mpc_loss_scratch_traffic_bytes(int B, int N);
int fc_mpc_loss(const float* X, const float* u0, const float* Z, const float* wpack, int B, int N,
                float alpha, long long B_global, int with_grad, float* cost, float* command,
                float* error, float* pred, float* du0, float* gl, void* workspace,
                size_t workspace_bytes, void* stream);
/* Same with `enable_noise=True` of MPCLoss.forward (UL/Functions.py:1400-1402, :1438-1440): noise_std * N(0,1) is added
 * to every surrogate output before the cost and the next controller call (the reference uses 0.01).  The normals
 * come from a counter-based Philox4x32-10 generator keyed by noise_seed with counter (trajectory, window), so a call is
 * reproducible for a given seed and independent of the kernel chosen; they are NOT torch.randn's stream.
 * noise_std = 0 is fc_mpc_loss.                                                                            */
int fc_mpc_loss_noise(const float* X, const float* u0, const float* Z, const float* wpack, int B, int N,
                      float alpha, long long B_global, int with_grad, float* cost, float* command,
                      float* error, float* pred, float* du0, float* gl, void* workspace,
                      size_t workspace_bytes, float noise_std, unsigned long long noise_seed, void* stream);
/* Controllers with hidden-layer repeats, FNNModel(width_dim > 1) (UL/Functions.py:261-289: fc_int + ReLU applied
 * width_dim - 1 times with shared weights).  fnn_int_w [50][50], fnn_int_b [50] device pointers; gl_wide [2560] receives
 * d loss / d fc_int.weight (2500, row-major) then d loss / d fc_int.bias (50).  Runs in the one-tile tcgen05 kernel.
 * width_dim = 1 is fc_mpc_loss_noise (fnn_int_*, gl_wide may be NULL).                                          */
size_t fc_mpc_loss_wide_workspace_bytes(int B, int N, int with_grad, int width_dim);
int fc_mpc_loss_wide(const float* X, const float* u0, const float* Z, const float* wpack, const float* fnn_int_w,
                     const float* fnn_int_b, int width_dim, int B, int N, float alpha, long long B_global,
                     int with_grad, float* cost, float* command, float* error, float* pred, float* du0, float* gl,
                     float* gl_wide, void* workspace, size_t workspace_bytes, float noise_std,
                     unsigned long long noise_seed, void* stream);

/* Training-sample construction on the device (replaces DataLoader collation of `SequenceDataset.__getitem__`,
 * UL/Functions.py:109-132, over the per-trajectory slices of `Data.get_individual_dataset`, :479-516).  Tables
 * Xtab [M][3], ytab [M], Ztab [M][5] = the concatenated per-trajectory datasets, M = n_traj * t_traj (device pointers);
 * idx [B] global sample indices (device, int64, 0 <= idx < M; not checked).  Sample g = trajectory g / t_traj, step
 * i = g % t_traj:  X[b] = Xtab[g];  Z[b] = rows max(i-lookback+1, 0)..i of the trajectory, front-padded with its
 * row 0;  y[b] = ytab of the trajectory's next step (the last one for the final step).  Bit-exact gather.       */
int fc_build_windows(const float* Xtab, const float* ytab, const float* Ztab, long long M, int t_traj, int lookback,
                     const long long* idx, long long B, float* X, float* y, float* Z, void* stream);

/* LSTM shadow roll-out of the closed loop (replaces the per-step `simulator_make_step` loop of
 * NeuralNetwork.loop, UL/Functions.py:969-1011 and :1196-1231): T windowed-LSTM inferences per trajectory, batched
 * over B trajectories, forward only.  The window starts as ten copies of row0[b] (scaled: x/scale_in, u_0/scale_in);
 * after window m the surrogate output y[b][m][0..3] (scaled) is logged and [y * ratio, u[b][m+1]] becomes the
 * newest row, ratio[q] = scale_out[q] / scale_in[q] (HOST array of 4 floats).  row0 [B][5], u [B][T] (scaled
 * commands; u[b][0] is not read), y [B][T][4] are device pointers.                                         */
size_t fc_lstm_shadow_workspace_bytes(int B, int T);
int fc_lstm_shadow_rollout(const float* row0, const float* u, const float* ratio, const float* wpack, int B, int T,
                           float* y, void* workspace, size_t workspace_bytes, void* stream);

/* ---- closed-loop deployment: replaces the body of NeuralNetwork.loop (Functions.py:1157-1237)
 * = FeasibilityRecovery.NN_make_step (:1596-1604) + simulator.make_step on the plant of
 * template_model.py:10-161, integrated with the fixed-step RK4 of Ruge_Kuta (:1759-1775) ----------
 *   x0 [B,5] initial state (y, y_dot, p1, p2, z), row-major; ref [n_ref,B] reference, entry
 *   min(t / steps_per_ref, n_ref-1) is used at step t; scale_in [3] MaxAbs scales of the controller
 *   input (y_dot, z, ref) -- the reference is scaled by scale_in[0] like NN_make_step does
 *   (:1597-1598); scale_out [1]; fnn_* = controller weights (raw state_dict layout, float32: the
 *   controller is evaluated in float32 in both variants, Functions.py:1601); T steps of ts seconds,
 *   `substeps` RK4 sub-steps per step (the reference scheme uses 4).
 *   Outputs (each may be NULL): meas [T+1,5,B] measurements (y, y_dot, floored p1, floored p2, z;
 *   template_model.py:152-156; meas[0] = x0), u [T,B] commands, x_final [B,5] raw final state.
 *   Time-major structure-of-arrays output so that a warp writes 128-byte lines.
 *   fc_closed_loop_rk4 integrates in float32, fc_closed_loop_rk4_f64 in float64 (reference dtype). */
int fc_closed_loop_rk4(const float* x0, const float* ref, int n_ref, int steps_per_ref, int B, int T,
                       float ts, int substeps, const float* scale_in, const float* scale_out,
                       const float* fnn_inp_w, const float* fnn_inp_b, const float* fnn_out_w,
                       float* meas, float* u, float* x_final, void* stream);
int fc_closed_loop_rk4_f64(const double* x0, const double* ref, int n_ref, int steps_per_ref, int B,
                           int T, double ts, int substeps, const double* scale_in,
                           const double* scale_out, const float* fnn_inp_w, const float* fnn_inp_b,
                           const float* fnn_out_w, double* meas, double* u, double* x_final,
                           void* stream);
/* The same with the process / measurement noise of NeuralNetwork.loop (UL/Functions.py:1176-1183, do-mpc
 * Simulator.make_step(u0, v0, w0)): x_next = integrate(x, u) + w0, y = measurement(x_next) + v0, the controller reads the
 * noisy measurement.  process_std, meas_std: HOST arrays of 5 floats (NULL = zeros); normals from the counter-based
 * generator of fc_mpc_loss_noise, keyed by seed.                                                            */
int fc_closed_loop_rk4_noise(const float* x0, const float* ref, int n_ref, int steps_per_ref, int B, int T, float ts,
                             int substeps, const float* scale_in, const float* scale_out, const float* fnn_inp_w,
                             const float* fnn_inp_b, const float* fnn_out_w, float* meas, float* u, float* x_final,
                             const float* process_std, const float* meas_std, unsigned long long seed, void* stream);
int fc_closed_loop_rk4_f64_noise(const double* x0, const double* ref, int n_ref, int steps_per_ref, int B, int T,
                                 double ts, int substeps, const double* scale_in, const double* scale_out,
                                 const float* fnn_inp_w, const float* fnn_inp_b, const float* fnn_out_w, double* meas,
                                 double* u, double* x_final, const float* process_std, const float* meas_std,
                                 unsigned long long seed, void* stream);
/* General form: f64 = 0 / 1 selects the plant precision (x0, ref, scale_in, scale_out, meas, u, x_final are float or
 * double arrays accordingly); fnn_int_w [50][50], fnn_int_b [50], width_dim: hidden-layer repeats of FNNModel
 * (UL/Functions.py:261-289; width_dim <= 1: none, pointers may be NULL); noise as above.                      */
int fc_closed_loop_rk4_ex(int f64, const void* x0, const void* ref, int n_ref, int steps_per_ref, int B, int T, double ts,
                          int substeps, const void* scale_in, const void* scale_out, const float* fnn_inp_w,
                          const float* fnn_inp_b, const float* fnn_out_w, const float* fnn_int_w, const float* fnn_int_b,
                          int width_dim, void* meas, void* u, void* x_final, const float* process_std,
                          const float* meas_std, unsigned long long seed, void* stream);

/* ---- surrogate training path: replaces, inside NeuralNetwork.train_model of the surrogate
 * (UL/Model_NN/Functions.py:520-569), `output = model(X, device)` (:551, LSTMModel.forward :313-340) and what
 * `loss.backward()` (:560) leaves in .grad of the eight surrogate parameters.  nn.MSELoss (UL/Model_NN/Main.py:227)
 * stays the caller's: the backward call takes d loss/d output.
 *   fc_lstm_train_pack        raw nn.LSTM weights (state_dict layout, float32) -> kernel images (fc_lstm_train_pack_floats()
 *                             floats); call after every optimizer step.
 *   fc_lstm_window_fwd        X [B,10,5] -> out [B,4]; save != 0 records the cell activations and hidden sequences
 *                             in `workspace` (fc_lstm_window_workspace_bytes(B, 1) bytes) for the backward call;
 *                             save == 0 (inference) needs only the scratch fc_lstm_window_workspace_bytes(B, 0) reports
 *                             (a few bytes, or the per-CTA buffers of the 80-sample forward used for large B).
 *   fc_lstm_window_bwd        d_out [B,4] + the workspace of the matching forward call -> the eight gradient tensors
 *                             (overwritten, state_dict shapes: g_ih0 [200,5], g_hh* / g_ih1 / g_ih2 [200,50],
 *                             g_fc_w [4,50], g_fc_b [4]).                                                          */
/* Path behind fc_lstm_window_fwd / _bwd: 0 = automatic = 2 = the tensor-core path (the pair kernel in training mode -- its
 * replica mode with 32-sample tiles for B <= 32 x #SMs -- for the forward and the data-gradient reverse sweep + a tcgen05
 * weight-gradient kernel contracting over the samples; measured faster at every batch size), 1 = the FP32 FFMA kernels.
 * Thread-local; also
 * FC_LSTM_TRAIN=ffma|tc before the first call.  The workspace size depends on the path: query it after selecting.   */
int fc_lstm_train_select_path(int mode);
/* The path (1 or 2) the calling thread's current selection takes for a batch of B samples.  A backward call that runs on
 * another thread than its forward (autograd worker threads) selects this value first, so that both use one workspace layout. */
int fc_lstm_train_path_for(int B);
size_t fc_lstm_train_pack_floats(void);
int fc_lstm_train_pack(const float* w_ih0, const float* w_hh0, const float* w_ih1, const float* w_hh1, const float* w_ih2,
                       const float* w_hh2, float* pack, void* stream);
size_t fc_lstm_window_workspace_bytes(int B, int save);
int fc_lstm_window_fwd(const float* X, const float* pack, const float* fc_w, const float* fc_b, int B, int save, float* out,
                       void* workspace, size_t workspace_bytes, void* stream);
int fc_lstm_window_bwd(const float* X, const float* d_out, const float* pack, const float* fc_w, int B, void* workspace,
                       size_t workspace_bytes, float* g_ih0, float* g_hh0, float* g_ih1, float* g_hh1, float* g_ih2,
                       float* g_hh2, float* g_fc_w, float* g_fc_b, void* stream);

/* ---- optimizer update: torch.optim.AdamW.step (UL/Main.py:195, UL/Model_NN/Main.py:230; not amsgrad,
 * not maximize) for `count` <= 8 tensors in ONE launch.  params / grads / exp_avg / exp_avg_sq / numel are HOST arrays
 * of `count` device pointers / element counts; `step` counts from 1; the gradient is read as grad * grad_scale.      */
int fc_adamw_step(int count, float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                  const int* numel, int step, float lr, float beta1, float beta2, float eps, float weight_decay,
                  float grad_scale, void* stream);

/* ---- controller forward / backward for the u_0 path of a training step: `output = model(X)` (UL/Functions.py:643,
 * FNNModel.forward :261-289, width_dim = 1, ReLU, Hardtanh) and its part of loss.backward() (:655).
 *   fc_fnn_forward   X [B,3] -> u [B] (= [B,1]); raw state_dict weights fc_inp.weight [50,3], fc_inp.bias [50],
 *                    fc_out.weight [1,50]; nothing is saved.
 *   fc_fnn_backward  du [B] -> g_flat [250] = [d fc_inp.weight 150 | d fc_inp.bias 50 | d fc_out.weight 50] (overwritten);
 *                    workspace of fc_fnn_backward_workspace_bytes() bytes.                                            */
int fc_fnn_forward(const float* X, const float* inp_w, const float* inp_b, const float* out_w, long long B, float* u, void* stream);
size_t fc_fnn_backward_workspace_bytes(void);
int fc_fnn_backward(const float* X, const float* du, const float* inp_w, const float* inp_b, const float* out_w, long long B,
                    float* g_flat, void* workspace, size_t workspace_bytes, void* stream);

/* ---- measurement helper: register-resident FFMA loop used by bench.py to measure the FP32
 * roofline denominator on the device it runs on; writes achieved FLOP/s to *flops_host.           */
int fc_fp32_peak(int iters, double* flops_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FORGING_B200_H_ */
