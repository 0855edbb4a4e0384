"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product; the product package never imports it.

CPU restatement of the reference's MPC-loss hot path, used as the checker for the CUDA kernels:

* ``lstm_window_forward``      restates ``LSTMModel.forward``   (UL/Functions.py:353-379):
  3 stacked bias-free LSTM layers run from a zero (h, c) state over a 10-row window, PyTorch gate
  order i,f,g,o, read-out ``fc`` on the last hidden state.
* ``fnn_forward``              restates ``FNNModel.forward``    (UL/Functions.py:261-289).
* ``mpc_loss_forward``         restates ``MPCLoss.forward``     (UL/Functions.py:1353-1472).
* ``mpc_loss_forward_backward`` additionally restates what ``loss.backward()`` produces at
  UL/Functions.py:655 for the live controller parameters and for ``output_controller`` -- a
  hand-derived reverse sweep (no autograd), following the dataflow in SURVEY.md section 8(a).
* ``mpc_loss_torch``           the same forward written with torch ops so that autograd gives an
  independent gradient and so that ``bench.py``'s ``cpu_baseline`` leg can time a CPU path that
  dispatches the same ATen kernels as the reference (``nn.LSTM``-free: explicit gate matmuls).

Parity pin: ``oracle/make_golden.py`` runs the *unmodified* reference (imported through
``oracle/ref_shim.py``) on seeded inputs and commits inputs/outputs/gradients under
``tests/golden``; ``tests/test_oracle_mpc.py`` checks every function here against those vectors
(fp32 and fp64).  Arithmetic is numpy in the dtype of the inputs (float32 or float64).
"""
from __future__ import annotations

import numpy as np

LOOKBACK = 10      # hard-coded look-back of the loss, UL/Functions.py:1434
HIDDEN = 50
LAYERS = 3
P1_MAX = 2.122366  # UL/Functions.py:1411
P2_MAX = 1.036233


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


# ----------------------------------------------------------------------------------------------
# weights container
# ----------------------------------------------------------------------------------------------
def weights_from_state_dicts(lstm_sd: dict, fnn_sd: dict, dtype=np.float64) -> dict:
    """Collect numpy weights from the two ``state_dict`` layouts."""
    def a(x):
        return np.asarray(x.detach().cpu().numpy() if hasattr(x, "detach") else x, dtype=dtype)
    w = {
        "w_ih": [a(lstm_sd[f"lstm.weight_ih_l{l}"]) for l in range(LAYERS)],
        "w_hh": [a(lstm_sd[f"lstm.weight_hh_l{l}"]) for l in range(LAYERS)],
        "fc_w": a(lstm_sd["fc.weight"]), "fc_b": a(lstm_sd["fc.bias"]),
        "inp_w": a(fnn_sd["fc_inp.weight"]), "inp_b": a(fnn_sd["fc_inp.bias"]),
        "out_w": a(fnn_sd["fc_out.weight"]),
    }
    if "fc_int.weight" in fnn_sd:
        w["int_w"] = a(fnn_sd["fc_int.weight"])
        w["int_b"] = a(fnn_sd["fc_int.bias"])
    return w


# ----------------------------------------------------------------------------------------------
# surrogate (LSTMModel.forward, UL/Functions.py:353-379)
# ----------------------------------------------------------------------------------------------
def lstm_window_forward(w: dict, window: np.ndarray, keep: bool = False):
    """window [B,10,5] -> x [B,4].  With ``keep`` also returns the per-cell activations
    (i,f,g,o,c_prev,c) needed by the reverse sweep."""
    B, L, _ = window.shape
    H = HIDDEN
    dt = window.dtype
    seq = window
    saved = []
    for l in range(LAYERS):
        w_ih, w_hh = w["w_ih"][l], w["w_hh"][l]
        h = np.zeros((B, H), dt)
        c = np.zeros((B, H), dt)
        outs = np.empty((B, L, H), dt)
        layer_saved = []
        for t in range(L):
            gates = seq[:, t, :] @ w_ih.T + h @ w_hh.T              # [B,200], rows i|f|g|o
            i = _sigmoid(gates[:, 0:H])
            f = _sigmoid(gates[:, H:2 * H])
            g = np.tanh(gates[:, 2 * H:3 * H])
            o = _sigmoid(gates[:, 3 * H:4 * H])
            c_prev = c
            c = f * c_prev + i * g
            h = o * np.tanh(c)
            outs[:, t, :] = h
            if keep:
                layer_saved.append((i, f, g, o, c_prev, c))
        saved.append(layer_saved)
        seq = outs
    x = seq[:, -1, :] @ w["fc_w"].T + w["fc_b"]
    if keep:
        return x, (saved, seq[:, -1, :])
    return x


def lstm_window_backward(w: dict, saved, gx: np.ndarray, t_min: int = 0) -> np.ndarray:
    """Data gradient of one window: gx [B,4] -> d(window) [B,10,5].

    Only the gradients of window rows ``t >= t_min`` are produced (rows below are constants of
    the batch, UL/Functions.py:1395); time steps below ``t_min`` cannot reach those rows, so the
    sweep stops there (exactly equal to the full sweep on the rows it returns)."""
    cells, _ = saved
    B = gx.shape[0]
    H = HIDDEN
    dt = gx.dtype
    L = LOOKBACK
    d_seq = np.zeros((B, L, H), dt)
    d_seq[:, L - 1, :] = gx @ w["fc_w"]                                # through fc
    for l in reversed(range(LAYERS)):
        w_ih, w_hh = w["w_ih"][l], w["w_hh"][l]
        in_dim = w_ih.shape[1]
        d_in = np.zeros((B, L, in_dim), dt)
        dh_rec = np.zeros((B, H), dt)
        dc = np.zeros((B, H), dt)
        for t in range(L - 1, t_min - 1, -1):
            i, f, g, o, c_prev, c = cells[l][t]
            dh = d_seq[:, t, :] + dh_rec
            tc = np.tanh(c)
            do = dh * tc
            dct = dc + dh * o * (1.0 - tc * tc)
            di = dct * g
            dg = dct * i
            df = dct * c_prev
            dc = dct * f
            dgate = np.concatenate(
                [di * i * (1.0 - i), df * f * (1.0 - f), dg * (1.0 - g * g), do * o * (1.0 - o)], axis=1)
            d_in[:, t, :] = dgate @ w_ih
            dh_rec = dgate @ w_hh
        d_seq = d_in
    return d_seq                                                      # [B,10,5]


# ----------------------------------------------------------------------------------------------
# controller (FNNModel.forward, UL/Functions.py:261-289)
# ----------------------------------------------------------------------------------------------
def lstm_shadow_rollout(w: dict, row0: np.ndarray, u: np.ndarray, ratio: np.ndarray):
    """LSTM shadow roll-out of the closed loop, restating ``NeuralNetwork.simulator_make_step``
    (UL/Functions.py:969-1011) as driven by ``NeuralNetwork.loop`` (:1196-1231), in the scaled domain: the window
    starts as ten copies of ``row0`` [B,5]; after window m the surrogate output y_m [B,4] is logged and
    ``[y_m * ratio, u[:, m+1]]`` (ratio = scale_out / scale_in) becomes the newest row.  Returns y [B,T,4]."""
    B, T = u.shape
    window = np.repeat(row0[:, None, :], LOOKBACK, axis=1).astype(row0.dtype)
    out = np.empty((B, T, 4), row0.dtype)
    for m in range(T):
        y = lstm_window_forward(w, window)
        out[:, m] = y
        nxt = np.concatenate((y * ratio[None, :], (u[:, m + 1:m + 2] if m + 1 < T else np.zeros((B, 1), row0.dtype))), axis=1)
        window = np.concatenate((window[:, 1:], nxt[:, None, :]), axis=1)
    return out


def fnn_forward(w: dict, x: np.ndarray, width_dim: int = 1, keep: bool = False):
    pre = [x @ w["inp_w"].T + w["inp_b"]]
    act = [np.maximum(pre[0], 0)]
    for _ in range(width_dim - 1):
        pre.append(act[-1] @ w["int_w"].T + w["int_b"])
        act.append(np.maximum(pre[-1], 0))
    v = act[-1] @ w["out_w"].T                                         # [B,1]
    u = np.clip(v, -1.0, 1.0)                                          # nn.Hardtanh
    if keep:
        return u, (x, pre, act, v)
    return u


def fnn_backward(w: dict, kept, gu: np.ndarray, grads: dict, width_dim: int = 1) -> np.ndarray:
    """gu [B,1] -> d(input) [B,3]; accumulates parameter gradients into ``grads``."""
    x, pre, act, v = kept
    dv = gu * ((v > -1.0) & (v < 1.0))                                 # hardtanh_backward
    grads["out_w"] += dv.T @ act[-1]
    da = dv @ w["out_w"]
    for k in range(width_dim - 1, 0, -1):
        dp = da * (pre[k] > 0)
        grads["int_w"] += dp.T @ act[k - 1]
        grads["int_b"] += dp.sum(0)
        da = dp @ w["int_w"]
    dp = da * (pre[0] > 0)
    grads["inp_w"] += dp.T @ x
    grads["inp_b"] += dp.sum(0)
    return dp @ w["inp_w"]


def _constraint(x):
    """UL/Functions.py:1411 / :1449."""
    r = lambda v: np.maximum(v, 0)
    return r(-x[:, 1]) + r(-x[:, 2]) + r(x[:, 1] - P1_MAX) + r(x[:, 2] - P2_MAX)


# ----------------------------------------------------------------------------------------------
# loss (MPCLoss.forward, UL/Functions.py:1353-1472)
# ----------------------------------------------------------------------------------------------
def philox_normal4(seed: int, B: int, N: int) -> np.ndarray:
    """The kernels' noise for ``enable_noise`` (forging_control_b200/csrc/fc_layout.h::philox_normal4): four standard
    normals per (trajectory b, window m) from Philox4x32-10 (key = seed, counter = (b, m, 0, 0)) + Box-Muller in
    float32.  Returns [B,N,4] float64.  (The reference draws ``0.01*torch.randn_like(x)``, UL/Functions.py:1401; its
    stream cannot be reproduced, the distribution is the same.)"""
    b = np.repeat(np.arange(B, dtype=np.uint64)[:, None], N, axis=1)
    m = np.repeat(np.arange(N, dtype=np.uint64)[None, :], B, axis=0)
    c = [b.copy(), m.copy(), np.zeros_like(b), np.zeros_like(b)]
    k0, k1 = np.uint64(seed & 0xffffffff), np.uint64((seed >> 32) & 0xffffffff)
    M32 = np.uint64(0xffffffff)
    for _ in range(10):
        p0, p1 = np.uint64(0xD2511F53) * c[0], np.uint64(0xCD9E8D57) * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & M32, p1 & M32, ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & M32, p0 & M32]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & M32, (k1 + np.uint64(0xBB67AE85)) & M32
    f32 = np.float32
    s = f32(2.3283064365386963e-10)
    u = [(ci.astype(f32) + f32(0.5)) * s for ci in c]
    clip = lambda v: np.clip(v, f32(1e-30), f32(0.99999994))
    r0 = np.sqrt(f32(-2.0) * np.log(clip(u[0])))
    r1 = np.sqrt(f32(-2.0) * np.log(clip(u[2])))
    t0, t1 = f32(6.2831853071795865) * u[1], f32(6.2831853071795865) * u[3]
    return np.stack((r0 * np.cos(t0), r0 * np.sin(t0), r1 * np.cos(t1), r1 * np.sin(t1)), axis=-1).astype(np.float64)


def mpc_loss_forward(w, X, u0, Z, N, alpha, width_dim=1, keep=False, noise=None):
    """X [B,3], u0 [B] (= output_controller.squeeze()), Z [B,10,5].  ``noise`` [B,N,4] (already scaled by the
    standard deviation) is added to the surrogate output of every window (enable_noise, UL/Functions.py:1400-1402).

    Returns dict(loss, cost[B], command[B], error[B], prediction[B,N]) (+ tape when ``keep``)."""
    B = X.shape[0]
    dt = X.dtype
    ref = X[:, -1]
    rows = np.empty((B, LOOKBACK + N, 5), dt)          # rows[:, 9+k] = [x_k, u_k]; last row holds x_N
    rows[:, :LOOKBACK] = Z
    rows[:, LOOKBACK - 1, 4] = u0
    cost = np.zeros((N, B), dt)
    cmd = np.zeros((N, B), dt)
    err = np.zeros((N, B), dt)
    pred = np.empty((B, N), dt)
    pred[:, 0] = u0
    tape = {"lstm": [], "fnn": [None]}
    u_prev = rows[:, LOOKBACK - 2, 4]
    u_cur = u0
    for m in range(N):
        res = lstm_window_forward(w, rows[:, m:m + LOOKBACK], keep=keep)
        x = res[0] if keep else res
        if noise is not None:
            x = x + noise[:, m].astype(dt)
        if keep:
            tape["lstm"].append(res[1])
        cmd[m] = alpha * np.square(u_prev - u_cur)
        err[m] = np.square(x[:, 0] - ref)
        cost[m] = err[m] + cmd[m] + _constraint(x)
        rows[:, LOOKBACK + m, :4] = x
        if m + 1 < N:
            inp = np.stack((x[:, 0], x[:, 3], ref), axis=1)
            r = fnn_forward(w, inp, width_dim, keep=keep)
            u_next = (r[0] if keep else r)[:, 0]
            if keep:
                tape["fnn"].append(r[1])
            rows[:, LOOKBACK + m, 4] = u_next
            pred[:, m + 1] = u_next
            u_prev, u_cur = u_cur, u_next
        else:
            rows[:, LOOKBACK + m, 4] = 0
    out = {
        "cost": cost.sum(0) / N, "command": cmd.sum(0) / N, "error": err.sum(0) / N,
        "prediction": pred,
    }
    out["loss"] = out["cost"].mean()
    if keep:
        tape["rows"] = rows
        return out, tape
    return out


def mpc_loss_forward_backward(w, X, u0, Z, N, alpha, width_dim=1, prune=True, noise=None):
    """Forward + hand-derived reverse sweep.  Returns (forward dict, grads dict) with
    ``grads`` = d loss / d {u0 [B], inp_w, inp_b, out_w, (int_w, int_b)}."""
    out, tape = mpc_loss_forward(w, X, u0, Z, N, alpha, width_dim, keep=True, noise=noise)
    B = X.shape[0]
    dt = X.dtype
    rows = tape["rows"]
    ref = X[:, -1]
    s = dt.type(1.0) / dt.type(N * B)
    g_rows = np.zeros((B, LOOKBACK + N, 5), dt)
    grads = {"inp_w": np.zeros_like(w["inp_w"]), "inp_b": np.zeros_like(w["inp_b"]),
             "out_w": np.zeros_like(w["out_w"])}
    if "int_w" in w:
        grads["int_w"] = np.zeros_like(w["int_w"])
        grads["int_b"] = np.zeros_like(w["int_b"])
    u = lambda k: rows[:, LOOKBACK - 1 + k, 4]           # u_k, k = -1 .. N-1

    def g_command(k):
        """d/du_k of the command-rate terms, UL/Functions.py:1405,1446 (scaled by s)."""
        g = -2 * alpha * (u(k - 1) - u(k)) * s
        if k + 1 <= N - 1:
            g = g + 2 * alpha * (u(k) - u(k + 1)) * s
        return g

    for m in range(N - 1, -1, -1):
        x = rows[:, LOOKBACK + m, :4]                    # x_{m+1}
        gx = np.zeros((B, 4), dt)
        gx[:, 0] = 2 * (x[:, 0] - ref) * s
        gx[:, 1] = s * ((x[:, 1] > P1_MAX).astype(dt) - (x[:, 1] < 0).astype(dt))
        gx[:, 2] = s * ((x[:, 2] > P2_MAX).astype(dt) - (x[:, 2] < 0).astype(dt))
        k = m + 1
        if k <= N - 1:                                   # row 9+k = [x_k, u_k] was fed back
            gu = g_rows[:, LOOKBACK - 1 + k, 4] + g_command(k)
            d_in = fnn_backward(w, tape["fnn"][k], gu[:, None], grads, width_dim)
            gx[:, 0] += d_in[:, 0]
            gx[:, 3] += d_in[:, 1]
            gx += g_rows[:, LOOKBACK - 1 + k, :4]
        t_min = max(0, LOOKBACK - 1 - m) if prune else 0
        g_rows[:, m:m + LOOKBACK] += lstm_window_backward(w, tape["lstm"][m], gx, t_min)
    grads["u0"] = g_rows[:, LOOKBACK - 1, 4] + g_command(0)
    return out, grads


def kink_margin(w, X, u0, Z, N, alpha, width_dim=1, noise=None) -> np.ndarray:
    """Per trajectory: the smallest distance of any argument of a non-smooth operation of the roll-out from its kink --
    ReLU pre-activations of the controller (UL/Functions.py:273-283), the Hardtanh saturation |v| = 1 (:287) and the four
    pressure-constraint ReLUs (:1411, :1449).  d loss/d u0 of a trajectory is discontinuous there, so an implementation
    that is accurate to eps on the VALUES may legitimately return a different per-trajectory gradient than the fp64
    arbiter exactly for the trajectories whose margin is below eps (and for no others)."""
    out, tape = mpc_loss_forward(w, X, u0, Z, N, alpha, width_dim, keep=True, noise=noise)
    rows = tape["rows"]
    B = X.shape[0]
    margin = np.full(B, np.inf)
    for m in range(N):
        x = rows[:, LOOKBACK + m, :4]
        for v in (x[:, 1], x[:, 2], x[:, 1] - P1_MAX, x[:, 2] - P2_MAX):
            margin = np.minimum(margin, np.abs(v))
        if m + 1 < N:
            _, pre, _, v = tape["fnn"][m + 1]
            for pk in pre:
                margin = np.minimum(margin, np.abs(pk).min(axis=1))
            margin = np.minimum(margin, np.abs(np.abs(v[:, 0]) - 1.0))
    return margin


# ----------------------------------------------------------------------------------------------
# torch restatement (autograd); used for cross-checks and for bench.py's cpu_baseline leg
# ----------------------------------------------------------------------------------------------
def mpc_loss_torch(tw: dict, X, u0, Z, N: int, alpha: float, width_dim: int = 1):
    """Same algorithm with torch ops.  ``tw`` = dict of torch tensors with the keys of
    ``weights_from_state_dicts``; controller tensors may require grad.  X [B,3], u0 [B], Z [B,10,5].
    Returns (loss, cost, command, error, prediction[B,N])."""
    import torch

    H = HIDDEN

    def surrogate(win):
        seq = win
        B = win.shape[0]
        for l in range(LAYERS):
            h = win.new_zeros(B, H)
            c = win.new_zeros(B, H)
            pre_in = seq @ tw["w_ih"][l].t()                       # [B,10,200]
            outs = []
            for t in range(LOOKBACK):
                gates = pre_in[:, t] + h @ tw["w_hh"][l].t()
                i, f, g, o = gates.split(H, dim=1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
                h = torch.sigmoid(o) * torch.tanh(c)
                outs.append(h)
            seq = torch.stack(outs, dim=1)
        return seq[:, -1] @ tw["fc_w"].t() + tw["fc_b"]

    def controller(inp):
        a = torch.relu(inp @ tw["inp_w"].t() + tw["inp_b"])
        for _ in range(width_dim - 1):
            a = torch.relu(a @ tw["int_w"].t() + tw["int_b"])
        return torch.clamp(a @ tw["out_w"].t(), -1.0, 1.0)

    def constraint(x):
        return (torch.relu(-x[:, 1]) + torch.relu(-x[:, 2]) + torch.relu(x[:, 1] - P1_MAX)
                + torch.relu(x[:, 2] - P2_MAX))

    ref = X[:, -1]
    win = torch.cat((Z[:, :-1], torch.cat((Z[:, -1, :4], u0[:, None]), dim=1)[:, None]), dim=1)
    u_prev, u_cur = Z[:, -2, 4], u0
    cost, cmd, err, pred = [], [], [], [u0]
    for m in range(N):
        x = surrogate(win)
        cmd.append(alpha * (u_prev - u_cur) ** 2)
        err.append((x[:, 0] - ref) ** 2)
        cost.append(err[-1] + cmd[-1] + constraint(x))
        if m + 1 < N:
            u_next = controller(torch.stack((x[:, 0], x[:, 3], ref), dim=1))
            win = torch.cat((win[:, 1:], torch.cat((x, u_next), dim=1)[:, None]), dim=1)
            pred.append(u_next[:, 0])
            u_prev, u_cur = u_cur, u_next[:, 0]
    cost_v = torch.stack(cost).sum(0) / N
    return (cost_v.mean(), cost_v, torch.stack(cmd).sum(0) / N, torch.stack(err).sum(0) / N,
            torch.stack(pred, dim=1))
