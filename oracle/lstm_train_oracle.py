"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product; the product package never imports it.

CPU restatement of the reference's SURROGATE TRAINING step and of the optimizer
update both training paths use (8f-2); checker for ``fc_lstm_window_fwd`` / ``fc_lstm_window_bwd`` /
``fc_adamw_step``:

* ``lstm_mse_forward_backward``  restates one iteration of ``NeuralNetwork.train_model``
  (UL/Model_NN/Functions.py:520-569): ``output = LSTMModel.forward(X)`` (:313-340, identical to
  UL/Functions.py:353-379), ``loss = nn.MSELoss()(output, y.squeeze())`` (UL/Model_NN/Main.py:227), and what
  ``loss.backward()`` (:560) leaves in ``.grad`` of the eight surrogate parameters -- a hand-derived reverse
  sweep with the weight gradients (no autograd).
* ``adamw_step``                 restates ``torch.optim.AdamW.step`` (decoupled weight decay, bias-corrected
  moments; UL/Main.py:195 with the default ``weight_decay=0.01``, UL/Model_NN/Main.py:230 with 0.0).

Parity pin: ``oracle/make_golden_surrogate.py`` runs the unmodified reference ``LSTMModel`` + ``nn.MSELoss`` +
``torch.optim.AdamW`` (fp32 and fp64) and commits ``tests/golden/surrogate_train_cases.npz``;
``tests/test_oracle_surrogate.py`` checks both functions against it.
"""
from __future__ import annotations

import numpy as np

from mpc_loss_oracle import HIDDEN, LAYERS, LOOKBACK, lstm_window_forward

GRAD_KEYS = tuple(f"lstm.weight_{k}_l{l}" for l in range(LAYERS) for k in ("ih", "hh")) + ("fc.weight", "fc.bias")


def lstm_mse_forward_backward(w: dict, X: np.ndarray, y: np.ndarray, d_out: np.ndarray | None = None):
    """X [B,10,5], y [B,4] -> (loss, out [B,4], grads keyed like the ``state_dict``).
    With ``d_out`` given, it replaces the MSE gradient (arbitrary upstream gradient of ``out``)."""
    B = X.shape[0]
    H, L = HIDDEN, LOOKBACK
    dt = X.dtype
    out, (cells, h_last) = lstm_window_forward(w, X, keep=True)
    diff = out - y
    loss = float(np.mean(diff.astype(np.float64) ** 2))
    gx = (2.0 / diff.size) * diff if d_out is None else np.asarray(d_out, dt)
    grads = {"fc.weight": gx.T @ h_last, "fc.bias": gx.sum(axis=0)}
    # hidden sequences per layer (input of the layer above / recurrent input of the own layer)
    hs = [np.stack([o * np.tanh(c) for (_, _, _, o, _, c) in cells[l]], axis=1) for l in range(LAYERS)]   # [B,10,50]
    d_seq = np.zeros((B, L, H), dt)
    d_seq[:, L - 1, :] = gx @ w["fc_w"]
    for l in reversed(range(LAYERS)):
        w_ih, w_hh = w["w_ih"][l], w["w_hh"][l]
        inp = X if l == 0 else hs[l - 1]
        g_ih = np.zeros_like(w_ih)
        g_hh = np.zeros_like(w_hh)
        d_in = np.zeros((B, L, w_ih.shape[1]), dt)
        dh_rec = np.zeros((B, H), dt)
        dc = np.zeros((B, H), dt)
        for t in range(L - 1, -1, -1):
            i, f, g, o, c_prev, c = cells[l][t]
            dh = d_seq[:, t, :] + dh_rec
            tc = np.tanh(c)
            dct = dc + dh * o * (1.0 - tc * tc)
            dgate = np.concatenate([dct * g * i * (1.0 - i), dct * c_prev * f * (1.0 - f), dct * i * (1.0 - g * g),
                                    dh * tc * o * (1.0 - o)], axis=1)
            dc = dct * f
            g_ih += dgate.T @ inp[:, t, :]
            if t > 0:
                g_hh += dgate.T @ hs[l][:, t - 1, :]
            d_in[:, t, :] = dgate @ w_ih
            dh_rec = dgate @ w_hh
        grads[f"lstm.weight_ih_l{l}"] = g_ih
        grads[f"lstm.weight_hh_l{l}"] = g_hh
        d_seq = d_in
    return loss, out, grads


def adamw_step(p, g, m, v, step: int, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2):
    """One ``torch.optim.AdamW`` update (``step`` = 1 for the first); returns (p, m, v)."""
    p = p * (1.0 - lr * weight_decay)
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = np.sqrt(v) / np.sqrt(bc2) + eps
    return p - (lr / bc1) * (m / denom), m, v
