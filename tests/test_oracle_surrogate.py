"""Oracle pin for the surrogate-training path and the optimizer update (8f-2): the numpy
restatement in ``oracle/lstm_train_oracle.py`` against outputs of the UNMODIFIED reference
(``tests/golden/surrogate_train_cases.npz``, made by ``oracle/make_golden_surrogate.py``)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lstm_train_oracle as T  # noqa: E402
import mpc_loss_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def cases():
    return np.load(os.path.join(ROOT, "tests", "golden", "surrogate_train_cases.npz"))


def _weights(sd, dtype):
    return O.weights_from_state_dicts(sd, {"fc_inp.weight": np.zeros((50, 3)), "fc_inp.bias": np.zeros(50),
                                           "fc_out.weight": np.zeros((1, 50))}, dtype)


def _rel(a, b):
    return np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("tag,dtype,tol", [("f64", np.float64, 1e-11), ("f32", np.float32, 2e-5)])
def test_shipped_weights_step_matches_reference(cases, tag, dtype, tol):
    W = np.load(os.path.join(ROOT, "tests", "golden", "weights.npz"))
    sd = {k[5:]: W[k] for k in W.files if k.startswith("lstm/")}
    X, y = cases["shipped_b37/X"].astype(dtype), cases["shipped_b37/y"][:, 0, :].astype(dtype)
    loss, out, grads = T.lstm_mse_forward_backward(_weights(sd, dtype), X, y)
    assert abs(loss - float(cases[f"shipped_b37/{tag}/avg_loss"])) <= max(tol, 1e-6) * abs(loss)
    assert _rel(out, cases[f"shipped_b37/{tag}/out"]) <= max(tol, 1e-6)
    for k in T.GRAD_KEYS:
        assert _rel(grads[k], cases[f"shipped_b37/{tag}/grad/{k}"]) <= tol, k


def test_three_adamw_steps_from_fresh_weights_match_reference(cases):
    sd = {k: cases[f"fresh_b256x3/init/{k}"].astype(np.float64) for k in T.GRAD_KEYS}
    m = {k: np.zeros_like(v) for k, v in sd.items()}
    v2 = {k: np.zeros_like(v) for k, v in sd.items()}
    losses = []
    for b in range(3):
        X, y = cases[f"fresh_b256x3/X{b}"].astype(np.float64), cases[f"fresh_b256x3/y{b}"][:, 0, :].astype(np.float64)
        loss, _, grads = T.lstm_mse_forward_backward(_weights(sd, np.float64), X, y)
        losses.append(loss)
        for k in T.GRAD_KEYS:
            sd[k], m[k], v2[k] = T.adamw_step(sd[k], grads[k], m[k], v2[k], b + 1, lr=1e-3, weight_decay=0.0)
    assert abs(np.mean(losses) - float(cases["fresh_b256x3/f64/avg_loss"])) <= 1e-10
    for k in T.GRAD_KEYS:
        assert _rel(grads[k], cases[f"fresh_b256x3/f64/grad/{k}"]) <= 1e-6, k       # stored as float32
        delta_ref = cases[f"fresh_b256x3/f64/after/{k}"].astype(np.float64) - cases[f"fresh_b256x3/init/{k}"]
        delta = sd[k] - cases[f"fresh_b256x3/init/{k}"]
        assert np.abs(delta - delta_ref).max() <= 1e-4 * np.abs(delta_ref).max(), k   # 'after' stored as float32


def test_adamw_restatement_matches_torch(cases):
    p = cases["adamw_ctl/p0"].astype(np.float32)
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    for s in range(5):
        p, m, v = T.adamw_step(p, cases["adamw_ctl/grads"][s], m, v, s + 1, lr=np.float32(1e-3))
        p, m, v = p.astype(np.float32), m.astype(np.float32), v.astype(np.float32)
        assert _rel(p, cases[f"adamw_ctl/p{s + 1}"]) <= 1e-6
