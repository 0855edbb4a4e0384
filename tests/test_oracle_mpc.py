"""CPU: the oracle restatement of the MPC loss against the golden vectors produced by the
unmodified reference (oracle/make_golden.py).  Pins the oracle before it is used as the checker."""
import numpy as np
import pytest
import torch

import mpc_loss_oracle as O
from conftest import case_names, rel_max, state_dicts

ALPHA = 20.0   # UL/Main.py:192
CASES = ["n1_b3", "n2_b5", "n5_b16", "n10_b15", "n10_b33_init", "n10_b40_trace", "n25_b9", "n12_b130_trace",
         "n6_b7_w2", "n10_b12_wide"]


def _total_grads(w, X, g, width_dim):
    """reference .grad = our roll-out gradients + the u0 path through model(X) (Functions.py:643)."""
    grads = {k: np.zeros_like(v) for k, v in g.items() if k != "u0"}
    kept = O.fnn_forward(w, X, width_dim, keep=True)[1]
    O.fnn_backward(w, kept, g["u0"][:, None], grads, width_dim)
    return {k: grads[k] + g[k] for k in grads}


def test_fixture_is_complete(golden_cases):
    assert case_names(golden_cases) == sorted(CASES)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_oracle_matches_reference(golden_cases, golden_weights, name, prec):
    C = golden_cases
    N, B, wd = (int(v) for v in C[f"{name}/meta"])
    dt = np.float64 if prec == "f64" else np.float32
    tol = 1e-12 if prec == "f64" else 1e-5          # fp32: two fp32 evaluation orders of the same maths
    lstm, fnn = state_dicts(golden_weights, str(C[f"{name}/ctl"]))
    w = O.weights_from_state_dicts(lstm, fnn, dt)
    X, Z = C[f"{name}/X"].astype(dt), C[f"{name}/Z"].astype(dt)
    u0 = C[f"{name}/{prec}/u0"].astype(dt)
    # the controller forward itself
    assert rel_max(O.fnn_forward(w, X, wd)[:, 0], u0) < (1e-13 if prec == "f64" else 1e-6)
    out, g = O.mpc_loss_forward_backward(w, X, u0, Z, N, dt(ALPHA), wd)
    ref = lambda k: C[f"{name}/{prec}/{k}"]
    assert abs(out["loss"] - ref("loss")) / abs(ref("loss")) < tol
    for k in ("cost", "command", "error", "prediction"):
        assert rel_max(out[k], ref(k)) < tol, k
    tot = _total_grads(w, X, g, wd)
    for ours, theirs in (("inp_w", "fc_inp.weight"), ("inp_b", "fc_inp.bias"), ("out_w", "fc_out.weight")):
        assert rel_max(tot[ours], ref("grad/" + theirs)) < tol, theirs
    if wd > 1:
        assert rel_max(tot["int_w"], ref("grad/fc_int.weight")) < tol
        assert rel_max(tot["int_b"], ref("grad/fc_int.bias")) < tol
    else:
        assert f"{name}/{prec}/grad/fc_int.weight" not in C.files     # stays None in the reference


@pytest.mark.parametrize("name", ["n5_b16", "n12_b130_trace"])
def test_pruned_sweep_equals_full_sweep(golden_cases, golden_weights, name):
    C = golden_cases
    N, B, wd = (int(v) for v in C[f"{name}/meta"])
    lstm, fnn = state_dicts(golden_weights, str(C[f"{name}/ctl"]))
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    X, Z, u0 = C[f"{name}/X"].astype(np.float64), C[f"{name}/Z"].astype(np.float64), C[f"{name}/f64/u0"]
    _, g1 = O.mpc_loss_forward_backward(w, X, u0, Z, N, ALPHA, wd, prune=True)
    _, g2 = O.mpc_loss_forward_backward(w, X, u0, Z, N, ALPHA, wd, prune=False)
    for k in g1:
        assert np.array_equal(g1[k], g2[k]), k


@pytest.mark.parametrize("name", ["n2_b5", "n10_b15"])
def test_torch_restatement_matches_reference(golden_cases, golden_weights, name):
    C = golden_cases
    N, B, wd = (int(v) for v in C[f"{name}/meta"])
    lstm, fnn = state_dicts(golden_weights, str(C[f"{name}/ctl"]))
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    tw = {k: ([torch.tensor(a) for a in v] if isinstance(v, list) else torch.tensor(v)) for k, v in w.items()}
    for k in ("inp_w", "inp_b", "out_w"):
        tw[k].requires_grad_()
    X = torch.tensor(C[f"{name}/X"].astype(np.float64))
    Z = torch.tensor(C[f"{name}/Z"].astype(np.float64))
    u0 = torch.clamp(torch.relu(X @ tw["inp_w"].t() + tw["inp_b"]) @ tw["out_w"].t(), -1, 1)[:, 0]
    loss, cost, cmd, err, pred = O.mpc_loss_torch(tw, X, u0, Z, N, ALPHA, wd)
    loss.backward()
    assert abs(loss.item() - C[f"{name}/f64/loss"]) / abs(C[f"{name}/f64/loss"]) < 1e-12
    assert rel_max(pred.detach().numpy(), C[f"{name}/f64/prediction"]) < 1e-12
    assert rel_max(tw["inp_w"].grad.numpy(), C[f"{name}/f64/grad/fc_inp.weight"]) < 1e-11
    assert rel_max(tw["out_w"].grad.numpy(), C[f"{name}/f64/grad/fc_out.weight"]) < 1e-11


@pytest.mark.reference
def test_oracle_against_live_reference():
    """Only where /root/reference is mounted: run the reference itself on fresh inputs."""
    import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference mount absent")
    import os
    F = ref_shim.load_reference_functions()
    lstm_sd = torch.load(os.path.join(ref_shim.MNN_DIR, "results/model_NN.pt"), map_location="cpu")
    fnn_sd = torch.load(os.path.join(ref_shim.UL_DIR, "results/NN_controller_N_10_5.pt"), map_location="cpu")
    torch.set_default_dtype(torch.float64)
    try:
        sim = F.LSTMModel(5, 50, 4, 3); sim.load_state_dict(lstm_sd); sim = sim.double()
        ctl = F.FNNModel(3, 50, 1, 1, torch.nn.ReLU, bias=True); ctl.load_state_dict(fnn_sd); ctl = ctl.double()
        g = torch.Generator().manual_seed(7)
        X = torch.rand(11, 3, generator=g, dtype=torch.float64) * 2 - 1
        Z = torch.rand(11, 10, 5, generator=g, dtype=torch.float64) * 2 - 1
        u0 = ctl(X)
        loss, feats = F.MPCLoss(7, 20.0)(sim, ctl, X, u0, Z, "cpu")
        loss.backward()
    finally:
        torch.set_default_dtype(torch.float32)
    w = O.weights_from_state_dicts(lstm_sd, fnn_sd, np.float64)
    out, gr = O.mpc_loss_forward_backward(w, X.numpy(), u0.detach().numpy()[:, 0], Z.numpy(), 7, 20.0)
    assert abs(out["loss"] - loss.item()) / abs(loss.item()) < 1e-12
    tot = _total_grads(w, X.numpy(), gr, 1)
    assert rel_max(tot["inp_w"], ctl.fc_inp.weight.grad.numpy()) < 1e-11
