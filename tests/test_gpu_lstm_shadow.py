"""GPU: LSTM shadow roll-out of the closed loop -- `fc_lstm_shadow_rollout`, the forward-only mode of
the pair kernel -- against the oracle restatement of `NeuralNetwork.simulator_make_step` as driven by
`NeuralNetwork.loop` (UL/Functions.py:969-1011, :1196-1231), against the stock nn.LSTM stepping the same recursion,
and through the `NeuralNetwork.loop` API."""
import numpy as np
import pytest
import torch

import mpc_loss_oracle as O
from conftest import rel_max, state_dicts

pytestmark = pytest.mark.gpu
fb = pytest.importorskip("forging_control_b200")


def _sim(golden_weights, dev):
    lstm, fnn = state_dicts(golden_weights, "c0")
    sim = fb.LSTMModel(5, 50, 4, 3)
    sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()})
    return sim.to(dev), lstm, fnn


@pytest.mark.parametrize("B,T", [(3, 30), (300, 12), (1000, 5)])
def test_shadow_rollout_matches_oracle(golden_weights, B, T):
    dev = torch.device("cuda:0")
    sim, lstm, fnn = _sim(golden_weights, dev)
    rng = np.random.default_rng(B + T)
    row0 = rng.uniform(-1, 1, (B, 5)).astype(np.float32)
    u = rng.uniform(-1, 1, (B, T)).astype(np.float32)
    ratio = np.array([1.0, 0.95, 1.05, 1.0])
    y = fb.lstm_shadow_native(sim, row0, u, ratio)
    assert y.shape == (B, T, 4) and y.dtype == np.float32
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    ref = O.lstm_shadow_rollout(w, row0.astype(np.float64), u.astype(np.float64), ratio)
    assert rel_max(y, ref) < 1e-5          # same bar as the training path (fp32 roll-out vs fp64 arbiter)


def test_shadow_rollout_matches_stock_lstm_stepping(golden_weights):
    """The recursion exactly as the reference runs it: one LSTMModel.forward per step on a sliding window."""
    dev = torch.device("cuda:0")
    sim, _, _ = _sim(golden_weights, dev)
    rng = np.random.default_rng(7)
    B, T = 64, 20
    row0 = rng.uniform(-1, 1, (B, 5)).astype(np.float32)
    u = rng.uniform(-1, 1, (B, T)).astype(np.float32)
    ratio = np.ones(4)
    y = fb.lstm_shadow_native(sim, row0, u, ratio)
    # stock stepping on the CPU: cuDNN's LSTM uses TF32 tensor-core math on this GPU (~1e-3), the ATen CPU LSTM is fp32
    import copy
    ref_sim = copy.deepcopy(sim).cpu()
    cpu = torch.device("cpu")
    window = torch.tensor(row0)[:, None, :].repeat(1, 10, 1)
    with torch.no_grad():
        for m in range(T):
            out = ref_sim(window, cpu)
            assert rel_max(y[:, m], out.double().numpy()) < 1e-5, m
            nxt = torch.cat((out, torch.tensor(u[:, m + 1:m + 2] if m + 1 < T else np.zeros((B, 1), np.float32))), dim=1)
            window = torch.cat((window[:, 1:], nxt[:, None, :]), dim=1)


def test_loop_api_returns_lstm_shadow(golden_weights):
    """NeuralNetwork.loop with a surrogate: results_LSTM = the shadow prediction in physical units, first column the
    initial state (UL/Functions.py:1196-1231)."""
    from sklearn.preprocessing import MaxAbsScaler
    dev = torch.device("cuda:0")
    sim, lstm, fnn = _sim(golden_weights, dev)
    ctl = fb.FNNModel(3, 50, 1, 1)
    ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()})
    mk = lambda s: MaxAbsScaler().fit(np.asarray(s)[None, :])
    scalers = {"input": mk(golden_weights["scale/scaler_input"]), "output": mk(golden_weights["scale/scaler_output"]),
               "y_dot": mk(golden_weights["scale/scaler_input"][:1])}
    s_in = np.array([0.917128, 32e6, 32e6, 0.382353, 0.382353])
    model_scalers = {"input": mk(s_in), "output": mk(s_in[:4])}
    init = {"y": 0.0, "y_dot": 0.0, "p1": 2156275.6006012624, "p2": 2961363.827545376, "z": 0.0}
    _, res, res_lstm, _, _ = fb.NeuralNetwork.loop(
        N_traj=2, T_traj=40, Ts=1e-3, controller=ctl.to(dev), simulator=None, simulator_LSTM=sim, init_state=init,
        scalers=scalers, model_scalers=model_scalers, bias_work=300, bias_return=20 ** 6, lookback=10, bar_title="NN",
        process_std=np.zeros(5), meas_std=np.zeros(5))
    assert set(res_lstm) == {"y_dot", "p1", "p2", "z"} and res_lstm["p1"].shape == (2, 41)
    assert np.allclose(res_lstm["p1"][:, 0], init["p1"]) and np.allclose(res_lstm["z"][:, 0], 0.0)
    # oracle on the same applied commands
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    x0 = np.tile(np.array([[0.0, 0.0, init["p1"], init["p2"], 0.0]]), (2, 1))
    row0 = np.concatenate((x0[:, 1:5], res["u"][:, 0:1]), axis=1) / s_in
    ref = O.lstm_shadow_rollout(w, row0, res["u"] / s_in[4], np.ones(4)) * s_in[:4]
    for i, n in enumerate(("y_dot", "p1", "p2", "z")):
        assert np.abs(res_lstm[n][:, 1:] - ref[:, :, i]).max() <= 2e-5 * max(np.abs(ref[:, :, i]).max(), s_in[i] * 1e-2), n
