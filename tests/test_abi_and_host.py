"""CPU: the C-ABI library loads and exports every symbol include/forging_b200.h declares (no compute
calls without a GPU), and the host-side mirror of the reference interface behaves like the reference."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import forging_control_b200 as fb
from forging_control_b200 import _native
from conftest import REPO, state_dicts


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "forging_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(fc_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(_native.EXPORTS)
    _native.build()
    L = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _native.lib().fc_version() == 100
    assert _native.lib().fc_pack_floats() == 123456 + 120264 + 120264


def test_state_dict_layout_loads_shipped_checkpoints_strictly(golden_weights):
    lstm, fnn = state_dicts(golden_weights, "c0")
    sim = fb.LSTMModel(5, 50, 4, 3)
    sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()}, strict=True)
    ctl = fb.FNNModel(3, 50, 1, 1, torch.nn.ReLU, bias=True)
    ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()}, strict=True)
    assert list(ctl.state_dict()) == ["fc_inp.weight", "fc_inp.bias", "fc_int.weight", "fc_int.bias", "fc_out.weight"]
    assert [tuple(v.shape) for v in sim.state_dict().values()] == \
        [(200, 5), (200, 50), (200, 50), (200, 50), (200, 50), (200, 50), (4, 50), (4,)]


def test_models_match_reference_outputs_on_cpu(golden_cases, golden_weights):
    """model(X) (Functions.py:643) reproduces the reference's recorded u0; LSTMModel.forward the oracle."""
    import mpc_loss_oracle as O
    C = golden_cases
    lstm, fnn = state_dicts(golden_weights, "c0")
    ctl = fb.FNNModel(3, 50, 1, 1)
    ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()})
    u0 = ctl(torch.tensor(C["n10_b15/X"])).detach().numpy()[:, 0]
    assert np.abs(u0 - C["n10_b15/f32/u0"]).max() < 1e-6
    sim = fb.LSTMModel(5, 50, 4, 3)
    sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()})
    y = sim(torch.tensor(C["n10_b15/Z"]), "cpu").detach().numpy()
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    assert np.abs(y - O.lstm_window_forward(w, C["n10_b15/Z"].astype(np.float64))).max() < 1e-5


def test_no_cpu_fallback(golden_weights):
    lstm, fnn = state_dicts(golden_weights, "c0")
    sim, ctl = fb.LSTMModel(5, 50, 4, 3), fb.FNNModel(3, 50, 1, 1)
    X, Z = torch.zeros(4, 3), torch.zeros(4, 10, 5)
    with pytest.raises(RuntimeError, match="CUDA"):
        fb.MPCLoss(10, 20.0)(sim, ctl, X, ctl(X), Z, "cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        fb.MPCLoss(10, 20.0)(sim, ctl, X, ctl(X), Z, "cpu", enable_noise=True)
    with pytest.raises(NotImplementedError):
        fb.MPCLoss(10, 20.0)(fb.LSTMModel(5, 64, 4, 3), ctl, X, ctl(X), Z, "cpu")
    with pytest.raises(RuntimeError, match="CUDA"):          # width_dim > 1 is supported (one-tile kernel); still no CPU path
        fb.MPCLoss(10, 20.0)(sim, fb.FNNModel(3, 50, 1, 2), X, ctl(X), Z, "cpu")
    with pytest.raises(NotImplementedError):
        fb.MPCLoss(10, 20.0)(sim, fb.FNNModel(3, 64, 1, 1), X, ctl(X), Z, "cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        fb.closed_loop_device(ctl, torch.zeros(2, 5), torch.zeros(3, 2), 1e-3, [1, 1, 1], [1])


def test_tvp_fun_and_reference_table(golden_trace):
    t = golden_trace["time"][:, 0]
    mine = np.array([fb.NeuralNetwork.tvp_fun(float(tt), 0.3, 300, 20 ** 6) for tt in t])
    assert np.array_equal(mine, golden_trace["tvp_fun"])
    tab = fb.tvp_reference_table(2, 300, 1e-3, 0.3, 300, 20 ** 6)
    assert np.array_equal(tab.reshape(-1), golden_trace["tvp"][:, 0])


def test_nn_make_step_matches_reference(golden_trace, golden_weights):
    from sklearn.preprocessing import MaxAbsScaler
    _, fnn = state_dicts(golden_weights, "c0")
    ctl = fb.FNNModel(3, 50, 1, 1)
    ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()})
    mk = lambda s: MaxAbsScaler().fit(np.asarray(s)[None, :])
    scalers = {"input": mk(golden_weights["scale/scaler_input"]), "output": mk(golden_weights["scale/scaler_output"]),
               "y_dot": mk(golden_weights["scale/scaler_input"][:1])}
    g = golden_trace
    for k in (0, 1, 150, 299, 300, 451, 599):
        inp = np.array([[g["meas_prev"][k, 1], g["meas_prev"][k, 4], g["tvp"][k, 0]]])
        u, sol, _ = fb.FeasibilityRecovery.NN_make_step(inp, ctl, scalers, None, None, None)
        assert abs(u.item() - g["nn_make_step_u"][k]) < 1e-7 and sol == 0.0


def test_get_scaler():
    assert type(fb.Data.get_scaler("MaxAbs")).__name__ == "MaxAbsScaler"
    assert type(fb.Data.get_scaler("robust")).__name__ == "RobustScaler"
    with pytest.raises(ValueError):
        fb.Data.get_scaler("quantile")


def test_shard_bounds_cover_the_batch():
    for total, ws in ((10, 3), (4194304, 8), (7, 8), (120, 1)):
        spans = [fb.shard_bounds(total, ws, r) for r in range(ws)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_install_swaps_hot_path_into_a_reference_like_module():
    import types
    ref = types.SimpleNamespace(NeuralNetwork=type("NeuralNetwork", (), {}), FeasibilityRecovery=type("FR", (), {}))
    fb.install(ref)
    assert ref.MPCLoss is fb.MPCLoss and ref.FNNModel is fb.FNNModel and ref.LSTMModel is fb.LSTMModel
    assert ref.NeuralNetwork.loop is fb.NeuralNetwork.loop


def test_lstm_model_on_cpu_is_the_stock_module_and_device_adamw_refuses_cpu():
    """The surrogate kernels serve CUDA tensors only: CPU evaluation (Main.py:347-361 after ``.cpu()``) stays stock
    ``nn.LSTM``; the device optimizer has no CPU path."""
    import torch
    import forging_control_b200 as fb
    from forging_control_b200 import surrogate as S
    m = fb.LSTMModel(5, 50, 4, 3)
    x = torch.zeros(2, 10, 5)
    assert not S.supported(m, x)
    assert m(x, "cpu").shape == (2, 4)
    with pytest.raises(NotImplementedError):
        S.lstm_window(m, x)
    opt = S.DeviceAdamW(m.parameters(), lr=1e-3, weight_decay=0.0)
    m(x, "cpu").sum().backward()
    with pytest.raises(RuntimeError):
        opt.step()
    with pytest.raises(NotImplementedError):
        S.DeviceAdamW(m.parameters(), amsgrad=True)


def test_install_surrogate_swaps_the_training_step_and_host_loop_matches_the_reference_on_cpu():
    """``install_surrogate`` on a reference-like module; the mirrored ``train_model`` / ``validate_model`` /
    ``train_loop`` (Model_NN/Functions.py:520-612, 754-822) reproduce the golden epoch of the unmodified reference when
    fed CPU tensors (stock ``nn.LSTM`` carries the arithmetic there, so this pins the HOST logic: loop order,
    ``y.squeeze()``, averaging over batches, optimizer placement)."""
    import types
    import torch
    from forging_control_b200 import surrogate as S
    ref = types.SimpleNamespace(NeuralNetwork=type("NeuralNetwork", (), {}))
    fb.install_surrogate(ref)
    assert ref.LSTMModel is fb.LSTMModel and ref.NeuralNetwork.train_model is S.SurrogateNeuralNetwork.train_model
    C = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "surrogate_train_cases.npz"))
    keys = [k.split("/")[-1] for k in C.files if k.startswith("fresh_b256x3/init/")]
    m = fb.LSTMModel(5, 50, 4, 3)
    m.load_state_dict({k: torch.tensor(C[f"fresh_b256x3/init/{k}"]) for k in keys}, strict=True)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=0.0)
    loader = [(torch.tensor(C[f"fresh_b256x3/X{b}"]), torch.tensor(C[f"fresh_b256x3/y{b}"])) for b in range(3)]
    avg = ref.NeuralNetwork.train_model(loader, m, torch.nn.MSELoss(), opt, "cpu")
    assert abs(avg - float(C["fresh_b256x3/f32/avg_loss"])) <= 1e-6 * abs(avg)
    for k, p in m.named_parameters():
        init = C[f"fresh_b256x3/init/{k}"]
        d_ref = C[f"fresh_b256x3/f32/after/{k}"] - init
        assert np.abs((p.detach().numpy() - init) - d_ref).max() <= 1e-3 * np.abs(d_ref).max(), k
    v = S.SurrogateNeuralNetwork.validate_model(loader, m, torch.nn.MSELoss(), "cpu")
    assert np.isfinite(v) and v < avg                     # three AdamW steps later the same batches fit better
    _, vt, vv, _ = S.SurrogateNeuralNetwork.train_loop(m, loader[:1], loader[1:2], torch.nn.MSELoss(), opt, 2, "cpu")
    assert len(vt) == 2 and len(vv) == 2 and vt[1] < vt[0]


@pytest.mark.reference
def test_install_on_the_real_reference_module():
    """install() on the unmodified reference ``Functions`` module (mount or staged copy): the names Main.py:19 imports
    resolve to the CUDA-path classes afterwards, the reference's own train_loop / test / Data stay in place."""
    import os
    import sys
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference not mounted")
    R = ref_shim.load_reference_functions(fresh=True)
    before = (R.FNNModel, R.LSTMModel, R.MPCLoss, R.NeuralNetwork.train_loop, R.Data)
    fb.install(R)
    assert R.FNNModel is fb.FNNModel and R.LSTMModel is fb.LSTMModel and R.MPCLoss is fb.MPCLoss
    assert R.FNNModel is not before[0] and R.MPCLoss is not before[2]
    assert R.NeuralNetwork.train_loop is before[3] and R.Data is before[4]           # untouched
    assert R.NeuralNetwork.train_model is fb.NeuralNetwork.train_model
    assert R.FeasibilityRecovery.NN_make_step is fb.FeasibilityRecovery.NN_make_step
    # the shipped checkpoints load strictly into the swapped classes (Main.py:154-168)
    import torch
    sim = R.LSTMModel(5, 50, 4, 3)
    sim.load_state_dict(torch.load(os.path.join(ref_shim.MNN_DIR, "results", "model_NN.pt"), map_location="cpu"), strict=True)
    ctl = R.FNNModel(3, 50, 1, 1)
    ctl.load_state_dict(torch.load(os.path.join(ref_shim.UL_DIR, "results", "NN_controller_N_10_0.pt"), map_location="cpu"), strict=True)
