"""GPU (-m gpu): ``forging_control_b200.install`` on the REAL reference module (the staged, unmodified
``Unsupervised Learning/Functions.py`` under oracle/_ref, or the read-only mount), and the reference's own
``NeuralNetwork.train_loop`` (Functions.py:825-923) driven through the swapped classes on the device -- against the
same two epochs run by the untouched reference on CPU tensors."""
import contextlib
import copy
import os
import sys

import numpy as np
import pytest
import torch

from conftest import REPO, state_dicts

pytestmark = pytest.mark.gpu


def _reference_module():
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import stage_reference
    if stage_reference.staged():
        os.environ["FORGING_REFERENCE_ROOT"] = stage_reference.STAGE_ROOT
    elif not os.path.isfile(os.path.join(stage_reference.SOURCE_ROOT, stage_reference.FILES[0])):
        pytest.skip("reference neither staged (oracle/_ref) nor mounted")
    import importlib
    import ref_shim
    importlib.reload(ref_shim)          # picks up FORGING_REFERENCE_ROOT
    R = ref_shim.load_reference_functions(fresh=True)

    @contextlib.contextmanager
    def no_bar(*a, **k):                # alive_progress (third party, absent from the image) is only the progress bar
        yield lambda *aa, **kk: None
    R.alive_bar = no_bar
    return R


def _loaders(n=60, batch=15, seed=5):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, 3, generator=g) * 2 - 1
    y = torch.rand(n, 1, generator=g) * 2 - 1
    Z = torch.rand(n, 10, 5, generator=g) * 2 - 1
    ds = torch.utils.data.TensorDataset(X, y, Z)
    mk = lambda: torch.utils.data.DataLoader(ds, batch_size=batch, shuffle=False)
    return mk(), mk()


def test_install_on_the_real_reference_module_and_its_train_loop(golden_weights):
    import forging_control_b200 as fb
    R = _reference_module()
    lstm, fnn = state_dicts(golden_weights, "init")
    N, alpha, epochs = 10, 20.0, 2

    def build(mod, dev):
        sim = mod.LSTMModel(5, 50, 4, 3)
        sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()}, strict=True)      # Main.py:154,168
        ctl = mod.FNNModel(3, 50, 1, 1)                                                      # Main.py:188
        ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()}, strict=True)
        sim, ctl = sim.to(dev), ctl.to(dev)                                                  # Main.py:328-329
        lf = mod.MPCLoss(prediction_horizon=N, alpha=alpha)                                  # Main.py:192
        opt = torch.optim.AdamW(ctl.parameters(), lr=1e-3)                                   # Main.py:195
        return sim, ctl, lf, opt

    # 1. the untouched reference on CPU tensors
    tr, va = _loaders()
    sim, ctl, lf, opt = build(R, torch.device("cpu"))
    ref_cls = (R.FNNModel, R.LSTMModel, R.MPCLoss)
    _, t_ref, v_ref, _, feats_ref = R.NeuralNetwork.train_loop(ctl, sim, tr, va, lf, opt, epochs, torch.device("cpu"))
    w_ref = copy.deepcopy({k: v.detach().clone() for k, v in ctl.state_dict().items()})

    # 2. install() swaps the hot-path classes into THIS module object; Main.py's calls then reach the CUDA path
    fb.install(R)
    assert R.FNNModel is fb.FNNModel and R.LSTMModel is fb.LSTMModel and R.MPCLoss is fb.MPCLoss
    assert (R.FNNModel, R.LSTMModel, R.MPCLoss) != ref_cls
    assert R.NeuralNetwork.train_model is fb.NeuralNetwork.train_model and R.NeuralNetwork.loop is fb.NeuralNetwork.loop
    dev = torch.device("cuda:0")
    tr, va = _loaders()
    sim, ctl, lf, opt = build(R, dev)
    assert isinstance(ctl, fb.FNNModel) and isinstance(lf, fb.MPCLoss)
    out_ctl, t_new, v_new, _, feats = R.NeuralNetwork.train_loop(ctl, sim, tr, va, lf, opt, epochs, dev)   # the reference's own loop
    assert out_ctl is ctl
    assert feats["loss"].shape == feats_ref["loss"].shape and feats["prediction"].shape == feats_ref["prediction"].shape
    # same optimisation trajectory: epoch losses and the trained weights (eight AdamW steps from the same start)
    assert np.allclose(t_new, t_ref, rtol=2e-5), (t_new, t_ref)
    assert np.allclose(v_new, v_ref, rtol=2e-5, atol=1e-7), (v_new, v_ref)
    for k, v in ctl.state_dict().items():
        a, b = v.detach().cpu().numpy(), w_ref[k].numpy()
        assert np.abs(a - b).max() <= 2e-5 * max(np.abs(b).max(), 1e-3), k
    assert np.abs(feats["loss"].cpu().numpy() - feats_ref["loss"].detach().numpy()).max() <= 2e-5 * np.abs(feats_ref["loss"].detach().numpy()).max()
