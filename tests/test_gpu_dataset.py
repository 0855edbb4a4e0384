"""GPU: device-side sample construction (fc_build_windows / DeviceSequenceLoader) is bit-exact with the
reference's SequenceDataset items (golden fixture) and with the oracle at large sizes; the loader feeds train_model."""
import os

import numpy as np
import pytest
import torch

import dataset_oracle as D
from conftest import REPO, state_dicts

pytestmark = pytest.mark.gpu
fb = pytest.importorskip("forging_control_b200")


def _tables():
    G = np.load(os.path.join(REPO, "tests", "golden", "sequence_dataset.npz"))
    cols, tab = list(G["columns"]), G["table"]
    pick = lambda names: np.ascontiguousarray(tab[:, [cols.index(c) for c in names]])
    return G, pick(["y_dot", "z", "ref"]), pick(["u"]), pick(["y_dot", "p1", "p2", "z", "u"])


def test_loader_reproduces_reference_items_bit_exactly():
    G, X, y, Z = _tables()
    loader = fb.DeviceSequenceLoader(X, y, Z, int(G["t_traj"]), batch_size=16, shuffle=False, lookback=int(G["lookback"]))
    assert len(loader) == 5
    got = [tuple(t.cpu().numpy() for t in batch) for batch in loader]
    assert np.array_equal(np.concatenate([g[0] for g in got]), G["X"])
    assert np.array_equal(np.concatenate([g[1] for g in got]), G["y"])
    assert np.array_equal(np.concatenate([g[2] for g in got]), G["Z"])
    assert got[-1][0].shape == (75 - 64, 3)                  # ragged last batch


def test_large_random_gather_matches_oracle():
    rng = np.random.default_rng(0)
    n_traj, t_traj = 700, 301
    M = n_traj * t_traj
    X, y, Z = (rng.standard_normal((M, w)).astype(np.float32) for w in (3, 1, 5))
    idx = rng.integers(0, M, 200_000)
    dev = torch.device("cuda:0")
    gx, gy, gz = fb.build_windows(torch.tensor(X).to(dev), torch.tensor(y).to(dev), torch.tensor(Z).to(dev), t_traj,
                                  torch.tensor(idx, dtype=torch.int64).to(dev), 10)
    rx, ry, rz = D.sequence_items(X, y, Z, t_traj, idx, 10)
    assert np.array_equal(gx.cpu().numpy(), rx) and np.array_equal(gy.cpu().numpy(), ry) and np.array_equal(gz.cpu().numpy(), rz)


def test_shuffled_loader_is_a_permutation_and_feeds_train_model(golden_weights):
    G, X, y, Z = _tables()
    torch.manual_seed(0)
    loader = fb.DeviceSequenceLoader(X, y, Z, int(G["t_traj"]), batch_size=32, shuffle=True)
    xs = np.concatenate([b[0].cpu().numpy() for b in loader])
    assert xs.shape == G["X"].shape and np.array_equal(np.sort(xs, axis=0), np.sort(G["X"], axis=0))
    dev = torch.device("cuda:0")
    lstm, fnn = state_dicts(golden_weights, "c0")
    sim = fb.LSTMModel(5, 50, 4, 3); sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()})
    ctl = fb.FNNModel(3, 50, 1, 1); ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()})
    sim, ctl = sim.to(dev), ctl.to(dev)
    opt = torch.optim.AdamW(ctl.parameters(), lr=1e-3)
    avg, feats = fb.NeuralNetwork.train_model(loader, sim, ctl, fb.MPCLoss(5, 20.0), opt, dev)
    assert np.isfinite(avg) and feats["loss"].shape == (75,) and feats["prediction"].shape == (75 * 5,)
