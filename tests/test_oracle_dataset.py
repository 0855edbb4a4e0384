"""CPU: the oracle restatement of the reference's sample construction (SequenceDataset.__getitem__,
UL/Functions.py:109-132, per-trajectory slices of Data.get_individual_dataset, :479-516) against items produced by
the unmodified reference classes (tests/golden/sequence_dataset.npz, oracle/make_golden.py)."""
import os

import numpy as np

import dataset_oracle as D
from conftest import REPO


def _tables():
    G = np.load(os.path.join(REPO, "tests", "golden", "sequence_dataset.npz"))
    cols, tab = list(G["columns"]), G["table"]
    pick = lambda names: np.ascontiguousarray(tab[:, [cols.index(c) for c in names]])
    return G, pick(["y_dot", "z", "ref"]), pick(["u"]), pick(["y_dot", "p1", "p2", "z", "u"])


def test_oracle_items_equal_reference_items_bit_exactly():
    G, X, y, Z = _tables()
    x2, y2, z2 = D.sequence_items(X, y, Z, int(G["t_traj"]), np.arange(len(X)), int(G["lookback"]))
    assert np.array_equal(x2, G["X"]) and np.array_equal(y2, G["y"]) and np.array_equal(z2, G["Z"])


def test_padding_and_trajectory_boundaries():
    G, X, y, Z = _tables()
    t, L = int(G["t_traj"]), int(G["lookback"])
    # first sample of the second trajectory: window = 10 copies of that trajectory's first row, target = its second row
    x2, y2, z2 = D.sequence_items(X, y, Z, t, np.array([t, 2 * t - 1]), L)
    assert np.array_equal(z2[0], np.repeat(Z[t:t + 1], L, axis=0)) and y2[0, 0] == y[t + 1, 0]
    # last sample of a trajectory: target = the trajectory's own last target (no look into the next trajectory)
    assert y2[1, 0] == y[2 * t - 1, 0] and np.array_equal(z2[1], Z[2 * t - L:2 * t])
