"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatement of the reference's training-sample construction: ``SequenceDataset.__getitem__``
(UL/Functions.py:109-132) over the per-trajectory slices made by ``Data.get_individual_dataset`` (:479-516).
Pinned by ``tests/golden/sequence_dataset.npz`` (items of the unmodified reference classes, oracle/make_golden.py)."""
from __future__ import annotations

import numpy as np


def sequence_items(Xtab: np.ndarray, ytab: np.ndarray, Ztab: np.ndarray, t_traj: int, idx: np.ndarray, lookback: int = 10):
    """Tables [M,3], [M,1], [M,5] = the concatenated per-trajectory datasets (M = n_traj * t_traj); ``idx`` global sample
    indices.  Sample g lives in trajectory k = g // t_traj at local step i = g % t_traj:
      x = X[g];  z = rows max(i-lookback+1, 0)..i of the trajectory, front-padded with its row 0 (:117-123);
      y = y of the NEXT step of the trajectory, the last one for the final step (:126-129)."""
    idx = np.asarray(idx, dtype=np.int64)
    k, i = idx // t_traj, idx % t_traj
    X = Xtab[idx]
    y = ytab[k * t_traj + np.minimum(i + 1, t_traj - 1)]
    r = np.arange(lookback)[None, :]
    loc = np.maximum(i[:, None] - lookback + 1 + r, 0)
    Z = Ztab[k[:, None] * t_traj + loc]
    return X, y, Z
