"""Device-side training-sample construction.

The reference builds every ``(x, y, z)`` sample in Python, one ``SequenceDataset.__getitem__`` call per sample
(UL/Functions.py:109-132), and lets ``DataLoader`` collate them (UL/Main.py:300-308).  ``DeviceSequenceLoader`` keeps the
flat scaled tables on the GPU and builds a whole batch of look-back windows with one gather kernel
(``fc_build_windows``); it yields ``(X [B,3], y [B,1], Z [B,lookback,5])`` CUDA tensors, i.e. exactly what
``NeuralNetwork.train_model`` / ``validate_model`` iterate over.  Bit-exact with the reference samples."""
from __future__ import annotations

import numpy as np
import torch

from . import _native


def build_windows(Xtab: torch.Tensor, ytab: torch.Tensor, Ztab: torch.Tensor, t_traj: int, idx: torch.Tensor, lookback: int = 10):
    """Tables [M,3], [M] or [M,1], [M,5] float32 CUDA; ``idx`` [B] int64 CUDA.  Returns (X [B,3], y [B,1], Z [B,lookback,5])."""
    if Xtab.device.type != "cuda":
        raise RuntimeError("build_windows: CUDA tensors required (forging_control_b200 has no CPU fallback)")
    dev = Xtab.device
    M = Xtab.shape[0]
    for t, w in ((Xtab, 3), (Ztab, 5)):
        if t.dtype != torch.float32 or t.dim() != 2 or t.shape != (M, w) or t.device != dev:
            raise ValueError("build_windows: expected float32 tables Xtab [M,3], Ztab [M,5] on one device")
    ytab = ytab.reshape(-1)
    if ytab.dtype != torch.float32 or ytab.numel() != M or ytab.device != dev or idx.dtype != torch.int64 or idx.device != dev:
        raise ValueError("build_windows: expected ytab [M] float32 and idx [B] int64 on the tables' device")
    Xtab, ytab, Ztab, idx = Xtab.contiguous(), ytab.contiguous(), Ztab.contiguous(), idx.contiguous()
    B = idx.numel()
    X = torch.empty(B, 3, dtype=torch.float32, device=dev)
    y = torch.empty(B, 1, dtype=torch.float32, device=dev)
    Z = torch.empty(B, lookback, 5, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _native.lib().fc_build_windows(_native.ptr(Xtab), _native.ptr(ytab), _native.ptr(Ztab), M, int(t_traj), int(lookback),
                                            _native.ptr(idx), B, _native.ptr(X), _native.ptr(y), _native.ptr(Z),
                                            _native.stream_ptr(dev))
    _native.check(rc, "fc_build_windows")
    return X, y, Z


class DeviceSequenceLoader:
    """Drop-in for ``DataLoader(ConcatDataset([SequenceDataset, ...]), batch_size, shuffle)`` (UL/Main.py:300-308).

    ``features`` / ``target`` / ``recurrent`` are float32 arrays or tensors [M,3], [M] or [M,1], [M,5] holding the
    per-trajectory datasets back to back (``t_traj`` rows each).  ``shuffle`` draws a device permutation per epoch from
    torch's generator (``torch.manual_seed``)."""

    def __init__(self, features, target, recurrent, t_traj: int, batch_size: int, shuffle: bool = False, lookback: int = 10,
                 device="cuda", drop_last: bool = False):
        dev = torch.device(device)
        as_t = lambda a: torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a, dtype=torch.float32).to(dev).contiguous()
        self.X, self.y, self.Z = as_t(features), as_t(target).reshape(-1), as_t(recurrent)
        self.M = self.X.shape[0]
        if self.M % t_traj != 0:
            raise ValueError("DeviceSequenceLoader: the table length must be a multiple of t_traj")
        self.t_traj, self.batch_size, self.shuffle = int(t_traj), int(batch_size), bool(shuffle)
        self.lookback, self.drop_last, self.device = int(lookback), bool(drop_last), dev

    @classmethod
    def from_datasets(cls, datasets, batch_size: int, shuffle: bool = False, device="cuda"):
        """From a list of the reference's ``SequenceDataset`` objects (attributes ``X``, ``y``, ``Z``, ``lookback``)."""
        t = len(datasets[0])
        if any(len(d) != t for d in datasets):
            raise ValueError("DeviceSequenceLoader.from_datasets: datasets of equal length expected")
        return cls(torch.cat([d.X for d in datasets]), torch.cat([d.y for d in datasets]), torch.cat([d.Z for d in datasets]),
                   t, batch_size, shuffle, datasets[0].lookback, device)

    def __len__(self):
        return self.M // self.batch_size if self.drop_last else (self.M + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        order = torch.randperm(self.M, device=self.device) if self.shuffle else torch.arange(self.M, device=self.device)
        for b in range(len(self)):
            idx = order[b * self.batch_size:(b + 1) * self.batch_size]
            yield build_windows(self.X, self.y, self.Z, self.t_traj, idx, self.lookback)
