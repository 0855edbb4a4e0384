#!/usr/bin/env python
"""Throughput of the LSTM shadow roll-out (fc_lstm_shadow_rollout, forward-only mode of the pair kernel): windowed-LSTM
inferences per second, device-resident inputs, CUDA events, median of 5 after 2 warm-ups."""
import ctypes, json, os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import forging_control_b200 as fb
from forging_control_b200 import _native
from forging_control_b200.Functions import _workspace
W = np.load(os.path.join(REPO, "tests/golden/weights.npz"))
lstm = {k[5:]: W[k] for k in W.files if k.startswith("lstm/")}
dev = torch.device("cuda:0")
sim = fb.LSTMModel(5, 50, 4, 3); sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()}); sim = sim.to(dev)
wp = fb.pack_weights(sim, None)
L = _native.lib()
ratio = (ctypes.c_float * 4)(1.0, 1.0, 1.0, 1.0)
for B, T in ((128 * 148 * 2, 100), (1048576, 20)):
    g = torch.Generator().manual_seed(5)
    row0 = (torch.rand(B, 5, generator=g) * 2 - 1).to(dev); u = (torch.rand(B, T, generator=g) * 2 - 1).to(dev)
    y = torch.empty(B, T, 4, device=dev)
    nb = int(L.fc_lstm_shadow_workspace_bytes(B, T)); work = _workspace(dev, nb)
    def run():
        _native.check(L.fc_lstm_shadow_rollout(_native.ptr(row0), _native.ptr(u), ratio, _native.ptr(wp), B, T, _native.ptr(y),
                                               _native.ptr(work), nb, _native.stream_ptr(dev)), "fc_lstm_shadow_rollout")
    for _ in range(2): run()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(json.dumps({"B": B, "T": T, "ms": ms, "window_inferences_per_s": B * T / (ms * 1e-3),
                      "algorithmic_tflops": 1.0208e6 * B * T / (ms * 1e-3) / 1e12}), flush=True)
