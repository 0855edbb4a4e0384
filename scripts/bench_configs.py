#!/usr/bin/env python
"""Per-config timings of the fused MPC loss (BASELINE.json configs 1-3 + the headline) for the selectable kernels,
device-resident inputs, CUDA events, median of 5 after 3 warm-ups."""
import json, os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import forging_control_b200 as fb
from forging_control_b200 import _native
W = np.load(os.path.join(REPO, "tests/golden/weights.npz"))
lstm = {k[5:]: W[k] for k in W.files if k.startswith("lstm/")}
fnn = {k[len("fnn_c0/"):]: W[k] for k in W.files if k.startswith("fnn_c0/")}
dev = torch.device("cuda:0")
sim = fb.LSTMModel(5, 50, 4, 3); sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()})
ctl = fb.FNNModel(3, 50, 1, 1); ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()})
sim, ctl = sim.to(dev), ctl.to(dev)
wp = fb.pack_weights(sim, ctl)
L = _native.lib()
rows = []
for name, N, B in (("config1 Main.py batch", 10, 15), ("config2", 5, 4096), ("config3", 25, 65536), ("headline/config5 per GPU", 10, 524288)):
    g = torch.Generator().manual_seed(1)
    X = (torch.rand(B, 3, generator=g) * 2 - 1).to(dev); Z = (torch.rand(B, 10, 5, generator=g) * 2 - 1).to(dev)
    with torch.no_grad():
        u0 = ctl(X).reshape(-1).contiguous()
    for kname, mode in (("ffma", 1), ("tcgen05 one tile", 2), ("tcgen05 pair", 3), ("replica (32-row tiles)", 4), ("auto", 0)):
        L.fc_mpc_select_kernel(mode)
        for _ in range(3):
            fb.mpc_loss_native(wp, X, u0, Z, N, 20.0, True)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = fb.mpc_loss_native(wp, X, u0, Z, N, 20.0, True); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        rows.append({"config": name, "N": N, "B": B, "kernel": kname, "ms": ms, "trajectory_steps_per_s": B * N / (ms * 1e-3), "loss": r["gl"][250].item()})
        print(json.dumps(rows[-1]), flush=True)
L.fc_mpc_select_kernel(0)
