#!/usr/bin/env python
"""Small / mid-size batches of the fused MPC loss: one-tile tcgen05 kernel, pair kernel and the replica mode (32-trajectory
tiles) side by side, device-resident inputs, CUDA events, median of 7 after 3 warm-ups."""
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
for N in (10, 5):
    for B in (15, 128, 1024, 4096, 4736, 6000, 9472, 18944, 37888):
        g = torch.Generator().manual_seed(1)
        X = (torch.rand(B, 3, generator=g) * 2 - 1).to(dev); Z = (torch.rand(B, 10, 5, generator=g) * 2 - 1).to(dev)
        with torch.no_grad():
            u0 = ctl(X).reshape(-1).contiguous()
        row = {"N": N, "B": B}
        for kname, mode in (("tc", 2), ("pair", 3), ("replica", 4), ("auto", 0)):
            L.fc_mpc_select_kernel(mode)
            for _ in range(3):
                fb.mpc_loss_native(wp, X, u0, Z, N, 20.0, True)
            ts = []
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r = fb.mpc_loss_native(wp, X, u0, Z, N, 20.0, True); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            row[kname + "_ms"] = round(float(np.median(ts)), 4)
            row[kname + "_loss"] = r["gl"][250].item()
        row["best_Msteps_per_s"] = round(B * N / (min(row[k + "_ms"] for k in ("tc", "pair", "replica")) * 1e3), 2)
        print(json.dumps(row), flush=True)
L.fc_mpc_select_kernel(0)
