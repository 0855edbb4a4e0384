#!/usr/bin/env python
"""Sustained A/B of the fused MPC-loss kernel: B = 524 288, N = 10, 12 back-to-back launches after 3 warm-ups, CUDA
events around the timed block (power-capped clocks included).  The library is chosen with FC_LIB_PATH."""
import json, os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import forging_control_b200 as fb
W = np.load(os.path.join(REPO, "tests/golden/weights.npz"))
lstm = {k[5:]: W[k] for k in W.files if k.startswith("lstm/")}
fnn = {k[len("fnn_c0/"):]: W[k] for k in W.files if k.startswith("fnn_c0/")}
dev = torch.device("cuda:0")
sim = fb.LSTMModel(5, 50, 4, 3); sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()}); sim = sim.to(dev)
ctl = fb.FNNModel(3, 50, 1, 1); ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()}); ctl = ctl.to(dev)
wp = fb.pack_weights(sim, ctl)
B, N = int(os.environ.get('AB_B', 524288)), int(os.environ.get('AB_N', 10))
g = torch.Generator(device=dev).manual_seed(1234)
X = torch.rand(B, 3, generator=g, device=dev) * 2 - 1
Z = torch.rand(B, 10, 5, generator=g, device=dev) * 2 - 1
u0 = ctl(X).detach().reshape(-1).contiguous()
from forging_control_b200 import _native
if os.environ.get('AB_MODE'):
    _native.lib().fc_mpc_select_kernel(int(os.environ['AB_MODE']))
for _ in range(3):
    fb.mpc_loss_native(wp, X, u0, Z, N, 20.0, True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = int(os.environ.get('AB_K', 12))
e0.record()
for _ in range(K):
    r = fb.mpc_loss_native(wp, X, u0, Z, N, 20.0, True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({"lib": os.environ.get("FC_LIB_PATH", "default"), "mode": os.environ.get("AB_MODE", "auto"), "B": B, "N": N, "ms": ms, "Msteps_per_s": B * N / ms / 1e3, "loss": float(r["gl"][250])}))
