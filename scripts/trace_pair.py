#!/usr/bin/env python
"""Development aid: event trace of the pair kernel (library built with -DFC_TC_TRACE, chosen with FC_LIB_PATH).
Dumps the raw (clock, event) streams of warps 0 / 5 / 13 of CTA 0 to gpurun_out/trace_pair.npz."""
import ctypes, os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import forging_control_b200 as fb
from forging_control_b200 import _native
W = np.load(os.path.join(REPO, "tests/golden/weights.npz"))
lstm = {k[5:]: W[k] for k in W.files if k.startswith("lstm/")}
fnn = {k[len("fnn_c0/"):]: W[k] for k in W.files if k.startswith("fnn_c0/")}
dev = torch.device("cuda:0")
sim = fb.LSTMModel(5, 50, 4, 3); sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()}); sim = sim.to(dev)
ctl = fb.FNNModel(3, 50, 1, 1); ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()}); ctl = ctl.to(dev)
wp = fb.pack_weights(sim, ctl)
B, N = int(os.environ.get('AB_B', 37888)), int(os.environ.get('AB_N', 10))
g = torch.Generator(device=dev).manual_seed(1234)
X = torch.rand(B, 3, generator=g, device=dev) * 2 - 1
Z = torch.rand(B, 10, 5, generator=g, device=dev) * 2 - 1
u0 = ctl(X).detach().reshape(-1).contiguous()
for _ in range(3):
    r = fb.mpc_loss_native(wp, X, u0, Z, N, 20.0, True)
torch.cuda.synchronize()
L = _native.lib()
ev = np.zeros((3, 32768), dtype=np.int64); n = np.zeros(3, dtype=np.int32)
L.fc_debug_trace(ev.ctypes.data_as(ctypes.c_void_p), n.ctypes.data_as(ctypes.c_void_p))
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(REPO, "gpurun_out", "trace_pair.npz"), ev=ev, n=n)
print("events", n.tolist(), "loss", float(r["gl"][250]))
