import os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import forging_control_b200 as fb
from forging_control_b200 import _native
from conftest import state_dicts
W = np.load(os.path.join(REPO, "tests/golden/weights.npz"))
dev = torch.device("cuda:0")
lstm, fnn = state_dicts(W, "c0")
sim = fb.LSTMModel(5,50,4,3); sim.load_state_dict({k: torch.tensor(v) for k,v in lstm.items()})
ctl = fb.FNNModel(3,50,1,1); ctl.load_state_dict({k: torch.tensor(v) for k,v in fnn.items()})
sim, ctl = sim.to(dev), ctl.to(dev)
wp = fb.pack_weights(sim, ctl)
L = _native.lib()
for B, N in ((100, 1), (300, 2), (40000, 10), (524288, 10)):
    g = torch.Generator().manual_seed(B)
    X = (torch.rand(B, 3, generator=g) * 2 - 1).to(dev); Z = (torch.rand(B, 10, 5, generator=g) * 2 - 1).to(dev)
    with torch.no_grad(): u0 = ctl(X).reshape(-1).contiguous()
    L.fc_mpc_select_kernel(3)
    print("launch", B, N, flush=True)
    t = time.time()
    r = fb.mpc_loss_native(wp, X, u0, Z, N, 20.0, True); torch.cuda.synchronize()
    r = fb.mpc_loss_native(wp, X, u0, Z, N, 20.0, True); torch.cuda.synchronize()
    print("done", B, N, r["gl"][250].item(), f"{(time.time()-t)*500:.1f} ms", flush=True)
