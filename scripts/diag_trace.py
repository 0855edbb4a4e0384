"""Summed-gradient error of the selected kernel on the trace-derived inputs (the cancellation-heavy case) and on
uniform inputs, for calibration experiments (env FC_MPC_KERNEL, FC_TC_ACC_COMP)."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "oracle")); sys.path.insert(0, os.path.join(REPO, "tests"))
import mpc_loss_oracle as O
import forging_control_b200 as fb
from conftest import state_dicts
W = np.load(os.path.join(REPO, "tests/golden/weights.npz")); T = np.load(os.path.join(REPO, "tests/golden/trace_windows.npz"))
dev = torch.device("cuda:0")
def rel(a, b): return float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())
res = []
for tag, seed, N in (("c0", 5, 10), ("c3", 6, 10), ("c0", 7, 25)):
    lstm, fnn = state_dicts(W, tag)
    sim = fb.LSTMModel(5,50,4,3); sim.load_state_dict({k: torch.tensor(v) for k,v in lstm.items()})
    ctl = fb.FNNModel(3,50,1,1); ctl.load_state_dict({k: torch.tensor(v) for k,v in fnn.items()})
    sim, ctl = sim.to(dev), ctl.to(dev)
    rng = np.random.default_rng(seed); reps = 6
    X = (np.tile(T["X"], (reps,1)) * (1 + 0.01*rng.uniform(-1,1,(256*reps,3)))).astype(np.float32)
    Z = (np.tile(T["Z"], (reps,1,1)) * (1 + 0.01*rng.uniform(-1,1,(256*reps,10,5)))).astype(np.float32)
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    u0 = O.fnn_forward(w, X.astype(np.float64))[:,0].astype(np.float32)
    out, g = O.mpc_loss_forward_backward(w, X.astype(np.float64), u0.astype(np.float64), Z.astype(np.float64), N, 20.0)
    r = fb.mpc_loss_native(fb.pack_weights(sim, ctl), torch.tensor(X).to(dev), torch.tensor(u0).to(dev), torch.tensor(Z).to(dev), N, 20.0, True)
    gl = r["gl"].cpu().numpy()
    d = np.abs(r["du0"].cpu().numpy() - g["u0"]) / np.abs(g["u0"]).max()
    res.append((tag, N, rel(gl[:150].reshape(50,3), g["inp_w"]), rel(gl[150:200], g["inp_b"]), rel(gl[200:250], g["out_w"][0]), d.max(), np.median(d)))
print(os.environ.get("FC_MPC_KERNEL"), os.environ.get("FC_TC_ACC_COMP"), " | ".join(f"{t} N={n}: inp_w {a:.1e} inp_b {b:.1e} out_w {c:.1e} du0 max {m:.1e} med {md:.1e}" for t,n,a,b,c,m,md in res))
