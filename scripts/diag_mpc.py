import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "oracle")); sys.path.insert(0, os.path.join(REPO, "tests"))
import mpc_loss_oracle as O
import forging_control_b200 as fb
from conftest import state_dicts
np.set_printoptions(linewidth=200, precision=3)
W = np.load(os.path.join(REPO, "tests/golden/weights.npz")); C = np.load(os.path.join(REPO, "tests/golden/mpc_loss_cases.npz"))
dev = torch.device("cuda:0")
for name in sys.argv[1:] or ["n10_b33_init", "n25_b9", "n10_b12_wide"]:
    N, B, wd = (int(v) for v in C[f"{name}/meta"]); tag = str(C[f"{name}/ctl"])
    lstm, fnn = state_dicts(W, tag)
    sim = fb.LSTMModel(5,50,4,3); sim.load_state_dict({k: torch.tensor(v) for k,v in lstm.items()})
    ctl = fb.FNNModel(3,50,1,1); ctl.load_state_dict({k: torch.tensor(v) for k,v in fnn.items()})
    sim, ctl = sim.to(dev), ctl.to(dev)
    wp = fb.pack_weights(sim, ctl)
    X, Z, u0 = C[f"{name}/X"], C[f"{name}/Z"], C[f"{name}/f32/u0"]
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    out, g = O.mpc_loss_forward_backward(w, X.astype(np.float64), u0.astype(np.float64), Z.astype(np.float64), N, 20.0)
    for rep in range(2):
        for wg in (True, False):
            r = fb.mpc_loss_native(wp, torch.tensor(X).to(dev), torch.tensor(u0).to(dev), torch.tensor(Z).to(dev), N, 20.0, wg)
            cost = r["cost"].cpu().numpy(); pred = r["pred"].cpu().numpy()
            ec = np.abs(cost - out["cost"]) / np.abs(out["cost"]).max()
            ep = np.abs(pred - out["prediction"])
            print(name, "rep", rep, "with_grad", wg, "loss", r["gl"][250].item(), "oracle", out["loss"])
            print("  bad trajectories (cost):", np.nonzero(ec > 1e-5)[0], ec[ec > 1e-5])
            bad = np.argwhere(ep > 1e-5)
            first = {}
            for b, k in bad:
                first.setdefault(int(b), int(k))
            print("  first bad horizon step per trajectory (pred):", first)
            if wg:
                print("  du0 err", np.abs(r["du0"].cpu().numpy() - g["u0"]).max() / np.abs(g["u0"]).max())
