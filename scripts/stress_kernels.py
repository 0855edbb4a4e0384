"""Randomised cross-check of the kernels behind fc_mpc_loss (FFMA / one-tile tcgen05 / pair / replica / pair with the tanh
polynomial) on odd shapes:
every output of every kernel must agree with the FFMA kernel to fp32 round-off (kink flips excepted on du0)."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import forging_control_b200 as fb
from forging_control_b200 import _native
from conftest import state_dicts
W = np.load(os.path.join(REPO, "tests/golden/weights.npz"))
dev = torch.device("cuda:0")
L = _native.lib()
rng = np.random.default_rng(2025)
shapes = [(1, 1), (2, 3), (31, 2), (32, 3), (33, 2), (65, 5), (127, 2), (128, 1), (129, 4), (255, 11), (256, 2), (257, 3), (300, 25), (1000, 1), (18943, 2), (18944, 3),
          (18945, 2), (37887, 2), (37889, 3), (50000, 7)]
worst = 0.0
for i, (B, N) in enumerate(shapes):
    tag = ("c0", "c3", "init")[i % 3]
    lstm, fnn = state_dicts(W, tag)
    sim = fb.LSTMModel(5, 50, 4, 3); sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()})
    ctl = fb.FNNModel(3, 50, 1, 1); ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()})
    sim, ctl = sim.to(dev), ctl.to(dev)
    wp = fb.pack_weights(sim, ctl)
    g = torch.Generator().manual_seed(1000 + i)
    scale = 1.0 if i % 4 else 2.5
    X = ((torch.rand(B, 3, generator=g) * 2 - 1) * scale).to(dev); Z = ((torch.rand(B, 10, 5, generator=g) * 2 - 1) * scale).to(dev)
    with torch.no_grad(): u0 = ctl(X).reshape(-1).contiguous()
    wg = bool(i % 5 != 4)
    outs = {}
    for mode in (1, 2, 3, 4, 5):
        L.fc_mpc_select_kernel(mode)
        outs[mode] = {k: (v.clone() if v is not None else None) for k, v in fb.mpc_loss_native(wp, X, u0, Z, N, 20.0, wg).items()}
    torch.cuda.synchronize()
    ref = outs[1]
    msg = [f"B={B} N={N} {tag} wg={wg}"]
    for mode in (2, 3, 4, 5):
        o = outs[mode]
        e = {k: float((o[k] - ref[k]).abs().max() / ref[k].abs().max().clamp_min(1e-30)) for k in ("cost", "pred", "command", "error")}
        e["loss"] = float(abs(o["gl"][250] - ref["gl"][250]) / abs(ref["gl"][250]))
        if wg:
            d = (o["du0"] - ref["du0"]).abs() / ref["du0"].abs().max().clamp_min(1e-30)
            e["du0_med"] = float(d.median()); e["du0_bad"] = int((d > 1e-5).sum())
            e["gl"] = float((o["gl"][:250] - ref["gl"][:250]).abs().max() / ref["gl"][:250].abs().max().clamp_min(1e-30))
        bad = max(e["cost"], e["pred"], e["loss"])
        worst = max(worst, bad)
        msg.append(f"k{mode}: " + " ".join(f"{k} {v:.1e}" if isinstance(v, float) else f"{k} {v}" for k, v in e.items()))
        assert bad < 1e-5 and not any(torch.isnan(o[k]).any() for k in ("cost", "pred", "gl")), msg
    print(" | ".join(msg), flush=True)
L.fc_mpc_select_kernel(0)
print("worst forward disagreement", worst)
