"""Soak test of the kernels added in the second half of round 2: many random batch sizes through the replica mode of the MPC
loss and through the tensor-core surrogate training path (both instantiations, chunked backward), checking finiteness and
agreement with the FFMA kernels.  A hang shows as the caller's timeout."""
import os, sys, numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import forging_control_b200 as fb
from forging_control_b200 import _native
L = _native.lib(); dev = torch.device("cuda:0")
W = np.load(os.path.join(REPO, "tests/golden/weights.npz"))
lstm = {k[5:]: W[k] for k in W.files if k.startswith("lstm/")}; fnn = {k[len("fnn_c0/"):]: W[k] for k in W.files if k.startswith("fnn_c0/")}
sim = fb.LSTMModel(5, 50, 4, 3); sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()}); sim = sim.to(dev)
ctl = fb.FNNModel(3, 50, 1, 1); ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()}); ctl = ctl.to(dev)
wp = fb.pack_weights(sim, ctl)
rng = np.random.default_rng(7)
worst = 0.0
for i in range(60):
    B = int(rng.choice([rng.integers(1, 130), rng.integers(130, 4737), rng.integers(4737, 40000)])); N = int(rng.integers(1, 13))
    X = torch.rand(B, 3, device=dev) * 2 - 1; Z = torch.rand(B, 10, 5, device=dev) * 2 - 1
    u0 = ctl(X).detach().reshape(-1).contiguous()
    outs = {}
    for mode in (1, 0, 4 if B <= 20000 else 3):
        L.fc_mpc_select_kernel(mode); r = fb.mpc_loss_native(wp, X, u0, Z, N, 20.0, True); outs[mode] = (r["gl"].clone(), r["cost"].clone())
    torch.cuda.synchronize()
    ref = outs[1]
    for mode, o in outs.items():
        assert torch.isfinite(o[0]).all() and torch.isfinite(o[1]).all(), (B, N, mode)
        e = float((o[0][:251] - ref[0][:251]).abs().max() / ref[0][:251].abs().max()); worst = max(worst, e)
        assert e < 2e-5, (B, N, mode, e)
L.fc_mpc_select_kernel(0)
print("mpc soak ok, worst gl disagreement with the FFMA kernel", worst, flush=True)
worst = 0.0
for i in range(40):
    B = int(rng.choice([rng.integers(1, 200), rng.integers(200, 4737), rng.integers(4737, 90000)]))
    x = torch.rand(B, 10, 5, device=dev) * 2 - 1; d = torch.randn(B, 4, device=dev) * float(10.0 ** rng.uniform(-6, 2))
    gs = {}
    for mode in (1, 2):
        L.fc_lstm_train_select_path(mode)
        m = fb.LSTMModel(5, 50, 4, 3); m.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()}); m = m.to(dev)
        out = m(x, dev); out.backward(d); torch.cuda.synchronize()
        gs[mode] = [p.grad.clone() for p in m.parameters()]
    for a, b in zip(gs[2], gs[1]):
        assert torch.isfinite(a).all(), B
        e = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)); worst = max(worst, e)
        if e > 2e-6: print("  B", B, "tensor", tuple(a.shape), "disagreement", e, flush=True)
        assert e < 2e-5, (B, e)
L.fc_lstm_train_select_path(0)
print("training soak ok, worst gradient disagreement with the FFMA kernels", worst, flush=True)
