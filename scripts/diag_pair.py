"""Development check: the pair kernel (FC mode 3) against the one-tile tcgen05 kernel (mode 2) on the same inputs."""
import os, sys
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
def rel(a, b): return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
for B, N, wg in ((100, 3, True), (130, 5, True), (300, 10, True), (1000, 10, False), (20000, 10, True), (37888 * 2 + 77, 12, True)):
    g = torch.Generator(device="cpu").manual_seed(B)
    X = (torch.rand(B, 3, generator=g) * 2 - 1).to(dev); Z = (torch.rand(B, 10, 5, generator=g) * 2 - 1).to(dev)
    with torch.no_grad(): u0 = ctl(X).reshape(-1).contiguous()
    outs = {}
    for mode in (2, 3):
        L.fc_mpc_select_kernel(mode)
        outs[mode] = fb.mpc_loss_native(wp, X, u0, Z, N, 20.0, wg)
        torch.cuda.synchronize()
    a, b = outs[3], outs[2]
    msg = f"B={B} N={N} wg={wg}: loss {rel(a['gl'][250], b['gl'][250]):.1e} cost {rel(a['cost'], b['cost']):.1e} pred {rel(a['pred'], b['pred']):.1e}"
    if wg: msg += f" du0 {rel(a['du0'], b['du0']):.1e} gl {rel(a['gl'][:250], b['gl'][:250]):.1e}"
    print(msg, flush=True)
L.fc_mpc_select_kernel(0)
