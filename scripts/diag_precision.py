"""Where does the kernel's gradient error come from?  Compare, on the same inputs, (a) the CUDA kernel and
(b) a float32 torch CPU evaluation (what the reference does) against the fp64 oracle."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "oracle")); sys.path.insert(0, os.path.join(REPO, "tests"))
import mpc_loss_oracle as O
import forging_control_b200 as fb
from conftest import state_dicts
W = np.load(os.path.join(REPO, "tests/golden/weights.npz"))
dev = torch.device("cuda:0")
def rel(a, b): return float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())
for (N, B, tag, seed) in [(10, 17797, "c0", 1244), (10, 4096, "c0", 5), (10, 4096, "init", 6), (25, 2048, "c0", 7)]:
    lstm, fnn = state_dicts(W, tag)
    sim = fb.LSTMModel(5,50,4,3); sim.load_state_dict({k: torch.tensor(v) for k,v in lstm.items()})
    ctl = fb.FNNModel(3,50,1,1); ctl.load_state_dict({k: torch.tensor(v) for k,v in fnn.items()})
    sim, ctl = sim.to(dev), ctl.to(dev)
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(B, 3, generator=g) * 2 - 1; Z = torch.rand(B, 10, 5, generator=g) * 2 - 1
    w64 = O.weights_from_state_dicts(lstm, fnn, np.float64)
    u0 = O.fnn_forward(w64, X.double().numpy())[:, 0].astype(np.float32)
    out, gr = O.mpc_loss_forward_backward(w64, X.double().numpy(), u0.astype(np.float64), Z.double().numpy(), N, 20.0)
    r = fb.mpc_loss_native(fb.pack_weights(sim, ctl), X.to(dev), torch.tensor(u0).to(dev), Z.to(dev), N, 20.0, True)
    gl = r["gl"].cpu().numpy()
    # float32 torch CPU
    w32 = O.weights_from_state_dicts(lstm, fnn, np.float32)
    tw = {k: ([torch.tensor(a) for a in v] if isinstance(v, list) else torch.tensor(v)) for k, v in w32.items()}
    for k in ("inp_w", "inp_b", "out_w"): tw[k].requires_grad_()
    u0t = torch.tensor(u0, requires_grad=True)
    l32 = O.mpc_loss_torch(tw, X, u0t, Z, N, 20.0)
    l32[0].backward()
    d_k = np.abs(r["du0"].cpu().numpy() - gr["u0"]) / np.abs(gr["u0"]).max()
    d_t = np.abs(u0t.grad.numpy() - gr["u0"]) / np.abs(gr["u0"]).max()
    print(f"N={N} B={B} {tag}: loss rel kernel {abs(gl[250]-out['loss'])/out['loss']:.2e} torch32 {abs(l32[0].item()-out['loss'])/out['loss']:.2e}")
    print(f"   inp_w  kernel {rel(gl[:150].reshape(50,3), gr['inp_w']):.2e}  torch32 {rel(tw['inp_w'].grad.numpy(), gr['inp_w']):.2e}")
    print(f"   out_w  kernel {rel(gl[200:250], gr['out_w'][0]):.2e}  torch32 {rel(tw['out_w'].grad.numpy()[0], gr['out_w'][0]):.2e}")
    print(f"   du0    kernel max {d_k.max():.2e} #>1e-5 {(d_k>1e-5).sum()} median {np.median(d_k):.2e} | torch32 max {d_t.max():.2e} #>1e-5 {(d_t>1e-5).sum()} median {np.median(d_t):.2e}")
    print(f"   cost   kernel {rel(r['cost'].cpu().numpy(), out['cost']):.2e} torch32 {rel(l32[1].detach().numpy(), out['cost']):.2e}; pred kernel {rel(r['pred'].cpu().numpy(), out['prediction']):.2e} torch32 {rel(l32[4].detach().numpy(), out['prediction']):.2e}")
