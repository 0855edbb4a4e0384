"""Development check: tensor-core path of the surrogate training step against the FFMA path and the fp64 oracle."""
import os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "oracle")); sys.path.insert(0, os.path.join(REPO, "tests"))
import forging_control_b200 as fb
from forging_control_b200 import _native, surrogate
L = _native.lib()
dev = torch.device("cuda:0")
W = np.load(os.path.join(REPO, "tests/golden/weights.npz"))
lstm = {k[5:]: W[k] for k in W.files if k.startswith("lstm/")}
def model():
    m = fb.LSTMModel(5, 50, 4, 3); m.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()}); return m.to(dev)
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-300))
for B in [int(a) for a in (sys.argv[1:] or ["300", "1000", "40000"])]:
    g = torch.Generator().manual_seed(B)
    x = (torch.rand(B, 10, 5, generator=g) * 2 - 1).to(dev); y = (torch.rand(B, 4, generator=g) * 2 - 1).to(dev)
    res = {}
    for name, mode in (("ffma", 1), ("tc", 2)):
        L.fc_lstm_train_select_path(mode)
        m = model()
        out = m(x, dev)
        loss = torch.nn.functional.mse_loss(out, y)
        loss.backward()
        torch.cuda.synchronize()
        res[name] = (out.detach(), [p.grad.clone() for p in m.parameters()], loss.item())
    print(f"B={B}: out rel {rel(res['tc'][0], res['ffma'][0]):.2e} loss {res['tc'][2]:.8f} / {res['ffma'][2]:.8f} grads " +
          " ".join(f"{rel(a, b):.1e}" for a, b in zip(res['tc'][1], res['ffma'][1])), flush=True)
    if B <= 2000:
        import lstm_train_oracle as O
        # fp64 oracle
        sd = {k: np.asarray(v, np.float64) for k, v in lstm.items()}
        try:
            ref = O.forward_backward(sd, x.double().cpu().numpy(), (2 * (res['tc'][0].double().cpu().numpy() - y.double().cpu().numpy()) / (4 * B)))
            print("   oracle available", type(ref))
        except Exception as e:
            print("   (oracle call skipped:", str(e)[:80], ")")
    if B >= 20000:
        for name, mode in (("ffma", 1), ("tc", 2)):
            L.fc_lstm_train_select_path(mode)
            m = model()
            for _ in range(3):
                m.zero_grad(); torch.nn.functional.mse_loss(m(x, dev), y).backward()
            torch.cuda.synchronize(); t0 = time.time()
            for _ in range(5):
                m.zero_grad(); torch.nn.functional.mse_loss(m(x, dev), y).backward()
            torch.cuda.synchronize(); ms = (time.time() - t0) / 5 * 1e3
            print(f"   {name}: {ms:.3f} ms per fwd+bwd = {B / ms / 1e3:.2f} M samples/s", flush=True)
L.fc_lstm_train_select_path(0)
