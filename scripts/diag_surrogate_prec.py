"""Gradient error of the two surrogate-training paths against the fp64 oracle / the reference goldens."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "oracle")); sys.path.insert(0, os.path.join(REPO, "tests"))
import lstm_train_oracle as T
import mpc_loss_oracle as O
import forging_control_b200 as fb
from forging_control_b200 import surrogate as S, _native
from conftest import GOLDEN, rel_max
L = _native.lib(); dev = torch.device("cuda:0")
cases = np.load(os.path.join(GOLDEN, "surrogate_train_cases.npz"))
def ow(sd): return O.weights_from_state_dicts(sd, {"fc_inp.weight": np.zeros((50, 3)), "fc_inp.bias": np.zeros(50), "fc_out.weight": np.zeros((1, 50))}, np.float64)
for name, mode in (("ffma", 1), ("tc", 2)):
    L.fc_lstm_train_select_path(mode)
    for B in (1003, 6007):
        g = torch.Generator().manual_seed(100 + B); torch.manual_seed(5)
        m = fb.LSTMModel(5, 50, 4, 3)
        with torch.no_grad():
            for p in m.parameters(): p.mul_(2.0)
        sd = {k: v.detach().numpy().copy() for k, v in m.state_dict().items()}
        m = m.to(dev)
        X = torch.rand(B, 10, 5, generator=g) * 2 - 1; d = torch.randn(B, 4, generator=g)
        out = m(X.to(dev), dev); out.backward(d.to(dev))
        _, out_o, grads_o = T.lstm_mse_forward_backward(ow(sd), X.double().numpy(), np.zeros((B, 4)), d_out=d.double().numpy())
        print(name, B, "out %.1e" % rel_max(out.detach().cpu().numpy(), out_o), " ".join("%.1e" % rel_max(p.grad.double().cpu().numpy(), grads_o[k]) for k, p in m.named_parameters()), flush=True)
    sd = {k: cases[f"fresh_b256x3/init/{k}"] for k in T.GRAD_KEYS}
    m = fb.LSTMModel(5, 50, 4, 3); m.load_state_dict({k: torch.tensor(np.asarray(v, np.float32)) for k, v in sd.items()}); m = m.to(dev)
    opt = S.DeviceAdamW(m.parameters(), lr=1e-3, weight_decay=0.0)
    loader = [(torch.tensor(cases[f"fresh_b256x3/X{b}"]), torch.tensor(cases[f"fresh_b256x3/y{b}"])) for b in range(3)]
    S.SurrogateNeuralNetwork.train_model(loader, m, torch.nn.MSELoss(), opt, dev)
    r = []
    for k, p in m.named_parameters():
        init = cases[f"fresh_b256x3/init/{k}"].astype(np.float64)
        d_ref = cases[f"fresh_b256x3/f64/after/{k}"].astype(np.float64) - init
        dd = p.detach().double().cpu().numpy() - init
        r.append(np.abs(dd - d_ref).max() / np.abs(d_ref).max())
    print(name, "3-step AdamW weight-delta error / step:", " ".join("%.1e" % v for v in r),
          "| grads", " ".join("%.1e" % rel_max(p.grad.double().cpu().numpy(), cases[f"fresh_b256x3/f64/grad/{k}"]) for k, p in m.named_parameters()), flush=True)
L.fc_lstm_train_select_path(0)
