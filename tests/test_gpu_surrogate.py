"""GPU (-m gpu): parity of the surrogate-training kernels (fc_lstm_window_fwd / fc_lstm_window_bwd) and of the
optimizer kernel (fc_adamw_step), called through the C ABI via forging_control_b200, against
  * the golden vectors produced by the unmodified reference (tests/golden/surrogate_train_cases.npz:
    Model_NN NeuralNetwork.train_model + nn.MSELoss + torch.optim.AdamW, fp64 arbiter),
  * the fp64 oracle (oracle/lstm_train_oracle.py) on seeded inputs with ragged batches,
  * size-independent properties (linearity in the upstream gradient, shard sums) at 65 536 samples.
Tolerance: max|a-b| <= 1e-5 * max|b| per tensor (loss: plain relative), as on the MPC-loss path."""
import os

import numpy as np
import pytest
import torch

import lstm_train_oracle as T
import mpc_loss_oracle as O
import forging_control_b200 as fb
from forging_control_b200 import surrogate as S
from conftest import GOLDEN, rel_max, state_dicts

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def cases():
    return np.load(os.path.join(GOLDEN, "surrogate_train_cases.npz"))


@pytest.fixture(params=["ffma", "tensor-core"], autouse=True)
def path(request):
    """Every test runs against both implementations behind fc_lstm_window_fwd / _bwd: the FP32 FFMA kernels and the
    tensor-core path (pair kernel in training mode + tcgen05 weight-gradient kernel); fc_lstm_train_select_path."""
    from forging_control_b200 import _native
    L = _native.lib()
    assert L.fc_lstm_train_select_path({"ffma": 1, "tensor-core": 2}[request.param]) == 0
    yield request.param
    assert L.fc_lstm_train_select_path(0) == 0


def _model(sd, dev):
    m = fb.LSTMModel(5, 50, 4, 3)
    m.load_state_dict({k: torch.tensor(np.asarray(v, np.float32)) for k, v in sd.items()}, strict=True)
    return m.to(dev)


def _oracle_weights(sd):
    return O.weights_from_state_dicts(sd, {"fc_inp.weight": np.zeros((50, 3)), "fc_inp.bias": np.zeros(50),
                                           "fc_out.weight": np.zeros((1, 50))}, np.float64)


def _grads(m):
    return {k: p.grad.detach().double().cpu().numpy() for k, p in m.named_parameters()}


def test_training_step_matches_reference_golden(dev, cases, golden_weights):
    lstm, _ = state_dicts(golden_weights, "c0")
    m = _model(lstm, dev)
    X = torch.tensor(cases["shipped_b37/X"]).to(dev)
    y = torch.tensor(cases["shipped_b37/y"]).to(dev)
    out = m(X, dev)
    assert out.shape == (37, 4) and out.requires_grad
    loss = torch.nn.MSELoss()(out, y.squeeze())          # Model_NN/Functions.py:554
    loss.backward()                                      # :560
    assert abs(loss.item() - float(cases["shipped_b37/f64/avg_loss"])) <= TOL * abs(loss.item())
    assert rel_max(out.detach().cpu().numpy(), cases["shipped_b37/f64/out"]) <= TOL
    for k, g in _grads(m).items():
        assert rel_max(g, cases[f"shipped_b37/f64/grad/{k}"]) <= TOL, k
    with torch.no_grad():                                # inference mode of the same kernel (nothing recorded)
        out2 = m(X, dev)
    assert torch.equal(out2, out.detach())


def test_three_epoch_steps_with_device_adamw_match_reference(dev, cases):
    sd = {k: cases[f"fresh_b256x3/init/{k}"] for k in T.GRAD_KEYS}
    m = _model(sd, dev)
    opt = S.DeviceAdamW(m.parameters(), lr=1e-3, weight_decay=0.0)           # Model_NN/Main.py:230
    loader = [(torch.tensor(cases[f"fresh_b256x3/X{b}"]), torch.tensor(cases[f"fresh_b256x3/y{b}"])) for b in range(3)]
    avg = S.SurrogateNeuralNetwork.train_model(loader, m, torch.nn.MSELoss(), opt, dev)
    assert abs(avg - float(cases["fresh_b256x3/f64/avg_loss"])) <= TOL * abs(avg)
    for k, g in _grads(m).items():                        # gradients of the last batch, after two optimizer steps
        assert rel_max(g, cases[f"fresh_b256x3/f64/grad/{k}"]) <= 3e-5, k
    for k, p in m.named_parameters():
        init = cases[f"fresh_b256x3/init/{k}"].astype(np.float64)
        d_ref = cases[f"fresh_b256x3/f64/after/{k}"].astype(np.float64) - init
        d = p.detach().double().cpu().numpy() - init
        # the reference's own fp32 run is 5e-6 .. 5e-5 of the step away from its fp64 run (weights quantised to fp32)
        assert np.abs(d - d_ref).max() <= 2e-4 * np.abs(d_ref).max(), k
    st = opt.state_dict()["state"][0]
    assert set(st) == {"step", "exp_avg", "exp_avg_sq"} and float(st["step"]) == 3.0


# 6007: more 40-sample tiles than SMs -> 80-sample forward (FFMA path); tensor-core path: <= 4736 = 32 x 148 samples run the
# replica instantiation (32-sample tiles), 4737 and 6007 pairs of 128-sample tiles
@pytest.mark.parametrize("B", [1, 33, 40, 41, 1003, 4737, 6007])
def test_ragged_batches_match_fp64_oracle_with_an_arbitrary_upstream_gradient(dev, B):
    g = torch.Generator().manual_seed(100 + B)
    torch.manual_seed(5)
    m = fb.LSTMModel(5, 50, 4, 3)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(2.0)                                   # livelier gates than the default init
    sd = {k: v.detach().numpy().copy() for k, v in m.state_dict().items()}
    m = m.to(dev)
    X = torch.rand(B, 10, 5, generator=g) * 2 - 1
    d = torch.randn(B, 4, generator=g)
    out = m(X.to(dev), dev)
    out.backward(d.to(dev))
    _, out_o, grads_o = T.lstm_mse_forward_backward(_oracle_weights(sd), X.double().numpy(), np.zeros((B, 4)),
                                                    d_out=d.double().numpy())
    assert rel_max(out.detach().cpu().numpy(), out_o) <= TOL
    for k, gr in _grads(m).items():
        assert rel_max(gr, grads_o[k]) <= TOL, k


def test_full_size_properties(dev, golden_weights):
    """65 536 samples: linearity of the reverse sweep in d_out and additivity over shards of the batch."""
    lstm, _ = state_dicts(golden_weights, "c0")
    m = _model(lstm, dev)
    B = 65536
    g = torch.Generator(device=dev).manual_seed(9)
    X = torch.rand(B, 10, 5, generator=g, device=dev) * 2 - 1
    d1 = torch.randn(B, 4, generator=g, device=dev) / B
    d2 = torch.randn(B, 4, generator=g, device=dev) / B

    def run(x, d):
        m.zero_grad()
        out = m(x, dev)
        out.backward(d)
        return out.detach(), _grads(m)

    o1, g1 = run(X, d1)
    o2, g2 = run(X, d2)
    o3, g3 = run(X, d1 + 2.0 * d2)
    assert torch.equal(o1, o2) and torch.equal(o1, o3) and bool(torch.isfinite(o1).all())
    with torch.no_grad():                                 # inference mode of the 80-sample forward (per-CTA scratch)
        assert torch.equal(m(X, dev), o1)
    _, ga = run(X[:30000], d1[:30000])
    _, gb = run(X[30000:], d1[30000:])
    for k in g1:
        assert rel_max(g3[k], g1[k] + 2.0 * g2[k]) <= TOL, k
        assert rel_max(ga[k] + gb[k], g1[k]) <= TOL, k
    # against stock autograd through nn.LSTM on a slice (cuDNN computes in TF32 here: loose tolerance, sanity only)
    ref = torch.nn.LSTM(5, 50, 3, batch_first=True, bias=False).to(dev)
    ref.load_state_dict({k[5:]: v for k, v in m.state_dict().items() if k.startswith("lstm.")})
    h, _ = ref(X[:512])
    assert rel_max((h[:, -1] @ m.fc.weight.T + m.fc.bias).detach().cpu().numpy(), o1[:512].cpu().numpy()) <= 5e-3


def test_adamw_kernel_matches_torch(dev, cases):
    ps = [torch.nn.Parameter(torch.tensor(cases["adamw_ctl/p0"][a:b]).to(dev)) for a, b in ((0, 150), (150, 200), (200, 250))]
    opt = S.DeviceAdamW(ps, lr=1e-3)                     # UL/Main.py:195 (default weight decay 0.01)
    for s in range(5):
        gflat = torch.tensor(cases["adamw_ctl/grads"][s]).to(dev)
        for p, (a, b) in zip(ps, ((0, 150), (150, 200), (200, 250))):
            p.grad = gflat[a:b].clone()
        v0 = ps[0]._version
        opt.step()
        assert ps[0]._version > v0                        # the weight-image caches key on the version counter
        got = torch.cat([p.detach().reshape(-1) for p in ps]).cpu().numpy()
        assert rel_max(got, cases[f"adamw_ctl/p{s + 1}"]) <= 1e-6
    # 11 tensors (two launches) against the stock optimizer on the device, state_dict interchange
    torch.manual_seed(3)
    shapes = [(200, 5), (200, 50)] * 5 + [(4,)]
    a = [torch.nn.Parameter(torch.randn(s, device=dev) * 0.1) for s in shapes]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    oa, ob = S.DeviceAdamW(a, lr=3e-3, weight_decay=0.05, betas=(0.8, 0.99)), torch.optim.AdamW(
        b, lr=3e-3, weight_decay=0.05, betas=(0.8, 0.99))
    for s in range(4):
        for p, q in zip(a, b):
            p.grad = torch.randn_like(p) * 10.0 ** (s - 2)
            q.grad = p.grad.clone()
        oa.step()
        ob.step()
    for p, q in zip(a, b):
        assert rel_max(p.detach().cpu().numpy(), q.detach().cpu().numpy()) <= 1e-6
    ob.load_state_dict(oa.state_dict())


def test_controller_training_with_device_adamw_tracks_torch_adamw(dev, golden_weights):
    """UL train_model (MPC loss) driven by DeviceAdamW: same parameters as with torch.optim.AdamW after 3 steps (the
    packed controller weights must be refreshed after every raw-pointer update)."""
    lstm, fnn = state_dicts(golden_weights, "init")
    g = torch.Generator().manual_seed(11)
    loader = [(torch.rand(64, 3, generator=g) * 2 - 1, torch.zeros(64, 1), torch.rand(64, 10, 5, generator=g) * 2 - 1)
              for _ in range(3)]
    res = []
    for cls in (S.DeviceAdamW, torch.optim.AdamW):
        sim = _model(lstm, dev)
        ctl = fb.FNNModel(3, 50, 1, 1)
        ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()})
        ctl = ctl.to(dev)
        opt = cls(ctl.parameters(), lr=1e-3)
        fb.NeuralNetwork.train_model(loader, sim, ctl, fb.MPCLoss(10, 20.0), opt, dev)
        res.append({k: p.detach().cpu().numpy() for k, p in ctl.named_parameters()})
    for k in res[0]:
        d = res[1][k] - fnn[k]
        assert np.abs(res[0][k] - res[1][k]).max() <= 1e-3 * max(np.abs(d).max(), 1e-12), k


def test_c_abi_errors(dev):
    L = fb._native.lib()
    assert L.fc_lstm_window_workspace_bytes(0, 1) == 0
    assert L.fc_lstm_window_fwd(0, 0, 0, 0, 4, 0, 0, 0, 0, 0) == -3
    x = torch.zeros(4, 10, 5, device=dev)
    m = fb.LSTMModel(5, 50, 4, 3).to(dev)
    with pytest.raises(NotImplementedError):
        S.lstm_window(m, x.cpu())
    with pytest.raises(NotImplementedError):
        S.lstm_window(m, x[:, :9])
