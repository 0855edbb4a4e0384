"""GPU (-m gpu): parity of the fused sm_100a MPC-loss kernel, called through the C ABI
(fc_pack_weights / fc_mpc_loss via forging_control_b200), against
  * the golden vectors produced by the unmodified reference (tests/golden/mpc_loss_cases.npz),
  * the fp64 oracle on seeded inputs at config-sized batches,
  * size-independent properties at BASELINE.json's full sizes.
Tolerance (SURVEY.md section 8a / north_star): loss relative <= 1e-5; every other tensor
max|a-b| <= 1e-5 * max|b|."""
import numpy as np
import pytest
import torch

import mpc_loss_oracle as O
import forging_control_b200 as fb
from conftest import rel_max, state_dicts

pytestmark = pytest.mark.gpu
TOL = 1e-5
ALPHA = 20.0
W1_CASES = ["n1_b3", "n2_b5", "n5_b16", "n10_b15", "n10_b33_init", "n10_b40_trace", "n25_b9", "n12_b130_trace",
            "n10_b12_wide"]


def _models(W, tag, dev):
    lstm, fnn = state_dicts(W, tag)
    sim = fb.LSTMModel(5, 50, 4, 3)
    sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()}, strict=True)
    ctl = fb.FNNModel(3, 50, 1, 1)
    ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()}, strict=True)
    return sim.to(dev), ctl.to(dev), lstm, fnn


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(params=["ffma", "tcgen05", "pair", "replica", "pair-precise"], autouse=True)
def kernel(request):
    """Every parity test runs against all kernels behind fc_mpc_loss: the FP32 FFMA kernel, the one-tile tcgen05 kernel,
    the two-tile tcgen05 pair kernel, its replica mode (32-trajectory tiles for small batches) and its instantiation with
    the small-argument tanh polynomial (fc_mpc_select_kernel; the automatic choice is restored afterwards)."""
    from forging_control_b200 import _native
    L = _native.lib()
    assert L.fc_mpc_select_kernel({"ffma": 1, "tcgen05": 2, "pair": 3, "replica": 4, "pair-precise": 5}[request.param]) == 0
    yield request.param
    assert L.fc_mpc_select_kernel(0) == 0


@pytest.mark.parametrize("name", W1_CASES)
def test_module_api_matches_reference_golden(dev, golden_cases, golden_weights, name):
    """The call sequence of NeuralNetwork.train_model (Functions.py:640-655) through the drop-in API."""
    C = golden_cases
    N, B, wd = (int(v) for v in C[f"{name}/meta"])
    sim, ctl, _, _ = _models(golden_weights, str(C[f"{name}/ctl"]), dev)
    X, Z = torch.tensor(C[f"{name}/X"]).to(dev), torch.tensor(C[f"{name}/Z"]).to(dev)
    out = ctl(X)
    loss, feats = fb.MPCLoss(prediction_horizon=N, alpha=ALPHA)(sim, ctl, X, out, Z, dev)
    loss.backward()
    for prec in ("f32", "f64"):
        ref = lambda k: C[f"{name}/{prec}/{k}"]
        assert abs(loss.item() - ref("loss")) / abs(ref("loss")) < TOL
        assert rel_max(feats["loss"].cpu().numpy(), ref("cost")) < TOL
        assert rel_max(feats["command"].cpu().numpy(), ref("command")) < TOL
        assert rel_max(feats["error"].cpu().numpy(), ref("error")) < TOL
        assert rel_max(feats["prediction"].cpu().numpy().reshape(B, N), ref("prediction")) < TOL
        assert feats["prediction"].shape == (B * N,)
        for p, k in ((ctl.fc_inp.weight, "fc_inp.weight"), (ctl.fc_inp.bias, "fc_inp.bias"),
                     (ctl.fc_out.weight, "fc_out.weight")):
            assert rel_max(p.grad.cpu().numpy(), ref("grad/" + k)) < TOL, (prec, k)
    assert ctl.fc_int.weight.grad is None and sim.lstm.weight_hh_l0.grad is None


def _assert_flips_sit_on_kinks(w, X, u0, Z, N, d, width_dim=1, noise=None):
    """The arbiter's own justification of every per-trajectory outlier: a trajectory may differ in d loss/d u0 only if
    the fp64 oracle has an argument of a ReLU / Hardtanh / constraint kink within the tolerance of the path (1e-5) of
    that kink -- there the derivative jumps and a value error <= 1e-5 decides the branch.  Anything else is a bug."""
    flipped = np.nonzero(d > TOL)[0]
    if len(flipped):
        margin = O.kink_margin(w, X[flipped], u0[flipped], Z[flipped], N, ALPHA, width_dim,
                               None if noise is None else noise[flipped])
        assert margin.max() < TOL, (flipped, margin, d[flipped])


def _kernel_vs_oracle_without_flips(run_kernel, run_oracle, X, Z, make_u0, max_rounds=4):
    """Batch-summed gradients in max-norm at the tolerance of the path.  A trajectory sitting on a kink (see above) moves
    the sums by O(1/B) of a per-trajectory gradient, which may exceed 1e-5 of the partly cancelling sums; such
    trajectories are replaced IN PLACE by a copy of a regular one (indices, and with them the counter-based noise of
    every other trajectory, stay what they were) and both sides are evaluated again."""
    X, Z = X.clone(), Z.clone()
    for _ in range(max_rounds):
        u0 = make_u0(X)
        r = run_kernel(X, Z, u0)
        out, g = run_oracle(X, Z, u0)
        d = np.abs(r["du0"].cpu().numpy() - g["u0"]) / np.abs(g["u0"]).max()
        flipped = np.nonzero(d > TOL)[0]
        if len(flipped) == 0:
            return r, out, g
        assert len(flipped) <= max(3, 5e-4 * len(d)), (len(flipped), d.max())
        safe = int(np.argmin(d))
        X[flipped], Z[flipped] = X[safe], Z[safe]
    raise AssertionError("trajectories keep landing on kinks")


def _seeded(B, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 3, generator=g) * 2 - 1, torch.rand(B, 10, 5, generator=g) * 2 - 1


@pytest.mark.parametrize("N,B,tag", [(5, 4096, "c0"), (10, 1000, "init"), (25, 500, "c3"), (10, 17797, "c0")])
def test_native_call_matches_fp64_oracle(dev, golden_weights, N, B, tag):
    """BASELINE config[1] (N=5, B=4096) and ragged / multi-wave batches against the fp64 oracle."""
    sim, ctl, lstm, fnn = _models(golden_weights, tag, dev)
    X, Z = _seeded(B, 1234 + N)
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    u0 = O.fnn_forward(w, X.double().numpy())[:, 0]
    wp = fb.pack_weights(sim, ctl)
    r = fb.mpc_loss_native(wp, X.to(dev), torch.tensor(u0, dtype=torch.float32).to(dev), Z.to(dev), N, ALPHA, True)
    out, g = O.mpc_loss_forward_backward(w, X.double().numpy(), u0.astype(np.float32).astype(np.float64),
                                         Z.double().numpy(), N, ALPHA)
    gl = r["gl"].cpu().numpy()
    assert abs(gl[250] - out["loss"]) / abs(out["loss"]) < TOL
    assert rel_max(r["cost"].cpu().numpy(), out["cost"]) < TOL
    assert rel_max(r["command"].cpu().numpy(), out["command"]) < TOL
    assert rel_max(r["error"].cpu().numpy(), out["error"]) < TOL
    assert rel_max(r["pred"].cpu().numpy(), out["prediction"]) < TOL
    # d loss/d u0 is a PER-TRAJECTORY gradient: ReLU / Hardtanh / constraint kinks make it discontinuous,
    # so a trajectory that sits within float32 rounding of a kink may legitimately take the other branch
    # than the fp64 arbiter (SURVEY.md section 7 "hard parts").  Require all but <= 0.05 % of the
    # trajectories within tolerance; the batch-summed gradients below are compared in max-norm.
    d = np.abs(r["du0"].cpu().numpy() - g["u0"]) / np.abs(g["u0"]).max()
    n_flip = int((d > TOL).sum())
    assert n_flip <= 5e-4 * B, (n_flip, d.max())
    _assert_flips_sit_on_kinks(w, X.double().numpy(), u0.astype(np.float32).astype(np.float64), Z.double().numpy(), N, d)
    if n_flip:
        # A flipped trajectory moves the batch-summed gradients by O(1/B) of a per-trajectory gradient,
        # which can exceed 1e-5 of the (partly cancelling) sums -- torch float32 vs float64 shows the same
        # effect (scripts/diag_precision.py).  Remove the flipped trajectories from BOTH sides: the kernel
        # is re-run without them, the oracle sums are corrected by linearity (sum_b = B * mean_b).
        keep = d <= TOL
        B2 = int(keep.sum())
        kt = torch.tensor(keep)
        r = fb.mpc_loss_native(wp, X[kt].contiguous().to(dev), torch.tensor(u0[keep], dtype=torch.float32).to(dev),
                               Z[kt].contiguous().to(dev), N, ALPHA, True)
        gl = r["gl"].cpu().numpy()
        _, gf = O.mpc_loss_forward_backward(w, X[~kt].double().numpy(), u0[~keep].astype(np.float32).astype(np.float64),
                                            Z[~kt].double().numpy(), N, ALPHA)
        g = {k: (g[k] * B - gf[k] * n_flip) / B2 for k in ("inp_w", "inp_b", "out_w")}
    assert rel_max(gl[:150].reshape(50, 3), g["inp_w"]) < TOL
    assert rel_max(gl[150:200], g["inp_b"]) < TOL
    assert rel_max(gl[200:250], g["out_w"][0]) < TOL
    assert np.all(gl[251:] == 0)


def test_trace_derived_inputs(dev, golden_weights, trace_windows):
    """Realistic state distribution (windows cut from the reference's own closed-loop data), tiled
    with a seeded +-1 % jitter."""
    sim, ctl, lstm, fnn = _models(golden_weights, "c0", dev)
    rng = np.random.default_rng(5)
    reps = 6
    X = np.tile(trace_windows["X"], (reps, 1)) * (1 + 0.01 * rng.uniform(-1, 1, (256 * reps, 3)))
    Z = np.tile(trace_windows["Z"], (reps, 1, 1)) * (1 + 0.01 * rng.uniform(-1, 1, (256 * reps, 10, 5)))
    X, Z = X.astype(np.float32), Z.astype(np.float32)
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    u0 = O.fnn_forward(w, X.astype(np.float64))[:, 0].astype(np.float32)
    r = fb.mpc_loss_native(fb.pack_weights(sim, ctl), torch.tensor(X).to(dev), torch.tensor(u0).to(dev),
                           torch.tensor(Z).to(dev), 10, ALPHA, True)
    out, g = O.mpc_loss_forward_backward(w, X.astype(np.float64), u0.astype(np.float64), Z.astype(np.float64), 10, ALPHA)
    gl = r["gl"].cpu().numpy()
    assert abs(gl[250] - out["loss"]) / abs(out["loss"]) < TOL
    assert rel_max(r["du0"].cpu().numpy(), g["u0"]) < TOL
    assert rel_max(gl[:150].reshape(50, 3), g["inp_w"]) < TOL
    assert rel_max(gl[200:250], g["out_w"][0]) < TOL


def test_forward_only_mode_and_no_grad(dev, golden_weights):
    sim, ctl, _, _ = _models(golden_weights, "c0", dev)
    X, Z = _seeded(333, 9)
    X, Z = X.to(dev), Z.to(dev)
    lf = fb.MPCLoss(10, ALPHA)
    loss_g, feats_g = lf(sim, ctl, X, ctl(X), Z, dev)
    with torch.no_grad():
        loss_n, feats_n = lf(sim, ctl, X, ctl(X), Z, dev)
    assert torch.equal(loss_g.detach(), loss_n) and torch.equal(feats_g["loss"], feats_n["loss"])
    assert not loss_n.requires_grad and loss_g.requires_grad


def test_upstream_gradient_scaling(dev, golden_weights):
    sim, ctl, _, _ = _models(golden_weights, "init", dev)
    X, Z = _seeded(240, 3)
    X, Z = X.to(dev), Z.to(dev)
    loss, _ = fb.MPCLoss(6, ALPHA)(sim, ctl, X, ctl(X), Z, dev)
    loss.backward()
    g1 = ctl.fc_inp.weight.grad.clone()
    ctl.zero_grad()
    loss, _ = fb.MPCLoss(6, ALPHA)(sim, ctl, X, ctl(X), Z, dev)
    (3.0 * loss).backward()
    assert rel_max(ctl.fc_inp.weight.grad.cpu().numpy(), 3.0 * g1.cpu().numpy()) < 1e-6


def test_full_size_properties(dev, golden_weights):
    """BASELINE config[2] size (N=25, B=65536): properties that do not need the oracle at that size.
    (i) a batch made of 512 tiled copies of a 128-trajectory block gives, for every copy, the costs of
    the block, and the loss / gradients of the block (mean over copies = block mean); (ii) the
    128-block itself is checked against the fp64 oracle; (iii) loss == mean(cost) and
    cost >= command + error (the constraint term is non-negative)."""
    N, B0, reps = 25, 128, 512
    sim, ctl, lstm, fnn = _models(golden_weights, "c0", dev)
    X0, Z0 = _seeded(B0, 77)
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    u0_0 = torch.tensor(O.fnn_forward(w, X0.double().numpy())[:, 0], dtype=torch.float32)
    wp = fb.pack_weights(sim, ctl)
    X, Z, u0 = X0.repeat(reps, 1).to(dev), Z0.repeat(reps, 1, 1).to(dev), u0_0.repeat(reps).to(dev)
    r = fb.mpc_loss_native(wp, X, u0, Z, N, ALPHA, True)
    cost = r["cost"].view(reps, B0)
    assert torch.equal(cost, cost[0:1].expand(reps, B0))                       # deterministic per trajectory
    assert torch.equal(r["pred"].view(reps, B0, N), r["pred"].view(reps, B0, N)[0:1].expand(reps, B0, N))
    du0 = r["du0"].view(reps, B0)
    assert torch.equal(du0, du0[0:1].expand(reps, B0))
    out, g = O.mpc_loss_forward_backward(w, X0.double().numpy(), u0_0.double().numpy(), Z0.double().numpy(), N, ALPHA)
    gl = r["gl"].cpu().numpy()
    assert abs(gl[250] - out["loss"]) / abs(out["loss"]) < TOL
    assert rel_max(cost[0].cpu().numpy(), out["cost"]) < TOL
    assert rel_max(du0[0].cpu().numpy() * reps, g["u0"]) < TOL                  # 1/(N*B) scaling
    assert rel_max(gl[:150].reshape(50, 3), g["inp_w"]) < TOL
    assert rel_max(gl[150:200], g["inp_b"]) < TOL
    assert rel_max(gl[200:250], g["out_w"][0]) < TOL
    assert abs(gl[250] - r["cost"].double().mean().item()) / gl[250] < 1e-6
    assert bool(torch.all(r["cost"] >= (r["command"] + r["error"]) * (1 - 1e-6)))


def test_sharded_call_sums_to_full_batch(dev, golden_weights):
    """The multi-GPU contract on one device: two shards evaluated with B_global sum to the full batch."""
    sim, ctl, _, _ = _models(golden_weights, "c0", dev)
    X, Z = _seeded(1001, 21)
    X, Z = X.to(dev), Z.to(dev)
    u0 = ctl(X).detach().reshape(-1)
    wp = fb.pack_weights(sim, ctl)
    full = fb.mpc_loss_native(wp, X, u0, Z, 10, ALPHA, True)
    lo, hi = fb.shard_bounds(1001, 2, 0)
    a = fb.mpc_loss_native(wp, X[lo:hi].contiguous(), u0[lo:hi].contiguous(), Z[lo:hi].contiguous(), 10, ALPHA, True, 1001)
    b = fb.mpc_loss_native(wp, X[hi:].contiguous(), u0[hi:].contiguous(), Z[hi:].contiguous(), 10, ALPHA, True, 1001)
    s = (a["gl"] + b["gl"]).cpu().numpy()
    assert rel_max(s[:251], full["gl"].cpu().numpy()[:251]) < 2e-6
    assert torch.equal(torch.cat((a["du0"], b["du0"])), full["du0"])


def test_training_loop_descends(dev, golden_weights):
    """NeuralNetwork.train_model end to end (DataLoader -> H2D -> loss -> backward -> AdamW)."""
    from torch.utils.data import DataLoader, TensorDataset
    sim, ctl, _, _ = _models(golden_weights, "init", dev)
    X, Z = _seeded(600, 11)
    loader = DataLoader(TensorDataset(X, torch.zeros(600, 1), Z), batch_size=150, shuffle=False)
    opt = torch.optim.AdamW(ctl.parameters(), lr=1e-3)
    lf = fb.MPCLoss(10, ALPHA)
    first, feats = fb.NeuralNetwork.train_model(loader, sim, ctl, lf, opt, dev)
    for _ in range(4):
        last, feats = fb.NeuralNetwork.train_model(loader, sim, ctl, lf, opt, dev)
    assert last < first
    assert feats["loss"].shape == (600,) and feats["prediction"].shape == (6000,)


def test_c_abi_error_codes(dev, golden_weights):
    from forging_control_b200 import _native
    L = _native.lib()
    sim, ctl, _, _ = _models(golden_weights, "c0", dev)
    wp = fb.pack_weights(sim, ctl)
    X, Z = _seeded(10, 1)
    X, Z = X.to(dev), Z.to(dev)
    u0 = torch.zeros(10, device=dev)
    o = [torch.empty(10, device=dev) for _ in range(4)]
    pred, gl, work = torch.empty(10, 10, device=dev), torch.empty(256, device=dev), torch.empty(64, dtype=torch.uint8, device=dev)
    p = _native.ptr
    rc = L.fc_mpc_loss(p(X), p(u0), p(Z), p(wp), 10, 10, 20.0, 10, 1, p(o[0]), p(o[1]), p(o[2]), p(pred), p(o[3]),
                       p(gl), p(work), 64, 0)
    assert rc == -5 and b"workspace" in L.fc_last_error()
    rc = L.fc_mpc_loss(p(X), p(u0), p(Z), p(wp), 0, 10, 20.0, 10, 1, p(o[0]), p(o[1]), p(o[2]), p(pred), p(o[3]),
                       p(gl), p(work), 64, 0)
    assert rc == -1
    rc = L.fc_mpc_loss(0, p(u0), p(Z), p(wp), 10, 10, 20.0, 10, 1, p(o[0]), p(o[1]), p(o[2]), p(pred), p(o[3]),
                       p(gl), p(work), 64, 0)
    assert rc == -3


def test_enable_noise_matches_oracle_with_the_same_noise(dev, golden_weights):
    """enable_noise=True (UL/Functions.py:1400-1402, :1438-1440) through fc_mpc_loss_noise: the counter-based generator
    is restated in the oracle (philox_normal4), so the noisy roll-out is checked like the clean one; runs for the three
    kernels (autouse fixture) with the same seed -> same noise."""
    sim, ctl, lstm, fnn = _models(golden_weights, "c0", dev)
    B, N, seed, std = 700, 6, 0xC0FFEE1234, 0.01
    g = torch.Generator().manual_seed(11)
    X = (torch.rand(B, 3, generator=g) * 2 - 1)
    Z = (torch.rand(B, 10, 5, generator=g) * 2 - 1)
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    noise = std * O.philox_normal4(seed, B, N)
    wp = fb.pack_weights(sim, ctl)

    def make_u0(Xc):
        with torch.no_grad():
            return ctl(Xc.to(dev)).reshape(-1).contiguous()

    def run_kernel(Xc, Zc, u0):
        return fb.mpc_loss_native(wp, Xc.to(dev), u0, Zc.to(dev), N, ALPHA, True, None, std, seed)

    def run_oracle(Xc, Zc, u0):
        return O.mpc_loss_forward_backward(w, Xc.double().numpy(), u0.double().cpu().numpy(), Zc.double().numpy(), N, ALPHA, noise=noise)

    # first the batch as given: values at the tolerance, per-trajectory outliers only on kinks of the arbiter
    u0 = make_u0(X)
    r = run_kernel(X, Z, u0)
    out, gr = run_oracle(X, Z, u0)
    clean, _ = O.mpc_loss_forward_backward(w, X.double().numpy(), u0.double().cpu().numpy(), Z.double().numpy(), N, ALPHA)
    gl = r["gl"].cpu().numpy()
    assert abs(gl[250] - out["loss"]) / abs(out["loss"]) < TOL
    assert abs(out["loss"] - clean["loss"]) / abs(clean["loss"]) > 1e-5
    assert rel_max(r["cost"].cpu().numpy(), out["cost"]) < TOL
    assert rel_max(r["pred"].cpu().numpy(), out["prediction"]) < TOL
    d = np.abs(r["du0"].cpu().numpy() - gr["u0"]) / np.abs(gr["u0"]).max()
    assert (d > TOL).sum() <= 3
    _assert_flips_sit_on_kinks(w, X.double().numpy(), u0.double().cpu().numpy(), Z.double().numpy(), N, d, noise=noise)
    # then the batch-summed gradients at the tolerance of the path
    r, out, gr = _kernel_vs_oracle_without_flips(run_kernel, run_oracle, X, Z, make_u0)
    gl = r["gl"].cpu().numpy()
    assert rel_max(gl[:150].reshape(50, 3), gr["inp_w"]) < TOL and rel_max(gl[150:200], gr["inp_b"]) < TOL
    assert rel_max(gl[200:250], gr["out_w"][0]) < TOL


def test_enable_noise_module_api_is_seeded_by_torch(dev, golden_weights):
    sim, ctl, _, _ = _models(golden_weights, "c0", dev)
    g = torch.Generator().manual_seed(3)
    X = (torch.rand(64, 3, generator=g) * 2 - 1).to(dev)
    Z = (torch.rand(64, 10, 5, generator=g) * 2 - 1).to(dev)
    loss_fn = fb.MPCLoss(prediction_horizon=5, alpha=ALPHA)
    def run(seed):
        torch.manual_seed(seed)
        loss, feats = loss_fn(sim, ctl, X, ctl(X), Z, dev, enable_noise=True)
        loss.backward()
        return loss.item()
    clean = loss_fn(sim, ctl, X, ctl(X), Z, dev)[0].item()
    a, b, c = run(1), run(1), run(2)
    assert a == b and a != c and a != clean
    assert abs(a - clean) / abs(clean) < 0.2         # 1 % noise on the surrogate outputs: a perturbation, not a different loss


def test_width_dim_2_matches_reference_golden(dev, golden_cases, golden_weights):
    """FNNModel(width_dim=2) (UL/Functions.py:261-289: one weight-shared fc_int + ReLU repeat): the reference's own
    outputs and gradients, fc_int included, through the drop-in API.  Every kernel selection routes this controller
    to the one-tile tcgen05 kernel."""
    C, name = golden_cases, "n6_b7_w2"
    N, B, wd = (int(v) for v in C[f"{name}/meta"])
    assert wd == 2
    lstm, fnn = state_dicts(golden_weights, "w2")
    sim = fb.LSTMModel(5, 50, 4, 3)
    sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()}, strict=True)
    ctl = fb.FNNModel(3, 50, 1, 2)
    ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()}, strict=True)
    sim, ctl = sim.to(dev), ctl.to(dev)
    X, Z = torch.tensor(C[f"{name}/X"]).to(dev), torch.tensor(C[f"{name}/Z"]).to(dev)
    loss, feats = fb.MPCLoss(prediction_horizon=N, alpha=ALPHA)(sim, ctl, X, ctl(X), Z, dev)
    loss.backward()
    for prec in ("f32", "f64"):
        ref = lambda k: C[f"{name}/{prec}/{k}"]
        assert abs(loss.item() - ref("loss")) / abs(ref("loss")) < TOL
        assert rel_max(feats["loss"].cpu().numpy(), ref("cost")) < TOL
        assert rel_max(feats["prediction"].cpu().numpy().reshape(B, N), ref("prediction")) < TOL
        for prm, k in ((ctl.fc_inp.weight, "fc_inp.weight"), (ctl.fc_inp.bias, "fc_inp.bias"), (ctl.fc_int.weight, "fc_int.weight"),
                       (ctl.fc_int.bias, "fc_int.bias"), (ctl.fc_out.weight, "fc_out.weight")):
            assert rel_max(prm.grad.cpu().numpy(), ref("grad/" + k)) < TOL, (prec, k)


@pytest.mark.parametrize("width,B,N", [(2, 300, 5), (3, 1000, 4)])
def test_wide_controller_matches_oracle(dev, golden_weights, width, B, N):
    """width_dim 2 and 3 at multi-tile batch sizes against the fp64 oracle (which is pinned on the reference's width_dim=2
    case, tests/test_oracle_mpc.py)."""
    lstm, fnn = state_dicts(golden_weights, "w2")
    sim = fb.LSTMModel(5, 50, 4, 3)
    sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()}, strict=True)
    ctl = fb.FNNModel(3, 50, 1, width)
    ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()}, strict=True)
    sim, ctl = sim.to(dev), ctl.to(dev)
    g = torch.Generator().manual_seed(width * 100 + B)
    X = (torch.rand(B, 3, generator=g) * 2 - 1)
    Z = (torch.rand(B, 10, 5, generator=g) * 2 - 1)
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    wp = fb.pack_weights(sim, ctl)
    iw, ib = ctl.fc_int.weight.detach().contiguous(), ctl.fc_int.bias.detach().contiguous()

    def make_u0(Xc):
        with torch.no_grad():
            return ctl(Xc.to(dev)).reshape(-1).contiguous()

    def run_kernel(Xc, Zc, u0):
        return fb.mpc_loss_native(wp, Xc.to(dev), u0, Zc.to(dev), N, ALPHA, True, None, 0.0, 0, iw, ib, width)

    def run_oracle(Xc, Zc, u0):
        return O.mpc_loss_forward_backward(w, Xc.double().numpy(), u0.double().cpu().numpy(), Zc.double().numpy(), N, ALPHA, width_dim=width)

    u0 = make_u0(X)
    r = run_kernel(X, Z, u0)
    out, gr = run_oracle(X, Z, u0)
    gl = r["gl"].cpu().numpy()
    assert abs(gl[250] - out["loss"]) / abs(out["loss"]) < TOL
    assert rel_max(r["cost"].cpu().numpy(), out["cost"]) < TOL
    assert rel_max(r["pred"].cpu().numpy(), out["prediction"]) < TOL
    d = np.abs(r["du0"].cpu().numpy() - gr["u0"]) / np.abs(gr["u0"]).max()
    assert (d > TOL).sum() <= 3
    _assert_flips_sit_on_kinks(w, X.double().numpy(), u0.double().cpu().numpy(), Z.double().numpy(), N, d, width_dim=width)
    r, out, gr = _kernel_vs_oracle_without_flips(run_kernel, run_oracle, X, Z, make_u0)
    gl, gw = r["gl"].cpu().numpy(), r["gl_wide"].cpu().numpy()
    for got, key in ((gl[:150].reshape(50, 3), "inp_w"), (gl[150:200], "inp_b"), (gl[200:250], "out_w"),
                     (gw[:2500].reshape(50, 50), "int_w"), (gw[2500:2550], "int_b")):
        assert rel_max(got, np.asarray(gr[key]).reshape(got.shape)) < TOL, key


def test_fused_controller_forward_backward_matches_oracle(dev, golden_weights):
    """FNNModel.forward on CUDA = fc_fnn_forward / fc_fnn_backward (Functions.py:261-289, the u_0 path of :643/:655):
    against the fp64 oracle with an arbitrary upstream gradient, ragged batch, saturated and dead units included."""
    for tag, B in (("c0", 1), ("init", 1000), ("c3", 20001)):
        _, fnn = state_dicts(golden_weights, tag)
        ctl = fb.FNNModel(3, 50, 1, 1)
        fnn = dict(fnn)
        if tag != "c0":
            fnn["fc_out.weight"] = fnn["fc_out.weight"] * 8.0  # make the Hardtanh saturate for part of the batch
        ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()}, strict=True)
        ctl = ctl.to(dev)
        g = torch.Generator().manual_seed(B)
        X = (torch.rand(B, 3, generator=g) * 6 - 3)
        d = torch.randn(B, 1, generator=g)
        u = ctl(X.to(dev))
        assert u.shape == (B, 1) and u.grad_fn is not None and type(u.grad_fn).__name__.startswith("_FusedFNN")
        u.backward(d.to(dev))
        w = O.weights_from_state_dicts(state_dicts(golden_weights, tag)[0], fnn, np.float64)
        u_o, kept = O.fnn_forward(w, X.double().numpy(), 1, keep=True)
        grads = {k: np.zeros_like(w[k]) for k in ("inp_w", "inp_b", "out_w")}
        O.fnn_backward(w, kept, d.double().numpy(), grads, 1)
        assert rel_max(u.detach().cpu().numpy(), np.asarray(u_o).reshape(B, 1)) <= TOL
        assert B == 1 or 0 < int((u.detach().abs() >= 1).sum()) < B
        for name, p in (("inp_w", ctl.fc_inp.weight), ("inp_b", ctl.fc_inp.bias), ("out_w", ctl.fc_out.weight)):
            assert rel_max(p.grad.cpu().numpy(), grads[name].reshape(p.shape)) <= TOL, (tag, name)
        assert ctl.fc_int.weight.grad is None
