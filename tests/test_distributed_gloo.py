"""CPU, world_size 2, gloo: the data-parallel path.  Each rank evaluates its contiguous shard with
the mean taken over the GLOBAL batch (the oracle stands in for the kernel here -- it is the checker
of the sharding arithmetic, not a product path) and ONE all-reduce over the flat [grads | loss] buffer
must reproduce the single-process result on the concatenated batch."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, REPO


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    for p in (REPO, os.path.join(REPO, "oracle")):
        sys.path.insert(0, p)
    import mpc_loss_oracle as O
    import forging_control_b200 as fb
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W = np.load(os.path.join(GOLDEN, "weights.npz"))
    C = np.load(os.path.join(GOLDEN, "mpc_loss_cases.npz"))
    lstm = {k[5:]: W[k] for k in W.files if k.startswith("lstm/")}
    fnn = {k[len("fnn_init/"):]: W[k] for k in W.files if k.startswith("fnn_init/")}
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    name = "n10_b33_init"
    X, Z, u0 = C[f"{name}/X"].astype(np.float64), C[f"{name}/Z"].astype(np.float64), C[f"{name}/f64/u0"]
    N, B = 10, len(X)
    lo, hi = fb.shard_bounds(B, world, rank)
    # shard evaluated with the global mean: scale the shard-mean results by b_local / B
    out, g = O.mpc_loss_forward_backward(w, X[lo:hi], u0[lo:hi], Z[lo:hi], N, 20.0)
    frac = (hi - lo) / B
    ctl = fb.FNNModel(3, 50, 1, 1).double()
    ctl.fc_inp.weight.grad = torch.tensor(g["inp_w"] * frac)
    ctl.fc_inp.bias.grad = torch.tensor(g["inp_b"] * frac)
    ctl.fc_out.weight.grad = torch.tensor(g["out_w"] * frac)
    loss = torch.tensor(out["loss"] * frac)
    params = [ctl.fc_inp.weight, ctl.fc_inp.bias, ctl.fc_out.weight]
    total = fb.allreduce_loss_and_grads(loss, params)
    if rank == 0:
        q.put((total.item(), [p.grad.numpy().copy() for p in params]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_process():
    import sys
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import mpc_loss_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    loss, grads = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    W = np.load(os.path.join(GOLDEN, "weights.npz"))
    C = np.load(os.path.join(GOLDEN, "mpc_loss_cases.npz"))
    lstm = {k[5:]: W[k] for k in W.files if k.startswith("lstm/")}
    fnn = {k[len("fnn_init/"):]: W[k] for k in W.files if k.startswith("fnn_init/")}
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    name = "n10_b33_init"
    out, g = O.mpc_loss_forward_backward(w, C[f"{name}/X"].astype(np.float64), C[f"{name}/f64/u0"],
                                         C[f"{name}/Z"].astype(np.float64), 10, 20.0)
    assert abs(loss - out["loss"]) / abs(out["loss"]) < 1e-12
    for a, b in zip(grads, (g["inp_w"], g["inp_b"], g["out_w"])):
        assert np.abs(a - b).max() / np.abs(b).max() < 1e-12
    assert abs(loss - float(C[f"{name}/f64/loss"])) / abs(loss) < 1e-12      # == the reference on the full batch


def _surrogate_worker(rank, world, port, q):
    import sys
    for p in (REPO, os.path.join(REPO, "oracle")):
        sys.path.insert(0, p)
    import forging_control_b200 as fb
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    C = np.load(os.path.join(GOLDEN, "surrogate_train_cases.npz"))
    W = np.load(os.path.join(GOLDEN, "weights.npz"))
    m = fb.LSTMModel(5, 50, 4, 3).double()            # CPU tensors: the stock nn.LSTM path carries the host-side logic
    m.load_state_dict({k[5:]: torch.tensor(W[k]).double() for k in W.files if k.startswith("lstm/")})
    X, y = torch.tensor(C["shipped_b37/X"]).double(), torch.tensor(C["shipped_b37/y"]).double()
    lo, hi = fb.shard_bounds(len(X), world, rank)
    total = fb.sharded_surrogate_step(m, torch.nn.MSELoss(), X[lo:hi], y[lo:hi], "cpu", global_batch=len(X))
    if rank == 0:
        q.put((total.item(), {k: p.grad.numpy().copy() for k, p in m.named_parameters()}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_surrogate_step_equals_the_reference_full_batch_step():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_surrogate_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    loss, grads = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    C = np.load(os.path.join(GOLDEN, "surrogate_train_cases.npz"))
    assert abs(loss - float(C["shipped_b37/f64/avg_loss"])) <= 1e-12 * abs(loss)
    for k, g in grads.items():
        ref = C[f"shipped_b37/f64/grad/{k}"]
        assert np.abs(g - ref).max() <= 1e-11 * np.abs(ref).max(), k


class _OracleLoss:
    """Checker standing in for the fused kernel on CPU tensors (tests only): the oracle's torch restatement of
    MPCLoss.forward with the kernel's sharding convention (mean over ``global_batch``)."""

    def __init__(self, tw, N, alpha, width_dim):
        self.tw, self.N, self.alpha, self.width_dim, self.global_batch = tw, N, alpha, width_dim, None

    def __call__(self, simulator, controller, X, u0, Z, device):
        import mpc_loss_oracle as O
        tw = dict(self.tw, inp_w=controller.fc_inp.weight, inp_b=controller.fc_inp.bias, out_w=controller.fc_out.weight,
                  int_w=controller.fc_int.weight, int_b=controller.fc_int.bias)
        loss = O.mpc_loss_torch(tw, X, u0[:, 0], Z, self.N, self.alpha, self.width_dim)[0]
        return loss * (X.shape[0] / float(self.global_batch)), {}


def _wide_worker(rank, world, port, q, n_use):
    import sys
    for p in (REPO, os.path.join(REPO, "oracle")):
        sys.path.insert(0, p)
    import mpc_loss_oracle as O
    import forging_control_b200 as fb
    from forging_control_b200.distributed import _trainable
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W = np.load(os.path.join(GOLDEN, "weights.npz"))
    C = np.load(os.path.join(GOLDEN, "mpc_loss_cases.npz"))
    lstm = {k[5:]: W[k] for k in W.files if k.startswith("lstm/")}
    fnn = {k[len("fnn_w2/"):]: W[k] for k in W.files if k.startswith("fnn_w2/")}
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    tw = {k: ([torch.tensor(a) for a in v] if isinstance(v, list) else torch.tensor(v)) for k, v in w.items()}
    ctl = fb.FNNModel(3, 50, 1, 2).double()
    ctl.load_state_dict({k: torch.tensor(v).double() for k, v in fnn.items()})
    name = "n6_b7_w2"
    X, Z = torch.tensor(C[f"{name}/X"][:n_use]).double(), torch.tensor(C[f"{name}/Z"][:n_use]).double()
    lo, hi = fb.shard_bounds(len(X), world, rank)
    bucket = fb.FlatGradBucket(_trainable(ctl))
    lf = _OracleLoss(tw, 6, 20.0, 2)
    total = fb.sharded_training_step(lf, None, ctl, X[lo:hi], Z[lo:hi], "cpu", global_batch=len(X), bucket=bucket)
    # the plain (bucket-less) path must issue the same collective and give the same answer
    total2 = fb.sharded_training_step(lf, None, ctl, X[lo:hi], Z[lo:hi], "cpu", global_batch=len(X))
    if rank == 0:
        q.put((total.item(), total2.item(), {k: p.grad.numpy().copy() for k, p in ctl.named_parameters() if p.grad is not None}))
    dist.barrier()
    dist.destroy_process_group()


def _run_wide(n_use):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_wide_worker, args=(r, 2, port, q, n_use)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_two_rank_step_reduces_the_hidden_layer_gradients_of_a_wide_controller():
    """width_dim = 2: fc_int.* gradients are part of the reduced buffer (they stayed rank-local before); pinned on the
    reference's own full-batch gradients of the same case."""
    loss, loss2, grads = _run_wide(7)
    C = np.load(os.path.join(GOLDEN, "mpc_loss_cases.npz"))
    assert abs(loss - float(C["n6_b7_w2/f64/loss"])) <= 1e-12 * abs(loss) and abs(loss - loss2) <= 1e-14
    for k in ("fc_inp.weight", "fc_inp.bias", "fc_int.weight", "fc_int.bias", "fc_out.weight"):
        ref = C[f"n6_b7_w2/f64/grad/{k}"]
        assert np.abs(grads[k] - ref).max() <= 1e-11 * np.abs(ref).max(), k


def test_empty_shard_joins_the_same_collective():
    """One trajectory over two ranks: rank 1 owns nothing and contributes zeros instead of hanging the all-reduce."""
    loss, loss2, grads = _run_wide(1)
    assert np.isfinite(loss) and abs(loss - loss2) <= 1e-14
    assert set(grads) == {"fc_inp.weight", "fc_inp.bias", "fc_int.weight", "fc_int.bias", "fc_out.weight"}
