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
