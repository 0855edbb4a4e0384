"""Data-parallel sharding of the hot path over the GPUs of one box.

Trajectories are independent through the forward roll-out and the reverse sweep; the only coupling
is the mean over the batch.  Every rank runs the fused kernel on a contiguous shard with
``global_batch`` = the total batch (so its partial loss / gradients are already scaled by
``1/B_global``) and ONE all-reduce(sum) over the flat ``[controller gradients | loss]`` buffer
(251 floats; NCCL over NVLink on GPUs, gloo in the CPU tests) makes every rank hold the
single-GPU result.  The closed-loop roll-out needs no collective at all.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous split of ``total`` trajectories: the first ``total % world_size`` ranks get one more."""
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def flatten_grads(params) -> torch.Tensor:
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])


def allreduce_loss_and_grads(loss: torch.Tensor, params, group=None) -> torch.Tensor:
    """Sum the flat ``[grads | loss]`` buffer over ranks in one collective and scatter the result
    back into ``p.grad``.  Returns the global loss (0-d tensor).  The buffer layout depends on ``params`` only: a
    parameter without a gradient on this rank (e.g. an empty shard) contributes zeros, so every rank issues the same
    collective."""
    params = [p for p in params if p.requires_grad]
    flat = torch.cat([flatten_grads(params), loss.detach().reshape(1).to(params[0].dtype)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[off:off + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat[off]


class FlatGradBucket:
    """One flat ``[grads of every trainable parameter | loss]`` device buffer whose slices ARE the parameters' ``.grad``
    tensors: autograd accumulates into it in place, so a data-parallel step is ``zero()`` (one memset), backward, one
    tiny copy of the loss and ONE all-reduce -- no ``cat`` / ``copy_`` glue around the collective."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradBucket: no trainable parameters")
        p0 = self.params[0]
        self.flat = torch.zeros(sum(p.numel() for p in self.params) + 1, dtype=p0.dtype, device=p0.device)
        self.attach()

    def attach(self):
        off = 0
        for p in self.params:
            n = p.numel()
            v = self.flat[off:off + n].view_as(p)
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v
            off += n

    def zero(self):
        self.attach()               # optimizer.zero_grad(set_to_none=True) drops the views
        self.flat.zero_()

    def allreduce(self, loss: torch.Tensor, group=None) -> torch.Tensor:
        self.flat[-1].copy_(loss.detach())
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self.flat[-1]


def _trainable(controller):
    """Every parameter the MPC loss produces a gradient for: ``fc_inp.*``, ``fc_out.weight`` and, for
    ``width_dim > 1`` controllers, the weight-shared hidden layer ``fc_int.*`` (FNNModel.forward, Functions.py:277-283)."""
    ps = [controller.fc_inp.weight, controller.fc_inp.bias, controller.fc_out.weight]
    if int(getattr(controller, "width_dim", 1)) > 1:
        ps += [controller.fc_int.weight, controller.fc_int.bias]
    return [p for p in ps if p is not None and p.requires_grad]


def sharded_training_step(loss_function, simulator, controller, X, Z, device, global_batch: int, group=None, bucket=None):
    """One data-parallel MPC-loss step on this rank's shard (X [b,3], Z [b,10,5] already on
    ``device``).  Leaves the globally reduced gradients in ``controller.parameters()`` and returns the
    global loss.  Mirrors the body of ``NeuralNetwork.train_model`` (Functions.py:640-655).  An empty shard
    (``b == 0``, more ranks than trajectories) contributes zeros to the same collective.  ``bucket`` (a
    ``FlatGradBucket`` over ``_trainable(controller)``) removes the flatten / scatter copies around the all-reduce."""
    loss_function.global_batch = int(global_batch)
    params = _trainable(controller)
    if bucket is not None:
        bucket.zero()
    else:
        for p in controller.parameters():
            p.grad = None
    if X.shape[0] > 0:
        u0 = controller(X)
        loss, _ = loss_function(simulator, controller, X, u0, Z, device)
        loss.backward()
    else:
        loss = torch.zeros((), dtype=torch.float32, device=params[0].device)
    if bucket is not None:
        return bucket.allreduce(loss, group)
    return allreduce_loss_and_grads(loss, params, group)


def sharded_surrogate_step(model, loss_function, X, y, device, global_batch: int, group=None):
    """One data-parallel surrogate-training step on this rank's shard (X [b,10,5], y [b,1,4] on
    ``device``): body of the surrogate ``train_model`` (Model_NN/Functions.py:548-560) with the batch mean of
    ``loss_function`` (``nn.MSELoss``) re-weighted by ``b / global_batch`` and ONE all-reduce(sum) over the flat
    ``[51 204 gradients | loss]`` buffer.  Leaves the global gradients in ``model.parameters()``, returns the global loss."""
    for p in model.parameters():
        p.grad = None
    loss = loss_function(model(X, device), y.squeeze(1)) * (X.shape[0] / float(global_batch))
    loss.backward()
    return allreduce_loss_and_grads(loss, list(model.parameters()), group)
