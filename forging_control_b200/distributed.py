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
    return torch.cat([p.grad.reshape(-1) for p in params])


def allreduce_loss_and_grads(loss: torch.Tensor, params, group=None) -> torch.Tensor:
    """Sum the flat ``[grads | loss]`` buffer over ranks in one collective and scatter the result
    back into ``p.grad``.  Returns the global loss (0-d tensor)."""
    params = [p for p in params if p.grad is not None]
    flat = torch.cat([p.grad.reshape(-1) for p in params] + [loss.detach().reshape(1)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat[off]


def sharded_training_step(loss_function, simulator, controller, X, Z, device, global_batch: int, group=None):
    """One data-parallel MPC-loss step on this rank's shard (X [b,3], Z [b,10,5] already on
    ``device``).  Leaves the globally reduced gradients in ``controller.parameters()`` and returns the
    global loss.  Mirrors the body of ``NeuralNetwork.train_model`` (Functions.py:640-655)."""
    loss_function.global_batch = int(global_batch)
    for p in controller.parameters():
        p.grad = None
    u0 = controller(X)
    loss, _ = loss_function(simulator, controller, X, u0, Z, device)
    loss.backward()
    params = [controller.fc_inp.weight, controller.fc_inp.bias, controller.fc_out.weight]
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
