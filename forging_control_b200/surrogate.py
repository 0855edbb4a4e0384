"""Surrogate training path and device-side optimizer.

Mirror of the training part of the reference ``Unsupervised Learning/Model_NN/Functions.py`` (citations relative to
that directory):

* ``lstm_window(model, x)``            what ``LSTMModel.forward`` (Functions.py:313-340) computes, as ONE forward
  kernel (``fc_lstm_window_fwd``) whose autograd ``backward`` is ONE reverse-sweep kernel with the weight gradients
  (``fc_lstm_window_bwd``) instead of cuDNN's LSTM (which runs TF32 on this GPU);
  ``forging_control_b200.LSTMModel.forward`` calls it for CUDA inputs of the reference configuration.
* ``SurrogateNeuralNetwork``           ``NeuralNetwork.train_model`` / ``validate_model`` / ``train_loop``
  (Functions.py:520-612, 754-822) with the reference signatures; the running loss stays on the device (one read-back
  per epoch instead of ``loss.item()`` per step, :566).
* ``DeviceAdamW``                      ``torch.optim.AdamW`` (Main.py:230; UL/Main.py:195) whose ``step`` is one launch
  of ``fc_adamw_step`` over all parameters; same hyper-parameters, same ``state_dict`` layout.

There is no CPU fallback: CUDA tensors of the supported configuration run the sm_100a kernels and everything else
raises (``lstm_window``) or is refused (``DeviceAdamW`` on CPU parameters).
"""
from __future__ import annotations

import ctypes
import logging
import weakref
from time import time

import torch
import torch.nn as nn

from . import _native

logger = logging.getLogger(__name__)

LOOKBACK = 10
_PACK_CACHE: dict = {}


def _lstm_params(model):
    l = model.lstm
    return [l.weight_ih_l0, l.weight_hh_l0, l.weight_ih_l1, l.weight_hh_l1, l.weight_ih_l2, l.weight_hh_l2,
            model.fc.weight, model.fc.bias]


def supported(model, x) -> bool:
    """True when (model, x) is the reference surrogate configuration on a CUDA device."""
    l = getattr(model, "lstm", None)
    fc = getattr(model, "fc", None)
    return (l is not None and fc is not None and x.is_cuda and x.dim() == 3 and x.shape[1] == LOOKBACK and x.shape[2] == 5
            and (l.input_size, l.hidden_size, l.num_layers) == (5, 50, 3) and not l.bias and l.batch_first
            and not l.bidirectional and tuple(fc.weight.shape) == (4, 50) and fc.bias is not None
            and l.weight_ih_l0.is_cuda and l.weight_ih_l0.dtype == torch.float32)


def _pack(ws) -> torch.Tensor:
    """Kernel weight images of the six LSTM matrices, re-packed when a parameter changed (``_version``).  A re-pack
    always writes a FRESH buffer: an autograd graph that still holds the previous images for its backward keeps them."""
    dev = ws[0].device
    ident = (dev.index, id(ws[0]))
    key = tuple((w.data_ptr(), w._version) for w in ws[:6])
    slot = _PACK_CACHE.get(ident)
    if slot is not None and slot[0] == key and all(r() is w for r, w in zip(slot[2], ws[:6])):
        return slot[1]
    L = _native.lib()
    buf = torch.empty(int(L.fc_lstm_train_pack_floats()), dtype=torch.float32, device=dev)
    wc = [w.detach().contiguous() for w in ws[:6]]
    with torch.cuda.device(dev):
        rc = L.fc_lstm_train_pack(*[_native.ptr(w) for w in wc], _native.ptr(buf), _native.stream_ptr(dev))
    _native.check(rc, "fc_lstm_train_pack")
    for k in [k for k, v in _PACK_CACHE.items() if any(r() is None for r in v[2])]:    # models that are gone
        del _PACK_CACHE[k]
    _PACK_CACHE[ident] = (key, buf, [weakref.ref(w) for w in ws[:6]])
    return buf


class _LstmWindow(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, save, *ws):
        dev = x.device
        B = x.shape[0]
        L = _native.lib()
        xc = x.detach().to(torch.float32).contiguous()
        pack = _pack(ws)
        fc_w, fc_b = ws[6].detach().contiguous(), ws[7].detach().contiguous()
        out = torch.empty(B, 4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            # save: records + hidden sequences for the backward call; large batches also need a per-CTA scratch without it
            nbytes = int(L.fc_lstm_window_workspace_bytes(B, int(save)))
            if nbytes == 0:
                raise RuntimeError("fc_lstm_window_workspace_bytes failed: " + L.fc_last_error().decode())
            work = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            rc = L.fc_lstm_window_fwd(_native.ptr(xc), _native.ptr(pack), _native.ptr(fc_w), _native.ptr(fc_b), B, int(save),
                                      _native.ptr(out), _native.ptr(work), nbytes, _native.stream_ptr(dev))
        _native.check(rc, "fc_lstm_window_fwd")
        ctx.saved = save
        ctx.path = int(L.fc_lstm_train_path_for(B))      # the backward may run on an autograd worker thread: same path there
        if save:
            ctx.save_for_backward(xc, pack, fc_w, work)
            ctx.shapes = [w.shape for w in ws]
        return out

    @staticmethod
    def backward(ctx, d_out):
        if not ctx.saved:
            raise RuntimeError("lstm_window: backward called but the forward ran without gradients")
        xc, pack, fc_w, work = ctx.saved_tensors
        dev = xc.device
        B = xc.shape[0]
        L = _native.lib()
        d = d_out.detach().to(torch.float32).contiguous()
        grads = [torch.empty(s, dtype=torch.float32, device=dev) for s in ctx.shapes]
        with torch.cuda.device(dev):
            prev = int(L.fc_lstm_train_path_for(B))
            L.fc_lstm_train_select_path(ctx.path)
            try:
                rc = L.fc_lstm_window_bwd(_native.ptr(xc), _native.ptr(d), _native.ptr(pack), _native.ptr(fc_w), B,
                                          _native.ptr(work), work.numel(), *[_native.ptr(g) for g in grads],
                                          _native.stream_ptr(dev))
            finally:
                if prev != ctx.path:
                    L.fc_lstm_train_select_path(0)
        _native.check(rc, "fc_lstm_window_bwd")
        return (None, None, *grads)


def lstm_window(model: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """``LSTMModel.forward`` on the sm_100a kernels: x [B,10,5] CUDA -> [B,4], differentiable w.r.t. the eight
    surrogate parameters (not w.r.t. x: the reference never needs it on this path)."""
    if not supported(model, x):
        raise NotImplementedError("lstm_window: supports LSTMModel(5, 50, 4, 3, bias=False) with CUDA float32 weights and "
                                  "x [B,10,5] on the same CUDA device only (no CPU path)")
    if x.requires_grad:
        raise NotImplementedError("lstm_window: gradients w.r.t. the input window are not provided on the training path")
    if x.shape[0] == 0:
        raise ValueError("lstm_window: empty batch")
    ws = _lstm_params(model)
    save = torch.is_grad_enabled() and any(w.requires_grad for w in ws)     # grad mode is off inside Function.forward
    return _LstmWindow.apply(x, bool(save), *ws)


# ----------------------------------------------------------------------------------------------
# optimizer
# ----------------------------------------------------------------------------------------------
class DeviceAdamW(torch.optim.AdamW):
    """``torch.optim.AdamW`` with ``step()`` as one ``fc_adamw_step`` launch per <= 8 parameter tensors (the controller has
    3 live tensors, the surrogate 8).  Constructor, hyper-parameters, ``param_groups`` and ``state_dict`` are torch's
    (state entries ``step`` / ``exp_avg`` / ``exp_avg_sq``), so checkpoints interchange with the stock optimizer."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, maximize=False):
        if amsgrad or maximize:
            raise NotImplementedError("DeviceAdamW: amsgrad / maximize are not used by the reference and not implemented")
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _native.lib()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            beta1, beta2 = group["betas"]
            step_no = None
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.grad.is_sparse:
                    raise RuntimeError("DeviceAdamW: parameters must be contiguous float32 CUDA tensors (no CPU path)")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                s = int(st["step"].item())
                if step_no is None:
                    step_no = s
                elif s != step_no:
                    raise RuntimeError("DeviceAdamW: parameters of one group must share the step count")
            dev = ps[0].device
            for i in range(0, len(ps), 8):
                chunk = ps[i:i + 8]
                n = len(chunk)
                gs = [p.grad.contiguous() for p in chunk]
                arr = ctypes.c_void_p * n
                with torch.cuda.device(dev):
                    rc = L.fc_adamw_step(n, arr(*[p.data_ptr() for p in chunk]), arr(*[g.data_ptr() for g in gs]),
                                         arr(*[self.state[p]["exp_avg"].data_ptr() for p in chunk]),
                                         arr(*[self.state[p]["exp_avg_sq"].data_ptr() for p in chunk]),
                                         (ctypes.c_int * n)(*[p.numel() for p in chunk]), step_no, float(group["lr"]),
                                         float(beta1), float(beta2), float(group["eps"]), float(group["weight_decay"]),
                                         float(grad_scale), _native.stream_ptr(dev))
                _native.check(rc, "fc_adamw_step")
                for p in chunk:      # updated through the raw pointer: bump the version counter the weight-image caches key on
                    torch.autograd.graph.increment_version(p)
        return loss


# ----------------------------------------------------------------------------------------------
# training orchestration (surrogate variant of NeuralNetwork)
# ----------------------------------------------------------------------------------------------
class SurrogateNeuralNetwork:
    """``NeuralNetwork`` of ``Model_NN/Functions.py`` (training part), same signatures."""

    @staticmethod
    def train_model(data_loader, model, loss_function, optimizer, device):
        """One epoch, Functions.py:520-569."""
        model.train()
        running = None
        for X, y in data_loader:
            X, y = X.to(device, non_blocking=True), y.to(device, non_blocking=True)
            optimizer.zero_grad()
            output = model(X, device)
            loss = loss_function(output, y.squeeze())
            loss.backward()
            optimizer.step()
            running = loss.detach().double() if running is None else running + loss.detach().double()
        return (running.item() if running is not None else 0.0) / len(data_loader)

    @staticmethod
    def validate_model(data_loader, model, loss_function, device):
        """Functions.py:572-612."""
        model.eval()
        total = None
        with torch.no_grad():
            for X, y in data_loader:
                X, y = X.to(device), y.to(device)
                v = loss_function(model(X, device), y.squeeze()).double()
                total = v if total is None else total + v
        return (total.item() if total is not None else 0.0) / len(data_loader)

    @staticmethod
    def train_loop(model, train_loader, val_loader, loss_function, optimizer, n_epochs, device):
        """Functions.py:754-822."""
        vec_t, vec_v = [], []
        t0 = time()
        for epoch in range(n_epochs):
            t_loss = SurrogateNeuralNetwork.train_model(train_loader, model, loss_function, optimizer, device)
            v_loss = SurrogateNeuralNetwork.validate_model(val_loader, model, loss_function, device)
            vec_t.append(t_loss)
            vec_v.append(v_loss)
            logger.info(f"[{100 * (epoch + 1) / n_epochs:.1f}%] Training loss: {t_loss:.4f},  Validation loss: {v_loss:.4f}")
        comp_time = time() - t0
        logger.info(f"Total time: {comp_time:.2f}s.")
        return model, vec_t, vec_v, comp_time
