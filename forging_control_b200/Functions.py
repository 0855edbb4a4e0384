"""Drop-in mirror of the hot-path part of the reference ``Unsupervised Learning/Functions.py``.

Same class names, constructor arguments, ``state_dict`` layout and call signatures as the
reference (file:line citations are relative to ``/root/reference/Unsupervised Learning/``):

* ``FNNModel``   (Functions.py:215-289)   controller 3 -> 50 -> (50)* -> 1, ReLU, Hardtanh
* ``LSTMModel``  (Functions.py:295-379)   surrogate ``nn.LSTM(5,50,3,batch_first,bias=False)`` + ``Linear(50,4)``
* ``MPCLoss``    (Functions.py:1336-1472) horizon roll-out loss -- here ONE fused sm_100a kernel that
  also runs the reverse-time sweep, so ``loss.backward()`` only scales pre-computed gradients
* ``NeuralNetwork.train_model`` / ``train_loop`` / ``validate_model`` (Functions.py:594-717, 825-923)
* ``NeuralNetwork.tvp_fun`` (Functions.py:926-966), ``NeuralNetwork.loop`` (Functions.py:1014-1289)
  -- the closed loop runs all ``N_traj`` trajectories as one batch of the RK4 plant kernel
* ``FeasibilityRecovery.NN_make_step`` (Functions.py:1560-1613), ``Data.get_scaler`` (:407-442)

Everything the hot path does not touch (datasets, plotting, NMPC teacher, IPOPT feasibility
recovery) stays in the reference; ``forging_control_b200.install(reference_module)`` swaps the
classes above into an imported reference module so that ``Main.py`` runs unchanged.

There is no CPU fallback: ``MPCLoss`` and ``NeuralNetwork.loop`` require CUDA tensors / a CUDA
device and raise otherwise.
"""
from __future__ import annotations

import logging
import random
import weakref
from time import time

import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _native
from .closed_loop import ClosedLoopTimer, closed_loop_rollout, tvp_reference_table

logger = logging.getLogger(__name__)

LOOKBACK = 10  # hard-coded in the reference loss, Functions.py:1434


def _maxabs_scale(scaler, what: str) -> np.ndarray:
    """``scale_`` of a fitted ``MaxAbsScaler`` (the reference's scalers, SL/results/scaler_*.pkl).  The reference calls
    ``transform`` / ``inverse_transform`` (Functions.py:1596-1604, :1198); the kernels take the per-feature divisor,
    which is the whole transform only for MaxAbs.  MinMax / Standard / Robust scalers (``Data.get_scaler``) carry an
    offset and a multiplier instead and are rejected rather than silently mis-scaled."""
    from sklearn.preprocessing import MaxAbsScaler
    if not isinstance(scaler, MaxAbsScaler):
        raise NotImplementedError(f"{what}: only MaxAbsScaler is supported by the CUDA path (got {type(scaler).__name__}); "
                                  "MinMax / Standard / Robust scalers have an offset the kernels do not apply")
    return np.asarray(scaler.scale_, dtype=np.float64)


# ----------------------------------------------------------------------------------------------
# controller
# ----------------------------------------------------------------------------------------------
class FNNModel(nn.Module):
    """Feed-forward controller, Functions.py:215-289 (same parameters / ``state_dict`` keys:
    ``fc_inp.*``, ``fc_int.*`` (weight-shared hidden layer, unused when ``width_dim == 1``),
    ``fc_out.weight``)."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, width_dim: int,
                 activation_fn=nn.ReLU, bias=True):
        super().__init__()
        self.width_dim = width_dim
        self.activation = activation_fn()
        self.constraint = nn.Hardtanh()
        self.fc_inp = nn.Linear(input_dim, hidden_dim, bias=bias)
        self.fc_int = nn.Linear(hidden_dim, hidden_dim, bias=bias)
        self.fc_out = nn.Linear(hidden_dim, output_dim, bias=False)
        for layer in (self.fc_inp, self.fc_int, self.fc_out):
            nn.init.xavier_normal_(layer.weight)
        nn.init.zeros_(self.fc_inp.bias)
        nn.init.zeros_(self.fc_int.bias)

    def _native_ok(self, x) -> bool:
        return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[0] > 0 and x.shape[1] == 3
                and not x.requires_grad
                and self.width_dim == 1 and isinstance(self.activation, nn.ReLU) and isinstance(self.constraint, nn.Hardtanh)
                and (self.constraint.min_val, self.constraint.max_val) == (-1.0, 1.0)
                and tuple(self.fc_inp.weight.shape) == (50, 3) and self.fc_inp.bias is not None
                and tuple(self.fc_out.weight.shape) == (1, 50) and self.fc_inp.weight.is_cuda
                and self.fc_inp.weight.dtype == torch.float32)

    def forward(self, x):
        if self._native_ok(x):
            # reference configuration on a CUDA device: one kernel forward, one kernel (+ reduce) backward (fc_fnn_*)
            return _FusedFNN.apply(x, self.fc_inp.weight, self.fc_inp.bias, self.fc_out.weight)
        out = self.activation(self.fc_inp(x))
        for _ in range(self.width_dim - 1):
            out = self.activation(self.fc_int(out))
        return self.constraint(self.fc_out(out))


class _FusedFNN(torch.autograd.Function):
    """``FNNModel.forward`` for the reference configuration (Functions.py:261-289 with ``width_dim = 1``) as one kernel;
    the backward recomputes the hidden layer and reduces the 250 weight gradients in one kernel + a 1-block reduce."""

    @staticmethod
    def forward(ctx, x, inp_w, inp_b, out_w):
        dev = x.device
        xc = x.detach().contiguous()
        ws = [w.detach().contiguous() for w in (inp_w, inp_b, out_w)]
        u = torch.empty(x.shape[0], 1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _native.lib().fc_fnn_forward(_native.ptr(xc), *[_native.ptr(w) for w in ws], x.shape[0], _native.ptr(u),
                                              _native.stream_ptr(dev))
        _native.check(rc, "fc_fnn_forward")
        ctx.save_for_backward(xc, *ws)
        return u

    @staticmethod
    def backward(ctx, du):
        xc, inp_w, inp_b, out_w = ctx.saved_tensors
        dev = xc.device
        L = _native.lib()
        d = du.detach().to(torch.float32).reshape(-1).contiguous()
        g = torch.empty(250, dtype=torch.float32, device=dev)
        nbytes = int(L.fc_fnn_backward_workspace_bytes())
        work = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = L.fc_fnn_backward(_native.ptr(xc), _native.ptr(d), _native.ptr(inp_w), _native.ptr(inp_b), _native.ptr(out_w),
                                   xc.shape[0], _native.ptr(g), _native.ptr(work), nbytes, _native.stream_ptr(dev))
        _native.check(rc, "fc_fnn_backward")
        return None, g[0:150].reshape(50, 3), g[150:200], g[200:250].reshape(1, 50)


# ----------------------------------------------------------------------------------------------
# surrogate
# ----------------------------------------------------------------------------------------------
class LSTMModel(nn.Module):
    """LSTM press surrogate, Functions.py:295-379.  ``state_dict`` keys ``lstm.weight_ih_l{k}``,
    ``lstm.weight_hh_l{k}``, ``fc.weight``, ``fc.bias``.  Stand-alone ``forward`` (the surrogate-training step,
    Model_NN/Functions.py:551, and the shadow prediction, Functions.py:999) runs the windowed-LSTM kernels of
    ``surrogate.lstm_window`` for CUDA inputs of the reference configuration (differentiable w.r.t. the weights) and the
    stock ``nn.LSTM`` module otherwise (CPU tensors, other shapes -- the ``.cpu()`` evaluation of Main.py:347-361); inside
    ``MPCLoss`` the weights are consumed by the fused kernel instead."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, layer_dim: int, bias=False,
                 device: torch.device = "cpu"):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.layer_dim = layer_dim
        self.lstm = nn.LSTM(input_dim, hidden_dim, layer_dim, batch_first=True, bias=bias)
        self.fc = nn.Linear(hidden_dim, output_dim)

    def initialize_hidden_states(self, batch_size: int, device: torch.device):
        shape = (self.layer_dim, batch_size, self.hidden_dim)
        dt = self.lstm.weight_ih_l0.dtype       # the reference uses the default dtype; follow the weights (.double() models)
        return (torch.zeros(shape, device=device, dtype=dt).requires_grad_(),
                torch.zeros(shape, device=device, dtype=dt).requires_grad_())

    def forward(self, x: torch.Tensor, device: torch.device):
        from . import surrogate
        if surrogate.supported(self, x) and not x.requires_grad and x.size(0) > 0:
            # reference configuration on a CUDA device: forward / reverse-sweep kernels (fc_lstm_window_fwd / _bwd)
            return surrogate.lstm_window(self, x)
        h0, c0 = self.initialize_hidden_states(x.size(0), device)
        out, _ = self.lstm(x, (h0.detach(), c0.detach()))
        return self.fc(out[:, -1, :])


# ----------------------------------------------------------------------------------------------
# packed-weight cache (per device, keyed by parameter storage + version)
# ----------------------------------------------------------------------------------------------
_PACK_CACHE: dict = {}          # (device, ids of the parameter tensors) -> (versions, packed buffer, weak refs)
_PACK_CACHE_SLOTS = 16
_WORKSPACE: dict = {}           # (device, stream) -> scratch buffer


def _check_models(simulator: nn.Module, controller: nn.Module):
    lstm = getattr(simulator, "lstm", None)
    fc = getattr(simulator, "fc", None)
    if lstm is None or fc is None:
        raise TypeError("MPCLoss: simulator must be an LSTMModel (attributes .lstm / .fc)")
    if (lstm.input_size, lstm.hidden_size, lstm.num_layers) != (5, 50, 3) or lstm.bias or not lstm.batch_first \
            or lstm.bidirectional or tuple(fc.weight.shape) != (4, 50) or fc.bias is None:
        raise NotImplementedError(
            "MPCLoss (fused sm_100a kernel) supports the reference surrogate LSTMModel(5, 50, 4, 3, bias=False) only")
    for name in ("fc_inp", "fc_out"):
        if not hasattr(controller, name):
            raise TypeError("MPCLoss: controller must be an FNNModel")
    if tuple(controller.fc_inp.weight.shape) != (50, 3) or tuple(controller.fc_out.weight.shape) != (1, 50) \
            or controller.fc_inp.bias is None:
        raise NotImplementedError("MPCLoss (fused sm_100a kernel) supports FNNModel(3, 50, 1, width_dim, bias=True) only")
    if getattr(controller, "width_dim", 1) > 1 and (tuple(controller.fc_int.weight.shape) != (50, 50) or controller.fc_int.bias is None):
        raise NotImplementedError("MPCLoss (fused sm_100a kernel) supports the 50 x 50 hidden layer fc_int with bias only")
    check_controller_nonlinearities(controller, "MPCLoss (fused sm_100a kernel)")


def check_controller_nonlinearities(controller, what: str):
    """The kernels hard-code ReLU hidden units and the default ``nn.Hardtanh`` saturation [-1, 1]
    (FNNModel.__init__, Functions.py:239-259); anything else must not be evaluated silently as if it were that."""
    if not isinstance(getattr(controller, "activation", None), nn.ReLU):
        raise NotImplementedError(f"{what} supports the ReLU controller only")
    c = getattr(controller, "constraint", None)
    if not isinstance(c, nn.Hardtanh) or (c.min_val, c.max_val) != (-1.0, 1.0):
        raise NotImplementedError(f"{what} supports the default nn.Hardtanh() saturation [-1, 1] only")


_NO_CONTROLLER = {}


def _weight_tensors(simulator, controller):
    l = simulator.lstm
    if controller is None:      # surrogate-only calls (LSTM shadow roll-out): zero controller block
        dev = l.weight_ih_l0.device
        z = _NO_CONTROLLER.get(dev)
        if z is None:
            z = _NO_CONTROLLER[dev] = [torch.zeros(50, 3, device=dev), torch.zeros(50, device=dev), torch.zeros(1, 50, device=dev)]
        ctl = z
    else:
        ctl = [controller.fc_inp.weight, controller.fc_inp.bias, controller.fc_out.weight]
    return [l.weight_ih_l0, l.weight_hh_l0, l.weight_ih_l1, l.weight_hh_l1, l.weight_ih_l2, l.weight_hh_l2,
            simulator.fc.weight, simulator.fc.bias] + ctl


def pack_weights(simulator: nn.Module, controller: nn.Module) -> torch.Tensor:
    """Device buffer with the tiled weight layouts of the kernels (``fc_pack_weights``).  Cached per
    (device, identity of the eleven parameter tensors); re-packed only when a parameter changed (optimizer steps bump
    ``_version``) and then into a FRESH buffer, so a buffer handed out earlier (to a caller of this function, or to a
    kernel still running on another stream) is never overwritten."""
    ws = _weight_tensors(simulator, controller)
    dev = ws[0].device
    if dev.type != "cuda":
        raise RuntimeError("MPCLoss: models must live on a CUDA device (no CPU fallback); call .to('cuda')")
    for w in ws:
        if w.device != dev or w.dtype != torch.float32:
            raise RuntimeError("MPCLoss: all weights must be float32 tensors on the same CUDA device")
    # The entry is valid only for the very same (still alive) parameter tensors at the same version: data_ptr alone
    # is not enough, the caching allocator hands freed addresses to new models.
    ident = (dev.index,) + tuple(id(w) for w in ws)
    key = tuple((w.data_ptr(), w._version) for w in ws)
    slot = _PACK_CACHE.get(ident)
    if slot is not None and slot[0] == key and all(r() is w for r, w in zip(slot[2], ws)):
        return slot[1]
    L = _native.lib()
    buf = torch.empty(int(L.fc_pack_floats()), dtype=torch.float32, device=dev)
    ws_c = [w.detach().contiguous() for w in ws]
    with torch.cuda.device(dev):
        rc = L.fc_pack_weights(*[_native.ptr(w) for w in ws_c], _native.ptr(buf), _native.stream_ptr(dev))
    _native.check(rc, "fc_pack_weights")
    if len(_PACK_CACHE) >= _PACK_CACHE_SLOTS:          # drop entries of dead models first, then the oldest
        for k in [k for k, v in _PACK_CACHE.items() if any(r() is None for r in v[2])] or [next(iter(_PACK_CACHE))]:
            _PACK_CACHE.pop(k, None)
    _PACK_CACHE.pop(ident, None)
    _PACK_CACHE[ident] = (key, buf, [weakref.ref(w) for w in ws])
    return buf


def _workspace(dev, nbytes: int) -> torch.Tensor:
    """Scratch of the fused kernels, one buffer per (device, stream): two streams on one device never share it."""
    k = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    buf = _WORKSPACE.get(k)
    if buf is None or buf.numel() < nbytes:
        _WORKSPACE.pop(k, None)
        buf = None
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _WORKSPACE[k] = buf
    return buf


def mpc_loss_native(wpack, X, u0, Z, N: int, alpha: float, with_grad: bool, global_batch: int | None = None,
                    noise_std: float = 0.0, noise_seed: int = 0, int_w=None, int_b=None, width_dim: int = 1):
    """Thin wrapper over ``fc_mpc_loss`` / ``fc_mpc_loss_noise``.  X [B,3], u0 [B], Z [B,10,5] float32 CUDA,
    contiguous.  Returns dict(cost, command, error, pred [B,N], gl [256], du0 [B] or None)."""
    B = X.shape[0]
    dev = X.device
    L = _native.lib()
    f32 = dict(dtype=torch.float32, device=dev)
    out = {k: torch.empty(B, **f32) for k in ("cost", "command", "error")}
    out["pred"] = torch.empty(B, N, **f32)
    out["gl"] = torch.empty(256, **f32)
    out["du0"] = torch.empty(B, **f32) if with_grad else None
    if width_dim > 1:            # FNNModel with hidden-layer repeats: extra gradient block d fc_int.weight | d fc_int.bias
        out["gl_wide"] = torch.zeros(2560, **f32)
        with torch.cuda.device(dev):
            nbytes = int(L.fc_mpc_loss_wide_workspace_bytes(B, N, int(with_grad), int(width_dim)))
            if nbytes == 0:
                raise RuntimeError("fc_mpc_loss_wide_workspace_bytes failed: " + L.fc_last_error().decode())
            work = _workspace(dev, nbytes)
            rc = L.fc_mpc_loss_wide(_native.ptr(X), _native.ptr(u0), _native.ptr(Z), _native.ptr(wpack), _native.ptr(int_w),
                                    _native.ptr(int_b), int(width_dim), B, N, float(alpha),
                                    int(global_batch if global_batch is not None else B), int(with_grad),
                                    _native.ptr(out["cost"]), _native.ptr(out["command"]), _native.ptr(out["error"]),
                                    _native.ptr(out["pred"]), _native.ptr(out["du0"]), _native.ptr(out["gl"]),
                                    _native.ptr(out["gl_wide"]), _native.ptr(work), nbytes, float(noise_std),
                                    int(noise_seed) & (2 ** 64 - 1), _native.stream_ptr(dev))
        _native.check(rc, "fc_mpc_loss_wide")
        return out
    with torch.cuda.device(dev):
        nbytes = int(L.fc_mpc_loss_workspace_bytes(B, N, int(with_grad)))
        if nbytes == 0:
            raise RuntimeError("fc_mpc_loss_workspace_bytes failed: " + L.fc_last_error().decode())
        work = _workspace(dev, nbytes)
        rc = L.fc_mpc_loss_noise(_native.ptr(X), _native.ptr(u0), _native.ptr(Z), _native.ptr(wpack), B, N, float(alpha),
                                 int(global_batch if global_batch is not None else B), int(with_grad),
                                 _native.ptr(out["cost"]), _native.ptr(out["command"]), _native.ptr(out["error"]),
                                 _native.ptr(out["pred"]), _native.ptr(out["du0"]), _native.ptr(out["gl"]),
                                 _native.ptr(work), nbytes, float(noise_std), int(noise_seed) & (2 ** 64 - 1),
                                 _native.stream_ptr(dev))
    _native.check(rc, "fc_mpc_loss")
    return out


class _FusedMPCLoss(torch.autograd.Function):
    """The kernel computes loss AND d loss/d(u0, fc_inp.*, fc_out.weight) in one launch (the loss is
    linear in the upstream gradient, so ``backward`` only scales)."""

    @staticmethod
    def forward(ctx, u0, inp_w, inp_b, out_w, X, Z, wpack, N, alpha, with_grad, global_batch, noise_std=0.0, noise_seed=0,
                int_w=None, int_b=None, width_dim=1):
        wide = width_dim > 1
        res = mpc_loss_native(wpack, X, u0.detach().reshape(-1).contiguous(), Z, N, alpha, with_grad, global_batch,
                              noise_std, noise_seed,
                              int_w.detach().contiguous() if wide else None, int_b.detach().contiguous() if wide else None, width_dim)
        ctx.with_grad = with_grad
        ctx.u0_shape = u0.shape
        ctx.wide = wide
        if with_grad:
            if wide:
                ctx.save_for_backward(res["du0"], res["gl"], res["gl_wide"])
            else:
                ctx.save_for_backward(res["du0"], res["gl"])
        loss = res["gl"][250].clone()
        pred = res["pred"].reshape(-1)
        ctx.mark_non_differentiable(res["cost"], res["command"], res["error"], pred)
        return loss, res["cost"], res["command"], res["error"], pred

    @staticmethod
    def backward(ctx, g_loss, *unused):
        if not ctx.with_grad:
            raise RuntimeError("MPCLoss: backward called but the forward ran without gradients")
        du0, gl = ctx.saved_tensors[:2]
        g_u0 = (du0 * g_loss).reshape(ctx.u0_shape)
        g = gl * g_loss
        g_int_w = g_int_b = None
        if ctx.wide:
            gw = ctx.saved_tensors[2] * g_loss
            g_int_w, g_int_b = gw[0:2500].reshape(50, 50), gw[2500:2550]
        return (g_u0, g[0:150].reshape(50, 3), g[150:200], g[200:250].reshape(1, 50),
                None, None, None, None, None, None, None, None, None, g_int_w, g_int_b, None)


# ----------------------------------------------------------------------------------------------
# MPC LOSS
# ----------------------------------------------------------------------------------------------
class MPCLoss(nn.Module):
    """Loss that mimics the MPC cost, Functions.py:1336-1472 (same constructor and ``forward``
    signature, same ``(loss, loss_features)`` return).  ``global_batch`` (extension, default None)
    is the batch size the mean runs over when the batch is sharded across ranks.

    ``enable_noise=True`` adds ``NOISE_STD * N(0,1)`` to every surrogate output (Functions.py:1400-1402, :1438-1440).
    The normals come from the kernel's counter-based Philox generator, seeded per call from torch's CPU generator
    (so ``torch.manual_seed`` makes a run reproducible); they are not ``torch.randn``'s stream -- the reference is
    not bit-reproducible in this mode either."""

    NOISE_STD = 0.01     # Functions.py:1401

    def __init__(self, prediction_horizon=10, alpha=0.1):
        super().__init__()
        self.N = prediction_horizon
        self.alpha = alpha
        self.activation = nn.ReLU()
        self.global_batch = None

    def forward(self, simulator: nn.Module, controller: nn.Module, input_controller: torch.Tensor,
                output_controller: torch.Tensor, states: torch.Tensor, device: torch.device, enable_noise=False):
        _check_models(simulator, controller)
        X, Z, u0 = input_controller, states, output_controller
        if X.device.type != "cuda" or Z.device.type != "cuda" or u0.device.type != "cuda":
            raise RuntimeError("MPCLoss: inputs must be CUDA tensors (forging_control_b200 has no CPU fallback)")
        B = X.shape[0]
        if X.dim() != 2 or X.shape[1] != 3 or tuple(Z.shape) != (B, LOOKBACK, 5) or u0.numel() != B:
            raise ValueError(f"MPCLoss: expected input_controller [B,3], output_controller [B,1], states [B,10,5]; "
                             f"got {tuple(X.shape)}, {tuple(u0.shape)}, {tuple(Z.shape)}")
        X = X.detach().to(torch.float32).contiguous()
        Z = Z.detach().to(torch.float32).contiguous()
        wpack = pack_weights(simulator, controller)
        params = (controller.fc_inp.weight, controller.fc_inp.bias, controller.fc_out.weight)
        width_dim = int(getattr(controller, "width_dim", 1))
        wide = (controller.fc_int.weight, controller.fc_int.bias) if width_dim > 1 else (None, None)
        with_grad = torch.is_grad_enabled() and (u0.requires_grad or any(p.requires_grad for p in params) or
                                                 any(p is not None and p.requires_grad for p in wide))
        noise_std, noise_seed = 0.0, 0
        if enable_noise:
            noise_std, noise_seed = self.NOISE_STD, int(torch.randint(0, 2 ** 62, (1,)).item())
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                # the generator's counter is the LOCAL trajectory index: de-correlate the shards of a sharded batch
                noise_seed = (noise_seed + torch.distributed.get_rank() * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
        loss, cost, command, error, pred = _FusedMPCLoss.apply(
            u0, *params, X, Z, wpack, int(self.N), float(self.alpha), bool(with_grad), self.global_batch,
            float(noise_std), noise_seed, wide[0], wide[1], width_dim)
        return loss, {"loss": cost, "command": command, "error": error, "prediction": pred}


# ----------------------------------------------------------------------------------------------
# NEURAL NETWORK (training / closed-loop orchestration)
# ----------------------------------------------------------------------------------------------
class NeuralNetwork:
    """Static helpers with the reference's signatures (Functions.py:590-1330)."""

    @staticmethod
    def train_model(data_loader, simulator, model, loss_function, optimizer, device, enable_noise=False):
        """One epoch of controller training, Functions.py:594-676.  The running loss is accumulated on the device
        (float64) and read back once per epoch instead of the reference's per-step ``loss.item()`` (:661): the
        kernel launches of consecutive steps queue back to back."""
        model.train()
        running = None
        feats = {"loss": [], "command": [], "error": [], "prediction": []}
        for X, _, z in data_loader:
            X, z = X.to(device, non_blocking=True), z.to(device, non_blocking=True)
            optimizer.zero_grad()
            output = model(X)
            loss, loss_features = loss_function(simulator, model, X, output, z, device, enable_noise)
            for k in feats:
                feats[k].append(loss_features[k])
            loss.backward()
            optimizer.step()
            running = loss.detach().double() if running is None else running + loss.detach().double()
        out = {k: torch.cat(v, dim=0) for k, v in feats.items()}
        return (running.item() if running is not None else 0.0) / len(data_loader), out

    @staticmethod
    def validate_model(data_loader, model, loss_function, device):
        """Functions.py:679-717."""
        model.eval()
        total = None
        with torch.no_grad():
            for X, y, _ in data_loader:
                X, y = X.to(device), y.to(device)
                v = loss_function(model(X), y).double()
                total = v if total is None else total + v
        return (total.item() if total is not None else 0.0) / len(data_loader)

    @staticmethod
    def train_loop(controller, simulator, train_loader, val_loader, loss_function, optimizer, n_epochs, device,
                   enable_noise=False):
        """Functions.py:825-923: epoch loop, validation MSE, wall time, features of the last epoch."""
        mse = nn.MSELoss()
        vec_t, vec_v = [], []
        per_epoch = {"loss": [], "command": [], "error": [], "prediction": []}
        start = time()
        for epoch in range(n_epochs):
            t_loss, feats = NeuralNetwork.train_model(train_loader, simulator, controller, loss_function,
                                                      optimizer, device, enable_noise)
            v_loss = NeuralNetwork.validate_model(val_loader, controller, mse, device)
            vec_t.append(t_loss)
            vec_v.append(v_loss)
            for k in per_epoch:
                per_epoch[k].append(feats[k].detach())
            logger.info(f"[{100 * (epoch + 1) / n_epochs:.1f}%] Training loss: {t_loss:.4f},  Validation loss: {v_loss:.4f}")
        loss_features = {k: torch.stack(v, dim=0) for k, v in per_epoch.items()}   # [n_epochs, ...]
        comp_time = time() - start
        logger.info(f"Total time: {comp_time:.2f}s.")
        return controller, vec_t, vec_v, comp_time, loss_features

    @staticmethod
    def tvp_fun(t_now: float, ref_step: float, bias_work: int, bias_return: int, epsilon=10 ** (-7)):
        """Piece-wise constant seeded reference, Functions.py:926-966."""
        phase = (t_now + epsilon) % ref_step
        period = (t_now + epsilon) // ref_step
        if phase < ref_step / 2:
            random.seed(period + bias_work)
            return 0.8 * random.random() + 0.1
        random.seed(period + bias_return)
        return -0.8 * random.random() - 0.1

    @staticmethod
    def simulator_make_step(X: np.ndarray, model: nn.Module, scalers: dict, noise: np.ndarray):
        """LSTM shadow prediction for one window, Functions.py:969-1011 (stock ``nn.LSTM``)."""
        model.eval()
        with torch.no_grad():
            dev = next(model.parameters()).device
            y_star = model(torch.as_tensor(X).float().to(dev), dev).cpu() + torch.as_tensor(noise).float()
            return scalers["output"].inverse_transform(y_star)

    @staticmethod
    def loop(N_traj: int, T_traj: int, Ts: float, controller: nn.Module, simulator, simulator_LSTM: nn.Module,
             init_state: dict, scalers: dict, model_scalers: dict, bias_work: float, bias_return: float, lookback: int,
             bar_title: str, process_std: np.ndarray, meas_std: np.ndarray, feasibility=False, device=None,
             dtype=torch.float64, substeps: int = 4):
        """Closed-loop deployment, Functions.py:1014-1289.

        The ``N_traj`` trajectories of the reference restart from ``init_state`` and only differ by
        their reference segment, so they run as ONE batch of the RK4 plant kernel (scaler -> FNN ->
        Hardtanh -> inverse scaler -> 4 RK4 sub-steps of the press ODE per sample).  ``simulator`` (the
        do-mpc CVODES object of the reference) is accepted for signature compatibility and returned
        untouched.  ``process_std`` / ``meas_std`` add the reference's process and measurement noise (:1176-1183) from
        the kernel's counter-based generator, seeded from ``np.random`` (so ``np.random.seed`` makes a run
        reproducible; the reference's own ``np.random.normal`` stream is not reproduced).  The IPOPT feasibility
        branch is outside the hot path and raises.
        Extensions: ``device`` (default ``cuda``), ``dtype`` (plant precision, default float64 like the
        reference plant), ``substeps``."""
        if feasibility:
            raise NotImplementedError("NeuralNetwork.loop: the IPOPT feasibility-recovery branch is out of scope")
        noisy = np.any(np.asarray(process_std) != 0) or np.any(np.asarray(meas_std) != 0)
        noise_seed = int(np.random.randint(0, 2 ** 62, dtype=np.int64)) if noisy else 0
        dev = torch.device(device) if device is not None else torch.device("cuda")
        x0 = np.array([[init_state.get(k, 0) for k in ("y", "y_dot", "p1", "p2", "z")]] * N_traj, dtype=np.float64)
        T_ref = Ts * T_traj
        ref = tvp_reference_table(N_traj, T_traj, Ts, T_ref, bias_work, bias_return)      # [N_traj, T_traj]
        scale_in = _maxabs_scale(scalers["input"], "NeuralNetwork.loop").copy()
        scale_in[2] = _maxabs_scale(scalers["y_dot"], "NeuralNetwork.loop")[0]             # Functions.py:1597-1598
        scale_out = _maxabs_scale(scalers["output"], "NeuralNetwork.loop")
        timer = ClosedLoopTimer()
        timer.tic()
        meas, u = closed_loop_rollout(controller, x0, ref, Ts, scale_in, scale_out, substeps=substeps,
                                      device=dev, dtype=dtype, process_std=process_std if noisy else None,
                                      meas_std=meas_std if noisy else None, noise_seed=noise_seed)
        timer.toc(n_steps=N_traj * T_traj)
        names = ("y", "y_dot", "p1", "p2", "z")
        results = {n: meas[:, :, i] for i, n in enumerate(names)}
        results["ref"] = ref
        results["u"] = u
        results_LSTM = NeuralNetwork._lstm_shadow(simulator_LSTM, model_scalers, x0, meas, u, lookback)
        return simulator, results, results_LSTM, timer, 0.0

    @staticmethod
    def _lstm_shadow(simulator_LSTM, model_scalers, x0, meas, u, lookback):
        """Diagnostic LSTM prediction logged next to the plant (Functions.py:969-1011, :1196-1231): the window at
        step t holds the shadow's own previous predictions and the applied commands; it is never fed
        back to the controller, so it runs after the roll-out, batched over trajectories, in ONE launch of the
        forward-only windowed-LSTM kernel (``fc_lstm_shadow_rollout``)."""
        if simulator_LSTM is None or model_scalers is None:
            return {}
        if lookback != 10:
            raise NotImplementedError("LSTM shadow: the kernel is specialised for the reference's look-back of 10")
        B, T1, _ = meas.shape
        T = T1 - 1
        s_in = _maxabs_scale(model_scalers["input"], "LSTM shadow")
        s_out = _maxabs_scale(model_scalers["output"], "LSTM shadow")
        out = np.zeros((B, T + 1, 4))
        out[:, 0] = x0[:, 1:5]
        if T > 0:
            row0 = np.concatenate((x0[:, 1:5], u[:, 0:1]), axis=1) / s_in
            y = lstm_shadow_native(simulator_LSTM, row0, np.asarray(u, dtype=np.float64) / s_in[4], s_out / s_in[:4])
            out[:, 1:] = y.astype(np.float64) * s_out
        return {"y_dot": out[:, :, 0], "p1": out[:, :, 1], "p2": out[:, :, 2], "z": out[:, :, 3]}


def lstm_shadow_native(simulator_LSTM, row0, u_scaled, ratio):
    """``fc_lstm_shadow_rollout``: row0 [B,5], u_scaled [B,T] (scaled domain, numpy or tensors), ratio [4] ->
    y [B,T,4] float32 numpy (scaled surrogate outputs).  The surrogate must live on a CUDA device."""
    dev = next(simulator_LSTM.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("forging_control_b200: the LSTM shadow roll-out needs the surrogate on a CUDA device (no CPU path)")
    wpack = pack_weights(simulator_LSTM, None)
    row0_d = torch.as_tensor(np.ascontiguousarray(row0), dtype=torch.float32).to(dev).contiguous()
    u_d = torch.as_tensor(np.ascontiguousarray(u_scaled), dtype=torch.float32).to(dev).contiguous()
    B, T = u_d.shape
    y = torch.empty(B, T, 4, dtype=torch.float32, device=dev)
    L = _native.lib()
    r = (ctypes.c_float * 4)(*[float(v) for v in np.asarray(ratio).reshape(-1)[:4]])
    with torch.cuda.device(dev):
        nbytes = int(L.fc_lstm_shadow_workspace_bytes(B, T))
        if nbytes == 0:
            raise RuntimeError("fc_lstm_shadow_workspace_bytes failed: " + L.fc_last_error().decode())
        work = _workspace(dev, nbytes)
        rc = L.fc_lstm_shadow_rollout(_native.ptr(row0_d), _native.ptr(u_d), r, _native.ptr(wpack), B, T, _native.ptr(y),
                                      _native.ptr(work), nbytes, _native.stream_ptr(dev))
    _native.check(rc, "fc_lstm_shadow_rollout")
    return y.cpu().numpy()


# ----------------------------------------------------------------------------------------------
# FEASIBILITY RECOVERY (controller step only)
# ----------------------------------------------------------------------------------------------
class FeasibilityRecovery:
    @staticmethod
    def NN_make_step(X: np.ndarray, model: nn.Module, scalers: dict, x_init=None, warm_start=None, feasibility=None):
        """Scaler -> FNN -> inverse scaler for one measurement, Functions.py:1560-1613 (feasibility=None)."""
        if feasibility:
            raise NotImplementedError("NN_make_step: the IPOPT feasibility-recovery branch is out of scope")
        model.eval()
        with torch.no_grad():
            # the reference's own calls (Functions.py:1596-1604): any fitted sklearn scaler works here
            X = np.asarray(X, dtype=np.float64)
            X_new = scalers["input"].transform(X)
            X_new[0, -1] = scalers["y_dot"].transform(np.array([[X[0, -1]]]))[0, -1]
            dev = next(model.parameters()).device
            y_star = model(torch.as_tensor(X_new).float().to(dev)).double().cpu().numpy()
            output = scalers["output"].inverse_transform(y_star)
        return output, 0.0, warm_start


class Data:
    @staticmethod
    def get_scaler(scaler: str):
        """Functions.py:407-442."""
        from sklearn.preprocessing import MaxAbsScaler, MinMaxScaler, RobustScaler, StandardScaler
        table = {"minmax": MinMaxScaler, "standard": StandardScaler, "maxabs": MaxAbsScaler, "robust": RobustScaler}
        try:
            return table[scaler.lower()]()
        except KeyError:
            raise ValueError(f"Scaler '{scaler}' is not supported. Choose from {list(table.keys())}.") from None
