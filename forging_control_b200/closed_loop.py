"""Host side of the closed-loop deployment path (``NeuralNetwork.loop``,
``Unsupervised Learning/Functions.py:1014-1289``): reference tables and the batched launch of the
RK4 plant kernel ``fc_closed_loop_rk4[_f64]``."""
from __future__ import annotations

import random
from time import perf_counter

import numpy as np
import torch

from . import _native


class ClosedLoopTimer:
    """Minimal stand-in for ``do_mpc.tools.Timer`` as used at Functions.py:1160,1173: the reference
    times one controller call per step; here one kernel launch covers ``n_steps`` steps."""

    def __init__(self):
        self.t_list = []
        self._t0 = None

    def tic(self):
        self._t0 = perf_counter()

    def toc(self, n_steps: int = 1):
        dt = perf_counter() - self._t0
        self.t_list.append(dt / max(n_steps, 1))
        return dt

    def info(self):
        t = np.asarray(self.t_list)
        return {"mean": float(t.mean()) if len(t) else 0.0, "n": len(t)}


def tvp_reference_table(N_traj: int, T_traj: int, Ts: float, T_ref: float, bias_work, bias_return,
                        epsilon: float = 1e-7) -> np.ndarray:
    """``ref[idx, t] = tvp_fun((idx*T_traj + t)*Ts, T_ref, ...)`` exactly as Functions.py:1163-1164
    evaluates it (Python ``random`` seeded per half-period, Functions.py:953-964).  The phase / period arithmetic is
    vectorised (the same float64 ``%`` and ``//``); Python's generator is seeded once per DISTINCT (period, half)
    instead of once per element, so a million-trajectory table costs one ``random.seed`` per half-period."""
    t_now = (np.arange(N_traj, dtype=np.int64)[:, None] * T_traj + np.arange(T_traj, dtype=np.int64)[None, :]) * Ts
    phase = np.mod(t_now + epsilon, T_ref)              # == Python float % for positive operands
    period = np.floor_divide(t_now + epsilon, T_ref)    # == Python float //
    work = phase < T_ref / 2
    ref = np.empty((N_traj, T_traj))
    for is_work, bias, a, b in ((True, bias_work, 0.8, 0.1), (False, bias_return, -0.8, -0.1)):
        m = work == is_work
        vals, inv = np.unique(period[m], return_inverse=True)
        draws = np.empty(len(vals))
        for i, pv in enumerate(vals):
            random.seed(float(pv) + bias)
            draws[i] = a * random.random() + b
        ref[m] = draws[inv]
    return ref


def _controller_weights(controller, dev):
    for name in ("fc_inp", "fc_out"):
        if not hasattr(controller, name):
            raise TypeError("closed loop: controller must be an FNNModel")
    if tuple(controller.fc_inp.weight.shape) != (50, 3) or tuple(controller.fc_out.weight.shape) != (1, 50) \
            or controller.fc_inp.bias is None:
        raise NotImplementedError("closed loop kernel supports FNNModel(3, 50, 1, width_dim, bias=True) only")
    from .Functions import check_controller_nonlinearities
    check_controller_nonlinearities(controller, "closed loop kernel")
    f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    width = int(getattr(controller, "width_dim", 1))
    wide = (None, None)
    if width > 1:
        if tuple(controller.fc_int.weight.shape) != (50, 50) or controller.fc_int.bias is None:
            raise NotImplementedError("closed loop kernel supports the 50 x 50 hidden layer fc_int with bias only")
        wide = (f(controller.fc_int.weight), f(controller.fc_int.bias))
    return f(controller.fc_inp.weight), f(controller.fc_inp.bias), f(controller.fc_out.weight), wide, width


def closed_loop_device(controller, x0, ref, Ts, scale_in, scale_out, substeps=4, steps_per_ref=1,
                       want_meas=True, want_u=True, process_std=None, meas_std=None, noise_seed=0, T=None):
    """Device-resident launch.  ``x0`` [B,5] and ``ref`` [n_ref,B] are CUDA tensors of the same dtype
    (float32 or float64).  Returns (meas [T+1,5,B] or None, u [T,B] or None, x_final [B,5]); the
    number of steps is ``T`` (default ``n_ref * steps_per_ref``; a shorter ``T`` stops inside the last reference segment).  ``process_std`` / ``meas_std`` (5 values each) switch on the
    process / measurement noise of ``NeuralNetwork.loop`` (Functions.py:1176-1183) from the kernel's counter-based
    generator keyed by ``noise_seed``."""
    if x0.device.type != "cuda":
        raise RuntimeError("closed loop: CUDA tensors required (forging_control_b200 has no CPU fallback)")
    dev, dt = x0.device, x0.dtype
    if dt not in (torch.float32, torch.float64) or ref.dtype != dt or ref.device != dev:
        raise ValueError("closed loop: x0 and ref must share device and dtype (float32 or float64)")
    B = x0.shape[0]
    if tuple(x0.shape) != (B, 5) or ref.dim() != 2 or ref.shape[1] != B:
        raise ValueError(f"closed loop: expected x0 [B,5], ref [n_ref,B]; got {tuple(x0.shape)}, {tuple(ref.shape)}")
    n_ref = ref.shape[0]
    T = n_ref * steps_per_ref if T is None else int(T)
    if T < 0 or T > n_ref * steps_per_ref:
        raise ValueError(f"closed loop: T={T} outside [0, n_ref * steps_per_ref = {n_ref * steps_per_ref}]")
    x0, ref = x0.contiguous(), ref.contiguous()
    w_in, b_in, w_out, wide, width = _controller_weights(controller, dev)
    s_in = torch.as_tensor(np.asarray(scale_in, dtype=np.float64), dtype=dt, device=dev)
    s_out = torch.as_tensor(np.asarray(scale_out, dtype=np.float64), dtype=dt, device=dev)
    meas = torch.empty((T + 1, 5, B), dtype=dt, device=dev) if want_meas else None
    u = torch.empty((T, B), dtype=dt, device=dev) if want_u else None
    xf = torch.empty((B, 5), dtype=dt, device=dev)
    L = _native.lib()
    import ctypes
    as5 = lambda v: (ctypes.c_float * 5)(*[float(a) for a in (np.zeros(5) if v is None else np.asarray(v, dtype=np.float64).reshape(5))])
    with torch.cuda.device(dev):
        rc = L.fc_closed_loop_rk4_ex(int(dt == torch.float64), _native.ptr(x0), _native.ptr(ref), n_ref, steps_per_ref, B, T,
                                     float(Ts), int(substeps), _native.ptr(s_in), _native.ptr(s_out), _native.ptr(w_in),
                                     _native.ptr(b_in), _native.ptr(w_out), _native.ptr(wide[0]), _native.ptr(wide[1]), width,
                                     _native.ptr(meas), _native.ptr(u), _native.ptr(xf), as5(process_std), as5(meas_std),
                                     int(noise_seed) & (2 ** 64 - 1), _native.stream_ptr(dev))
    _native.check(rc, "fc_closed_loop_rk4")
    return meas, u, xf


def closed_loop_rollout(controller, x0, ref, Ts, scale_in, scale_out, substeps=4, device="cuda",
                        dtype=torch.float64, process_std=None, meas_std=None, noise_seed=0):
    """numpy in / numpy out convenience used by ``NeuralNetwork.loop``: ``x0`` [B,5], ``ref`` [B,T]
    (physical units).  Returns (meas [B,T+1,5], u [B,T]) as float64 numpy arrays."""
    dev = torch.device(device)
    x0_t = torch.as_tensor(np.ascontiguousarray(x0), dtype=dtype).to(dev)
    ref_t = torch.as_tensor(np.ascontiguousarray(np.asarray(ref).T), dtype=dtype).to(dev)
    meas, u, _ = closed_loop_device(controller, x0_t, ref_t, Ts, scale_in, scale_out, substeps,
                                    process_std=process_std, meas_std=meas_std, noise_seed=noise_seed)
    return (meas.permute(2, 0, 1).double().cpu().numpy(), u.t().double().cpu().numpy())
