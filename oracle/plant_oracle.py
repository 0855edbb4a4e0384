"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product; the product package never imports it.

CPU restatement (numpy, vectorised over trajectories, fp64 by default) of the closed-loop
deployment path:

* ``press_rhs``            the 5-state hydraulic forging-press ODE of ``UL/template_model.py:20-149``
  (constants :20-62, geometry :74-83, forging force :88-99, smooth pressure floor :104-112,
  valve flows :120-129, chamber volumes :132-133, friction :142, right-hand sides :145-149).
* ``measurement``          ``UL/template_model.py:152-156`` (floored pressures are what is logged).
* ``rk4_step``             classical RK4 with M=4 sub-steps per sample, the scheme of
  ``FeasibilityRecovery.Ruge_Kuta`` (``UL/Functions.py:1759-1775``).
* ``controller_step``      ``FeasibilityRecovery.NN_make_step`` (``UL/Functions.py:1596-1604``):
  MaxAbs scaling of (y_dot, z), reference scaled by the y_dot scaler, FNN evaluated in float32,
  inverse MaxAbs scaling of the command.
* ``tvp_reference``        ``NeuralNetwork.tvp_fun`` (``UL/Functions.py:926-966``).
* ``closed_loop``          the body of ``NeuralNetwork.loop`` (``UL/Functions.py:1128-1237``)
  without noise and without the LSTM shadow.

PARITY PIN.  The arithmetic of the reference plant lives in third-party code that is absent from
``/root/reference`` and from this image: do-mpc (``do_mpc.simulator.Simulator.make_step``) ->
CasADi ``integrator('cvodes')`` -> SUNDIALS CVODES (adaptive BDF, abstol 1e-5, reltol 1e-6,
``UL/template_simulator.py:19-24``); no version is pinned anywhere in the reference.  The
restated RHS is pinned against the reference's shipped CVODES closed-loop trace
(``tests/golden/closed_loop_trace.npz``, 600 one-step known answers) by
``tests/test_oracle_plant.py``; the controller step and ``tvp_fun`` are pinned against outputs of
the reference's own functions run in the build container (same fixture).
"""
from __future__ import annotations

import random

import numpy as np

# ---- constants, UL/template_model.py:20-62,88-92 ------------------------------------------
M_ = 90000.0
B_ = 25000.0
FT = 200000.0
D1 = 0.6
D2 = 0.5
A1 = np.pi * D1 ** 2 / 4
A2 = np.pi * D2 ** 2 / 4
G_ = 9.81
KB = 22e9
V1_0 = 0.3
V2_0 = 0.1
KL_1 = 8e-13
KL_2 = 14e-14
CD = 0.63
RHO = 858.0
D_ = 0.006
PS = 32e6
PT = 101325.0
MU = 0.3
K_ = 1.115
W0 = 0.2
H0 = 0.5
B0 = 0.1
A_ = 0.14 + 0.36 * (B0 / W0) - 0.054 * (B0 / W0) ** 2
T_ = 900.0
T1 = 0.005
M0 = 1200e6
M1 = -0.0025
M2 = -0.0587
M3 = 0.1165
M4 = -0.0065
FLOOR_EPS = 1e-6

INIT_STATE = np.array([0.0, 0.0, 2156275.6006012624, 2961363.827545376, 0.0])   # UL/Main.py:447-451
STATE_SCALE = np.array([0.02, 0.4, 32e6, 32e6, 0.15])                            # SURVEY.md 8c error scaling


def smooth_floor(p):
    return 0.5 * (p + np.sqrt(p * p + p.dtype.type(FLOOR_EPS)))


def press_rhs(x, u):
    """x [...,5] (y, y_dot, p1, p2, z), u [...] -> dx/dt [...,5]."""
    dt = x.dtype.type
    y, v, p1, p2, z = (x[..., i] for i in range(5))
    h1 = dt(H0) - y
    ratio = dt(H0) / h1
    w1 = dt(W0) * ratio ** dt(A_)
    b1 = dt(B0) * (1 + dt(0.67) * (ratio * dt(W0) / w1 - 1))
    forging = (y > 0) & (v >= 0)
    with np.errstate(all="ignore"):
        ys = np.where(forging, y, dt(1e-3))                    # CasADi if_else masks the dead branch
        vs = np.where(forging, v, dt(0.0))
        Kd = dt(K_) * (1 + dt(MU) * b1 / (2 * ys) + ys / (4 * b1))
        Ad = w1 * b1
        e = np.log(dt(H0) / (dt(H0) - ys))
        e_dot = vs / (dt(H0) - ys)
        Fd = Kd * Ad * dt(M0) * dt(np.exp(M1 * T_)) * e ** dt(M2) * e_dot ** dt(M3) * np.exp(dt(M4) / e)
    Fd = np.where(forging, Fd, dt(0.0))
    p1e = smooth_floor(p1)
    p2e = smooth_floor(p2)
    kv = dt(np.pi * D_ * CD) * z
    c = dt(2.0 / RHO)

    def flow(dp):
        return kv * np.sqrt(c * np.abs(dp)) * np.sign(dp)

    work = z >= 0
    qPB = np.where(work, flow(dt(PS) - p1e), flow(p1e - dt(PT)))
    qAT = np.where(work, flow(p2e - dt(PT)), flow(dt(PS) - p2e))
    V1 = dt(V1_0 / 2) + dt(A1) * y
    V2 = dt(V2_0 / 2) - dt(A2) * y
    Ft = np.where(np.abs(v) <= dt(0.5), dt(FT) * v / dt(0.5), dt(FT))
    out = np.empty_like(x)
    out[..., 0] = v
    out[..., 1] = (dt(3 * np.pi * D1 ** 2 / 4) * p1e - dt(np.pi * D2 ** 2 / 2) * p2e - dt(B_) * v - Ft - Fd) / dt(M_) + dt(G_)
    out[..., 2] = dt(KB) / V1 * (qPB / 3 - dt(A1) * v - dt(KL_1) * p1e)
    out[..., 3] = dt(KB) / V2 * (-qAT / 2 + dt(A2) * v - dt(KL_2) * p2e)
    out[..., 4] = (u - z) / dt(T1)
    return out


def measurement(x):
    y = x.copy()
    y[..., 2] = smooth_floor(x[..., 2])
    y[..., 3] = smooth_floor(x[..., 3])
    return y


def rk4_step(x, u, ts=1e-3, substeps=4, w=None):
    """``w`` [...,5]: process noise of do-mpc's ``set_rhs(..., process_noise=True)`` (UL/template_model.py:145-149),
    an additive term of the right-hand side held constant over the step: dx/dt = f(x, u) + w."""
    dt = x.dtype.type
    h = dt(ts / substeps)
    f = press_rhs if w is None else (lambda xx, uu: press_rhs(xx, uu) + w)
    for _ in range(substeps):
        k1 = f(x, u)
        k2 = f(x + h / 2 * k1, u)
        k3 = f(x + h / 2 * k2, u)
        k4 = f(x + h * k3, u)
        x = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    return x


def controller_step(fnn, scale_in, scale_out, y_dot, z, ref):
    """fnn = dict(inp_w [50,3], inp_b [50], out_w [1,50]) float32.  Inputs are physical units
    (fp64); returns the physical command (fp64), like NN_make_step."""
    f32 = np.float32
    xin = np.stack((y_dot / scale_in[0], z / scale_in[1], ref / scale_in[0]), axis=-1).astype(f32)
    hid = np.maximum(xin @ fnn["inp_w"].astype(f32).T + fnn["inp_b"].astype(f32), f32(0))
    for _ in range(int(fnn.get("width_dim", 1)) - 1):           # FNNModel.forward, UL/Functions.py:277-283: weight-shared repeats
        hid = np.maximum(hid @ fnn["int_w"].astype(f32).T + fnn["int_b"].astype(f32), f32(0))
    v = hid @ fnn["out_w"].astype(f32).T
    u_s = np.clip(v[..., 0], f32(-1), f32(1))
    return u_s.astype(np.float64) * scale_out[0]


def tvp_reference(t_now, ref_step, bias_work, bias_return, epsilon=1e-7):
    if ((t_now + epsilon) % ref_step) < ref_step / 2:
        random.seed((t_now + epsilon) // ref_step + bias_work)
        return 0.8 * random.random() + 0.1
    random.seed((t_now + epsilon) // ref_step + bias_return)
    return -0.8 * random.random() - 0.1


def closed_loop(fnn, scale_in, scale_out, x0, ref, ts=1e-3, substeps=4, dtype=np.float64,
                process_std=None, meas_std=None, normals=None, u_ulp_jitter=None):
    """x0 [B,5] raw initial state, ref [B,T] physical reference per step.
    Returns (meas [B,T+1,5], u [B,T]); meas[:,0] = x0 as given (Functions.py:1134-1138).

    Noise (NeuralNetwork.loop, UL/Functions.py:1176-1183 -> do-mpc ``Simulator.make_step(u0, v0, w0)``): all five
    states are declared ``process_noise=True`` (UL/template_model.py:145-149), so ``w0`` is a term of the right-hand
    side, ``x_next = integrate(dx/dt = f(x, u) + w0)`` with one draw per step (a rate: UL/Main.py:88-96 has 0.5 m/s on
    y and 5e7 Pa/s on the pressures), ``y = measurement(x_next) + v0``, the controller reads ``y``.  ``normals``
    [B, 3T, 4] are the standard normals of the kernels' generator (mpc_loss_oracle.philox_normal4(seed, B, 3T)):
    step k uses the 12 values normals[:, 3k:3k+3].reshape(B,12): w0 = process_std * [0:5], v0 = meas_std * [5:10].

    ``u_ulp_jitter`` (a numpy Generator): every command is moved by -1, 0 or +1 float32 ulp at random.  The float32
    controller of any two implementations differs by about that (summation order); the spread between a jittered and a
    plain run is the arbiter for how far two correct closed loops of this stiff, kinked plant may drift apart."""
    x = np.asarray(x0, dtype=dtype).copy()
    Bn, T = ref.shape
    meas = np.empty((Bn, T + 1, 5), dtype)
    us = np.empty((Bn, T), dtype)
    meas[:, 0] = x
    y = x
    for k in range(T):
        u = controller_step(fnn, scale_in, scale_out, y[:, 1].astype(np.float64), y[:, 4].astype(np.float64), ref[:, k])
        if u_ulp_jitter is not None:
            u = u * (1.0 + u_ulp_jitter.integers(-1, 2, size=u.shape) * 2.0 ** -23)
        us[:, k] = u
        if normals is not None:
            e = normals[:, 3 * k:3 * k + 3].reshape(Bn, 12)
            w = (np.asarray(process_std, np.float32)[None, :] * e[:, 0:5].astype(np.float32)).astype(dtype)
            x = rk4_step(x, u.astype(dtype), ts, substeps, w=w)
            y = measurement(x) + (np.asarray(meas_std, np.float32)[None, :] * e[:, 5:10].astype(np.float32)).astype(dtype)
        else:
            x = rk4_step(x, u.astype(dtype), ts, substeps)
            y = measurement(x)
        meas[:, k + 1] = y
    return meas, us
