"""CPU: the plant / controller-step / reference-generator restatement against the reference's
shipped CVODES closed-loop trace and against outputs of the reference's own functions
(tests/golden/closed_loop_trace.npz, produced by oracle/make_golden.py)."""
import numpy as np

import plant_oracle as P
from conftest import state_dicts


def _fnn(W, tag="c0"):
    _, f = state_dicts(W, tag)
    return {"inp_w": f["fc_inp.weight"], "inp_b": f["fc_inp.bias"], "out_w": f["fc_out.weight"]}


def test_tvp_reference_matches_reference_function(golden_trace):
    t = golden_trace["time"][:, 0]
    mine = np.array([P.tvp_reference(float(tt), 0.3, 300, 20 ** 6) for tt in t])
    assert np.array_equal(mine, golden_trace["tvp_fun"])          # reference's own tvp_fun output
    assert np.abs(mine - golden_trace["tvp"][:, 0]).max() == 0.0  # and what the closed loop recorded
    # known answers quoted in SURVEY.md section 4
    for tt, v in ((0.0, 0.5781566283), (0.150, -0.2953094241), (0.300, 0.4189880826), (0.450, -0.6062606789)):
        assert abs(P.tvp_reference(tt, 0.3, 300, 20 ** 6) - v) < 1e-9


def test_controller_step_matches_nn_make_step(golden_trace, golden_weights):
    g = golden_trace
    si, so = golden_weights["scale/scaler_input"], golden_weights["scale/scaler_output"]
    u = P.controller_step(_fnn(golden_weights), si, so, g["meas_prev"][:, 1], g["meas_prev"][:, 4], g["tvp"][:, 0])
    assert np.abs(u - g["nn_make_step_u"]).max() < 1e-7           # the reference's NN_make_step
    assert np.abs(u - g["u"][:, 0]).max() < 2e-7                  # and the recorded commands (4.5e-8 measured)


def test_one_step_rk4_against_cvodes_trace(golden_trace):
    g = golden_trace
    x, u, y = g["x"], g["u"][:, 0], g["y"]
    for dt in (np.float64, np.float32):
        xn = P.rk4_step(x.astype(dt), u.astype(dt))
        err = np.abs(P.measurement(xn).astype(np.float64) - y) / P.STATE_SCALE
        assert np.median(err, 0).max() < 5e-7
        assert np.percentile(err, 99, 0).max() < 1e-4              # p2: 2.7e-5
        assert err.max() < 6e-3                                    # isolated switching-event rows (163, 456)
        assert (err.max(1) > 1e-4).sum() <= 6


def test_single_substep_is_unstable_and_eight_is_closer(golden_trace):
    g = golden_trace
    x, u, y = g["x"], g["u"][:, 0], g["y"]
    e1 = (np.abs(P.measurement(P.rk4_step(x.copy(), u, substeps=1)) - y) / P.STATE_SCALE).max()
    e4 = (np.abs(P.measurement(P.rk4_step(x.copy(), u, substeps=4)) - y) / P.STATE_SCALE).max()
    e8 = (np.abs(P.measurement(P.rk4_step(x.copy(), u, substeps=8)) - y) / P.STATE_SCALE).max()
    assert e1 > 0.1 and e8 < e4 < 6e-3


def test_closed_loop_replay_of_recorded_trajectory(golden_trace, golden_weights):
    g = golden_trace
    si, so = golden_weights["scale/scaler_input"], golden_weights["scale/scaler_output"]
    meas, us = P.closed_loop(_fnn(golden_weights), si, so, P.INIT_STATE[None], g["tvp"][None, :300, 0])
    err = np.abs(meas[0, 1:] - g["y"][:300]) / P.STATE_SCALE
    assert err.max() < 1e-2 and np.median(err) < 1e-4
    m32, _ = P.closed_loop(_fnn(golden_weights), si, so, P.INIT_STATE[None], g["tvp"][None, :300, 0], dtype=np.float32)
    assert (np.abs(m32 - meas) / P.STATE_SCALE).max() < 1e-4       # fp32 vs fp64 RK4 at matched step


def test_rhs_branches():
    x = np.array([[0.01, 0.2, 5e6, 3e6, 0.05], [0.01, -0.2, 5e6, 3e6, -0.05], [0.0, 0.0, -1e5, 3e6, 0.0],
                  [0.01, 0.9, 5e6, 3e6, 0.05], [0.01, -0.9, 5e6, 3e6, 0.05]])
    d = P.press_rhs(x, np.zeros(5))
    assert np.all(np.isfinite(d))
    assert d[0, 1] < d[1, 1]                      # forging force only when y>0 and y_dot>=0
    assert abs(P.smooth_floor(np.array([-1e5]))[0]) < 1e-10


def test_process_noise_is_a_rate_held_over_the_step(golden_weights):
    """do-mpc's ``set_rhs(..., process_noise=True)`` (UL/template_model.py:145-149) adds w to the right-hand side:
    for the linear spool state dz/dt = (u - z)/T1 + w the exact one-step answer is known in closed form, and with the
    reference's own noise vector (UL/Main.py:88-96) the closed loop stays finite and bounded."""
    import mpc_loss_oracle as O
    x = P.INIT_STATE[None].copy()
    w = np.array([[0.0, 0.0, 0.0, 0.0, 2.0]])
    ts, u = 1e-3, np.array([0.01])
    z1 = P.rk4_step(x, u, ts, 4, w=w)[0, 4]
    a = np.exp(-ts / P.T1)
    assert abs(z1 - (u[0] + w[0, 4] * P.T1) * (1 - a)) < 1e-9          # z' = (u + w T1 - z)/T1 from z = 0
    assert abs(P.rk4_step(x, u, ts, 4, w=np.array([[0.5, 0, 0, 0, 0.0]]))[0, 0] - P.rk4_step(x, u, ts, 4)[0, 0] - 0.5e-3) < 1e-6   # dy/dt = v + w
    si, so = golden_weights["scale/scaler_input"], golden_weights["scale/scaler_output"]
    B, T = 8, 60
    pstd = np.array([5e-1, 2e-0, 5e7, 5e7, 2e-0])
    normals = O.philox_normal4(1234, B, 3 * T)
    ref = np.full((B, T), 0.4)
    meas, us = P.closed_loop(_fnn(golden_weights), si, so, np.repeat(P.INIT_STATE[None], B, 0), ref, 1e-3, 4, np.float64,
                             pstd, np.zeros(5), normals)
    assert np.isfinite(meas).all() and np.abs(meas[:, :, 0]).max() < 0.1 and meas[:, :, 2:4].max() < 64e6
