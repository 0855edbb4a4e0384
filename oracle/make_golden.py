"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Generates the committed fixtures under ``tests/golden`` by RUNNING THE UNMODIFIED REFERENCE
(imported by path through ``oracle/ref_shim.py``) in the build container, plus data arrays
extracted from the reference's shipped result artefacts.  Run from the repo root:

    python oracle/make_golden.py

Fixtures (all small, numpy ``.npz``):

* ``weights.npz``          shipped checkpoints re-saved as arrays: surrogate
  ``Model_NN/results/model_NN.pt``, controllers ``results/NN_controller_N_10_{0,3}.pt`` and a
  fresh ``FNNModel`` after ``torch.manual_seed(0)`` (un-trained gradients), MaxAbs scaler vectors.
* ``mpc_loss_cases.npz``   for each case: inputs X, Z, the reference's fp32 and fp64 outputs of
  ``MPCLoss.forward`` (UL/Functions.py:1353-1472) and ``loss.backward()`` (:655): loss, cost,
  command, error, prediction, gradients of fc_inp.weight / fc_inp.bias / fc_out.weight.
* ``closed_loop_trace.npz`` arrays of ``results/forging_unsupervised_N_10.pkl`` (CVODES closed
  loop, 600 steps: _x,_u,_y,_tvp,_time) = known answers for the controller step and the plant
  step; ``tvp_fun`` values (UL/Functions.py:926-966) on the same time grid; the reference's own
  ``FeasibilityRecovery.NN_make_step`` (:1560-1613) outputs on the recorded measurements.
* ``sequence_dataset.npz``  every item of the reference's ``SequenceDataset`` (UL/Functions.py:66-132) over the
  per-trajectory slices of ``Data.get_individual_dataset`` (:479-516) for a small random table.
* ``trace_windows.npz``     scaled 10-row look-back windows cut from
  ``Model_NN/results/MPC_simulation.pkl`` (realistic state distribution).
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
warnings.simplefilter("ignore")


def _np(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def make_weights(F):
    lstm_sd = torch.load(os.path.join(ref_shim.MNN_DIR, "results/model_NN.pt"), map_location="cpu")
    out = {f"lstm/{k}": v for k, v in _np(lstm_sd).items()}
    ctl = {}
    for tag, idx in (("c0", 0), ("c3", 3)):
        sd = torch.load(os.path.join(ref_shim.UL_DIR, f"results/NN_controller_N_10_{idx}.pt"), map_location="cpu")
        ctl[tag] = sd
        out.update({f"fnn_{tag}/{k}": v for k, v in _np(sd).items()})
    torch.manual_seed(0)
    fresh = F.FNNModel(3, 50, 1, 1, torch.nn.ReLU, bias=True)
    ctl["init"] = fresh.state_dict()
    out.update({f"fnn_init/{k}": v for k, v in _np(fresh.state_dict()).items()})
    torch.manual_seed(1)
    fresh2 = F.FNNModel(3, 50, 1, 2, torch.nn.ReLU, bias=True)       # width_dim = 2 uses fc_int
    with torch.no_grad():
        fresh2.fc_int.bias.uniform_(-0.1, 0.1)
    ctl["w2"] = fresh2.state_dict()
    out.update({f"fnn_w2/{k}": v for k, v in _np(fresh2.state_dict()).items()})
    for name, path in (("scaler_model_input", os.path.join(ref_shim.MNN_DIR, "results/scaler_model_input.pkl")),
                       ("scaler_model_output", os.path.join(ref_shim.MNN_DIR, "results/scaler_model_output.pkl")),
                       ("scaler_input", os.path.join(ref_shim.SL_DIR, "results/scaler_input.pkl")),
                       ("scaler_output", os.path.join(ref_shim.SL_DIR, "results/scaler_output.pkl"))):
        out[f"scale/{name}"] = np.asarray(ref_shim.load_sklearn_scaler(path).scale_, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "weights.npz"), **out)
    return lstm_sd, ctl


def run_reference(F, lstm_sd, fnn_sd, width_dim, X, Z, N, alpha, dtype):
    torch.set_default_dtype(dtype)
    try:
        sim = F.LSTMModel(5, 50, 4, 3)
        sim.load_state_dict(lstm_sd)
        sim = sim.to(dtype)
        ctl = F.FNNModel(3, 50, 1, width_dim, torch.nn.ReLU, bias=True)
        ctl.load_state_dict(fnn_sd)
        ctl = ctl.to(dtype)
        Xt = torch.tensor(X, dtype=dtype)
        Zt = torch.tensor(Z, dtype=dtype)
        u0 = ctl(Xt)                                            # UL/Functions.py:643
        loss, feats = F.MPCLoss(N, alpha)(sim, ctl, Xt, u0, Zt, "cpu")   # :646
        loss.backward()                                         # :655
        res = {"loss": np.asarray(loss.item()), "u0": u0.detach().numpy()[:, 0]}
        res["cost"] = feats["loss"].detach().numpy()
        res["command"] = feats["command"].detach().numpy()
        res["error"] = feats["error"].detach().numpy()
        res["prediction"] = feats["prediction"].detach().numpy().reshape(len(X), N)
        for n, p in ctl.named_parameters():
            if p.grad is not None:
                res["grad/" + n] = p.grad.numpy().copy()
    finally:
        torch.set_default_dtype(torch.float32)
    return res


def trace_windows():
    arr = ref_shim.load_dompc_arrays(os.path.join(ref_shim.MNN_DIR, "results/MPC_simulation.pkl"))
    x, u, tvp = arr["_x"], arr["_u"], arr["_tvp"]
    rows = np.stack((x[:, 1], x[:, 2], x[:, 3], x[:, 4], u[:, 0]), axis=1)        # y_dot,p1,p2,z,u
    s_in = np.asarray(ref_shim.load_sklearn_scaler(
        os.path.join(ref_shim.MNN_DIR, "results/scaler_model_input.pkl")).scale_)
    s_c = np.asarray(ref_shim.load_sklearn_scaler(os.path.join(ref_shim.SL_DIR, "results/scaler_input.pkl")).scale_)
    rows_s = rows / s_in
    idx = np.arange(9, len(rows), 11)[:256]
    Z = np.stack([rows_s[i - 9:i + 1] for i in idx]).astype(np.float32)
    X = np.stack((x[idx, 1] / s_c[0], x[idx, 4] / s_c[1], tvp[idx, 0] / s_c[0]), axis=1).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "trace_windows.npz"), X=X, Z=Z)
    return X, Z


def make_cases(F, lstm_sd, ctl):
    rng = np.random.default_rng(20241018)
    Xt, Zt = trace_windows()
    cases = {}
    spec = [  # name, N, B, controller tag, width_dim, distribution
        ("n1_b3", 1, 3, "c0", 1, "uniform"),
        ("n2_b5", 2, 5, "c0", 1, "uniform"),
        ("n5_b16", 5, 16, "c0", 1, "uniform"),
        ("n10_b15", 10, 15, "c0", 1, "uniform"),          # the reference's own batch (Main.py:84,297)
        ("n10_b33_init", 10, 33, "init", 1, "uniform"),   # un-trained controller
        ("n10_b40_trace", 10, 40, "c3", 1, "trace"),
        ("n25_b9", 25, 9, "c0", 1, "uniform"),
        ("n12_b130_trace", 12, 130, "c0", 1, "trace"),    # > one 120-trajectory tile
        ("n6_b7_w2", 6, 7, "w2", 2, "uniform"),           # width_dim = 2 (fc_int live)
        ("n10_b12_wide", 10, 12, "init", 1, "wide"),      # large inputs: saturation / constraint kinks
    ]
    for name, N, B, tag, wd, dist in spec:
        if dist == "uniform":
            X = rng.uniform(-1, 1, (B, 3)).astype(np.float32)
            Z = rng.uniform(-1, 1, (B, 10, 5)).astype(np.float32)
        elif dist == "wide":
            X = rng.uniform(-3, 3, (B, 3)).astype(np.float32)
            Z = rng.uniform(-3, 3, (B, 10, 5)).astype(np.float32)
        else:
            sel = rng.choice(len(Xt), B, replace=False)
            X, Z = Xt[sel].copy(), Zt[sel].copy()
        cases[f"{name}/X"] = X
        cases[f"{name}/Z"] = Z
        cases[f"{name}/meta"] = np.array([N, B, wd], dtype=np.int64)
        cases[f"{name}/ctl"] = np.array(tag)
        for dn, dt in (("f32", torch.float32), ("f64", torch.float64)):
            res = run_reference(F, lstm_sd, ctl[tag], wd, X, Z, N, 20.0, dt)
            for k, v in res.items():
                cases[f"{name}/{dn}/{k}"] = v
        print(name, "loss f32", cases[f"{name}/f32/loss"], "f64", cases[f"{name}/f64/loss"])
    np.savez_compressed(os.path.join(OUT, "mpc_loss_cases.npz"), **cases)


def make_trace(F, ctl):
    arr = ref_shim.load_dompc_arrays(os.path.join(ref_shim.UL_DIR, "results/forging_unsupervised_N_10.pkl"))
    out = {k.strip("_"): np.asarray(v, dtype=np.float64) for k, v in arr.items() if k != "_aux"}
    t = out["time"][:, 0]
    out["tvp_fun"] = np.array([F.NeuralNetwork.tvp_fun(float(tt), 0.3, 300, 20 ** 6) for tt in t])
    # the reference's own controller step on the recorded measurements
    scalers = {"input": ref_shim.load_sklearn_scaler(os.path.join(ref_shim.SL_DIR, "results/scaler_input.pkl")),
               "output": ref_shim.load_sklearn_scaler(os.path.join(ref_shim.SL_DIR, "results/scaler_output.pkl"))}
    from sklearn.preprocessing import MaxAbsScaler
    ys = MaxAbsScaler()
    ys.fit(np.array([[scalers["input"].scale_[0]]]))
    scalers["y_dot"] = ys
    for s in scalers.values():
        if not hasattr(s, "clip"):
            s.clip = False
    model = F.FNNModel(3, 50, 1, 1, torch.nn.ReLU, bias=True)
    model.load_state_dict(ctl["c0"])
    # measurement fed to the controller at step k is y[k-1] (x0 returned by make_step), init state first
    y_prev = np.vstack((np.array([[0.0, 0.0, 2156275.6006012624, 2961363.827545376, 0.0]]), out["y"][:-1]))
    # the reference resets to the init state at the start of every trajectory (Functions.py:1131-1147)
    y_prev[300] = y_prev[0]
    u_ref = np.empty(len(t))
    for k in range(len(t)):
        inp = np.array([[y_prev[k, 1], y_prev[k, 4], out["tvp"][k, 0]]])
        u, _, _ = F.FeasibilityRecovery.NN_make_step(inp, model, scalers, None, None, None)
        u_ref[k] = np.asarray(u).item()
    out["nn_make_step_u"] = u_ref
    out["meas_prev"] = y_prev
    print("controller replay max|du| vs recorded:", np.abs(u_ref - out["u"][:, 0]).max())
    np.savez_compressed(os.path.join(OUT, "closed_loop_trace.npz"), **out)


def make_sequence_dataset(F):
    """Items of the reference's ``SequenceDataset`` (UL/Functions.py:66-132) over per-trajectory slices as built by
    ``Data.get_individual_dataset`` (:479-516): 3 trajectories x 25 rows, look-back 10 -> every (x, y, z) item."""
    import pandas as pd
    rng = np.random.default_rng(2024)
    n_traj, t_traj, lookback = 3, 25, 10
    cols = ["y_dot", "p1", "p2", "z", "u", "ref"]
    df = pd.DataFrame(rng.uniform(-1, 1, (n_traj * t_traj, len(cols))).astype(np.float32), columns=cols)
    target, features, model_features = ["u"], ["y_dot", "z", "ref"], ["y_dot", "p1", "p2", "z", "u"]
    datasets, _ = F.Data.get_individual_dataset(df, target, features, model_features, t_traj, lookback)
    xs, ys, zs = [], [], []
    for ds in datasets:
        for i in range(len(ds)):
            x, y, z = ds[i]
            xs.append(x.numpy()); ys.append(y.numpy()); zs.append(z.numpy())
    np.savez_compressed(os.path.join(OUT, "sequence_dataset.npz"), table=df.values.astype(np.float32), columns=np.array(cols),
                        t_traj=t_traj, lookback=lookback, X=np.stack(xs), y=np.stack(ys), Z=np.stack(zs))
    print("sequence_dataset.npz:", np.stack(xs).shape, np.stack(ys).shape, np.stack(zs).shape)


def main():
    os.makedirs(OUT, exist_ok=True)
    F = ref_shim.load_reference_functions()
    lstm_sd, ctl = make_weights(F)
    make_cases(F, lstm_sd, ctl)
    make_trace(F, ctl)
    make_sequence_dataset(F)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
