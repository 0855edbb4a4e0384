#!/usr/bin/env python
"""Closed-loop deployment benchmark (BASELINE.json config 4): FNN controller + press plant, RK4 with 4 sub-steps
at 1 ms, B parallel trajectories x T steps on one GPU; trajectory-steps/s with final-state-only output and with
the full measurement log; CPU comparator = the fp64 numpy oracle (restatement) on a bounded sample."""
import argparse, json, os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "oracle"))
import forging_control_b200 as fb
import plant_oracle as P      # CPU comparator only

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1048576)
ap.add_argument("--steps", type=int, default=2000)
ap.add_argument("--log-batch", type=int, default=131072)
args = ap.parse_args()
W = np.load(os.path.join(REPO, "tests/golden/weights.npz"))
fnn = {k[len("fnn_c0/"):]: W[k] for k in W.files if k.startswith("fnn_c0/")}
ctl = fb.FNNModel(3, 50, 1, 1); ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()})
si, so = W["scale/scaler_input"], W["scale/scaler_output"]
dev = torch.device("cuda:0")

def inputs(B, T, seed=4321):
    rng = np.random.default_rng(seed)
    x0 = np.tile(P.INIT_STATE, (B, 1))
    x0[:, 0] = rng.uniform(0, 0.02, B); x0[:, 1] = rng.uniform(-0.1, 0.1, B)
    x0[:, 2] = rng.uniform(1e6, 8e6, B); x0[:, 3] = rng.uniform(1e6, 8e6, B)
    n_seg = (T + 149) // 150
    seg = rng.uniform(0.1, 0.9, (B, n_seg)) * np.where(np.arange(n_seg) % 2 == 0, 1.0, -1.0)
    return x0, seg

out = {"workload": f"closed loop, B={args.batch} trajectories x T={args.steps} steps of 1 ms, RK4 x4 sub-steps"}
x0, seg = inputs(args.batch, args.steps)
for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
    B = args.batch if name == "f32" else args.batch // 8
    x0_t = torch.tensor(x0[:B], dtype=dt).to(dev); seg_t = torch.tensor(seg[:B].T.copy(), dtype=dt).to(dev)
    T = seg_t.shape[0] * 150
    fb.closed_loop_device(ctl, x0_t[:1024].contiguous(), seg_t[:, :1024].contiguous(), 1e-3, si, so, 4, 150, want_meas=False, want_u=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _, _, xf = fb.closed_loop_device(ctl, x0_t, seg_t, 1e-3, si, so, 4, 150, want_meas=False, want_u=False); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out[f"{name}_final_only"] = {"trajectory_steps_per_s": B * T / (ms * 1e-3), "ms": ms, "B": B, "T": T}
    Bl = min(args.log_batch, B)
    e0.record(); m, u, _ = fb.closed_loop_device(ctl, x0_t[:Bl].contiguous(), seg_t[:, :Bl].contiguous(), 1e-3, si, so, 4, 150); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out[f"{name}_full_log"] = {"trajectory_steps_per_s": Bl * T / (ms * 1e-3), "ms": ms, "B": Bl, "T": T,
                               "log_GB": (m.numel() * m.element_size() + u.numel() * u.element_size()) / 1e9,
                               "hbm_write_gbs": (m.numel() * m.element_size() + u.numel() * u.element_size()) / (ms * 1e-3) / 1e9}
    del m, u
# CPU comparator: numpy fp64 oracle, vectorised over trajectories, bounded sample
fo = {"inp_w": fnn["fc_inp.weight"], "inp_b": fnn["fc_inp.bias"], "out_w": fnn["fc_out.weight"]}
Bc, Tc = 4096, 150
ref = np.repeat(seg[:Bc, :1], Tc, axis=1)
t0 = time.perf_counter(); P.closed_loop(fo, si, so, x0[:Bc], ref); dt_ = time.perf_counter() - t0
out["cpu_baseline"] = {"trajectory_steps_per_s": Bc * Tc / dt_, "kind": "port (numpy fp64 restatement, single process)", "sample": f"B={Bc} x T={Tc}"}
print(json.dumps(out))
