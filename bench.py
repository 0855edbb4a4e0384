#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json:
    metric  = MPC-loss fwd+bwd trajectory-steps/s (horizon N=10)
    step    = one fused MPC-loss forward + reverse sweep (fc_mpc_loss) over one synthetic batch
    workload= config 5 of BASELINE.json per GPU: N=10, 4 194 304 / 8 = 524 288 trajectories per GPU
              (weak scaling: every rank owns 524 288 trajectories; no data-path collective, one
              all-reduce of the 251-float [gradients | loss] buffer per step when N_gpus > 1)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    (N > 1: launched by torchrun, one rank per GPU; RANK/LOCAL_RANK/WORLD_SIZE from the env)

`value`    : device-resident inputs, CUDA-event timed on the launching stream, max over ranks.
`e2e`      : the same metric through the public drop-in API (pinned host tensors -> H2D -> controller
             -> MPCLoss -> backward -> loss.item()), copies inside the timed region.
`roofline` : achieved = F_alg * trajectory-steps / kernel time with F_alg = 2 041 600 FLOP per
             trajectory-step.  Default kernel (tcgen05, fp16 hi/lo operands): bound "tensor", peak =
             measured dense bf16 GEMM peak (kind::f16 rate); the FP32-FFMA-equivalent fraction (peak measured in
             the same run by a register-resident FFMA loop, fc_fp32_peak) and HBM figures beside it.
             FC_MPC_KERNEL=ffma selects the FP32 FFMA kernel (bound "fp32").
`roofline.traffic`: measured (ncu capture of this tree under profiles/, tied to the kernel sources by hash) or computed
             from the selected kernel's workspace layout (fc_mpc_loss_scratch_traffic_bytes) -- never a constant.
`parity_check`: the BENCHMARKED launch against the fp64 oracle (first rows, a sub-batch through the same kernel, shard
             additivity at the full size) and, at N > 1, the NCCL-reduced gradient buffer against a single-GPU launch
             over the concatenated batch.  Outside every timed region.
`extras.closed_loop`: BASELINE config 4 (1 048 576 trajectories x 2000 steps, sharded over the ranks) at every N.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference (`Unsupervised Learning/Functions.py`: FNNModel ->
             MPCLoss.forward -> loss.backward(), stock nn.LSTM) imported from the staged copy oracle/_ref on the host
             cores, bounded sample per step, same --steps/--warmup; rows for the BASELINE.md section-3 configs at all
             threads and 1 thread.  Fallback (labelled kind "port"): the oracle's torch restatement.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

HORIZON = 10
ALPHA = 20.0
B_PER_GPU = 524288
F_ALG = 2041600.0            # FLOP per trajectory-step, fwd + bwd
ALG_BYTES_PER_TRAJ = 12 + 200 + 12 + 4 * HORIZON   # X + Z + three cost vectors + prediction
CPU_SAMPLE_B = 4096


def _golden_weights():
    W = np.load(os.path.join(REPO, "tests", "golden", "weights.npz"))
    lstm = {k[5:]: W[k] for k in W.files if k.startswith("lstm/")}
    fnn = {k[len("fnn_c0/"):]: W[k] for k in W.files if k.startswith("fnn_c0/")}
    return lstm, fnn


def _synthetic(B, seed):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(B, 3, generator=g) * 2 - 1
    Z = torch.rand(B, 10, 5, generator=g) * 2 - 1
    return X, Z


# ---------------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (staged copy, oracle/_ref) -- fallback: the oracle's torch port, labelled
# ---------------------------------------------------------------------------------------------------
REF_CONFIGS = ((10, 15), (10, 4096), (5, 4096), (25, 4096))      # BASELINE.md section 3 / SURVEY.md 8d


def _load_reference():
    """The reference's own ``Functions`` module, imported by path from the staged copy ``oracle/_ref`` (made by
    ``__graft_entry__.build()`` from the read-only mount; it travels to the GPU box) or from the mount itself.
    Returns (module, root) or (None, reason)."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import stage_reference
    root = None
    if stage_reference.staged():
        root = stage_reference.STAGE_ROOT
    elif os.path.isfile(os.path.join(stage_reference.SOURCE_ROOT, stage_reference.FILES[0])):
        root = stage_reference.SOURCE_ROOT
    if root is None:
        return None, "oracle/_ref not staged and /root/reference not mounted"
    os.environ["FORGING_REFERENCE_ROOT"] = root
    try:
        import ref_shim
        return ref_shim.load_reference_functions(), root
    except Exception as e:      # noqa: BLE001
        return None, "import of the staged reference failed: " + repr(e)[:200]


class _RefBench:
    """``output = model(X); loss, _ = MPCLoss(...)(simulator, model, X, output, z, device); loss.backward();
    loss.item()`` of the reference's train_model (Functions.py:640-661) on CPU tensors, stock code path."""

    def __init__(self):
        self.R, self.root = _load_reference()
        self.kind = "reference" if self.R is not None else "port"
        lstm, fnn = _golden_weights()
        if self.R is not None:
            R = self.R
            self.sim = R.LSTMModel(5, 50, 4, 3)
            self.ctl = R.FNNModel(3, 50, 1, 1)
            pt_l = os.path.join(self.root, "Unsupervised Learning", "Model_NN", "results", "model_NN.pt")
            pt_c = os.path.join(self.root, "Unsupervised Learning", "results", "NN_controller_N_10_0.pt")
            if os.path.isfile(pt_l) and os.path.isfile(pt_c):       # the shipped checkpoints, strict load (Main.py:168)
                self.sim.load_state_dict(torch.load(pt_l, map_location="cpu"))
                self.ctl.load_state_dict(torch.load(pt_c, map_location="cpu"))
            else:
                self.sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()})
                self.ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()})
        else:
            import mpc_loss_oracle as O      # bench.py's cpu_baseline / reference arm only
            self.O = O
            w = O.weights_from_state_dicts(lstm, fnn, np.float32)
            self.tw = {k: ([torch.tensor(a) for a in v] if isinstance(v, list) else torch.tensor(v)) for k, v in w.items()}
            for k in ("inp_w", "inp_b", "out_w"):
                self.tw[k].requires_grad_()

    def make_step(self, N, B):
        X, Z = _synthetic(B, 1234)
        if self.R is not None:
            lf = self.R.MPCLoss(prediction_horizon=N, alpha=ALPHA)
            sim, ctl = self.sim, self.ctl

            def step():
                for p in ctl.parameters():           # optimizer.zero_grad() of the reference holds controller params only
                    p.grad = None
                out = ctl(X)
                loss, _ = lf(sim, ctl, X, out, Z, torch.device("cpu"))
                loss.backward()
                return loss.item()
            return step
        tw, O = self.tw, self.O

        def step():
            for k in ("inp_w", "inp_b", "out_w"):
                tw[k].grad = None
            u0 = torch.clamp(torch.relu(X @ tw["inp_w"].t() + tw["inp_b"]) @ tw["out_w"].t(), -1.0, 1.0)[:, 0]
            loss = O.mpc_loss_torch(tw, X, u0, Z, N, ALPHA)[0]
            loss.backward()
            return loss.item()
        return step

    def time(self, N, B, steps, warmup, threads):
        torch.set_num_threads(threads)
        step = self.make_step(N, B)
        for _ in range(warmup):
            step()
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            loss = step()
            times.append(time.perf_counter() - t0)
        mean = sum(times) / len(times)
        return {"N": N, "B": B, "threads": threads, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * mean,
                "trajectory_steps_per_s": B * N / mean, "best": B * N / min(times), "loss": loss}

    def describe(self, r):
        what = ("UNMODIFIED reference Functions.py (FNNModel -> MPCLoss.forward -> loss.backward(), nn.LSTM), staged copy "
                "oracle/_ref" if self.kind == "reference" else "oracle torch port of MPCLoss fwd + backward (reference not staged)")
        return (f"{r['steps']} steps (+{r['warmup']} warm-up) of B={r['B']} trajectories, N={r['N']}, fp32, torch "
                f"{torch.__version__} CPU, {r['threads']} threads; {what}")


def cpu_reference_rows(rb, budget_s=60.0):
    """(N, B) in REF_CONFIGS x {all host threads, 1 thread}, bounded: one warm-up, then as
    many steps (1..3) as fit ~budget_s / 8 per row."""
    cores = os.cpu_count() or 1
    rows = []
    for threads in (cores, 1):
        for N, B in REF_CONFIGS:
            try:
                torch.set_num_threads(threads)
                step = rb.make_step(N, B)
                t0 = time.perf_counter()
                step()
                first = time.perf_counter() - t0
                n = int(max(1, min(3, (budget_s / 8.0) // max(first, 1e-3))))
                r = rb.time(N, B, n, 0, threads)
                rows.append({k: r[k] for k in ("N", "B", "threads", "steps", "ms_per_step", "trajectory_steps_per_s")})
            except Exception as e:      # noqa: BLE001
                rows.append({"N": N, "B": B, "threads": threads, "error": repr(e)[:200]})
    torch.set_num_threads(cores)
    return rows


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rb = _RefBench()
    r = rb.time(HORIZON, CPU_SAMPLE_B, args.steps, args.warmup, cores)
    rows = [] if args.no_cpu_baseline else cpu_reference_rows(rb)
    line = {
        "impl": "reference", "metric": "mpc_loss_fwd_bwd_trajectory_steps_per_s", "value": r["trajectory_steps_per_s"],
        "unit": "trajectory-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict(_workload_config(args.batch_per_gpu, args.horizon, int(os.environ.get("WORLD_SIZE", "1"))),
                       cpu_sample=f"each step = B={CPU_SAMPLE_B} trajectories of that workload (bounded sample; the host "
                                  f"CPU runs ~80 k trajectory-steps/s, the full 524288-trajectory step would take > 1 min)"),
        "loss": r["loss"],
        "cpu_baseline": {"value": r["trajectory_steps_per_s"], "unit": "trajectory-steps/s", "cores": cores, "kind": rb.kind,
                         "sample": rb.describe(r), "rows": rows},
        "e2e": {"value": r["trajectory_steps_per_s"], "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _workload_config(B, N, world):
    return {"workload": f"fused MPC-loss fwd+bwd, N={N}, {B} synthetic trajectories per GPU "
                        f"(BASELINE config 5: 4194304 / 8), U(-1,1) inputs, shipped surrogate + controller weights",
            "horizon": N, "alpha": ALPHA, "batch_per_gpu": B, "global_batch": B * world}


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region: NVML polled every 5 ms from a thread (the timed
    region of the default run is well under a second, too short for an `nvidia-smi -lms` child to start up);
    `nvidia-smi` is the fallback when NVML cannot be loaded."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index
        self.nvml, self.handle, self.run, self.thread = None, None, False, None
        self.sm, self.power, self.mask, self.mx = [], [], 0, None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while self.run:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.power.append(n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
                self.mask |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nvml is not None:
            self.run = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.run = False
            self.thread.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "power_w_max": max(self.power) if self.power else None, "samples": len(self.sm),
                    "reasons": sorted(name for bit, name in self.REASONS if self.mask & bit), "source": "nvml, 5 ms poll"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi -lms 100"}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
CL_TRAJ, CL_STEPS, CL_HALF = 1048576, 2000, 150          # BASELINE config 4 / SURVEY.md 8d
CL_LOG_TRAJ = 65536                                       # logged slice: full [T+1,5,B] + [T,B] output


def _source_hash():
    """sha256 (16 hex) over the kernel sources: ties an ncu-derived number under profiles/ to the tree it was taken on."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(REPO, "forging_control_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _oracle():
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import mpc_loss_oracle as O      # checker only (parity_check fields, outside every timed region)
    return O


def parity_headline(fb, L, wpack, X, u0, Z, res, N, B_global, lstm, fnn, n_rows=256, n_sub=2048):
    """parity_check of the BENCHMARKED launch (B = batch_per_gpu, many passes per CTA) against the fp64 oracle:
    (i) cost / d loss/d u0 of its first `n_rows` trajectories, (ii) the gradient buffer of an `n_sub`-row sub-batch run
    through the same kernel with the same global batch, (iii) shard additivity of the full-size gradient buffer
    (size-independent property: the sum over 8 contiguous shards equals the one-launch result)."""
    O = _oracle()
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    Xn, Zn, un = (t[:n_sub].double().cpu().numpy() for t in (X, Z, u0))
    out, g = O.mpc_loss_forward_backward(w, Xn, un, Zn, N, ALPHA)
    scale = n_sub / float(B_global)                       # oracle means over n_sub rows, the kernel over B_global
    chk = {"rows": n_rows, "sub_batch": n_sub,
           "cost": _rel(res["cost"][:n_rows].cpu().numpy(), out["cost"][:n_rows]),
           "du0": _rel(res["du0"][:n_rows].cpu().numpy(), g["u0"][:n_rows] * scale)}
    prev = None
    B = X.shape[0]
    if B > 128 * 148:
        L.fc_mpc_select_kernel(KERNEL_HEADLINE)           # the sub-batch must run through the benchmarked kernel
    try:
        r2 = fb.mpc_loss_native(wpack, X[:n_sub].contiguous(), u0[:n_sub].contiguous(), Z[:n_sub].contiguous(), N, ALPHA, True, B_global)
        gl = r2["gl"].double().cpu().numpy()
        want = np.concatenate((g["inp_w"].reshape(-1), g["inp_b"].reshape(-1), g["out_w"].reshape(-1))) * scale
        chk["gl_sub_batch"] = _rel(gl[:250], want)
        chk["loss_sub_batch"] = abs(gl[250] - out["loss"] * scale) / abs(out["loss"] * scale)
        # shard additivity at the full size
        acc = torch.zeros(251, dtype=torch.float64, device=X.device)
        for k in range(8):
            lo, hi = fb.shard_bounds(B, 8, k)
            rk = fb.mpc_loss_native(wpack, X[lo:hi].contiguous(), u0[lo:hi].contiguous(), Z[lo:hi].contiguous(), N, ALPHA, True, B_global)
            acc += rk["gl"][:251].double()
        chk["gl_shard_additivity"] = _rel(acc.cpu().numpy(), res["gl"][:251].double().cpu().numpy())
    finally:
        L.fc_mpc_select_kernel(0)
    chk["tolerance"] = 1e-5
    chk["status"] = "ok" if max(chk[k] for k in ("cost", "du0", "gl_sub_batch", "loss_sub_batch", "gl_shard_additivity")) <= 1e-5 else "FAIL"
    return chk


KERNEL_HEADLINE = 3     # fc_mpc_select_kernel id of the kernel the automatic choice takes at the headline size


def parity_nccl(fb, dist, wpack, X, u0, Z, N, world, rank, dev, n_sub=4096):
    """N > 1: every rank evaluates its first n_sub rows with global_batch = world * n_sub, ONE NCCL all-reduce sums the
    flat [gradients | loss] buffers; rank 0 then runs the concatenated world * n_sub batch in a single launch on its own
    GPU and the two 251-float buffers must agree to 1e-5 ("8-GPU == 1-GPU gradients")."""
    Bg = world * n_sub
    xs, us, zs = X[:n_sub].contiguous(), u0[:n_sub].contiguous(), Z[:n_sub].contiguous()
    r = fb.mpc_loss_native(wpack, xs, us, zs, N, ALPHA, True, Bg)
    gl = r["gl"][:251].clone()
    dist.all_reduce(gl)
    gx = [torch.empty_like(xs) for _ in range(world)]
    gu = [torch.empty_like(us) for _ in range(world)]
    gz = [torch.empty_like(zs) for _ in range(world)]
    dist.all_gather(gx, xs); dist.all_gather(gu, us); dist.all_gather(gz, zs)
    if rank != 0:
        return None
    one = fb.mpc_loss_native(wpack, torch.cat(gx), torch.cat(gu), torch.cat(gz), N, ALPHA, True, Bg)
    err = _rel(gl.double().cpu().numpy(), one["gl"][:251].double().cpu().numpy())
    return {"ranks": world, "rows_per_rank": n_sub, "gl_nccl_vs_single_gpu": err, "tolerance": 1e-5,
            "status": "ok" if err <= 1e-5 else "FAIL"}


def closed_loop_extra(fb, dist, dev, rank, world, fnn, scale_in, scale_out, with_cpu):
    """BASELINE config 4: 1 048 576 trajectories x 2000 steps of 1 ms (4 RK4 sub-steps), FNN controller + press plant,
    sharded contiguously over the ranks (no collective), float32 plant, final state only -- plus a logged slice with the
    full [T+1,5,B] + [T,B] output, a one-step parity check against the fp64 plant oracle and (N=1) the numpy port on
    the host cores."""
    ctl = fb.FNNModel(3, 50, 1, 1)
    ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()})
    ctl = ctl.to(dev)
    lo, hi = fb.shard_bounds(CL_TRAJ, world, rank)
    Bs = hi - lo
    n_ref = (CL_STEPS + CL_HALF - 1) // CL_HALF
    g = torch.Generator().manual_seed(4321)              # seeded perturbation of the reference init state
    base = torch.tensor([0.0, 0.0, 2156275.6006012624, 2961363.827545376, 0.0])
    r5 = torch.rand(CL_TRAJ, 5, generator=g)
    x0 = base.repeat(CL_TRAJ, 1)
    x0[:, 0] = 0.02 * r5[:, 0]; x0[:, 1] = -0.1 + 0.2 * r5[:, 1]
    x0[:, 2] = 1e6 + 7e6 * r5[:, 2]; x0[:, 3] = 1e6 + 7e6 * r5[:, 3]
    # Functions.py:953-964 pattern per trajectory: a work half-period +U(0.1, 0.9) m/s followed by a return half-period;
    # the return speed mirrors the work speed of its period so that the die comes back and y stays inside the
    # plant's valid stroke (0 <= y < H0 = 0.5 m) over all 2000 steps
    amp = 0.1 + 0.8 * torch.rand((n_ref + 1) // 2, CL_TRAJ, generator=g)
    ref = torch.stack([amp[k // 2] if k % 2 == 0 else -amp[k // 2] for k in range(n_ref)])
    x0_d, ref_d = x0[lo:hi].to(dev).contiguous(), ref[:, lo:hi].to(dev).contiguous()

    def run(xd, rd, log):
        return fb.closed_loop_device(ctl, xd, rd, 1e-3, scale_in, scale_out, 4, CL_HALF, want_meas=log, want_u=log, T=CL_STEPS)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), out

    ms, out = timed(lambda: run(x0_d, ref_d, False), 2)
    xf = out[2]
    finite = torch.isfinite(xf).all(dim=1)
    finite_frac = float(finite.double().mean().item())
    nl = min(CL_LOG_TRAJ, Bs)
    ms_log, out_log = timed(lambda: run(x0_d[:nl].contiguous(), ref_d[:, :nl].contiguous(), True), 1)
    log_bytes = (out_log[0].numel() + out_log[1].numel()) * 4
    ok = finite[:nl].cpu().numpy()
    same = _rel(out_log[2].cpu().numpy()[ok], xf[:nl].cpu().numpy()[ok]) if ok.any() else None   # logging does not change the states
    del out_log
    res = {"metric": "closed_loop_trajectory_steps_per_s", "config": f"BASELINE config 4: {CL_TRAJ} trajectories x {CL_STEPS} steps of 1 ms, "
           f"4 RK4 sub-steps, float32 plant, sharded contiguously over {world} rank(s) ({Bs} per rank), no collective",
           "value": CL_TRAJ * CL_STEPS / (ms * 1e-3), "unit": "trajectory-steps/s", "ms": ms, "n_gpus": world, "output": "final state only",
           "final_states_finite_fraction": finite_frac,
           "logged": {"trajectories_per_rank": nl, "value_per_gpu": nl * CL_STEPS / (ms_log * 1e-3), "ms": ms_log,
                      "bytes_written": log_bytes, "write_gbs": log_bytes / (ms_log * 1e-3) / 1e9,
                      "final_state_vs_unlogged_rel": same}}
    if rank == 0:
        sys.path.insert(0, os.path.join(REPO, "oracle"))
        import plant_oracle as P         # checker + CPU port (never on the product path)
        fn = {"inp_w": fnn["fc_inp.weight"], "inp_b": fnn["fc_inp.bias"], "out_w": fnn["fc_out.weight"]}
        xs, rs = x0[:64].double().numpy(), ref[:1, :64].double().numpy().T          # [64,1]
        m_ref, u_ref = P.closed_loop(fn, scale_in, scale_out, xs, rs)
        chk = {}
        for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
            m, u, _ = fb.closed_loop_device(ctl, torch.tensor(xs, dtype=dt).to(dev), torch.tensor(rs.T.copy(), dtype=dt).to(dev),
                                            1e-3, scale_in, scale_out, 4, 1)
            chk[name] = float((np.abs(m.permute(2, 0, 1).double().cpu().numpy() - m_ref) / P.STATE_SCALE).max())
        chk["tolerance"] = {"f64": 1e-6, "f32": 1e-4}
        chk["what"] = "one closed-loop step of 64 trajectories vs oracle/plant_oracle.py (fp64 RK4, M=4), max error scaled by [0.02,0.4,32e6,32e6,0.15]"
        chk["status"] = "ok" if chk["f64"] <= 1e-6 and chk["f32"] <= 1e-4 else "FAIL"
        res["parity_check"] = chk
        try:
            res["ncu"] = json.load(open(os.path.join(REPO, "profiles", "r02_closed_loop_ncu.json")))
        except Exception:
            res["ncu"] = None
        if with_cpu:
            Bc, Tc = 4096, 40
            xc = x0[:Bc].double().numpy()
            rc = np.repeat(ref[:1, :Bc].double().numpy().T, Tc, axis=1)
            t0 = time.perf_counter()
            P.closed_loop(fn, scale_in, scale_out, xc, rc)
            dt = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": Bc * Tc / dt, "unit": "trajectory-steps/s", "cores": os.cpu_count() or 1, "kind": "port",
                                   "sample": f"{Bc} trajectories x {Tc} steps, fp64 numpy restatement (oracle/plant_oracle.py; the reference "
                                             "plant is do-mpc/CasADi/CVODES, absent from the image)"}
    return res


def run_ours(args):
    import torch.distributed as dist
    import forging_control_b200 as fb
    from forging_control_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; forging_control_b200 has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch_per_gpu
    N = args.horizon
    B_global = B * world
    L = _native.lib()

    lstm, fnn = _golden_weights()
    sim = fb.LSTMModel(5, 50, 4, 3)
    sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()})
    ctl = fb.FNNModel(3, 50, 1, 1)
    ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()})
    sim, ctl = sim.to(dev), ctl.to(dev)
    X_h, Z_h = _synthetic(B, 1234 + rank)
    X_h, Z_h = X_h.pin_memory(), Z_h.pin_memory()
    X, Z = X_h.to(dev), Z_h.to(dev)
    with torch.no_grad():
        u0 = ctl(X).reshape(-1).contiguous()
    wpack = fb.pack_weights(sim, ctl)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- measured FP32 peak (same run, same device) ------------------------------------------------
    import ctypes
    flops = ctypes.c_double(0.0)
    _native.check(L.fc_fp32_peak(4096, ctypes.byref(flops), _native.stream_ptr(dev)), "fc_fp32_peak")
    fp32_peak = flops.value

    # ---- device-resident timed region ------------------------------------------------------------------
    def step_dev():
        r = fb.mpc_loss_native(wpack, X, u0, Z, N, ALPHA, True, B_global)
        if world > 1:
            dist.all_reduce(r["gl"][:251])
        return r

    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        res = step_dev()
        ev[i + 1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[args.steps])
    t = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = t.item()
    loss_val = res["gl"][250].item()

    # ---- end-to-end through the public API -----------------------------------------------------------
    lf = fb.MPCLoss(prediction_horizon=N, alpha=ALPHA)
    lf.global_batch = B_global
    bucket = fb.FlatGradBucket([ctl.fc_inp.weight, ctl.fc_inp.bias, ctl.fc_out.weight])

    def step_e2e():
        Xd = X_h.to(dev, non_blocking=True)
        Zd = Z_h.to(dev, non_blocking=True)
        bucket.zero()                       # the parameters' .grad are slices of one flat buffer
        out = ctl(Xd)
        loss, _ = lf(sim, ctl, Xd, out, Zd, dev)
        loss.backward()
        if world > 1:
            loss = bucket.allreduce(loss)   # ONE collective over [250 gradients | loss], no flatten / scatter copies
        return loss.item()

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        l_e2e = step_e2e()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item() / e2e_steps

    # ---- parity of the benchmarked launch (outside the timed regions) ------------------------------------
    parity = {}
    if not args.no_parity:
        if world > 1:
            parity["nccl"] = parity_nccl(fb, dist, wpack, X, u0, Z, N, world, rank, dev)
        if rank == 0:
            res1 = fb.mpc_loss_native(wpack, X, u0, Z, N, ALPHA, True, B_global)     # un-reduced buffers of this rank
            parity["headline"] = parity_headline(fb, L, wpack, X, u0, Z, res1, N, B_global, lstm, fnn)
    barrier()

    # ---- closed loop (BASELINE config 4), every N -----------------------------------------------------
    closed = None
    if not args.no_closed_loop:
        W = np.load(os.path.join(REPO, "tests", "golden", "weights.npz"))
        try:
            closed = closed_loop_extra(fb, dist, dev, rank, world, fnn, W["scale/scaler_input"], W["scale/scaler_output"],
                                       with_cpu=(world == 1 and not args.no_cpu_baseline))
        except Exception as e:      # noqa: BLE001
            closed = {"error": repr(e)[:300]}
    barrier()

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = B_global * N / (ms_per_step * 1e-3)
        kern_ms = float(np.mean(step_ms))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        achieved = F_ALG * B * N / (kern_ms * 1e-3) / 1e12
        kernel = os.environ.get("FC_MPC_KERNEL", "auto")
        use_tc = kernel in ("tc", "pair", "auto")
        kname = {"ffma": "FP32 FFMA (fc::mpc_loss_kernel)", "tc": "tcgen05 fp16 hi/lo split, one tile per CTA (fc::mpc_loss_tc_kernel)"}.get(
            kernel, "tcgen05 fp16 hi/lo split, two tiles per CTA (fc::mpc_loss_pair_kernel)")
        # kind::f16 runs at the bf16 rate; MEASURED_PEAKS.json holds the dense bf16 figure of this pool (sustained:
        # the kernel runs for 60+ ms under the power cap), fallback 1.4 PFLOP/s (B200_PROFILING.md)
        tc_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        fp32 = {"achieved": achieved, "peak": fp32_peak / 1e12, "frac": achieved / (fp32_peak / 1e12), "unit": "TFLOP/s",
                "peak_source": "fc_fp32_peak register-resident FFMA loop measured in this run (nominal 74.45)"}
        ncu = None
        for f in ("r02b_pair_kernel_ncu.json", "r02_pair_kernel_ncu.json"):      # newest capture whose source hash matches
            try:
                c = json.load(open(os.path.join(REPO, "profiles", f)))
            except Exception:
                continue
            if c.get("source_hash") == _source_hash():
                ncu = c
                break
        if use_tc:
            roof = {"bound": "tensor", "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s", "frac": achieved / tc_peak,
                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (kind::f16 runs at the bf16 rate)"
                                    if peaks else "fallback 1.4 PFLOP/s dense bf16"),
                    "note": "gate contraction on tcgen05 with fp16 hi/lo split operands (three kind::f16 MMAs per fp32-accurate "
                            "product: the tensor pipe executes ~3.3x the algorithmic FLOPs, which caps this fraction near 0.30). "
                            "What the event trace and ncu say limits the kernel (profiles/r02b_*): the cell update on the FP32/MUFU "
                            "pipes is bound by its instruction count (a sub-partition retires it at IPC ~0.7 whatever the warp "
                            "count; 46 instructions per forward cell unit after the round-2 trim), its critical path is the "
                            "18-unit warps (units split 16/16/18), the reverse sweep is co-bound by its 39-MMA chain; the record "
                            "traffic takes about half of the HBM peak (hbm.traffic_frac) and the tensor pipe is ~25 % active",
                    "fp32_equivalent": fp32}
        else:
            roof = dict(fp32, bound="fp32")
        # DRAM traffic per launch: measured = ncu --set full capture of THIS tree (profiles/r02_pair_kernel_ncu.json, tied
        # to the kernel sources by hash) scaled by trajectories; otherwise computed from the kernel's workspace layout
        computed = int(L.fc_mpc_loss_scratch_traffic_bytes(B, N)) + (12 + 200 + 12 + 4 * N) * B
        traffic, tsrc = computed, ("computed: activation records + inter-layer sequence scratch of the selected kernel's workspace "
                                   "layout (fc_mpc_loss_scratch_traffic_bytes, written once + read once) + algorithmic I/O; upper "
                                   "bound, the L2 absorbs part of the sequence scratch")
        if ncu and ncu.get("source_hash") == _source_hash() and ncu.get("horizon") == N and kernel in ("pair", "auto"):
            traffic = ncu["dram_bytes_per_trajectory"] * B
            tsrc = (f"measured: dram__bytes_read.sum + dram__bytes_write.sum of an ncu --set full capture of this tree "
                    f"({ncu.get('file')}, B={ncu.get('batch')}) scaled by trajectories; computed from the layout: {computed}")
        roof.update({"flop_per_trajectory_step": F_ALG, "kernel_ms": kern_ms, "traffic": traffic, "traffic_source": tsrc,
                     "algorithmic_bytes": ALG_BYTES_PER_TRAJ * B,
                     "hbm": {"algorithmic_gbs": ALG_BYTES_PER_TRAJ * B / (kern_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                             "traffic_gbs": traffic / (kern_ms * 1e-3) / 1e9, "traffic_frac": traffic / (kern_ms * 1e-3) / 1e9 / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}})
        line = {
            "metric": "mpc_loss_fwd_bwd_trajectory_steps_per_s", "value": value, "unit": "trajectory-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(_workload_config(B, N, world), kernel=kname, parallelism=f"dp{world}",
                           l2="inputs (105 MB Z + GBs of activation records) exceed the 126 MB L2"),
            "loss": loss_val,
            "e2e": {"value": B_global * N / (e2e_ms * 1e-3), "unit": "trajectory-steps/s",
                    "h2d_bytes_per_step": int(X_h.numel() * 4 + Z_h.numel() * 4), "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms, "loss": l_e2e,
                    "path": "pinned host -> .to(device) -> FNNModel(X) -> MPCLoss.forward -> loss.backward() -> "
                            "(N>1: one all-reduce of the flat [grads|loss] bucket) -> loss.item()"},
            "gpu_launches": 2 * args.steps,
            "gpu_launches_detail": "per step: the fused MPC-loss kernel + fc::mpc_finalize_kernel (NCCL all-reduce extra when n_gpus>1)",
            "roofline": roof,
            "clocks": clocks,
            "parity_check": parity,
            "extras": {"closed_loop": closed},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["extras"].update(_extras(dev, fp32_peak))
            rb = _RefBench()
            r = rb.time(N, CPU_SAMPLE_B, 12, 2, os.cpu_count() or 1)
            line["cpu_baseline"] = {"value": r["trajectory_steps_per_s"], "unit": "trajectory-steps/s", "cores": r["threads"],
                                    "kind": rb.kind, "sample": rb.describe(r)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _extras(dev, fp32_peak):
    """Secondary measurements of the widened rows and of the small-batch configurations, outside the timed
    regions of the headline metric.  Never fatal for the bench line.
      surrogate_train_step  LSTMModel.forward + MSELoss + backward + DeviceAdamW through the module API, B = 65 536
                            device-resident samples, CUDA events, median of 5 after 2 warm-ups; automatic path (tensor cores
                            for B >= 8192) and the FP32 FFMA kernels beside it
      small_batches         BASELINE configs 1 (Main.py's own batch of 15, N = 10), 2 (N = 5, B = 4096) and 3 (N = 25,
                            B = 65 536): one fused MPC-loss launch, automatic kernel choice and the one-tile kernel beside it"""
    out = {}
    try:
        import forging_control_b200 as fb
        from forging_control_b200 import _native
        L = _native.lib()
        B = 65536
        g = torch.Generator(device=dev).manual_seed(1)
        X = torch.rand(B, 10, 5, generator=g, device=dev) * 2 - 1
        y = torch.rand(B, 1, 4, generator=g, device=dev) * 2 - 1
        mse = torch.nn.MSELoss()
        res = {}
        for name, mode in (("auto", 0), ("ffma", 1)):
            L.fc_lstm_train_select_path(mode)
            torch.manual_seed(0)
            m = fb.LSTMModel(5, 50, 4, 3).to(dev)
            opt = fb.DeviceAdamW(m.parameters(), lr=1e-3, weight_decay=0.0)
            ts = []
            for i in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                opt.zero_grad()
                loss = mse(m(X, dev), y.squeeze())
                loss.backward()
                opt.step()
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    ts.append(e0.elapsed_time(e1))
            res[name] = (float(np.median(ts)), float(loss.item()), int(L.fc_lstm_window_workspace_bytes(B, 1)), int(L.fc_lstm_train_path_for(B)))
        L.fc_lstm_train_select_path(0)
        ms, lossv, wsb, path = res["auto"]
        flop = 3041200.0 * B
        out["surrogate_train_step"] = {
            "metric": "samples_per_s", "value": B / (ms * 1e-3), "ms_per_step": ms, "batch": B, "loss": lossv,
            "fp32_equivalent_frac": flop / (ms * 1e-3) / fp32_peak, "flop_per_sample": 3041200.0,
            "workspace_bytes_per_sample": wsb / B,
            "path": ("tensor cores: fc::lstm_train_pair_kernel (forward; forward with records + reverse sweep) + fc::lt2::dw_kernel "
                     "(tcgen05 weight gradients, MN-major operands)" if path == 2 else "FP32 FFMA kernels") +
                    " behind LSTMModel.forward -> nn.MSELoss -> backward -> DeviceAdamW.step",
            "ffma_kernels": {"value": B / (res["ffma"][0] * 1e-3), "ms_per_step": res["ffma"][0], "loss": res["ffma"][1],
                             "fp32_roofline_frac": flop / (res["ffma"][0] * 1e-3) / fp32_peak}}
    except Exception as e:      # noqa: BLE001
        out["surrogate_train_step"] = {"error": repr(e)[:300]}
    try:
        import forging_control_b200 as fb
        from forging_control_b200 import _native
        L = _native.lib()
        lstm, fnn = _golden_weights()
        sim = fb.LSTMModel(5, 50, 4, 3); sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()})
        ctl = fb.FNNModel(3, 50, 1, 1); ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()})
        sim, ctl = sim.to(dev), ctl.to(dev)
        wp = fb.pack_weights(sim, ctl)
        rows = {}
        for cname, N, B in (("config1_main_py_batch", 10, 15), ("config2", 5, 4096), ("config3", 25, 65536)):
            g = torch.Generator().manual_seed(1)
            X = (torch.rand(B, 3, generator=g) * 2 - 1).to(dev)
            Z = (torch.rand(B, 10, 5, generator=g) * 2 - 1).to(dev)
            with torch.no_grad():
                u0 = ctl(X).reshape(-1).contiguous()
            r = {"N": N, "B": B}
            for kname, mode in (("auto", 0), ("one_tile_kernel", 2)):
                L.fc_mpc_select_kernel(mode)
                for _ in range(3):
                    fb.mpc_loss_native(wp, X, u0, Z, N, ALPHA, True)
                ts = []
                for _ in range(7 if B < 10000 else 3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); o = fb.mpc_loss_native(wp, X, u0, Z, N, ALPHA, True); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                r[kname + "_ms"] = float(np.median(ts))
                r[kname + "_loss"] = float(o["gl"][250])
            L.fc_mpc_select_kernel(0)
            r["trajectory_steps_per_s"] = B * N / (r["auto_ms"] * 1e-3)
            rows[cname] = r
        rows["kernel"] = ("auto = replica mode of the pair kernel (32-trajectory tiles, fc::mpc_loss_replica_kernel) for B <= 32 x #SMs; "
                          "config3 (N = 25, B = 65 536, BASELINE configs[2]): the pair kernel, 256 tile pairs on 148 CTAs = two passes")
        out["small_batches"] = rows
    except Exception as e:      # noqa: BLE001
        out["small_batches"] = {"error": repr(e)[:300]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=B_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-closed-loop", action="store_true", help="skip extras.closed_loop (BASELINE config 4)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity_check fields")
    ap.add_argument("--horizon", type=int, default=HORIZON, help="prediction horizon N (headline metric: 10)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
