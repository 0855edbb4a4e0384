#!/usr/bin/env python
"""fc_build_windows (device-side SequenceDataset): achieved HBM GB/s on a training-step-sized batch
(B=524288 samples = the headline batch per GPU), algorithmic bytes = 4*(54 written + 54 read) + 8 (index) per sample."""
import json, os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import forging_control_b200 as fb
dev = torch.device("cuda:0")
n_traj, t_traj = 14000, 301
M = n_traj * t_traj
g = torch.Generator(device="cpu").manual_seed(0)
X = torch.randn(M, 3, generator=g).to(dev); y = torch.randn(M, generator=g).to(dev); Z = torch.randn(M, 5, generator=g).to(dev)
peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
for B in (4096, 524288, 4194304):
    idx = torch.randperm(M, device=dev)[:B].contiguous() if B <= M else torch.randint(0, M, (B,), device=dev)
    for _ in range(3): fb.build_windows(X, y, Z, t_traj, idx, 10)
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fb.build_windows(X, y, Z, t_traj, idx, 10); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts)); byts = B * (4 * 54 * 2 + 8)
    print(json.dumps({"B": B, "ms": ms, "samples_per_s": B / (ms * 1e-3), "algorithmic_GBs": byts / (ms * 1e-3) / 1e9,
                      "hbm_peak_GBs": peaks.get("hbm_gbs"), "frac": byts / (ms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6547.8),
                      "note": "includes the three torch.empty allocations of the outputs (caching allocator)"}), flush=True)
