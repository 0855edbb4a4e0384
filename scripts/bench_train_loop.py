#!/usr/bin/env python
"""End-to-end time per training step through the reference-facing `NeuralNetwork.train_model` (UL/Functions.py:594-676
mirror): host batches -> H2D -> controller -> MPCLoss -> backward -> AdamW, for the reference's own batch size
(UL/Main.py: B=15, N=10) and BASELINE config 2 (B=4096, N=5)."""
import json, os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import forging_control_b200 as fb
W = np.load(os.path.join(REPO, "tests/golden/weights.npz"))
lstm = {k[5:]: W[k] for k in W.files if k.startswith("lstm/")}
fnn = {k[len("fnn_c0/"):]: W[k] for k in W.files if k.startswith("fnn_c0/")}
dev = torch.device("cuda:0")
sim = fb.LSTMModel(5, 50, 4, 3); sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()}); sim = sim.to(dev)
for name, B, N, steps in (("config1 (Main.py)", 15, 10, 200), ("config2", 4096, 5, 100)):
    ctl = fb.FNNModel(3, 50, 1, 1); ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()}); ctl = ctl.to(dev)
    opt = torch.optim.AdamW(ctl.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(7)
    loader = [((torch.rand(B, 3, generator=g) * 2 - 1).pin_memory(), torch.zeros(B, 1).pin_memory(), (torch.rand(B, 10, 5, generator=g) * 2 - 1).pin_memory())
              for _ in range(steps)]
    loss_fn = fb.MPCLoss(prediction_horizon=N, alpha=20.0)
    fb.NeuralNetwork.train_model(loader[:5], sim, ctl, loss_fn, opt, dev)      # warm-up
    torch.cuda.synchronize(); t = time.perf_counter()
    avg, feats = fb.NeuralNetwork.train_model(loader, sim, ctl, loss_fn, opt, dev)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(json.dumps({"config": name, "B": B, "N": N, "steps": steps, "ms_per_step": 1e3 * dt / steps, "avg_loss": avg,
                      "trajectory_steps_per_s": B * N * steps / dt}), flush=True)
