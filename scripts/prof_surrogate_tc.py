"""Kernel-time breakdown of the surrogate training step (tensor-core path) with torch.profiler."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import forging_control_b200 as fb
from forging_control_b200 import _native
L = _native.lib()
dev = torch.device("cuda:0")
W = np.load(os.path.join(REPO, "tests/golden/weights.npz"))
lstm = {k[5:]: W[k] for k in W.files if k.startswith("lstm/")}
m = fb.LSTMModel(5, 50, 4, 3); m.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()}); m = m.to(dev)
for B in [int(a) for a in (sys.argv[1:] or ["65536"])]:
    x = (torch.rand(B, 10, 5, device=dev) * 2 - 1); y = (torch.rand(B, 4, device=dev) * 2 - 1)
    for mode in (2,):
        L.fc_lstm_train_select_path(mode)
        for _ in range(2):
            m.zero_grad(); torch.nn.functional.mse_loss(m(x, dev), y).backward()
        torch.cuda.synchronize()
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                m.zero_grad(); torch.nn.functional.mse_loss(m(x, dev), y).backward()
            torch.cuda.synchronize()
        print(f"=== B={B} mode={mode}")
        rows = [(e.key[:70], e.count, e.self_device_time_total / 3e3) for e in prof.key_averages() if e.self_device_time_total > 0]
        for k, c, ms in sorted(rows, key=lambda r: -r[2])[:10]:
            print(f"{ms:9.3f} ms/step  x{c // 3:<3d} {k}")
L.fc_lstm_train_select_path(0)
