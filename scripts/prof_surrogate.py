#!/usr/bin/env python
"""One surrogate-training step (B = 2 tiles per SM) for ncu: pack + fc_lstm_window_fwd + MSE + fc_lstm_window_bwd + AdamW."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import forging_control_b200 as fb
from forging_control_b200 import surrogate as S
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 40 * 2
torch.manual_seed(0)
m = fb.LSTMModel(5, 50, 4, 3).to(dev)
opt = S.DeviceAdamW(m.parameters(), lr=1e-3, weight_decay=0.0)
g = torch.Generator(device=dev).manual_seed(1)
X = torch.rand(B, 10, 5, generator=g, device=dev) * 2 - 1
y = torch.rand(B, 1, 4, generator=g, device=dev) * 2 - 1
for _ in range(3):
    opt.zero_grad(); loss = torch.nn.MSELoss()(m(X, dev), y.squeeze()); loss.backward(); opt.step()
torch.cuda.synchronize()
print("loss", loss.item())
