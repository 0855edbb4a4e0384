#!/usr/bin/env python
"""Surrogate-training step: samples/s of LSTMModel.forward + MSELoss + backward (weight gradients) +
AdamW through the module API, CUDA events, median of 7 after 3 warm-ups; per-kernel times of fc_lstm_window_fwd /
fc_lstm_window_bwd through the C ABI; roofline fraction against the FFMA peak measured in the same run; the path the
reference runs (stock nn.LSTM + MSELoss + AdamW) on the host cores and on the same GPU beside it.
Algorithmic FLOPs per sample: forward 1 020 400 + data gradients 1 000 400 + weight gradients 1 020 400 = 3 041 200."""
import ctypes, json, os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import forging_control_b200 as fb
from forging_control_b200 import _native, surrogate as S
F_ALG = 3041200.0
dev = torch.device("cuda:0")
L = _native.lib()
peak = ctypes.c_double(0.0)
_native.check(L.fc_fp32_peak(20000, ctypes.byref(peak), _native.stream_ptr(dev)), "fc_fp32_peak")


def ev(fn, warm=3, reps=7):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


class Stock(torch.nn.Module):          # the reference LSTMModel verbatim in behaviour (nn.LSTM + fc), for the side-by-side
    def __init__(s):
        super().__init__(); s.lstm = torch.nn.LSTM(5, 50, 3, batch_first=True, bias=False); s.fc = torch.nn.Linear(50, 4)
    def forward(s, x, device):
        out, _ = s.lstm(x); return s.fc(out[:, -1, :])


for B in (256, 4096, 65536, 524288):
    torch.manual_seed(0)
    m = fb.LSTMModel(5, 50, 4, 3).to(dev)
    opt = S.DeviceAdamW(m.parameters(), lr=1e-3, weight_decay=0.0)
    mse = torch.nn.MSELoss()
    g = torch.Generator(device=dev).manual_seed(1)
    X = torch.rand(B, 10, 5, generator=g, device=dev) * 2 - 1
    y = torch.rand(B, 1, 4, generator=g, device=dev) * 2 - 1

    def step(model=m, o=opt):
        o.zero_grad(); loss = mse(model(X, dev), y.squeeze()); loss.backward(); o.step(); return loss
    ms = ev(step)
    # kernels alone through the C ABI
    ws = S._lstm_params(m); pack = S._pack(ws)
    nb = int(L.fc_lstm_window_workspace_bytes(B, 1)); work = torch.empty(nb, dtype=torch.uint8, device=dev)
    out = torch.empty(B, 4, device=dev); d = torch.randn(B, 4, device=dev) / B
    grads = [torch.empty_like(w) for w in ws]
    st = _native.stream_ptr(dev)
    f = lambda: _native.check(L.fc_lstm_window_fwd(X.data_ptr(), pack.data_ptr(), ws[6].data_ptr(), ws[7].data_ptr(), B, 1,
                                                   out.data_ptr(), work.data_ptr(), nb, st), "fwd")
    b = lambda: _native.check(L.fc_lstm_window_bwd(X.data_ptr(), d.data_ptr(), pack.data_ptr(), ws[6].data_ptr(), B, work.data_ptr(),
                                                   nb, *[t.data_ptr() for t in grads], st), "bwd")
    ms_f, ms_b = ev(f), ev(b)
    # stock nn.LSTM on the same GPU (cuDNN, TF32) and on the host cores (what the reference runs)
    sg = Stock().to(dev); og = torch.optim.AdamW(sg.parameters(), lr=1e-3, weight_decay=0.0)
    ms_cudnn = ev(lambda: step(sg, og))
    rec = {"B": B, "ms_step": ms, "samples_per_s": B / (ms * 1e-3), "ms_fwd_kernel": ms_f, "ms_bwd_kernel": ms_b,
           "kernel_tflops": F_ALG * B / ((ms_f + ms_b) * 1e-3) / 1e12, "fp32_peak_tflops": peak.value / 1e12,
           "roofline_frac_fp32": F_ALG * B / ((ms_f + ms_b) * 1e-3) / peak.value,
           "workspace_MB": nb / 1e6, "stock_cudnn_tf32_ms_step": ms_cudnn}
    if B <= 4096:
        sc = Stock(); oc = torch.optim.AdamW(sc.parameters(), lr=1e-3, weight_decay=0.0)
        Xc, yc = X.cpu(), y.cpu()
        def cstep():
            oc.zero_grad(); l = mse(sc(Xc, "cpu"), yc.squeeze()); l.backward(); oc.step()
        cstep(); t0 = time.perf_counter(); n = 5
        for _ in range(n): cstep()
        rec["cpu_reference_ms_step"] = (time.perf_counter() - t0) / n * 1e3
        rec["cpu_threads"] = torch.get_num_threads()
    print(json.dumps(rec), flush=True)
