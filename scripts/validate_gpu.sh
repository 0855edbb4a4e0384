cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu3.log 2>&1; tail -3 gpurun_out/r02b_pytest_gpu3.log
python scripts/stress_kernels.py > gpurun_out/r02b_stress.log 2>&1; tail -3 gpurun_out/r02b_stress.log
python __graft_entry__.py smoke 2>&1 | tail -2
# memcheck of the new kernels on small problems (compute-sanitizer is closed on this pool: prints a notice)
cat > /tmp/mc.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.environ["GRAFT_REPO_ROOT"])
import forging_control_b200 as fb
from forging_control_b200 import _native
L = _native.lib(); dev = torch.device("cuda:0")
W = np.load("tests/golden/weights.npz")
lstm = {k[5:]: W[k] for k in W.files if k.startswith("lstm/")}; fnn = {k[len("fnn_c0/"):]: W[k] for k in W.files if k.startswith("fnn_c0/")}
sim = fb.LSTMModel(5, 50, 4, 3); sim.load_state_dict({k: torch.tensor(v) for k, v in lstm.items()}); sim = sim.to(dev)
ctl = fb.FNNModel(3, 50, 1, 1); ctl.load_state_dict({k: torch.tensor(v) for k, v in fnn.items()}); ctl = ctl.to(dev)
wp = fb.pack_weights(sim, ctl)
for B, N, mode in ((45, 3, 4), (300, 2, 3), (300, 2, 5)):
    X = torch.rand(B, 3, device=dev) * 2 - 1; Z = torch.rand(B, 10, 5, device=dev) * 2 - 1
    u0 = ctl(X).detach().reshape(-1).contiguous()
    L.fc_mpc_select_kernel(mode); r = fb.mpc_loss_native(wp, X, u0, Z, N, 20.0, True); torch.cuda.synchronize(); print("mpc", B, N, mode, float(r["gl"][250]))
L.fc_mpc_select_kernel(0)
L.fc_lstm_train_select_path(2)
m = fb.LSTMModel(5, 50, 4, 3).to(dev)
x = torch.rand(300, 10, 5, device=dev) * 2 - 1; y = torch.rand(300, 4, device=dev)
loss = torch.nn.functional.mse_loss(m(x, dev), y); loss.backward(); torch.cuda.synchronize(); print("train tc", loss.item(), float(m.lstm.weight_hh_l1.grad.abs().max()))
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/mc.py > gpurun_out/r02b_memcheck.log 2>&1; tail -6 gpurun_out/r02b_memcheck.log
