"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Generates ``tests/golden/surrogate_train_cases.npz`` by RUNNING THE UNMODIFIED REFERENCE surrogate-training step
(``UL/Model_NN/Functions.py`` imported by path through ``oracle/ref_shim.py``) in the build container:
``NeuralNetwork.train_model`` (:520-569) with ``LSTMModel(5, 50, 4, 3)`` (:255-340), ``nn.MSELoss()`` and
``torch.optim.AdamW(lr=1e-3, weight_decay=0.0)`` (UL/Model_NN/Main.py:224-230), in fp32 and fp64.  Run from the
repo root:

    python oracle/make_golden_surrogate.py

Cases:
* ``shipped_b37``  shipped surrogate weights (``results/model_NN.pt``), one batch of 37 samples (ragged tile),
  U(-1,1) inputs: loss, output, the eight gradient tensors, weights after one AdamW step.
* ``fresh_b256x3`` fresh ``LSTMModel`` after ``torch.manual_seed(2)``, three batches of 256 (the reference batch
  size): average loss returned by ``train_model``, gradients of the last batch, weights after three AdamW steps.
* ``adamw_ctl``    five ``torch.optim.AdamW`` steps (lr 1e-3, default weight decay 0.01, UL/Main.py:195) on a
  250-float controller-sized parameter set with seeded gradients.
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
KEYS = tuple(f"lstm.weight_{k}_l{l}" for l in range(3) for k in ("ih", "hh")) + ("fc.weight", "fc.bias")


def run(M, sd, batches, dtype, lr=1e-3):
    torch.set_default_dtype(dtype)
    try:
        model = M.LSTMModel(5, 50, 4, 3)
        model.load_state_dict(sd)
        model = model.to(dtype)
        opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=0.0)           # Model_NN/Main.py:230
        loader = [(torch.tensor(X, dtype=dtype), torch.tensor(y, dtype=dtype)) for X, y in batches]
        avg = M.NeuralNetwork.train_model(loader, model, torch.nn.MSELoss(), opt, "cpu")   # :520-569
        res = {"avg_loss": np.asarray(avg)}
        for n, p in model.named_parameters():
            res["grad/" + n] = p.grad.numpy().copy()           # gradients of the LAST batch
            res["after/" + n] = p.detach().numpy().copy()
    finally:
        torch.set_default_dtype(torch.float32)
    return res


def forward_only(M, sd, X, dtype):
    torch.set_default_dtype(dtype)
    try:
        model = M.LSTMModel(5, 50, 4, 3)
        model.load_state_dict(sd)
        model = model.to(dtype)
        with torch.no_grad():
            return model(torch.tensor(X, dtype=dtype), "cpu").numpy()
    finally:
        torch.set_default_dtype(torch.float32)


def main():
    M = ref_shim.load_reference_surrogate_functions()
    out = {}
    g = torch.Generator().manual_seed(4242)

    sd = torch.load(os.path.join(ref_shim.MNN_DIR, "results/model_NN.pt"), map_location="cpu")
    X = (torch.rand(37, 10, 5, generator=g) * 2 - 1).numpy()
    y = (torch.rand(37, 1, 4, generator=g) * 2 - 1).numpy()
    out["shipped_b37/X"], out["shipped_b37/y"] = X, y
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        r = run(M, sd, [(X, y)], dt)
        out[f"shipped_b37/{tag}/out"] = forward_only(M, sd, X, dt)
        for k, v in r.items():
            if tag == "f64" and k.startswith("after/"):
                continue                                        # keep the fixture small
            out[f"shipped_b37/{tag}/{k}"] = v

    torch.manual_seed(2)
    fresh = M.LSTMModel(5, 50, 4, 3).state_dict()
    for k in KEYS:
        out[f"fresh_b256x3/init/{k}"] = fresh[k].numpy().copy()
    batches = []
    for b in range(3):
        Xb = (torch.rand(256, 10, 5, generator=g) * 2 - 1).numpy()
        yb = (torch.rand(256, 1, 4, generator=g) * 2 - 1).numpy()
        batches.append((Xb, yb))
        out[f"fresh_b256x3/X{b}"], out[f"fresh_b256x3/y{b}"] = Xb, yb
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        r = run(M, fresh, batches, dt)
        for k, v in r.items():
            if tag == "f32" and k.startswith("grad/"):
                continue                                        # keep the fixture small: fp64 arbiter only
            out[f"fresh_b256x3/{tag}/{k}"] = v.astype(np.float32) if (tag == "f64" and v.ndim > 0) else v

    # torch.optim.AdamW on controller-sized parameters (UL/Main.py:195: default weight_decay = 0.01)
    torch.manual_seed(7)
    ps = [torch.nn.Parameter(torch.randn(50, 3) * 0.3), torch.nn.Parameter(torch.randn(50) * 0.1),
          torch.nn.Parameter(torch.randn(1, 50) * 0.3)]
    out["adamw_ctl/p0"] = np.concatenate([p.detach().numpy().ravel() for p in ps])
    opt = torch.optim.AdamW(ps, lr=1e-3)
    grads = []
    for s in range(5):
        gs = [torch.randn(p.shape, generator=g) * (10.0 ** (s - 3)) for p in ps]
        for p, gg in zip(ps, gs):
            p.grad = gg.clone()
        opt.step()
        grads.append(np.concatenate([gg.numpy().ravel() for gg in gs]))
        out[f"adamw_ctl/p{s + 1}"] = np.concatenate([p.detach().numpy().ravel() for p in ps])
    out["adamw_ctl/grads"] = np.stack(grads)

    path = os.path.join(OUT, "surrogate_train_cases.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
