import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
for p in (REPO, os.path.join(REPO, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the read-only reference mount /root/reference")


@pytest.fixture(scope="session")
def golden_weights():
    return np.load(os.path.join(GOLDEN, "weights.npz"))


@pytest.fixture(scope="session")
def golden_cases():
    return np.load(os.path.join(GOLDEN, "mpc_loss_cases.npz"))


@pytest.fixture(scope="session")
def golden_trace():
    return np.load(os.path.join(GOLDEN, "closed_loop_trace.npz"))


@pytest.fixture(scope="session")
def trace_windows():
    return np.load(os.path.join(GOLDEN, "trace_windows.npz"))


def case_names(cases):
    return sorted({k.split("/")[0] for k in cases.files})


def state_dicts(W, tag):
    """(lstm_sd, fnn_sd) numpy dicts for controller `tag` from tests/golden/weights.npz."""
    lstm = {k[len("lstm/"):]: W[k] for k in W.files if k.startswith("lstm/")}
    pre = f"fnn_{tag}/"
    fnn = {k[len(pre):]: W[k] for k in W.files if k.startswith(pre)}
    return lstm, fnn


def rel_max(a, b):
    """max|a-b| / max|b|  (the tolerance definition of SURVEY.md section 8a)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
