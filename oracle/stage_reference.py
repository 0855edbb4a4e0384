"""TEST / BENCH INFRASTRUCTURE ONLY (oracle side) -- never imported by the product package.

Recipe that stages the UNMODIFIED reference files of the hot path into the git-ignored directory ``oracle/_ref/`` so
that they travel to the GPU box with the snapshot (``/root/reference`` does not exist there):

    oracle/_ref/Unsupervised Learning/Functions.py                      (MPCLoss, FNNModel, LSTMModel, NeuralNetwork ...)
    oracle/_ref/Unsupervised Learning/results/NN_controller_N_10_0.pt   (shipped controller)
    oracle/_ref/Unsupervised Learning/Model_NN/results/model_NN.pt      (shipped LSTM surrogate)
    oracle/_ref/Unsupervised Learning/Model_NN/Functions.py             (surrogate-training variant)

The files are byte-for-byte copies made at ``build()`` time (``__graft_entry__.build`` calls ``stage()`` when
``/root/reference`` is mounted); nothing under ``oracle/_ref/`` is committed (``.gitignore``) and nothing in the
product package reads it.  ``bench.py --impl reference`` and ``cpu_baseline`` import the staged ``Functions.py``
through ``oracle/ref_shim.py`` with ``FORGING_REFERENCE_ROOT=oracle/_ref`` (stub modules stand in for casadi / do_mpc /
plotly / alive_progress, exactly as for the golden fixtures), so the CPU arm times the reference's own
``model(X) -> MPCLoss.forward -> loss.backward()``.
"""
from __future__ import annotations

import filecmp
import os
import shutil

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE_ROOT = os.path.join(REPO, "oracle", "_ref")
SOURCE_ROOT = "/root/reference"
FILES = (
    "Unsupervised Learning/Functions.py",
    "Unsupervised Learning/results/NN_controller_N_10_0.pt",
    "Unsupervised Learning/Model_NN/results/model_NN.pt",
    "Unsupervised Learning/Model_NN/Functions.py",
)


def staged() -> bool:
    return all(os.path.isfile(os.path.join(STAGE_ROOT, f)) for f in FILES[:3])


def stage(source_root: str = SOURCE_ROOT) -> bool:
    """Copy the files above from the mounted reference; returns False (and stages nothing) when it is not mounted."""
    if not os.path.isfile(os.path.join(source_root, FILES[0])):
        return False
    for rel in FILES:
        src, dst = os.path.join(source_root, rel), os.path.join(STAGE_ROOT, rel)
        if not os.path.isfile(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    return staged()


if __name__ == "__main__":
    print("staged" if stage() else "reference not mounted: nothing staged")
