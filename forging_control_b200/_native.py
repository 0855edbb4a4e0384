"""ctypes binding of ``libforging_b200.so`` (C ABI declared in ``include/forging_b200.h``).

The shared library holds the hand-written sm_100a kernels.  There is no CPU fallback: if the
library has not been built (``python -c "import __graft_entry__ as g; g.build()"``) or a call
fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_REPO_DIR = os.path.dirname(_PKG_DIR)
LIB_PATH = os.environ.get("FC_LIB_PATH", os.path.join(_PKG_DIR, "libforging_b200.so"))   # override: development builds
SOURCES = [os.path.join(_PKG_DIR, "csrc", f) for f in
           ("fc_api.cu", "fc_mpc_kernel.inl", "fc_layout.h", "fc_plant.cuh", "fc_mpc_tc_kernel.inl", "fc_tc_layout.h",
            "fc_mpc_pair_kernel.inl", "fc_pair_layout.h", "fc_lstm_train.cuh", "fc_fnn.cuh", "fc_lstm_train_tc.cuh")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

# every symbol include/forging_b200.h declares
EXPORTS = ("fc_last_error", "fc_version", "fc_pack_floats", "fc_pack_weights",
           "fc_mpc_loss_workspace_bytes", "fc_mpc_loss", "fc_mpc_select_kernel", "fc_mpc_loss_scratch_traffic_bytes", "fc_closed_loop_rk4",
           "fc_closed_loop_rk4_f64", "fc_fp32_peak", "fc_lstm_shadow_workspace_bytes", "fc_lstm_shadow_rollout",
           "fc_mpc_loss_noise", "fc_closed_loop_rk4_noise", "fc_closed_loop_rk4_f64_noise",
           "fc_build_windows", "fc_mpc_loss_wide_workspace_bytes", "fc_mpc_loss_wide",
           "fc_closed_loop_rk4_ex", "fc_lstm_train_select_path", "fc_lstm_train_path_for", "fc_lstm_train_pack_floats", "fc_lstm_train_pack", "fc_lstm_window_workspace_bytes",
           "fc_lstm_window_fwd", "fc_lstm_window_bwd", "fc_adamw_step", "fc_fnn_forward", "fc_fnn_backward_workspace_bytes",
           "fc_fnn_backward")

_c_float_p = ctypes.c_void_p   # raw device pointers are passed as integers
_lib = None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources in-tree for sm_100a with nvcc (cross-compiles without a GPU)."""
    if not force and os.path.isfile(LIB_PATH):
        newest = max(os.path.getmtime(s) for s in SOURCES + [os.path.join(_REPO_DIR, "include", "forging_b200.h")])
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, SOURCES[0]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """Load the native library (once).  Raises if it is missing: no fallback path exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'). forging_control_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f32, f64, sz = (ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float,
                                  ctypes.c_double, ctypes.c_size_t)
    L.fc_last_error.restype = ctypes.c_char_p
    L.fc_last_error.argtypes = []
    L.fc_version.restype = i32
    L.fc_pack_floats.restype = sz
    L.fc_pack_weights.restype = i32
    L.fc_pack_weights.argtypes = [vp] * 12 + [vp]
    L.fc_mpc_loss_workspace_bytes.restype = sz
    L.fc_mpc_loss_workspace_bytes.argtypes = [i32, i32, i32]
    L.fc_mpc_loss_scratch_traffic_bytes.restype = sz
    L.fc_mpc_loss_scratch_traffic_bytes.argtypes = [i32, i32]
    L.fc_mpc_select_kernel.restype = i32
    L.fc_mpc_select_kernel.argtypes = [i32]
    L.fc_mpc_loss.restype = i32
    L.fc_mpc_loss.argtypes = [vp, vp, vp, vp, i32, i32, f32, i64, i32, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.fc_mpc_loss_noise.restype = i32
    L.fc_mpc_loss_noise.argtypes = [vp, vp, vp, vp, i32, i32, f32, i64, i32, vp, vp, vp, vp, vp, vp, vp, sz, f32,
                                    ctypes.c_ulonglong, vp]
    L.fc_closed_loop_rk4.restype = i32
    L.fc_closed_loop_rk4.argtypes = [vp, vp, i32, i32, i32, i32, f32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.fc_closed_loop_rk4_f64.restype = i32
    L.fc_closed_loop_rk4_f64.argtypes = [vp, vp, i32, i32, i32, i32, f64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.fc_lstm_shadow_workspace_bytes.restype = sz
    L.fc_lstm_shadow_workspace_bytes.argtypes = [i32, i32]
    L.fc_lstm_shadow_rollout.restype = i32
    L.fc_lstm_shadow_rollout.argtypes = [vp, vp, ctypes.POINTER(f32), vp, i32, i32, vp, vp, sz, vp]
    fp5 = ctypes.POINTER(f32)
    L.fc_closed_loop_rk4_noise.restype = i32
    L.fc_closed_loop_rk4_noise.argtypes = [vp, vp, i32, i32, i32, i32, f32, i32, vp, vp, vp, vp, vp, vp, vp, vp, fp5, fp5,
                                           ctypes.c_ulonglong, vp]
    L.fc_closed_loop_rk4_f64_noise.restype = i32
    L.fc_closed_loop_rk4_f64_noise.argtypes = [vp, vp, i32, i32, i32, i32, f64, i32, vp, vp, vp, vp, vp, vp, vp, vp, fp5, fp5,
                                               ctypes.c_ulonglong, vp]
    L.fc_mpc_loss_wide_workspace_bytes.restype = sz
    L.fc_mpc_loss_wide_workspace_bytes.argtypes = [i32, i32, i32, i32]
    L.fc_mpc_loss_wide.restype = i32
    L.fc_mpc_loss_wide.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, sz, f32,
                                   ctypes.c_ulonglong, vp]
    L.fc_closed_loop_rk4_ex.restype = i32
    L.fc_closed_loop_rk4_ex.argtypes = [i32, vp, vp, i32, i32, i32, i32, f64, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp,
                                        fp5, fp5, ctypes.c_ulonglong, vp]
    L.fc_build_windows.restype = i32
    L.fc_build_windows.argtypes = [vp, vp, vp, i64, i32, i32, vp, i64, vp, vp, vp, vp]
    L.fc_lstm_train_select_path.restype = i32
    L.fc_lstm_train_select_path.argtypes = [i32]
    L.fc_lstm_train_path_for.restype = i32
    L.fc_lstm_train_path_for.argtypes = [i32]
    L.fc_lstm_train_pack_floats.restype = sz
    L.fc_lstm_train_pack_floats.argtypes = []
    L.fc_lstm_train_pack.restype = i32
    L.fc_lstm_train_pack.argtypes = [vp] * 6 + [vp, vp]
    L.fc_lstm_window_workspace_bytes.restype = sz
    L.fc_lstm_window_workspace_bytes.argtypes = [i32, i32]
    L.fc_lstm_window_fwd.restype = i32
    L.fc_lstm_window_fwd.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, sz, vp]
    L.fc_lstm_window_bwd.restype = i32
    L.fc_lstm_window_bwd.argtypes = [vp, vp, vp, vp, i32, vp, sz] + [vp] * 8 + [vp]
    L.fc_adamw_step.restype = i32
    L.fc_adamw_step.argtypes = [i32, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp),
                                ctypes.POINTER(i32), i32, f32, f32, f32, f32, f32, f32, vp]
    L.fc_fnn_forward.restype = i32
    L.fc_fnn_forward.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    L.fc_fnn_backward_workspace_bytes.restype = sz
    L.fc_fnn_backward_workspace_bytes.argtypes = []
    L.fc_fnn_backward.restype = i32
    L.fc_fnn_backward.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, sz, vp]
    L.fc_fp32_peak.restype = i32
    L.fc_fp32_peak.argtypes = [i32, ctypes.POINTER(f64), vp]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().fc_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (fc_status {rc}): {msg}")


def ptr(t) -> int:
    """Raw device pointer of a torch tensor (or 0 for None)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream
