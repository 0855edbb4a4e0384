"""CPU: the CUDA kernel SOURCE of the fused MPC loss (forging_control_b200/csrc/fc_mpc_kernel.inl)
compiled by g++ into a thread-block emulation (tests/emu/fc_emu.cpp: one OS thread per CUDA thread,
std::barrier for __syncthreads) and checked against the oracle.  This validates tiling, index
arithmetic and the reverse-sweep dataflow without a GPU; the GPU tests (-m gpu) validate the real
kernel.  The emulation library is a test artefact and is never loaded by the product package."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import mpc_loss_oracle as O
from conftest import REPO, rel_max, state_dicts

EMU_SRC = os.path.join(REPO, "tests", "emu", "fc_emu.cpp")
EMU_LIB = os.path.join(REPO, "tests", "emu", "libfc_emu.so")
DEPS = [EMU_SRC] + [os.path.join(REPO, "forging_control_b200", "csrc", f) for f in
                    ("fc_mpc_kernel.inl", "fc_layout.h", "fc_mpc_tc_kernel.inl", "fc_tc_layout.h",
                     "fc_mpc_pair_kernel.inl", "fc_pair_layout.h")]
FP = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def emu():
    if not os.path.isfile(EMU_LIB) or os.path.getmtime(EMU_LIB) < max(os.path.getmtime(d) for d in DEPS):
        subprocess.run(["g++", "-O2", "-std=c++20", "-shared", "-fPIC", "-pthread", "-o", EMU_LIB, EMU_SRC], check=True)
    return ctypes.CDLL(EMU_LIB)


def _p(a):
    return a.ctypes.data_as(FP)


def _pack(emu, lstm, fnn):
    names = [("l", "lstm.weight_ih_l0"), ("l", "lstm.weight_hh_l0"), ("l", "lstm.weight_ih_l1"), ("l", "lstm.weight_hh_l1"),
             ("l", "lstm.weight_ih_l2"), ("l", "lstm.weight_hh_l2"), ("l", "fc.weight"), ("l", "fc.bias"),
             ("f", "fc_inp.weight"), ("f", "fc_inp.bias"), ("f", "fc_out.weight")]
    arrs = [np.ascontiguousarray((lstm if s == "l" else fnn)[n], dtype=np.float32) for s, n in names]
    out = np.zeros(emu.fc_emu_pack_floats(), np.float32)
    emu.fc_emu_pack_weights(*[_p(a) for a in arrs], _p(out))
    return out


def _run(emu, wp, X, u0, Z, N, alpha, with_grad=1, grid=2, b_global=None):
    B = len(X)
    o = {k: np.zeros(B, np.float32) for k in ("cost", "command", "error", "du0")}
    o["pred"] = np.zeros((B, N), np.float32)
    o["gl"] = np.zeros(256, np.float32)
    emu.fc_emu_mpc_loss(_p(X), _p(u0), _p(Z), _p(wp), B, N, ctypes.c_float(alpha),
                        ctypes.c_longlong(b_global or B), with_grad, grid, _p(o["cost"]), _p(o["command"]),
                        _p(o["error"]), _p(o["pred"]), _p(o["du0"]), _p(o["gl"]))
    return o


@pytest.mark.parametrize("name", ["n1_b3", "n2_b5", "n5_b16", "n10_b15", "n10_b40_trace", "n12_b130_trace", "n10_b12_wide"])
def test_emulated_kernel_matches_oracle(emu, golden_cases, golden_weights, name):
    C = golden_cases
    N, B, wd = (int(v) for v in C[f"{name}/meta"])
    lstm, fnn = state_dicts(golden_weights, str(C[f"{name}/ctl"]))
    wp = _pack(emu, lstm, fnn)
    X, Z = C[f"{name}/X"], C[f"{name}/Z"]
    u0 = np.ascontiguousarray(C[f"{name}/f32/u0"])
    o = _run(emu, wp, X, u0, Z, N, 20.0)
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    out, g = O.mpc_loss_forward_backward(w, X.astype(np.float64), u0.astype(np.float64), Z.astype(np.float64), N, 20.0)
    tol = 1e-5
    assert abs(o["gl"][250] - out["loss"]) / abs(out["loss"]) < tol
    assert abs(o["gl"][250] - C[f"{name}/f32/loss"]) / abs(C[f"{name}/f32/loss"]) < tol    # the reference itself
    for k, ok in (("cost", "cost"), ("command", "command"), ("error", "error"), ("pred", "prediction")):
        assert rel_max(o[k], out[ok]) < tol, k
    assert rel_max(o["du0"], g["u0"]) < tol
    assert rel_max(o["gl"][:150].reshape(50, 3), g["inp_w"]) < tol
    assert rel_max(o["gl"][150:200], g["inp_b"]) < tol
    assert rel_max(o["gl"][200:250], g["out_w"][0]) < tol


def test_emulated_forward_only_and_sharding(emu, golden_cases, golden_weights):
    """with_grad=0 gives the same forward; two shards with B_global sum to the full-batch result."""
    C, name = golden_cases, "n5_b16"
    N, B, wd = (int(v) for v in C[f"{name}/meta"])
    lstm, fnn = state_dicts(golden_weights, "c0")
    wp = _pack(emu, lstm, fnn)
    X, Z, u0 = C[f"{name}/X"], C[f"{name}/Z"], np.ascontiguousarray(C[f"{name}/f32/u0"])
    full = _run(emu, wp, X, u0, Z, N, 20.0)
    fwd = _run(emu, wp, X, u0, Z, N, 20.0, with_grad=0)
    assert np.array_equal(full["cost"], fwd["cost"]) and np.array_equal(full["pred"], fwd["pred"])
    assert np.all(fwd["gl"][:250] == 0)
    a = _run(emu, wp, np.ascontiguousarray(X[:9]), np.ascontiguousarray(u0[:9]), np.ascontiguousarray(Z[:9]), N, 20.0, b_global=B)
    b = _run(emu, wp, np.ascontiguousarray(X[9:]), np.ascontiguousarray(u0[9:]), np.ascontiguousarray(Z[9:]), N, 20.0, b_global=B)
    assert rel_max(a["gl"][:251] + b["gl"][:251], full["gl"][:251]) < 2e-6
    assert rel_max(np.concatenate((a["du0"], b["du0"])), full["du0"]) < 1e-6


# ---- tcgen05 kernel sources (one-tile and pair): TMEM = array, MMA = fp16 hi/lo dot products executed by the issuing
# thread, mbarriers = counters.  Checks operand layouts, the hi/lo split, the per-tile hand-shakes and the dataflow;
# the truncating hardware accumulator is not modelled (its compensation is switched off in the emulation).
def _pack_v(emu, lstm, fnn, variant):
    names = [("l", "lstm.weight_ih_l0"), ("l", "lstm.weight_hh_l0"), ("l", "lstm.weight_ih_l1"), ("l", "lstm.weight_hh_l1"),
             ("l", "lstm.weight_ih_l2"), ("l", "lstm.weight_hh_l2"), ("l", "fc.weight"), ("l", "fc.bias"),
             ("f", "fc_inp.weight"), ("f", "fc_inp.bias"), ("f", "fc_out.weight")]
    arrs = [np.ascontiguousarray((lstm if s == "l" else fnn)[n], dtype=np.float32) for s, n in names]
    out = np.zeros(getattr(emu, f"fc_emu_pack_floats_{variant}")(), np.float32)
    getattr(emu, f"fc_emu_pack_weights_{variant}")(*[_p(a) for a in arrs], _p(out))
    return out


def _run_v(emu, variant, wp, X, u0, Z, N, alpha, with_grad=1, grid=2):
    B = len(X)
    o = {k: np.zeros(B, np.float32) for k in ("cost", "command", "error", "du0")}
    o["pred"] = np.zeros((B, N), np.float32)
    o["gl"] = np.zeros(256, np.float32)
    getattr(emu, f"fc_emu_mpc_loss_{variant}")(_p(X), _p(u0), _p(Z), _p(wp), B, N, ctypes.c_float(alpha), ctypes.c_longlong(B),
                                               with_grad, grid, _p(o["cost"]), _p(o["command"]), _p(o["error"]), _p(o["pred"]),
                                               _p(o["du0"]), _p(o["gl"]))
    return o


def _check_v(o, out, g, tol=1e-5):
    assert abs(o["gl"][250] - out["loss"]) / abs(out["loss"]) < tol
    for k, ok in (("cost", "cost"), ("command", "command"), ("error", "error"), ("pred", "prediction")):
        assert rel_max(o[k], out[ok]) < tol, k
    assert rel_max(o["du0"], g["u0"]) < tol
    assert rel_max(o["gl"][:150].reshape(50, 3), g["inp_w"]) < tol
    assert rel_max(o["gl"][150:200], g["inp_b"]) < tol
    assert rel_max(o["gl"][200:250], g["out_w"][0]) < tol


@pytest.mark.parametrize("variant,name", [("tc", "n1_b3"), ("tc", "n2_b5"), ("pair", "n1_b3"), ("pair", "n2_b5"),
                                          ("replica", "n1_b3"), ("replica", "n2_b5")])
def test_emulated_tcgen05_kernels_match_oracle(emu, golden_cases, golden_weights, variant, name):
    C = golden_cases
    N, B, wd = (int(v) for v in C[f"{name}/meta"])
    lstm, fnn = state_dicts(golden_weights, str(C[f"{name}/ctl"]))
    X, Z = C[f"{name}/X"], C[f"{name}/Z"]
    u0 = np.ascontiguousarray(C[f"{name}/f32/u0"])
    o = _run_v(emu, variant, _pack_v(emu, lstm, fnn, variant), X, u0, Z, N, 20.0)
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    out, g = O.mpc_loss_forward_backward(w, X.astype(np.float64), u0.astype(np.float64), Z.astype(np.float64), N, 20.0)
    _check_v(o, out, g)


@pytest.mark.parametrize("variant", ["pair"])
def test_emulated_pair_kernel_two_tiles_and_odd_tail(emu, golden_cases, golden_weights, variant):
    """B = 300 on one emulated CTA: a pass with two tiles in flight, then a pass with a single (ragged) tile."""
    C, name, rep = golden_cases, "n2_b5", 60
    N = int(C[f"{name}/meta"][0])
    lstm, fnn = state_dicts(golden_weights, str(C[f"{name}/ctl"]))
    rng = np.random.default_rng(0)
    X = np.ascontiguousarray(np.tile(C[f"{name}/X"], (rep, 1)), dtype=np.float32)
    Z = np.tile(C[f"{name}/Z"], (rep, 1, 1))
    Z = np.ascontiguousarray(Z * (1 + 0.05 * rng.standard_normal(Z.shape)), dtype=np.float32)
    u0 = np.ascontiguousarray(np.tile(C[f"{name}/f32/u0"], rep))
    o = _run_v(emu, variant, _pack_v(emu, lstm, fnn, variant), X, u0, Z, N, 20.0, grid=1)
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    out, g = O.mpc_loss_forward_backward(w, X.astype(np.float64), u0.astype(np.float64), Z.astype(np.float64), N, 20.0)
    _check_v(o, out, g)


def test_emulated_replica_kernel_several_tiles_and_ragged_tail(emu, golden_cases, golden_weights):
    """Replica mode of the pair-kernel source (32-trajectory tiles): B = 75 on two emulated CTAs = tiles 0 and 2 on CTA 0
    (the second one ragged, 11 trajectories), tile 1 on CTA 1."""
    C, name, rep = golden_cases, "n2_b5", 15
    N = int(C[f"{name}/meta"][0])
    lstm, fnn = state_dicts(golden_weights, str(C[f"{name}/ctl"]))
    rng = np.random.default_rng(1)
    X = np.ascontiguousarray(np.tile(C[f"{name}/X"], (rep, 1)), dtype=np.float32)
    Z = np.tile(C[f"{name}/Z"], (rep, 1, 1))
    Z = np.ascontiguousarray(Z * (1 + 0.05 * rng.standard_normal(Z.shape)), dtype=np.float32)
    u0 = np.ascontiguousarray(np.tile(C[f"{name}/f32/u0"], rep))
    o = _run_v(emu, "replica", _pack_v(emu, lstm, fnn, "replica"), X, u0, Z, N, 20.0, grid=2)
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    out, g = O.mpc_loss_forward_backward(w, X.astype(np.float64), u0.astype(np.float64), Z.astype(np.float64), N, 20.0)
    _check_v(o, out, g)


@pytest.mark.parametrize("variant", ["pair", "replica"])
def test_emulated_lstm_shadow_rollout_matches_oracle(emu, golden_weights, variant):
    """Forward-only shadow mode of the two-tile kernels (fc_lstm_shadow_rollout) against the oracle restatement of
    simulator_make_step / loop (UL/Functions.py:969-1011, :1196-1231)."""
    lstm, fnn = state_dicts(golden_weights, "c0")
    wp = _pack_v(emu, lstm, fnn, variant)
    rng = np.random.default_rng(3)
    B, T = 7, 4
    row0 = np.ascontiguousarray(rng.uniform(-1, 1, (B, 5)), dtype=np.float32)
    u = np.ascontiguousarray(rng.uniform(-1, 1, (B, T)), dtype=np.float32)
    ratio = np.array([1.0, 0.9, 1.1, 1.0], np.float32)
    y = np.zeros((B, T, 4), np.float32)
    getattr(emu, f"fc_emu_lstm_shadow_{variant}")(_p(row0), _p(u), _p(ratio), _p(wp), B, T, 1, _p(y))
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    ref = O.lstm_shadow_rollout(w, row0.astype(np.float64), u.astype(np.float64), ratio.astype(np.float64))
    assert rel_max(y, ref) < 1e-5


@pytest.mark.parametrize("variant", ["ffma", "pair", "replica"])
def test_emulated_enable_noise_matches_oracle_with_the_same_noise(emu, golden_cases, golden_weights, variant):
    """enable_noise (UL/Functions.py:1400-1402): the kernels' counter-based generator restated in the oracle
    (philox_normal4) gives the same roll-out, costs and gradients."""
    C, name = golden_cases, "n2_b5"
    N, B, wd = (int(v) for v in C[f"{name}/meta"])
    lstm, fnn = state_dicts(golden_weights, str(C[f"{name}/ctl"]))
    X, Z = C[f"{name}/X"], C[f"{name}/Z"]
    u0 = np.ascontiguousarray(C[f"{name}/f32/u0"])
    seed, std = 0x1234567890ABCDEF, 0.01
    emu.fc_emu_set_noise.argtypes = [ctypes.c_float, ctypes.c_ulonglong]
    emu.fc_emu_set_noise(std, seed)
    try:
        if variant == "ffma":
            o = _run(emu, _pack(emu, lstm, fnn), X, u0, Z, N, 20.0)
        else:
            o = _run_v(emu, variant, _pack_v(emu, lstm, fnn, variant), X, u0, Z, N, 20.0)
    finally:
        emu.fc_emu_set_noise(0.0, 0)
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    noise = std * O.philox_normal4(seed, B, N)
    out, g = O.mpc_loss_forward_backward(w, X.astype(np.float64), u0.astype(np.float64), Z.astype(np.float64), N, 20.0, noise=noise)
    _check_v(o, out, g)
    clean, _ = O.mpc_loss_forward_backward(w, X.astype(np.float64), u0.astype(np.float64), Z.astype(np.float64), N, 20.0)
    assert abs(out["loss"] - clean["loss"]) / abs(clean["loss"]) > 1e-4      # the noise does change the result


def test_emulated_wide_controller_matches_reference_and_oracle(emu, golden_cases, golden_weights):
    """FNNModel(width_dim=2) in the one-tile tcgen05 kernel source (fwd_glue_wide / bwd_glue_wide): the roll-out and all
    controller gradients, fc_int included, against the oracle on the reference's own width_dim=2 case."""
    C, name = golden_cases, "n6_b7_w2"
    N, B, wd = (int(v) for v in C[f"{name}/meta"])
    lstm, fnn = state_dicts(golden_weights, "w2")
    X, Z = C[f"{name}/X"], C[f"{name}/Z"]
    u0 = np.ascontiguousarray(C[f"{name}/f32/u0"])
    wp = _pack_v(emu, lstm, fnn, "tc")
    int_w = np.ascontiguousarray(fnn["fc_int.weight"], dtype=np.float32)
    int_b = np.ascontiguousarray(fnn["fc_int.bias"], dtype=np.float32)
    o = {k: np.zeros(B, np.float32) for k in ("cost", "command", "error", "du0")}
    o["pred"] = np.zeros((B, N), np.float32)
    o["gl"] = np.zeros(256, np.float32)
    glw = np.zeros(2560, np.float32)
    emu.fc_emu_mpc_loss_tc_wide(_p(X), _p(u0), _p(Z), _p(wp), _p(int_w), _p(int_b), wd, B, N, ctypes.c_float(20.0),
                                ctypes.c_longlong(B), 1, 1, _p(o["cost"]), _p(o["command"]), _p(o["error"]), _p(o["pred"]),
                                _p(o["du0"]), _p(o["gl"]), _p(glw))
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    out, g = O.mpc_loss_forward_backward(w, X.astype(np.float64), u0.astype(np.float64), Z.astype(np.float64), N, 20.0, width_dim=wd)
    _check_v(o, out, g)
    assert rel_max(glw[:2500].reshape(50, 50), g["int_w"]) < 1e-5
    assert rel_max(glw[2500:2550], np.asarray(g["int_b"]).reshape(-1)) < 1e-5
    assert abs(o["gl"][250] - C[f"{name}/f32/loss"]) / abs(C[f"{name}/f32/loss"]) < 1e-5     # the reference itself


# ---- surrogate training mode of the pair-kernel source (MpcParams::train): forward, and the operand-format scratch the
# reverse sweep hands to the tcgen05 weight-gradient kernel (gate gradients, hidden sequences, features), decoded on the
# host and contracted over the samples exactly as fc::lt2::dw_kernel does -> the eight gradient tensors of the oracle.
def _decode_image(img, pieces, stage_rows=32, tile=128):
    """[hi | lo][stage][piece][stage_rows][8 halves] (fp16 bit patterns in float32 words) -> float64 [tile][pieces * 8]."""
    h = img.view(np.float16).astype(np.float64).reshape(2, tile // stage_rows, pieces, stage_rows, 8)
    v = h[0] + h[1]                                          # hi + lo
    return v.transpose(0, 2, 1, 3).reshape(tile, pieces * 8)


@pytest.mark.parametrize("replica", [0, 1])
def test_emulated_training_mode_forward_and_weight_gradient_operands(emu, golden_weights, replica):
    import lstm_train_oracle as T
    lstm, fnn = state_dicts(golden_weights, "c0")
    wp = _pack_v(emu, lstm, fnn, "pair")
    rng = np.random.default_rng(11)
    rows = 32 if replica else 128
    B = 150 if not replica else 40                           # one emulated CTA, two tiles (two passes in replica mode), the last ragged
    ntile = -(-B // rows)
    X = np.ascontiguousarray(rng.uniform(-1, 1, (B, 10, 5)), dtype=np.float32)
    dy = np.ascontiguousarray(rng.standard_normal((B, 4)) / B, dtype=np.float32)
    fc_w = np.ascontiguousarray(lstm["fc.weight"], dtype=np.float32)
    fc_b = np.ascontiguousarray(lstm["fc.bias"], dtype=np.float32)
    emu.fc_emu_train_tile_floats.restype = ctypes.c_long
    tile_floats = int(emu.fc_emu_train_tile_floats()) // (4 if replica else 1)
    y = np.zeros((B, 4), np.float32)
    hlast = np.zeros((B, 50), np.float32)
    ws = np.zeros(ntile * tile_floats, np.float32)
    scale = 2.0 ** 16
    for mode in (1, 2):
        emu.fc_emu_lstm_train(mode, _p(X), _p(dy), _p(wp), _p(fc_w), _p(fc_b), B, ctypes.c_float(scale), 1, _p(y), _p(hlast), _p(ws), replica)
    w = O.weights_from_state_dicts(lstm, fnn, np.float64)
    _, out_o, grads_o = T.lstm_mse_forward_backward(w, X.astype(np.float64), np.zeros((B, 4)), d_out=dy.astype(np.float64))
    assert rel_max(y, out_o) < 1e-5
    assert rel_max(dy.astype(np.float64).T @ hlast.astype(np.float64), grads_o["fc.weight"]) < 1e-5
    # layout constants of fc_pair_layout.h (kTr*), scaled with the samples of a tile
    hs_slot, ft_slot, dg_slot = 2 * 7 * rows * 4, 2 * rows * 4, 2 * 26 * rows * 4
    hs_off, ft_off = 0, 3 * 11 * hs_slot
    dg_off = ft_off + 10 * ft_slot
    assert tile_floats == dg_off + 30 * dg_slot
    g_ih = [np.zeros((200, 5)), np.zeros((200, 50)), np.zeros((200, 50))]
    g_hh = [np.zeros((200, 50)) for _ in range(3)]
    perm = np.array([(m % 4) * 50 + m // 4 for m in range(200)])           # operand row m = unit * 4 + gate -> PyTorch row
    for tile in range(ntile):
        base = tile * tile_floats
        for l in range(3):
            for t in range(10):
                dg = _decode_image(ws[base + dg_off + (l * 10 + t) * dg_slot:][:dg_slot], 26, tile=rows)[:, :200] / scale
                rec = _decode_image(ws[base + hs_off + (l * 11 + t) * hs_slot:][:hs_slot], 7, tile=rows)[:, :50] / 1024.0      # h_{t-1}, slot 0 = 0
                if l == 0:
                    inp = _decode_image(ws[base + ft_off + t * ft_slot:][:ft_slot], 1, tile=rows)[:, :5] / 1024.0
                else:
                    inp = _decode_image(ws[base + hs_off + ((l - 1) * 11 + t + 1) * hs_slot:][:hs_slot], 7, tile=rows)[:, :50] / 1024.0
                g_ih[l][perm] += dg.T @ inp
                g_hh[l][perm] += dg.T @ rec
    for l in range(3):
        assert rel_max(g_ih[l], grads_o[f"lstm.weight_ih_l{l}"]) < 1e-5, l
        assert rel_max(g_hh[l], grads_o[f"lstm.weight_hh_l{l}"]) < 1e-5, l
