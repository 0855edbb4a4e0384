"""TEST INFRASTRUCTURE ONLY (oracle side) -- never imported by the product package.

Imports the *unmodified* reference ``Unsupervised Learning/Functions.py`` from the read-only
mount ``/root/reference`` inside the build container.  The reference imports third-party modules
that are absent from this image at module import time (casadi, do_mpc, plotly, alive_progress;
``Functions.py:24-42``) and evaluates annotations such as ``do_mpc.simulator.Simulator`` at
``def`` time (``Functions.py:1014, 1615, 1790``).  We register permissive stub modules for those
names in ``sys.modules`` so that the torch-only part of the file (``FNNModel`` ``:215-289``,
``LSTMModel`` ``:295-379``, ``MPCLoss`` ``:1336-1472``, ``NeuralNetwork.tvp_fun`` ``:926-966``,
``FeasibilityRecovery.NN_make_step`` ``:1560-1613``) can execute.  Nothing from the reference is
copied; it is loaded by path.

This module only works where ``/root/reference`` exists (the build container).  It is used by
``oracle/make_golden.py`` to produce the committed fixtures in ``tests/golden`` and by the
``reference-live`` CPU tests, which skip when the mount is absent (e.g. on the GPU box).
"""
from __future__ import annotations

import importlib.util
import os
import pickle
import sys
import types

REFERENCE_ROOT = os.environ.get("FORGING_REFERENCE_ROOT", "/root/reference")
UL_DIR = os.path.join(REFERENCE_ROOT, "Unsupervised Learning")
SL_DIR = os.path.join(REFERENCE_ROOT, "Supervised Learning")
MNN_DIR = os.path.join(UL_DIR, "Model_NN")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(UL_DIR, "Functions.py"))


class _AnyMeta(type):
    """Metaclass whose attribute lookups manufacture further dummy classes."""

    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _AnyMeta(name, (), {})


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _AnyMeta(name, (), {})


_STUBS = [
    "casadi", "casadi.tools", "plotly", "plotly.graph_objects", "plotly.subplots", "plotly.io",
    "plotly.io._base_renderers", "plotly.io._renderers", "alive_progress", "do_mpc",
    "do_mpc.tools", "do_mpc.simulator", "do_mpc.controller", "do_mpc.controller._mpc",
    "do_mpc.tools._timer", "do_mpc.data", "do_mpc.model",
]


def _install_stubs():
    for name in _STUBS:
        if name in sys.modules:
            continue
        mod = _StubModule(name)
        mod.__path__ = []  # behave like a package so that sub-imports resolve
        sys.modules[name] = mod
    sys.modules["casadi"].__all__ = ["MX"]
    sys.modules["casadi"].MX = _AnyMeta("MX", (), {})
    sys.modules["casadi.tools"].__all__ = []

    class BrowserRenderer:  # the reference subclasses it (Functions.py, TitleBrowserRenderer)
        def __init__(self, *a, **k):
            pass

    sys.modules["plotly.io._base_renderers"].BrowserRenderer = BrowserRenderer
    sys.modules["plotly.io._base_renderers"].open_html_in_browser = lambda *a, **k: None
    sys.modules["plotly.io._renderers"].renderers = {}


_CACHE = {}


def load_reference_functions(fresh: bool = False):
    """Return the reference ``Functions`` module (UL variant), imported by path.  ``fresh`` imports a private copy of
    the module object (for tests that patch it, e.g. ``forging_control_b200.install``)."""
    if "ul" in _CACHE and not fresh:
        return _CACHE["ul"]
    if not reference_available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    cwd = os.getcwd()
    # the module opens 'my_log.log' for writing at import (Functions.py:59-60): do it in /tmp
    os.makedirs("/tmp/forging_ref_import", exist_ok=True)
    os.chdir("/tmp/forging_ref_import")
    try:
        spec = importlib.util.spec_from_file_location(
            "forging_reference_UL_Functions" + ("_fresh" if fresh else ""), os.path.join(UL_DIR, "Functions.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    if not fresh:
        _CACHE["ul"] = mod
    return mod


def load_reference_surrogate_functions():
    """Return the reference ``Model_NN/Functions.py`` module (surrogate training variant), imported by path."""
    if "mnn" in _CACHE:
        return _CACHE["mnn"]
    if not os.path.isfile(os.path.join(MNN_DIR, "Functions.py")):
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    cwd = os.getcwd()
    os.makedirs("/tmp/forging_ref_import", exist_ok=True)
    os.chdir("/tmp/forging_ref_import")
    try:
        spec = importlib.util.spec_from_file_location(
            "forging_reference_MNN_Functions", os.path.join(MNN_DIR, "Functions.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    _CACHE["mnn"] = mod
    return mod


class _Dummy:
    """Stand-in for every non-numpy class found in the do-mpc result pickles."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__["_state"] = state

    def __call__(self, *a, **k):
        return _Dummy()


class _PermissiveUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in ("numpy", "builtins", "collections", "copyreg", "_codecs"):
            return super().find_class(module, name)
        return type(name, (_Dummy,), {})


def load_dompc_arrays(path):
    """Extract the ``simulator`` data arrays (_x,_u,_y,_tvp,_time,_aux) of a do-mpc result pickle."""
    with open(path, "rb") as fh:
        obj = _PermissiveUnpickler(fh).load()
    sim = obj["simulator"]
    state = sim.__dict__.get("_state", sim.__dict__)
    if not isinstance(state, dict):
        state = state.__dict__
    out = {}
    for k in ("_x", "_u", "_y", "_tvp", "_time", "_aux"):
        if k in state:
            out[k] = state[k]
    return out


def load_sklearn_scaler(path):
    with open(path, "rb") as fh:
        return pickle.load(fh)
