"""Import the REAL reference (read-only, /root/reference) inside this build container.

Only `make_golden.py` and the `test_oracle_vs_reference*` tests use this file, and only
when /root/reference exists (it does not exist on the GPU box).  The reference pulls in
matplotlib / dgl / open3d at import time (utils.py:12-20, planners.py:18,
env/flex_rewards.py:4); none of them is on the hot path, so inert stub modules are put
into sys.modules first.  `torch.cuda.Event` / `torch.cuda.synchronize` are used
unconditionally by `ptcl_model_rollout` (planners.py:341-358) -> replaced by wall-clock
stand-ins while the CPU reference runs.
"""
import contextlib
import os
import sys
import time
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("PILE_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "model", "gnn_dyn.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _install_stubs():
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.patches = _stub("matplotlib.patches", Ellipse=object)
        mpl.transforms = _stub("matplotlib.transforms")
        mpl.font_manager = _stub("matplotlib.font_manager")
    try:
        import dgl  # noqa: F401
    except Exception:
        dgl = _stub("dgl")
        dgl.geometry = _stub("dgl.geometry", farthest_point_sampler=None)
    try:
        import open3d  # noqa: F401
    except Exception:
        _stub("open3d")


_REF = {}


def load_reference():
    """-> dict(gnn_dyn=<module>, planners=<module>, flex_rewards=<module>, utils=<module>)."""
    if _REF:
        return _REF
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    _REF["gnn_dyn"] = importlib.import_module("model.gnn_dyn")
    _REF["utils"] = importlib.import_module("utils")
    _REF["flex_rewards"] = importlib.import_module("env.flex_rewards")
    _REF["planners"] = importlib.import_module("planners")
    return _REF


class _WallEvent:
    def __init__(self, enable_timing=True):
        self.t = 0.0

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


@contextlib.contextmanager
def cpu_cuda_shims():
    """Let planners.ptcl_model_rollout run without a GPU."""
    ev, sync = torch.cuda.Event, torch.cuda.synchronize
    torch.cuda.Event = _WallEvent
    torch.cuda.synchronize = lambda *a, **k: None
    try:
        yield
    finally:
        torch.cuda.Event, torch.cuda.synchronize = ev, sync


@contextlib.contextmanager
def in_reference_cwd():
    """utils.gen_goal_shape opens 'env/target_shapes/...' relative to cwd (utils.py:567)."""
    old = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        yield
    finally:
        os.chdir(old)
