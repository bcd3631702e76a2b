"""Importing the UNMODIFIED reference Python package (test infrastructure only).

The reference's host code is Python.  In the CPU container it is imported from where it lies (/root/reference); on
the GPU box that path does not exist, so `stage()` (called by oracle/build_ref.py next to the extension build)
mirrors the reference's `jmodt/**/*.py` files into `oracle/_ref/py/` — git-ignored, outputs only under oracle/_ref/,
shipped to the GPU box with the snapshot like the compiled extensions.  No reference source is tracked by this repo.

Two ways of importing it:

    import_reference(ops="reference")   the reference's own `jmodt/ops/*.py` wrappers over ITS compiled CUDA
                                        extensions (oracle/_ref/*.so): the golden generators use this
    import_reference(ops="dropin")      `jmodt_b200.dropin.install()` first, so the reference's detection /
                                        tracking code runs over THIS package's operators: the drop-in tests use this

Only tests/, tests/golden/make_golden_*.py, bench.py's CPU legs and __graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("JMODT_REFERENCE", "/root/reference")
STAGED = os.path.join(HERE, "_ref", "py")

EASYDICT_SHIM = '''"""Minimal stand-in for the `easydict` package (not installed here): attribute access on nested dicts."""


class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in {**(d or {}), **kw}.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    __setitem__ = __setattr__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)
'''


def stage(force: bool = False) -> bool:
    """Mirror the reference's python files into oracle/_ref/py (no-op without /root/reference)."""
    src = os.path.join(REF, "jmodt")
    if not os.path.isdir(src):
        return os.path.isdir(os.path.join(STAGED, "jmodt"))
    dst = os.path.join(STAGED, "jmodt")
    if os.path.isdir(dst) and not force:
        return True
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    for d, _, files in os.walk(src):
        for f in files:
            if f.endswith(".py"):
                rel = os.path.relpath(os.path.join(d, f), src)
                os.makedirs(os.path.dirname(os.path.join(dst, rel)), exist_ok=True)
                shutil.copyfile(os.path.join(d, f), os.path.join(dst, rel))
    with open(os.path.join(STAGED, "easydict.py"), "w") as fh:
        fh.write(EASYDICT_SHIM)
    return True


def python_root() -> str | None:
    """Directory to put on sys.path so that `import jmodt` finds the reference: the reference checkout when it is
    present (CPU container), else the staged mirror (GPU box)."""
    if os.path.isdir(os.path.join(REF, "jmodt")):
        return REF
    if os.path.isdir(os.path.join(STAGED, "jmodt")):
        return STAGED
    return None


def available() -> bool:
    return python_root() is not None


def _stub(name: str, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def import_reference(ops: str = "reference") -> None:
    """Make `import jmodt...` work in this process (idempotent).  See the module docstring for `ops`."""
    assert ops in ("reference", "dropin")
    root = python_root()
    if root is None:
        raise ImportError("reference python not available (no /root/reference and no oracle/_ref/py)")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    try:
        import easydict  # noqa: F401
    except ImportError:
        shim_dir = STAGED if os.path.exists(os.path.join(STAGED, "easydict.py")) else None
        if shim_dir is None:
            import tempfile
            shim_dir = tempfile.mkdtemp()
            with open(os.path.join(shim_dir, "easydict.py"), "w") as fh:
                fh.write(EASYDICT_SHIM)
        sys.path.insert(0, shim_dir)
    if root not in sys.path:
        sys.path.insert(0, root)
    # third-party packages the reference imports at module level but which are not on the hot path and not installed
    # here (data_association.py:3 -> ortools CBC solver, kalman.py:2 -> filterpy): stubbed so the modules import
    for name in ("ortools", "ortools.linear_solver", "filterpy", "filterpy.kalman", "tensorboardX"):
        try:
            importlib.import_module(name)
        except ImportError:
            _stub(name)
    if "pywraplp" not in sys.modules.get("ortools.linear_solver").__dict__:
        sys.modules["ortools.linear_solver"].pywraplp = types.SimpleNamespace()
    if "KalmanFilter" not in sys.modules.get("filterpy.kalman").__dict__:
        sys.modules["filterpy.kalman"].KalmanFilter = object
    if ops == "dropin":
        from jmodt_b200 import dropin
        dropin.install()
        return
    for pkg, name in (("jmodt.ops.pointnet2", "pointnet2_cuda"), ("jmodt.ops.roipool3d", "roipool3d_cuda"),
                      ("jmodt.ops.iou3d", "iou3d_cuda")):
        full = pkg + "." + name
        if full in sys.modules:
            continue
        so = os.path.join(HERE, "_ref", name + ".so")
        if os.path.exists(so):
            import torch  # noqa: F401  (the extensions link against libtorch)
            spec = importlib.util.spec_from_file_location(name, so)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        else:
            mod = types.ModuleType(name)
        sys.modules[full] = mod


def set_eval_cfg(post_nms_top_n: int = 100):
    """The reference's global cfg as tools/eval.py sets it for the joint model (README.md:96-107 / the shipped
    yaml): RPN + RCNN enabled, LI-Fusion on, TEST-mode proposal counts."""
    from jmodt.config import cfg
    cfg.RPN.ENABLED = True
    cfg.RCNN.ENABLED = True
    cfg.RPN.FIXED = False
    cfg.LI_FUSION.ENABLED = True
    cfg.TEST.RPN_POST_NMS_TOP_N = post_nms_top_n
    return cfg
