import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _load_ref():
    """The UNMODIFIED reference extensions built by oracle/build_ref.py (None if unavailable)."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not all(os.path.exists(os.path.join(ref_dir, n + ".so"))
               for n in ("pointnet2_cuda", "roipool3d_cuda", "iou3d_cuda")):
        try:
            from oracle import build_ref
            build_ref.build()
        except Exception:
            return None
    if not os.path.exists(os.path.join(ref_dir, "pointnet2_cuda.so")):
        return None
    import importlib.util
    import types

    import torch  # noqa: F401  (the extensions link against libtorch)
    ns = types.SimpleNamespace()
    for name in ("pointnet2_cuda", "roipool3d_cuda", "iou3d_cuda"):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ref_dir, name + ".so"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        setattr(ns, name, mod)
    return ns


@pytest.fixture(scope="session")
def ref_ext():
    ns = _load_ref()
    if ns is None:
        pytest.skip("oracle/_ref (compiled reference) not available")
    return ns


@pytest.fixture(scope="session")
def cref():
    from oracle import cref as c
    c.build()
    return c


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from jmodt_b200 import _lib
    _lib.lib()  # raises if the library was not built: GPU tests must run the native code
    return torch.device("cuda:0")


def clustered_cloud(rng, b, n, spread=1.0, dup_frac=0.05):
    """(b,n,3) float32 cloud with clusters and exact duplicates (exercises tie rules)."""
    k = max(1, n // 64)
    ctr = rng.uniform(-10, 10, (b, k, 3)) * spread
    which = rng.integers(0, k, (b, n))
    pts = np.take_along_axis(ctr, which[..., None].repeat(3, -1), 1) + rng.normal(0, 0.5, (b, n, 3))
    nd = int(n * dup_frac)
    if nd:
        for bi in range(b):
            src = rng.integers(0, n, nd)
            dst = rng.integers(0, n, nd)
            pts[bi, dst] = pts[bi, src]
    return pts.astype(np.float32)
