"""jmodt_b200 — B200 (sm_100a) implementation of the JMODT per-frame hot path.

Sub-packages mirror the reference's operator API (`jmodt/ops/{pointnet2,roipool3d,iou3d}`);
`jmodt_b200.dropin.install()` aliases them under the reference's module paths so
`jmodt/detection` and `jmodt/tracking` import them unchanged.  Every op runs a hand-written
CUDA kernel from `jmodt_b200/csrc` through the C ABI in `include/jmodt_b200.h`; there is no
CPU or eager fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
