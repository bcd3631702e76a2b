"""Make `import jmodt.ops.pointnet2.pointnet2_utils` (etc.) resolve to jmodt_b200.

The reference imports its operators as
    from jmodt.ops.pointnet2 import pointnet2_cuda            (pointnet2_utils.py:7)
    from jmodt.ops.roipool3d import roipool3d_cuda            (roipool3d_utils.py:4)
    from jmodt.ops.iou3d import iou3d_cuda                    (iou3d_utils.py:3)
    import jmodt.ops.pointnet2.pytorch_utils as pt_utils      (rcnn.py:5, rpn.py:5)
    from jmodt.ops.pointnet2.pointnet2_modules import ...     (rcnn.py:8, backbone.py:6)
    from jmodt.ops.iou3d import iou3d_utils                   (proposal_target_layer.py:6)
    from jmodt.ops.roipool3d import roipool3d_utils           (proposal_target_layer.py:7)
`install()` registers this package's modules under those names in sys.modules.
"""
from __future__ import annotations

import importlib
import sys
import types

_MAP = {
    "jmodt.ops.pointnet2": "jmodt_b200.pointnet2",
    "jmodt.ops.pointnet2.pointnet2_cuda": "jmodt_b200.pointnet2.pointnet2_cuda",
    "jmodt.ops.pointnet2.pointnet2_utils": "jmodt_b200.pointnet2.pointnet2_utils",
    "jmodt.ops.pointnet2.pointnet2_modules": "jmodt_b200.pointnet2.pointnet2_modules",
    "jmodt.ops.pointnet2.pytorch_utils": "jmodt_b200.pointnet2.pytorch_utils",
    "jmodt.ops.roipool3d": "jmodt_b200.roipool3d",
    "jmodt.ops.roipool3d.roipool3d_cuda": "jmodt_b200.roipool3d.roipool3d_cuda",
    "jmodt.ops.roipool3d.roipool3d_utils": "jmodt_b200.roipool3d.roipool3d_utils",
    "jmodt.ops.iou3d": "jmodt_b200.iou3d",
    "jmodt.ops.iou3d.iou3d_cuda": "jmodt_b200.iou3d.iou3d_cuda",
    "jmodt.ops.iou3d.iou3d_utils": "jmodt_b200.iou3d.iou3d_utils",
}


def boxes_dist_gpu(boxes_a, boxes_b):
    """Drop-in for `jmodt.tracking.data_association.boxes_dist_gpu` (data_association.py:10-28): assign it over the
    reference function (`data_association.boxes_dist_gpu = dropin.boxes_dist_gpu`) — that module also imports ortools,
    so it is patched rather than aliased."""
    from .association import boxes_dist_gpu as f
    return f(boxes_a, boxes_b)


def install() -> None:
    """Alias the reference's operator module paths to this package (idempotent)."""
    for parent in ("jmodt", "jmodt.ops"):
        if parent not in sys.modules:
            try:
                importlib.import_module(parent)
            except ImportError:
                mod = types.ModuleType(parent)
                mod.__path__ = []  # namespace-like
                sys.modules[parent] = mod
    for ref_name, our_name in _MAP.items():
        mod = importlib.import_module(our_name)
        sys.modules[ref_name] = mod
        parent, _, leaf = ref_name.rpartition(".")
        setattr(sys.modules[parent], leaf, mod)
