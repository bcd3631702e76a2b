"""Drop-in for the reference package `jmodt.ops.iou3d`."""
from . import iou3d_cuda, iou3d_utils  # noqa: F401
