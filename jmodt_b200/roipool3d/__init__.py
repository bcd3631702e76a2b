"""Drop-in for the reference package `jmodt.ops.roipool3d`."""
from . import roipool3d_cuda, roipool3d_utils  # noqa: F401
