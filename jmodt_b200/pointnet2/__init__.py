"""Drop-in for the reference package `jmodt.ops.pointnet2` (same module and symbol names)."""
from . import pointnet2_cuda, pointnet2_utils, pytorch_utils, pointnet2_modules  # noqa: F401
