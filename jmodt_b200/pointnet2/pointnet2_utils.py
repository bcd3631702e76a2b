"""Python operator API of the reference's `jmodt/ops/pointnet2/pointnet2_utils.py`, backed by
the sm_100a kernels in libjmodt_b200.so.

Same public names, argument order, dtypes and layouts as the reference (file:line cited per
symbol) so `jmodt/detection` and `jmodt/tracking` can import this module unchanged:
index tensors are int32 CUDA tensors, features are fp32 channel-first (B, C, N), coordinates
are contiguous (B, N, 3).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _lib


def _scatter_add_sorted(src: torch.Tensor, idx_flat: torch.Tensor, n_targets: int, rep: int = 1,
                        weight: torch.Tensor = None) -> torch.Tensor:
    """out[b, c, t] = sum over the positions q with idx_flat[b, q] == t, in ascending q, of
    src[b, c, q // rep] * (weight[b, q] if weight is given).  The backward of gather / grouping / interpolation with a
    FIXED summation order (the reference's atomicAdd kernels are order-nondeterministic): the index list is sorted
    once (stable), the runs are summed sequentially by csrc/interpolate.cu:segmented_scatter_kernel."""
    B, C, L_src = src.shape
    Lq = idx_flat.shape[1]
    sorted_idx, order = torch.sort(idx_flat.long(), dim=1, stable=True)
    bounds = torch.arange(n_targets + 1, device=src.device, dtype=torch.long).unsqueeze(0).expand(B, -1).contiguous()
    seg_off = torch.searchsorted(sorted_idx.contiguous(), bounds).to(torch.int32).contiguous()
    order = order.to(torch.int32).contiguous()
    out = torch.empty((B, C, n_targets), dtype=torch.float32, device=src.device)
    src = src.contiguous()
    w = None if weight is None else weight.reshape(B, Lq).contiguous()
    st = _lib.stream_and_device(src)
    _lib.check(_lib.lib().jmb_segmented_scatter_add(B, C, L_src, Lq, n_targets, rep, src.data_ptr(), order.data_ptr(),
                                                    seg_off.data_ptr(), _lib.ptr(w), out.data_ptr(), st),
               "segmented_scatter_add")
    return out

from . import pointnet2_cuda


def _new(ref: torch.Tensor, shape, dtype):
    return torch.empty(shape, dtype=dtype, device=ref.device)


class farthestPointSampling(Function):
    """reference pointnet2_utils.py:10-36"""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        output = _new(xyz, (B, npoint), torch.int32)
        # the kernel keeps the running min-distances in registers; no (B, N) temp tensor
        pointnet2_cuda.farthest_point_sampling_wrapper(B, N, npoint, xyz, None, output)
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(ctx, a=None):
        return None, None


farthest_point_sample = farthestPointSampling.apply


class GatherOperation(Function):
    """reference pointnet2_utils.py:39-73"""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        B, npoint = idx.size()
        _, C, N = features.size()
        output = _new(features, (B, C, npoint), torch.float32)
        pointnet2_cuda.gather_points_wrapper(B, C, N, npoint, features, idx, output)
        ctx.for_backwards = (idx, C, N)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        grad_features = _scatter_add_sorted(grad_out, idx.view(B, npoint), N)      # fixed summation order
        return grad_features, None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    """reference pointnet2_utils.py:76-105 — returns (sqrt(dist2), idx)"""

    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        assert unknown.is_contiguous()
        assert known.is_contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = _new(unknown, (B, N, 3), torch.float32)
        idx = _new(unknown, (B, N, 3), torch.int32)
        pointnet2_cuda.three_nn_wrapper(B, N, m, unknown, known, dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    """reference pointnet2_utils.py:108-153"""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        assert weight.is_contiguous()
        B, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        output = _new(features, (B, c, n), torch.float32)
        pointnet2_cuda.three_interpolate_wrapper(B, c, m, n, features.float(), idx, weight, output)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        grad_features = _scatter_add_sorted(grad_out, idx.view(B, n * 3), m, rep=3, weight=weight)   # fixed summation order
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    """reference pointnet2_utils.py:156-197"""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        B, nfeatures, nsample = idx.size()
        _, C, N = features.size()
        output = _new(features, (B, C, nfeatures, nsample), torch.float32)
        pointnet2_cuda.group_points_wrapper(B, C, N, nfeatures, nsample, features.float(), idx, output)
        ctx.for_backwards = (idx, N)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        grad_features = _scatter_add_sorted(grad_out.reshape(B, C, npoint * nsample), idx.view(B, npoint * nsample), N)
        return grad_features, None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    """reference pointnet2_utils.py:200-228"""

    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        assert new_xyz.is_contiguous()
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        idx = _new(xyz, (B, npoint, nsample), torch.int32)  # fully written by the kernel
        pointnet2_cuda.ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


def ball_query_msg2(radius_a: float, nsample_a: int, radius_b: float, nsample_b: int, xyz: torch.Tensor,
                    new_xyz: torch.Tensor):
    """Two ball queries over the same centres in ONE scan of the cloud (what PointnetSAModuleMSG needs per level,
    reference pointnet2_modules.py:41-42); returns the same two index tensors as two `ball_query` calls."""
    from .. import _lib
    assert new_xyz.is_contiguous() and xyz.is_contiguous()
    B, N, _ = xyz.size()
    npoint = new_xyz.size(1)
    idx_a = _new(xyz, (B, npoint, nsample_a), torch.int32)
    idx_b = _new(xyz, (B, npoint, nsample_b), torch.int32)
    st = _lib.stream_and_device(xyz)
    L = _lib.lib()
    if N >= GRID_MIN_POINTS:
        # large cloud: hashed cell list, a centre visits 27 cells instead of all N points (identical result)
        ws_bytes = L.jmb_ball_query_grid_workspace_bytes(B, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xyz.device)
        _lib.check(L.jmb_ball_query_msg2_grid(B, N, npoint, radius_a, nsample_a, radius_b, nsample_b, new_xyz.data_ptr(),
                                              xyz.data_ptr(), idx_a.data_ptr(), idx_b.data_ptr(), ws.data_ptr(), ws_bytes,
                                              st), "ball_query_msg2_grid")
        _lib.launch_count += 1
        return idx_a, idx_b
    _lib.check(L.jmb_ball_query_msg2(B, N, npoint, radius_a, nsample_a, radius_b, nsample_b,
                                     new_xyz.data_ptr(), xyz.data_ptr(), idx_a.data_ptr(), idx_b.data_ptr(),
                                     st), "ball_query_msg2")
    return idx_a, idx_b


GRID_MIN_POINTS = 8192      # clouds at least this large go through the cell-list ball query (RPN level 0)


class QueryAndGroup(nn.Module):
    """reference pointnet2_utils.py:231-264"""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)  # (B, 3, npoint, nsample)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return grouped_xyz
        grouped_features = grouping_operation(features, idx)
        if self.use_xyz:
            return torch.cat([grouped_xyz, grouped_features], dim=1)  # (B, 3 + C, npoint, nsample)
        return grouped_features


class GroupAll(nn.Module):
    """reference pointnet2_utils.py:267-290"""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        grouped_xyz = xyz.transpose(1, 2).contiguous().unsqueeze(2)  # (B, 3, 1, N)
        if features is None:
            return grouped_xyz
        grouped_features = features.unsqueeze(2)
        if self.use_xyz:
            return torch.cat([grouped_xyz, grouped_features], dim=1)  # (B, 3 + C, 1, N)
        return grouped_features
