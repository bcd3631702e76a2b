"""PointNet++ set-abstraction / feature-propagation modules with the reference's constructor
signatures, forward contracts and state_dict keys (reference
jmodt/ops/pointnet2/pointnet2_modules.py:11-164).

forward() composes the sm_100a ops of this package (FPS, gather, ball query, grouping,
three_nn, three_interpolate) with the SharedMLP stacks.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils
from . import pytorch_utils as pt_utils


class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None
        self.pool_method = 'max_pool'

    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None, new_xyz=None):
        """
        :param xyz: (B, N, 3), features: (B, C, N)
        :return: new_xyz (B, npoint, 3), new_features (B, sum_k mlps[k][-1], npoint), idx (B, npoint) | None
        """
        idx = None
        if new_xyz is None and self.npoint is not None:
            idx = pointnet2_utils.farthest_point_sample(xyz, self.npoint)
            new_xyz = pointnet2_utils.gather_operation(
                xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()

        pooled = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            grouped = mlp(grouper(xyz, new_xyz, features))  # (B, mlp[-1], npoint, nsample)
            if self.pool_method == 'max_pool':
                grouped = F.max_pool2d(grouped, kernel_size=[1, grouped.size(3)])
            elif self.pool_method == 'avg_pool':
                grouped = F.avg_pool2d(grouped, kernel_size=[1, grouped.size(3)])
            else:
                raise NotImplementedError
            pooled.append(grouped.squeeze(-1))  # (B, mlp[-1], npoint)
        return new_xyz, torch.cat(pooled, dim=1), idx


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Set abstraction with multi-scale grouping (pointnet2_modules.py:66-99)."""

    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int], mlps: List[List[int]],
                 bn: bool = True, use_xyz: bool = True, pool_method='max_pool', instance_norm=False):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for radius, nsample, mlp_spec in zip(radii, nsamples, mlps):
            self.groupers.append(pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz)
                                 if npoint is not None else pointnet2_utils.GroupAll(use_xyz))
            if use_xyz:
                mlp_spec[0] += 3  # in place, like the reference (:95-97): callers rely on it
            self.mlps.append(pt_utils.SharedMLP(mlp_spec, bn=bn, instance_norm=instance_norm))
        self.pool_method = pool_method


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set abstraction (pointnet2_modules.py:102-121)."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pool_method='max_pool', instance_norm=False):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz,
                         pool_method=pool_method, instance_norm=instance_norm)


class PointnetFPModule(nn.Module):
    """Feature propagation (pointnet2_modules.py:124-164)."""

    def __init__(self, *, mlp: List[int], bn: bool = True, activation=nn.ReLU(inplace=True)):
        super().__init__()
        self.mlp = pt_utils.SharedMLP(mlp, bn=bn, activation=activation)

    def forward(self, unknown: torch.Tensor, known: torch.Tensor, unknow_feats: torch.Tensor,
                known_feats: torch.Tensor) -> torch.Tensor:
        """
        :param unknown: (B, n, 3), known: (B, m, 3), unknow_feats: (B, C1, n), known_feats: (B, C2, m)
        :return: (B, mlp[-1], n)
        """
        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
            interpolated = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        new_features = interpolated if unknow_feats is None else torch.cat([interpolated, unknow_feats], dim=1)
        return self.mlp(new_features.unsqueeze(-1)).squeeze(-1)
