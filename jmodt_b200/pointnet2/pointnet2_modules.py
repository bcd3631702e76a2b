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
from .. import runtime, tc


def _use_fused(module: nn.Module) -> bool:
    """Inference (eval mode, no autograd) runs the fused tcgen05 path; training keeps the differentiable
    composition of the individual ops."""
    return bool(getattr(module, "fused", True)) and not module.training and not torch.is_grad_enabled()


def pack_shared_mlp(mlp) -> list:
    """SharedMLP -> list of tc.PackedLayer with eval-mode BatchNorm folded in (pytorch_utils.py:36-102)."""
    packed = []
    for blk in mlp:
        assert not hasattr(blk, "in"), "instance norm is not supported by the fused path"
        bn = blk.bn.bn if hasattr(blk, "bn") else None
        w, b = tc.fold_conv_bn(blk.conv, bn)
        packed.append(tc.PackedLayer(w, b, relu=hasattr(blk, "activation")))
    return packed


class SAPlan:
    """Coordinate-only part of a set-abstraction layer: FPS indices (B, npoint) | None, centres (B, npoint, 3) | None,
    one neighbour list (B, npoint, nsample) int32 per scale (None for GroupAll)."""
    __slots__ = ("idx", "new_xyz", "nbr", "filled")

    def __init__(self, idx, new_xyz, nbr):
        self.idx, self.new_xyz, self.nbr = idx, new_xyz, nbr

    def tensors(self):
        return [t for t in (self.idx, self.new_xyz, *(self.nbr or ())) if t is not None]


class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None
        self.pool_method = 'max_pool'

    def plan(self, xyz: torch.Tensor, new_xyz=None):
        """Everything of the layer that depends on coordinates only — FPS indices, the sampled centres and the
        ball-query neighbour lists (pointnet2_modules.py:36-50, pointnet2_utils.py:236-252).  A caller that knows
        xyz early can run this on a side stream while the previous level's features are still being computed and
        pass the result to forward(plan=...)."""
        idx = None
        if new_xyz is None and self.npoint is not None:
            idx = pointnet2_utils.farthest_point_sample(xyz, self.npoint)
            new_xyz = pointnet2_utils.gather_operation(
                xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        nbr = [None] * len(self.groupers)
        qg = [isinstance(g, pointnet2_utils.QueryAndGroup) for g in self.groupers]
        if len(self.groupers) == 2 and all(qg):
            ga, gb = self.groupers      # both scales of an MSG level query the same centres: one scan of the cloud
            nbr[0], nbr[1] = pointnet2_utils.ball_query_msg2(ga.radius, ga.nsample, gb.radius, gb.nsample, xyz, new_xyz)
        else:
            for gi, g in enumerate(self.groupers):
                if qg[gi]:
                    nbr[gi] = pointnet2_utils.ball_query(g.radius, g.nsample, xyz, new_xyz)
        return SAPlan(idx, new_xyz, nbr)

    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None, new_xyz=None, plan=None):
        """
        :param xyz: (B, N, 3), features: (B, C, N)
        :return: new_xyz (B, npoint, 3), new_features (B, sum_k mlps[k][-1], npoint), idx (B, npoint) | None
        """
        if _use_fused(self) and self.pool_method == 'max_pool':
            if plan is None:
                plan = self.plan(xyz, new_xyz)
            return plan.new_xyz, self._forward_fused(xyz, features, plan.new_xyz, plan.nbr), plan.idx

        idx = None
        if plan is not None:
            idx, new_xyz = plan.idx, plan.new_xyz
        elif new_xyz is None and self.npoint is not None:
            idx = pointnet2_utils.farthest_point_sample(xyz, self.npoint)
            new_xyz = pointnet2_utils.gather_operation(
                xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()

        pooled = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            with pt_utils.torch_layers():    # op-by-op composition: the SharedMLP runs its torch forward (autograd / reference)
                grouped = mlp(grouper(xyz, new_xyz, features))  # (B, mlp[-1], npoint, nsample)
            if self.pool_method == 'max_pool':
                grouped = F.max_pool2d(grouped, kernel_size=[1, grouped.size(3)])
            elif self.pool_method == 'avg_pool':
                grouped = F.avg_pool2d(grouped, kernel_size=[1, grouped.size(3)])
            else:
                raise NotImplementedError
            pooled.append(grouped.squeeze(-1))  # (B, mlp[-1], npoint)
        return new_xyz, torch.cat(pooled, dim=1), idx


    def train(self, mode: bool = True):
        self._packed = None          # weights may change: re-pack on the next fused forward
        return super().train(mode)

    # ---- fused inference path: grouping fused into the first layer, max-pool into the last ----------
    def pack(self):
        return [pack_shared_mlp(m) for m in self.mlps]

    def _forward_fused(self, xyz, features, new_xyz, nbr):
        packed = tc.packed_for(self, self.pack)
        if features is not None:
            features = features.contiguous()
        fuse = getattr(self, "fuse_chain", True)
        c_in = 0 if features is None else features.shape[1]
        one_kernel = [isinstance(g, pointnet2_utils.QueryAndGroup) and fuse and
                      tc.sa_fused_supported(layers, c_in, new_xyz.shape[1], g.nsample)
                      for g, layers in zip(self.groupers, packed)]

        def scale(gi):
            grouper, layers = self.groupers[gi], packed[gi]
            if isinstance(grouper, pointnet2_utils.QueryAndGroup):
                assert grouper.use_xyz, "the fused path groups xyz with the features"
                idx = nbr[gi]
                pool = grouper.nsample
                if one_kernel[gi]:
                    return tc.sa_fused(layers, xyz, features, idx, new_xyz)   # first-layer GEMM over the points + one kernel
                l0 = layers[0]
                if (features is not None and len(layers) > 1 and l0.relu and l0._w32 is not None and l0.M % 4 == 0
                        and idx.shape[1] * idx.shape[2] > xyz.shape[1]):
                    # wide levels: the first layer over the POINTS, then the per-neighbour finish (tc.hoisted_first_layer)
                    h = tc.hoisted_first_layer(l0, xyz, features, idx, new_xyz)
                else:
                    h = tc.grouped_first_layer(l0, xyz, features, idx, new_xyz, grouper.nsample,
                                               pool=pool if len(layers) == 1 else 0)
            else:  # GroupAll
                pool = xyz.shape[1]
                assert pool <= 128 and 128 % pool == 0, "GroupAll fused path: point count must divide 128"
                h = tc.grouped_first_layer(layers[0], xyz, features, None, None, 0,
                                           pool=pool if len(layers) == 1 else 0)
            for i, layer in enumerate(layers[1:]):
                h = tc.mlp_layer(layer, h, pool=pool if i == len(layers) - 2 else 0)
            return h

        # the scales of a level are independent chains of small launches: run them on forked streams
        outs = runtime.parallel(*[(lambda gi=gi: scale(gi)) for gi in range(len(self.groupers))])
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)


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

    interp_after_first_layer = True      # fused path without skip features: see forward()

    def train(self, mode: bool = True):
        self._packed = None
        self._first_linear = None
        return super().train(mode)

    @staticmethod
    def plan(unknown: torch.Tensor, known: torch.Tensor):
        """Coordinate-only part (pointnet2_modules.py:148-151): the three nearest known points of every unknown
        point and their normalised inverse-distance weights."""
        dist, idx = pointnet2_utils.three_nn(unknown, known)
        dist_recip = 1.0 / (dist + 1e-8)
        weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
        return idx, weight

    def forward(self, unknown: torch.Tensor, known: torch.Tensor, unknow_feats: torch.Tensor,
                known_feats: torch.Tensor, plan=None) -> torch.Tensor:
        """
        :param unknown: (B, n, 3), known: (B, m, 3), unknow_feats: (B, C1, n), known_feats: (B, C2, m)
        :return: (B, mlp[-1], n)
        """
        fused = _use_fused(self)
        packed = None
        if fused:
            def build():
                self._first_linear = None
                return pack_shared_mlp(self.mlp)
            packed = tc.packed_for(self, build)
        if known is not None:
            idx, weight = plan if plan is not None else self.plan(unknown, known)
            if fused and self.interp_after_first_layer and packed[0].relu and packed[0]._w32 is not None and \
                    known_feats.shape[2] < unknown.shape[1]:
                # The first layer is linear up to its ReLU, and so is the interpolation (weights sum to one):
                #     relu(W . [interp(f); skip] + b) = relu(interp(W_a . f) + (W_b . skip + b)).
                # Running W_a on the m KNOWN points and interpolating its (narrower) output does n/m times fewer
                # FLOPs, gathers fewer channels and needs no concat (level 0, no skip features: 16 384 -> 4 096
                # columns, 256 -> 128 channels).  Same result to fp32 rounding.
                c2 = known_feats.shape[1]
                lin = getattr(self, "_first_linear", None)
                if lin is None or lin[0].K != c2:
                    w, b = packed[0]._w32, packed[0].bias[: packed[0].M]
                    if unknow_feats is None:
                        lin = (tc.PackedLayer(w, b, relu=False), None)
                    else:
                        lin = (tc.PackedLayer(w[:, :c2].contiguous(), None, relu=False),
                               tc.PackedLayer(w[:, c2:].contiguous(), b, relu=False))
                    self._first_linear = lin
                if unknow_feats is None:
                    h = tc.mlp_layer(lin[0], known_feats.contiguous())
                    h = torch.relu_(pointnet2_utils.three_interpolate(h, idx, weight))
                else:
                    ha, hb = runtime.parallel(lambda: tc.mlp_layer(lin[0], known_feats.contiguous()),
                                              lambda: tc.mlp_layer(lin[1], unknow_feats.contiguous()))
                    h = torch.relu_(pointnet2_utils.three_interpolate(ha, idx, weight).add_(hb))
                for layer in packed[1:]:
                    h = tc.mlp_layer(layer, h)
                return h
            interpolated = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        new_features = interpolated if unknow_feats is None else torch.cat([interpolated, unknow_feats], dim=1)
        if fused:
            h = new_features.contiguous()
            for layer in packed:
                h = tc.mlp_layer(layer, h)
            return h
        with pt_utils.torch_layers():
            return self.mlp(new_features.unsqueeze(-1)).squeeze(-1)
