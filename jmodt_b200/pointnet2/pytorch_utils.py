"""`SharedMLP`, `Conv1d`, `Conv2d`, `FC`, `BatchNorm1d/2d` with the reference's constructor
signatures and — importantly — its parameter names (reference
jmodt/ops/pointnet2/pytorch_utils.py:6-236), so reference checkpoints load key for key:

    <name>layer{i}.conv.weight / .conv.bias / .bn.bn.{weight,bias,running_mean,running_var}

(`load_checkpoint` in the reference uses strict=False, train_utils.py:31-47, which would
silently skip a renamed key.)  These modules only hold parameters and define the unfused
forward; the fused sm_100a kernels read the same parameters through
`jmodt_b200.fused.fold_shared_mlp`.
"""
from __future__ import annotations

from typing import List, Tuple

import torch.nn as nn


def _norm_act_conv(seq: nn.Sequential, *, conv, norm, act, inorm, preact: bool, name: str):
    """Registers the sub-modules in the order the reference does (pytorch_utils.py:82-102)."""
    tail = [("bn", norm), ("activation", act), ("in", inorm if norm is None else None)]
    if preact:
        for key, mod in tail:
            if mod is not None:
                seq.add_module(name + key, mod)
    seq.add_module(name + "conv", conv)
    if not preact:
        for key, mod in tail:
            if mod is not None:
                seq.add_module(name + key, mod)


class _BNBase(nn.Sequential):
    def __init__(self, in_size, batch_norm=None, name=""):
        super().__init__()
        self.add_module(name + "bn", batch_norm(in_size))
        nn.init.constant_(self[0].weight, 1.0)
        nn.init.constant_(self[0].bias, 0)


class BatchNorm1d(_BNBase):
    def __init__(self, in_size: int, *, name: str = ""):
        super().__init__(in_size, batch_norm=nn.BatchNorm1d, name=name)


class BatchNorm2d(_BNBase):
    def __init__(self, in_size: int, name: str = ""):
        super().__init__(in_size, batch_norm=nn.BatchNorm2d, name=name)


class _ConvBase(nn.Sequential):
    def __init__(self, in_size, out_size, kernel_size, stride, padding, activation, bn, init, conv=None,
                 batch_norm=None, bias=True, preact=False, name="", instance_norm=False,
                 instance_norm_func=None):
        super().__init__()
        bias = bias and (not bn)  # pytorch_utils.py:58
        conv_unit = conv(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding, bias=bias)
        init(conv_unit.weight)
        if bias:
            nn.init.constant_(conv_unit.bias, 0)
        width = in_size if preact else out_size
        norm = batch_norm(width) if bn else None
        inorm = instance_norm_func(width, affine=False, track_running_stats=False) if instance_norm else None
        _norm_act_conv(self, conv=conv_unit, norm=norm, act=activation, inorm=inorm, preact=preact, name=name)


class Conv1d(_ConvBase):
    def __init__(self, in_size: int, out_size: int, *, kernel_size: int = 1, stride: int = 1, padding: int = 0,
                 activation=nn.ReLU(inplace=True), bn: bool = False, init=nn.init.kaiming_normal_,
                 bias: bool = True, preact: bool = False, name: str = "", instance_norm=False):
        super().__init__(in_size, out_size, kernel_size, stride, padding, activation, bn, init, conv=nn.Conv1d,
                         batch_norm=BatchNorm1d, bias=bias, preact=preact, name=name,
                         instance_norm=instance_norm, instance_norm_func=nn.InstanceNorm1d)


class Conv2d(_ConvBase):
    def __init__(self, in_size: int, out_size: int, *, kernel_size: Tuple[int, int] = (1, 1),
                 stride: Tuple[int, int] = (1, 1), padding: Tuple[int, int] = (0, 0),
                 activation=nn.ReLU(inplace=True), bn: bool = False, init=nn.init.kaiming_normal_,
                 bias: bool = True, preact: bool = False, name: str = "", instance_norm=False):
        super().__init__(in_size, out_size, kernel_size, stride, padding, activation, bn, init, conv=nn.Conv2d,
                         batch_norm=BatchNorm2d, bias=bias, preact=preact, name=name,
                         instance_norm=instance_norm, instance_norm_func=nn.InstanceNorm2d)


class SharedMLP(nn.Sequential):
    """Stack of 1x1 Conv2d [+BN] [+ReLU] (pytorch_utils.py:6-33)."""

    def __init__(self, args: List[int], *, bn: bool = False, activation=nn.ReLU(inplace=True),
                 preact: bool = False, first: bool = False, name: str = "", instance_norm: bool = False):
        super().__init__()
        for i in range(len(args) - 1):
            plain = first and preact and i == 0  # the very first pre-activation layer has no BN/act
            self.add_module(
                name + "layer{}".format(i),
                Conv2d(args[i], args[i + 1], bn=bn and not plain, activation=None if plain else activation,
                       preact=preact, instance_norm=instance_norm))


class FC(nn.Sequential):
    """pytorch_utils.py:201-236"""

    def __init__(self, in_size: int, out_size: int, *, activation=nn.ReLU(inplace=True), bn: bool = False,
                 init=None, preact: bool = False, name: str = ""):
        super().__init__()
        fc = nn.Linear(in_size, out_size, bias=not bn)
        if init is not None:
            init(fc.weight)
        if not bn:
            nn.init.constant_(fc.bias, 0)
        norm = BatchNorm1d(in_size if preact else out_size) if bn else None
        if preact:
            if norm is not None:
                self.add_module(name + "bn", norm)
            if activation is not None:
                self.add_module(name + "activation", activation)
        self.add_module(name + "fc", fc)
        if not preact:
            if norm is not None:
                self.add_module(name + "bn", norm)
            if activation is not None:
                self.add_module(name + "activation", activation)
