"""`SharedMLP`, `Conv1d`, `Conv2d`, `FC`, `BatchNorm1d/2d` with the reference's constructor
signatures and — importantly — its parameter names (reference
jmodt/ops/pointnet2/pytorch_utils.py:6-236), so reference checkpoints load key for key:

    <name>layer{i}.conv.weight / .conv.bias / .bn.bn.{weight,bias,running_mean,running_var}

(`load_checkpoint` in the reference uses strict=False, train_utils.py:31-47, which would
silently skip a renamed key.)

forward(): in eval mode with autograd off and CUDA input, every 1x1 conv / Linear block (+ folded eval-mode
BatchNorm, + ReLU) runs on the tcgen05 layer kernel (`jmodt_b200.tc.mlp_layer`, csrc/tc_gemm.cu), so reference code
that calls these modules directly — `rcnn.py:178-196`, `rpn.py:81-82`, `tracker.py:86,106,109` — reaches the
tensor cores through the drop-in.  Training (or CPU tensors, or a block the kernel does not cover) takes the
module's own torch forward, which is what autograd needs.  The fused set-abstraction / feature-propagation paths
read the same parameters through `jmodt_b200.tc.fold_conv_bn`.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn as nn


_torch_only = 0


class torch_layers:
    """Context manager: inside it every block runs its torch (cuDNN / cuBLAS) forward even in eval mode — the
    reference composition the parity tests compare the tensor-core path with."""

    def __enter__(self):
        global _torch_only
        _torch_only += 1

    def __exit__(self, *exc):
        global _torch_only
        _torch_only -= 1


def _on_tensor_cores(module: nn.Module, x) -> bool:
    """Inference on the device: eval mode, autograd off, fp32 CUDA input."""
    return (_torch_only == 0 and not module.training and not torch.is_grad_enabled() and isinstance(x, torch.Tensor)
            and x.is_cuda and x.dtype == torch.float32 and x.dim() >= 2 and x.numel() > 0)


def _tc_linear(packed, x: torch.Tensor) -> torch.Tensor:
    """y[b, :, ...] = act(W . x[b, :, ...] + bias) on the tcgen05 layer kernel; x (B, C, *spatial) -> (B, M, *spatial).

    A batch of single-column problems — (G, C, 1) is how the reference feeds its cls / reg / link / start-end heads
    (`rcnn.py:195-196`, `tracker.py:86,106,109`) — is ONE (C, G) channel-first problem, not G tiles of one column: the
    result is returned as the transposed view (G, M, 1) of the (M, G) output, which the next layer of the stack
    recognises and consumes without a copy."""
    from .. import tc
    B, C = x.shape[0], x.shape[1]
    spatial = x.shape[2:]
    n = 1
    for d in spatial:
        n *= d
    if n == 1 and B > 1:
        xt = x.reshape(B, C).t()                      # (C, G): free when x is the previous layer's transposed view
        y = tc.mlp_layer(packed, xt.contiguous().unsqueeze(0))[0]          # (M, G)
        return y.t().reshape(B, packed.M, *spatial)   # a view: (G, M) has strides (1, G)
    y = tc.mlp_layer(packed, x.reshape(B, C, n).contiguous())
    return y.view(B, packed.M, *spatial)


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
        # conv -> [BN] -> [ReLU] with a 1x1 kernel is one tensor-core layer; anything else keeps the torch forward
        one = all(k == 1 for k in conv_unit.kernel_size) and all(v == 1 for v in conv_unit.stride) and \
            all(v == 0 for v in conv_unit.padding)
        self._tc_ok = bool(one and not preact and not instance_norm and
                           (activation is None or isinstance(activation, nn.ReLU)))
        self._tc_names = (name + "conv", name + "bn" if bn else None, activation is not None)

    def _pack(self):
        from .. import tc
        conv_name, bn_name, relu = self._tc_names
        bn = getattr(self, bn_name).bn if bn_name else None
        w, b = tc.fold_conv_bn(getattr(self, conv_name), bn)
        return tc.PackedLayer(w, b, relu=relu)

    def forward(self, x):
        if self._tc_ok and _on_tensor_cores(self, x) and x.dim() >= 3:
            from .. import tc
            return _tc_linear(tc.packed_for(self, self._pack), x)
        return super().forward(x)


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
        self._tc_ok = bool(not preact and (activation is None or isinstance(activation, nn.ReLU)))
        self._tc_names = (name + "fc", name + "bn" if bn else None, activation is not None)

    def _pack(self):
        from .. import tc
        fc_name, bn_name, relu = self._tc_names
        bn = getattr(self, bn_name).bn if bn_name else None
        w, b = tc.fold_conv_bn(getattr(self, fc_name), bn)
        return tc.PackedLayer(w, b, relu=relu)

    def forward(self, x):
        if self._tc_ok and _on_tensor_cores(self, x) and x.dim() == 2 and x.shape[0] > 1:
            from .. import tc
            return _tc_linear(tc.packed_for(self, self._pack), x.unsqueeze(-1)).squeeze(-1)      # (B, in) -> (B, out)
        return super().forward(x)
