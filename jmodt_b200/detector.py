"""RPN backbone with LI-Fusion, RPN heads, proposal layer and the PointRCNN wrapper — the callers of the
hot path, mirrored from the reference with the same sub-module
names / state_dict keys:

    PointNet2MSG, BasicBlock, IALayer, AttentionFusion, feature_gather   jmodt/detection/modeling/backbone.py
    RPN                                                                  jmodt/detection/modeling/rpn.py
    ProposalLayer, decode_bbox_target                                    jmodt/detection/layers/proposal_layer.py,
                                                                         jmodt/utils/bbox_transform.py:27-260
    PointRCNN                                                            jmodt/detection/modeling/point_rcnn.py

Inference only.  Point-cloud ops and every 1x1-conv / Linear layer on the point path run on this package's
sm_100a kernels, and so does the image decoder (backbone.py:187-196: deconvolutions + 1x1 conv + BatchNorm + ReLU +
sampling, evaluated at the sampled pixels only: csrc/image_decode.cu).  The 3x3 image convolutions
(backbone.py:15-30,170) run as implicit GEMMs on the same tcgen05 kernel as the 1x1 layers (BasicBlock._forward_tc;
`pt_utils.torch_layers()` selects cuDNN)."""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, box_utils, runtime, tc
from .head import RCNN, HeadConfig, affinity, affinity_batched
from .iou3d import iou3d_cuda
from .pointnet2 import pytorch_utils as pt_utils
from .pointnet2.pointnet2_modules import PointnetFPModule, PointnetSAModuleMSG, SAPlan


@dataclass
class RpnConfig:
    """cfg.RPN / cfg.LI_FUSION / cfg.EVAL values used at inference (reference jmodt/config.py:34-98, 200-208)."""
    input_channels: int = 0
    sa_npoints: List[int] = field(default_factory=lambda: [4096, 1024, 256, 64])
    sa_radius: List[List[float]] = field(default_factory=lambda: [[0.1, 0.5], [0.5, 1.0], [1.0, 2.0], [2.0, 4.0]])
    sa_nsample: List[List[int]] = field(default_factory=lambda: [[16, 32]] * 4)
    sa_mlps: List[List[List[int]]] = field(default_factory=lambda: [[[16, 16, 32], [32, 32, 64]],
                                                                    [[64, 64, 128], [64, 96, 128]],
                                                                    [[128, 196, 256], [128, 196, 256]],
                                                                    [[256, 256, 512], [256, 384, 512]]])
    fp_mlps: List[List[int]] = field(default_factory=lambda: [[128, 128], [256, 256], [512, 512], [512, 512]])
    cls_fc: List[int] = field(default_factory=lambda: [128])
    reg_fc: List[int] = field(default_factory=lambda: [128])
    dp_ratio: float = 0.5
    use_bn: bool = True
    loc_scope: float = 3.0
    loc_bin_size: float = 0.5
    num_head_bin: int = 12
    score_thresh: float = 0.2
    li_fusion: bool = True
    img_channels: List[int] = field(default_factory=lambda: [3, 64, 128, 256, 512])
    point_channels: List[int] = field(default_factory=lambda: [96, 256, 512, 1024])
    deconv_reduce: List[int] = field(default_factory=lambda: [16, 16, 16, 16])
    deconv_kernels: List[int] = field(default_factory=lambda: [2, 4, 8, 16])
    img_features_channel: int = 128
    nms_type: str = "normal"
    pre_nms_top_n: int = 9000
    post_nms_top_n: int = 100
    nms_thresh: float = 0.8
    mean_size: tuple = (1.52563191462, 1.62856739989, 3.88311640418)

    @property
    def reg_channel(self) -> int:       # rpn.py:32-37 with LOC_XZ_FINE
        return int(self.loc_scope / self.loc_bin_size) * 2 * 4 + self.num_head_bin * 2 + 3 + 1


def feature_gather(feature_map: torch.Tensor, xy: torch.Tensor) -> torch.Tensor:
    """backbone.py:79-89 on the sm_100a kernels: feature_map (B,C,H,W), xy (B,N,2) in [-1,1] -> (B,C,N).  A
    channels-last map (what the image convolutions emit, `image_features(dense=False)`) is read as it lies."""
    B, C, H, W = feature_map.shape
    N = xy.shape[1]
    fm, g = feature_map.float(), xy.contiguous()
    out = torch.empty((B, C, N), dtype=torch.float32, device=fm.device)
    st = _lib.stream_and_device(fm)
    if not fm.is_contiguous() and fm.is_contiguous(memory_format=torch.channels_last) and C % 2 == 0:
        _lib.check(_lib.lib().jmb_feature_gather_nhwc(B, C, H, W, N, fm.data_ptr(), g.data_ptr(), out.data_ptr(), st),
                   "feature_gather")
        return out
    fm = fm.contiguous()
    _lib.check(_lib.lib().jmb_feature_gather(B, C, H, W, N, fm.data_ptr(), g.data_ptr(), out.data_ptr(), st),
               "feature_gather")
    return out


class BasicBlock(nn.Module):
    """backbone.py:15-30.  Training / CPU: cuDNN / torch.  Inference on the device: tcgen05 implicit GEMMs."""

    def __init__(self, in_channels, out_channels, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=2 * stride, padding=1, bias=False)

    def forward(self, x):
        if pt_utils._on_tensor_cores(self, x) and x.dim() == 4:
            return self._forward_tc(x)
        return self.conv2(self.relu(self.bn1(self.conv1(x))))

    def _forward_tc(self, x):
        """Inference on the device: both 3x3 convolutions as implicit GEMMs on the tcgen05 kernel (csrc/tc_gemm.cu,
        convolution mode: fp32-grade three-term bf16 products, BatchNorm folded into conv1, ReLU in its epilogue),
        channels-last in and out.  Returns a (B, C, H/2, W/2) tensor in torch.channels_last memory format."""
        P = tc.packed_for(self, lambda: (tc.pack_conv3x3(self.conv1, self.bn1, relu=True), tc.pack_conv3x3(self.conv2)))
        xn = x.permute(0, 2, 3, 1)                       # a view when x is channels-last
        if xn.shape[3] < P[0].cpad:                      # RGB input: one zero channel makes a pixel one 16-byte piece
            xn = F.pad(xn, (0, P[0].cpad - xn.shape[3]))
        y = tc.conv3x3(P[1], tc.conv3x3(P[0], xn.contiguous()))
        return y.permute(0, 3, 1, 2)


class IALayer(nn.Module):
    """backbone.py:33-58"""

    def __init__(self, channels):
        super().__init__()
        self.ic, self.pc = channels
        rc = self.pc // 4
        self.conv1 = nn.Sequential(nn.Conv1d(self.ic, self.pc, 1), nn.BatchNorm1d(self.pc), nn.ReLU())
        self.fc1 = nn.Linear(self.ic, rc)
        self.fc2 = nn.Linear(self.pc, rc)
        self.fc3 = nn.Linear(rc, 1)
        self._packed = None

    def train(self, mode: bool = True):
        self._packed = None
        return super().train(mode)

    def pack(self):
        w, b = tc.fold_conv_bn(self.conv1[0], self.conv1[1])
        return {"conv1": tc.PackedLayer(w, b, True),
                "fc1": tc.PackedLayer(self.fc1.weight, self.fc1.bias, False),
                "fc2": tc.PackedLayer(self.fc2.weight, self.fc2.bias, False),
                "fc3": tc.PackedLayer(self.fc3.weight, self.fc3.bias, False),
                # the attention weight's three Linear layers as one fp32 kernel (csrc/feature_gather.cu:ia_attention_kernel)
                "w12": torch.cat((self.fc1.weight, self.fc2.weight), dim=1).detach().float().contiguous(),
                "b12": (self.fc1.bias + self.fc2.bias).detach().float().contiguous(),
                "w3": self.fc3.weight.detach().float().reshape(-1).contiguous(),
                "b3": float(self.fc3.bias.detach().float().item())}

    @torch.no_grad()
    def forward(self, img_feas, point_feas):
        P = tc.packed_for(self, self.pack)
        img_feas, point_feas = img_feas.contiguous(), point_feas.contiguous()
        B, ic, N = img_feas.shape
        pc = point_feas.shape[1]

        def attention():
            att = torch.empty((B, 1, N), dtype=torch.float32, device=img_feas.device)
            st = _lib.stream_and_device(img_feas)
            _lib.check(_lib.lib().jmb_ia_attention(B, ic, pc, P["w3"].numel(), N, img_feas.data_ptr(), point_feas.data_ptr(),
                                                   P["w12"].data_ptr(), P["b12"].data_ptr(), P["w3"].data_ptr(), P["b3"],
                                                   att.data_ptr(), st), "ia_attention")
            return att
        if B * N >= 32768 and P["w3"].numel() <= 64:
            # many points, few reduced channels (level 0 and the final fusion): the attention weight as ONE fp32 SIMT
            # kernel next to the image projection, on forked streams (runtime.parallel)
            att, conv = runtime.parallel(attention, lambda: tc.mlp_layer(P["conv1"], img_feas))
            return conv * att
        # few points, wide layers (levels 2, 3): three independent projections on the tensor cores
        ri, rp, conv = runtime.parallel(lambda: tc.mlp_layer(P["fc1"], img_feas),      # (B, rc, N)
                                        lambda: tc.mlp_layer(P["fc2"], point_feas),
                                        lambda: tc.mlp_layer(P["conv1"], img_feas))
        att = torch.sigmoid(tc.mlp_layer(P["fc3"], torch.tanh(ri + rp)))   # (B, 1, N)
        return conv * att


class AttentionFusion(nn.Module):
    """backbone.py:61-76"""

    def __init__(self, img_in_channels, pc_in_channels, out_channels):
        super().__init__()
        self.IA_Layer = IALayer(channels=[img_in_channels, pc_in_channels])
        self.conv1 = nn.Conv1d(pc_in_channels + pc_in_channels, out_channels, 1)
        self.bn1 = nn.BatchNorm1d(out_channels)
        self._packed = None

    def train(self, mode: bool = True):
        self._packed = None
        return super().train(mode)

    @torch.no_grad()
    def forward(self, point_features, img_features):
        packed = tc.packed_for(self, lambda: tc.PackedLayer(*tc.fold_conv_bn(self.conv1, self.bn1), True))
        img_features = self.IA_Layer(img_features, point_features)
        return tc.mlp_layer(packed, torch.cat([point_features, img_features], dim=1).contiguous())


class GeometryPlan:
    """Output of PointNet2MSG.geometry(): one SAPlan per set-abstraction level and one (idx, weight) pair per
    feature-propagation level.  `tensors()` lists them in a fixed order; `copy_from()` overwrites this plan's tensors
    with another plan's (same shapes) — the hand-over between two pipelined steps."""

    def __init__(self, sa, fp, l0_features=None):
        self.sa, self.fp = sa, fp
        # features of set-abstraction level 0 when the cloud carries coordinates only: that level's MLPs see nothing but
        # relative coordinates, so it belongs to the coordinate-only stage and is computed with it, one step ahead
        self.l0_features = l0_features

    def tensors(self):
        out = []
        for pl in self.sa:
            out.extend(pl.tensors())
        for i in sorted(self.fp):
            out.extend(self.fp[i])
        if self.l0_features is not None:
            out.append(self.l0_features)
        return out

    def copy_from(self, other: "GeometryPlan"):
        torch._foreach_copy_(self.tensors(), other.tensors())


class PointNet2MSG(nn.Module):
    """backbone.py:92-198"""

    def __init__(self, input_channels=0, use_xyz=True, cfg: RpnConfig | None = None):
        super().__init__()
        self.cfg = cfg = cfg or RpnConfig()
        self.input_channels = input_channels
        self.SA_modules = nn.ModuleList()
        channel_in = input_channels
        skip_channel_list = [input_channels]
        channel_out = 0
        for k in range(len(cfg.sa_npoints)):
            mlps = [[channel_in] + list(m) for m in cfg.sa_mlps[k]]
            channel_out = sum(m[-1] for m in mlps)
            self.SA_modules.append(PointnetSAModuleMSG(npoint=cfg.sa_npoints[k], radii=cfg.sa_radius[k],
                                                       nsamples=cfg.sa_nsample[k], mlps=mlps, use_xyz=use_xyz,
                                                       bn=cfg.use_bn))
            skip_channel_list.append(channel_out)
            channel_in = channel_out
        if cfg.li_fusion:
            self.Img_Block = nn.ModuleList()
            self.Fusion_Conv = nn.ModuleList()
            self.DeConv = nn.ModuleList()
            for i in range(len(cfg.img_channels) - 1):
                self.Img_Block.append(BasicBlock(cfg.img_channels[i], cfg.img_channels[i + 1], stride=1))
                self.Fusion_Conv.append(AttentionFusion(cfg.img_channels[i + 1], cfg.point_channels[i],
                                                        cfg.point_channels[i]))
                self.DeConv.append(nn.ConvTranspose2d(cfg.img_channels[i + 1], cfg.deconv_reduce[i],
                                                      kernel_size=cfg.deconv_kernels[i], stride=cfg.deconv_kernels[i]))
            self.image_fusion_conv = nn.Conv2d(sum(cfg.deconv_reduce), cfg.img_features_channel // 4, kernel_size=1)
            self.image_fusion_bn = nn.BatchNorm2d(cfg.img_features_channel // 4)
            self.final_fusion_img_point = AttentionFusion(cfg.img_features_channel // 4, cfg.img_features_channel,
                                                          cfg.img_features_channel)
        self.FP_modules = nn.ModuleList()
        for k in range(len(cfg.fp_mlps)):
            pre_channel = cfg.fp_mlps[k + 1][-1] if k + 1 < len(cfg.fp_mlps) else channel_out
            self.FP_modules.append(PointnetFPModule(mlp=[pre_channel + skip_channel_list[k]] + list(cfg.fp_mlps[k])))

    @staticmethod
    def _break_up_pc(pc):
        xyz = pc[..., 0:3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        return xyz, features

    @torch.no_grad()
    def image_features(self, image, dense: bool = True):
        """The image stack (backbone.py:170,187-193): the four Img_Block maps (3x3 convolutions: cuDNN) and, with
        `dense`, the de-convolved + fused full-resolution map as the reference materialises it.  dense=False returns
        `(channels-last maps, None)`: forward() then evaluates the decoder only at the sampled pixels (decode_gather)."""
        maps, x = [], image
        if not dense:
            x = x.contiguous(memory_format=torch.channels_last)
        for blk in self.Img_Block:
            x = blk(x)
            maps.append(x)
        if not dense:
            return maps, None
        de = torch.cat([dc(m) for dc, m in zip(self.DeConv, maps)], dim=1)
        fused = F.relu(self.image_fusion_bn(self.image_fusion_conv(de)))
        return maps, fused

    def _decoder_pack(self):
        """Weights of csrc/image_decode.cu: the DeConv slices of each of the 16 x 16 pixel phases in 32-channel chunks (hi / lo planes), and
        image_fusion_conv with the BatchNorm (eval) affine and the DeConv biases folded in.  Rebuilt when the modules'
        parameters change."""
        mods = list(self.DeConv) + [self.image_fusion_conv, self.image_fusion_bn]
        tok = tuple(t for m in mods for t in tc.weights_token(m))
        d = self.__dict__
        if d.get("_dec_pack") is None or d.get("_dec_token") != tok:
            assert len(self.DeConv) == 4 and self.image_fusion_conv.out_channels == 32, \
                "decode_gather is built for four decoder levels and 32 fused channels"
            ph = torch.arange(16)
            chunks, biases = [], []
            for l, dc in enumerate(self.DeConv):
                s = 2 << l
                assert dc.kernel_size == (s, s) and dc.stride == (s, s) and dc.out_channels == 16 and \
                    dc.in_channels % 64 == 0, "decode_gather needs kernel = stride = 2^(level+1), 16 outputs"
                w = dc.weight.detach().float().permute(2, 3, 1, 0)            # (s, s, 16, C)
                w = w[ph % s][:, ph % s]                                       # (16, 16, 16, C): phase (py, px)
                C = w.shape[-1]
                chunks.append(w.reshape(256, 16, C // 32, 32).permute(0, 2, 1, 3))
                biases.append(dc.bias.detach().float() if dc.bias is not None else torch.zeros(16, device=w.device))
            w1, b1 = tc.fold_conv_bn(self.image_fusion_conv, self.image_fusion_bn)
            b1 = b1 + w1 @ torch.cat(biases)
            w = torch.cat(chunks, dim=1).contiguous()                         # (256, chunks, 16, 32)
            # the kernel multiplies on bf16 tensor-core instructions with a hi + lo split of both operands (fp32-grade, as
            # the layer kernels): two planes of bf16 pairs along the input channel, packed into 32-bit words
            hi = w.to(torch.bfloat16)
            lo = (w - hi.float()).to(torch.bfloat16)
            wexp = torch.stack([hi, lo], dim=2).contiguous().view(torch.int32)      # (256, chunks, 2, 16, 16)
            d["_dec_pack"] = (wexp, w1.contiguous(), b1.contiguous())
            d["_dec_token"] = tok
        return d["_dec_pack"]

    @torch.no_grad()
    def decode_gather(self, maps, xy, image_hw=None):
        """backbone.py:187-194 without the full-resolution maps: grid_sample(relu(bn(conv1x1(cat DeConv_i(maps_i)))), xy)
        evaluated at the four bilinear taps of every point only.  maps: the four Img_Block outputs, xy (B,N,2)."""
        wexp, w1, b1 = self._decoder_pack()
        ms = [m.float().contiguous(memory_format=torch.channels_last) for m in maps]
        B, N = xy.shape[0], xy.shape[1]
        H, W = image_hw if image_hw is not None else (ms[0].shape[2] * 2, ms[0].shape[3] * 2)
        for l, m in enumerate(ms):
            assert m.shape[2] == H >> (l + 1) and m.shape[3] == W >> (l + 1), "image maps must halve level by level"
        g = xy.contiguous()
        out = torch.empty((B, 32, N), dtype=torch.float32, device=g.device)
        ws = torch.empty(int(_lib.lib().jmb_decode_workspace_bytes(B, N)), dtype=torch.uint8, device=g.device)
        st = _lib.stream_and_device(g)
        _lib.check(_lib.lib().jmb_decode_gather(B, N, H, W, g.data_ptr(), *[m.data_ptr() for m in ms],
                                                *[m.shape[1] for m in ms], wexp.data_ptr(), w1.data_ptr(), b1.data_ptr(),
                                                ws.data_ptr(), out.data_ptr(), st), "decode_gather")
        return out

    def geometry(self, xyz: torch.Tensor) -> "GeometryPlan":
        """The coordinate-only stage of the whole backbone on the CURRENT stream: FPS + centres + both ball-query
        neighbour lists of the four set-abstraction levels, three_nn indices + interpolation weights of the four
        feature-propagation levels (pointnet2_modules.py:36-50,146-152).  It depends on the cloud alone, so a caller
        that streams frames can compute it for batch k + 1 while batch k's tensor-core stages run
        (`runtime.GeometryAhead`) and hand it to forward(..., geometry=...)."""
        sa, fp, xs = [], {}, [xyz.contiguous()]
        for m in self.SA_modules:
            plan = m.plan(xs[-1])
            sa.append(plan)
            xs.append(plan.new_xyz)
        for i in range(-1, -(len(self.FP_modules) + 1), -1):
            fp[i] = self.FP_modules[i].plan(xs[i - 1], xs[i])
        l0 = None
        if self.level0_with_geometry and self.input_channels == 0 and sa[0].nbr is not None and not self.training:
            # xyz-only input: no feature stage feeds this level.  Not event-timed: inside a captured graph the external
            # event-record nodes of this (side-stream) launch would be ordered with the caller's stage events
            was, tc.profiler.enabled = tc.profiler.enabled, False
            try:
                l0 = self.SA_modules[0](xs[0], None, plan=sa[0])[1]
            finally:
                tc.profiler.enabled = was
        return GeometryPlan(sa, fp, l0)

    # Opt-in (JMB_L0_WITH_GEOMETRY=1): measured 8.15-8.22 ms per step against 8.25 — the step is bound by total SM time, and
    # the level's two launches then share the SMs with the per-proposal stage of the previous batch
    level0_with_geometry = os.environ.get("JMB_L0_WITH_GEOMETRY", "0") == "1"

    overlap_geometry = True
    # Opt-in (JMB_L0_CHUNKS=8): consume the level-0 FPS output in prefixes while the sampler is still running.  The
    # gates (`wait_indices`) poll a buffer another stream's kernel fills, which needs that kernel to be making progress
    # concurrently — true on an otherwise idle B200 (the sampler holds 32 of 148 SMs) but not something CUDA guarantees,
    # so the default is 1: level 0 waits for the sampler's completion event like every other level.  The throughput path
    # (`PointRCNN.geometry` + `forward(..., geometry=...)`, bench.py) takes the whole coordinate stage off the critical
    # path instead by computing it one step ahead.
    l0_chunks = int(os.environ.get("JMB_L0_CHUNKS", "1"))

    def _geometry(self, xyz):
        """FPS, ball query and three_nn depend on coordinates only, the MLP stacks on features only.  In eval mode the
        coordinate chain of all levels (FPS0 -> FPS1 -> ..., the neighbour lists, the interpolation weights) runs on a
        high-priority side stream, one event per level, and the feature chain on the caller's stream waits for the
        level it needs: the narrow, latency-bound FPS kernels and the three_nn searches then overlap the tensor-core
        work instead of stalling it.  Level 0 goes one step further: its FPS (4 096 dependent iterations on 32 SMs,
        the longest kernel of the step) publishes indices as it goes, and the caller's stream consumes them in
        `l0_chunks` prefixes behind `wait_indices` gates (see forward()).  Same kernels, same results."""
        if not (self.overlap_geometry and not self.training and xyz.is_cuda):
            return None, None
        from .pointnet2 import pointnet2_cuda, pointnet2_utils
        main = torch.cuda.current_stream()
        side = getattr(self, "_geo_stream", None)
        if side is None or side.device != xyz.device:
            side = self._geo_stream = torch.cuda.Stream(device=xyz.device, priority=-1)
        side.wait_stream(main)
        sa_plans, fp_plans = [], {}
        with torch.cuda.stream(side):
            xs = [xyz]
            for li, sa in enumerate(self.SA_modules):
                if li == 0 and self.l0_chunks > 1 and sa.npoint % self.l0_chunks == 0:
                    B, N, _ = xyz.shape
                    idx = torch.full((B, sa.npoint), -1, dtype=torch.int32, device=xyz.device)
                    filled = torch.cuda.Event()     # the gates must not look at the buffer before it is reset
                    filled.record(side)
                    pointnet2_cuda.farthest_point_sampling_wrapper(B, N, sa.npoint, xyz, None, idx)
                    new_xyz = pointnet2_utils.gather_operation(
                        xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
                    p = SAPlan(idx, new_xyz, None)      # neighbour lists are built per chunk on the caller's stream
                    p.filled = filled
                else:
                    p = sa.plan(xs[-1])
                ev = torch.cuda.Event()
                ev.record(side)
                for t in p.tensors():
                    t.record_stream(main)
                sa_plans.append((p, ev))
                xs.append(p.new_xyz)
            for i in range(-1, -(len(self.FP_modules) + 1), -1):
                p = self.FP_modules[i].plan(xs[i - 1], xs[i])
                ev = torch.cuda.Event()
                ev.record(side)
                for t in p:
                    t.record_stream(main)
                fp_plans[i] = (p, ev)
        return sa_plans, fp_plans

    def _sa0_chunked(self, sa, xyz, features, plan):
        """Level-0 set abstraction on prefixes of the FPS output while the sampler is still running."""
        from .pointnet2 import pointnet2_cuda, pointnet2_utils
        step = sa.npoint // self.l0_chunks
        xyz_t = xyz.transpose(1, 2).contiguous()
        outs = []
        torch.cuda.current_stream().wait_event(plan.filled)
        for k0 in range(0, sa.npoint, step):
            pointnet2_cuda.wait_indices(plan.idx, k0, k0 + step)
            idx_c = plan.idx[:, k0:k0 + step].contiguous()
            ctr = pointnet2_utils.gather_operation(xyz_t, idx_c).transpose(1, 2).contiguous()
            outs.append(sa(xyz, features, plan=sa.plan(xyz, ctr))[1])
        return torch.cat(outs, dim=2)

    @torch.no_grad()
    def forward(self, pc, image=None, xy=None, image_maps=None, geometry=None):
        """image_maps = (maps, fused) from image_features() may be passed to skip the image stack (fused None: the
        decoder is evaluated at the sampled pixels only, image_decode.cu); geometry = the
        result of geometry(xyz) computed earlier (in stream order before this call) to skip the coordinate stage."""
        xyz, features = self._break_up_pc(pc)
        l_xyz, l_features, l_xy = [xyz], [features], [xy]
        if self.cfg.li_fusion and image_maps is None:
            image_maps = self.image_features(image, dense=False)
        if geometry is not None:
            sa_plans = [(pl, None) for pl in geometry.sa]
            fp_plans = {i: (pl, None) for i, pl in geometry.fp.items()}
        else:
            sa_plans, fp_plans = self._geometry(xyz)
        main = torch.cuda.current_stream() if xyz.is_cuda else None
        decoded = None
        if self.cfg.li_fusion and image_maps[1] is None:
            # the decoder needs only the image maps and the pixel coordinates and is consumed by the last fusion layer:
            # it runs on a background stream under the chain of small set-abstraction / propagation launches
            decoded = runtime.spawn(lambda: self.decode_gather(image_maps[0], xy))
        for i, sa in enumerate(self.SA_modules):
            if i == 0 and geometry is not None and geometry.l0_features is not None and features is None:
                plan_i = geometry.sa[0]
                li_xyz, li_features, li_index = plan_i.new_xyz, geometry.l0_features, plan_i.idx
            elif sa_plans is not None:
                plan_i, ev = sa_plans[i]
                if plan_i.nbr is None:     # level 0, consumed in prefixes while FPS runs
                    li_features = self._sa0_chunked(sa, l_xyz[i], l_features[i], plan_i)
                    main.wait_event(ev)
                    li_xyz, li_index = plan_i.new_xyz, plan_i.idx
                else:
                    if ev is not None:
                        main.wait_event(ev)
                    li_xyz, li_features, li_index = sa(l_xyz[i], l_features[i], plan=plan_i)
            else:
                li_xyz, li_features, li_index = sa(l_xyz[i], l_features[i])
            if self.cfg.li_fusion:
                li_xy = torch.gather(l_xy[i], 1, li_index.long().unsqueeze(-1).repeat(1, 1, 2))
                img_gather = feature_gather(image_maps[0][i], li_xy)
                li_features = self.Fusion_Conv[i](li_features, img_gather)
                l_xy.append(li_xy)
            l_xyz.append(li_xyz)
            l_features.append(li_features)
        for i in range(-1, -(len(self.FP_modules) + 1), -1):
            plan_i = None
            if fp_plans is not None:
                plan_i, ev = fp_plans[i]
                if ev is not None:
                    main.wait_event(ev)
            l_features[i - 1] = self.FP_modules[i](l_xyz[i - 1], l_xyz[i], l_features[i - 1], l_features[i],
                                                   plan=plan_i)
        if self.cfg.li_fusion:
            fused = feature_gather(image_maps[1], xy) if decoded is None else decoded.join()
            l_features[0] = self.final_fusion_img_point(l_features[0], fused)
        return l_xyz[0], l_features[0]


class RPN(nn.Module):
    """rpn.py:13-87"""

    def __init__(self, use_xyz=True, mode="TEST", cfg: RpnConfig | None = None):
        super().__init__()
        self.cfg = cfg = cfg or RpnConfig()
        self.backbone_net = PointNet2MSG(input_channels=cfg.input_channels, use_xyz=use_xyz, cfg=cfg)

        def head(hidden, c_out):
            layers, pre = [], cfg.fp_mlps[0][-1]
            for h in hidden:
                layers.append(pt_utils.Conv1d(pre, h, bn=cfg.use_bn))
                pre = h
            layers.append(pt_utils.Conv1d(pre, c_out, activation=None))
            if cfg.dp_ratio >= 0:
                layers.insert(1, nn.Dropout(cfg.dp_ratio))
            return nn.Sequential(*layers)

        self.rpn_cls_layer = head(cfg.cls_fc, 1)
        self.rpn_reg_layer = head(cfg.reg_fc, cfg.reg_channel)
        self.proposal_layer = ProposalLayer(mode=mode, cfg=cfg)
        nn.init.constant_(self.rpn_cls_layer[2].conv.bias, -np.log((1 - 0.01) / 0.01))    # rpn.py:62-64
        nn.init.normal_(self.rpn_reg_layer[-1].conv.weight, mean=0, std=0.001)
        self._packed = None

    def train(self, mode: bool = True):
        self._packed = None
        return super().train(mode)

    @torch.no_grad()
    def forward(self, input_data, image_maps=None, geometry=None):
        from .head import _pack_stack, run_stack
        if self.training:
            raise RuntimeError("jmodt_b200.detector.RPN is inference only: call .eval() first")
        packed = tc.packed_for(self, lambda: (_pack_stack(self.rpn_cls_layer), _pack_stack(self.rpn_reg_layer)))
        xyz, feats = self.backbone_net(input_data["pts_input"], input_data.get("img"), input_data.get("pts_xy"),
                                       image_maps=image_maps, geometry=geometry)
        rpn_cls, rpn_reg = runtime.parallel(
            lambda: run_stack(packed[0], feats, point_major_out=True),      # (B, N, 1): transposed by the epilogue
            lambda: run_stack(packed[1], feats, point_major_out=True))      # (B, N, 76)
        return {"rpn_cls": rpn_cls, "rpn_reg": rpn_reg, "backbone_xyz": xyz, "backbone_features": feats}


def decode_bbox_target(roi_box3d, pred_reg, loc_scope, loc_bin_size, num_head_bin, anchor_size, get_ry_fine=False):
    """bbox_transform.py:27-260 for the configuration the detector runs with: BBOX_AVG_BY_BIN=True,
    get_xz_fine=True, get_y_by_bin=False, RY_WITH_BIN=False (config.py:193-208); get_ry_fine=False for the RPN
    (proposal_layer.py:24-32), True for the RCNN head (tools/eval.py:109-116: heading bins over [-pi/4, pi/4]).
    roi_box3d (N, 3|7), pred_reg (N, C) -> (N, 7) [x, y, z, h, w, l, ry]."""
    per_loc_bin_num = int(loc_scope / loc_bin_size) * 2
    nb = per_loc_bin_num
    pred_x_bin = F.softmax(pred_reg[:, 0:nb], 1)
    pred_z_bin = F.softmax(pred_reg[:, nb:2 * nb], 1)
    centers = (torch.arange(nb, device=pred_reg.device).float() * loc_bin_size + loc_bin_size / 2 - loc_scope)
    pred_x_abs = centers + pred_reg[:, 2 * nb:3 * nb] * loc_bin_size
    pred_z_abs = centers + pred_reg[:, 3 * nb:4 * nb] * loc_bin_size
    pos_x = (pred_x_abs * pred_x_bin).sum(dim=1)
    pos_z = (pred_z_abs * pred_z_bin).sum(dim=1)
    start = 4 * nb
    pos_y = roi_box3d[:, 1] + pred_reg[:, start]
    start += 1
    ry_bin = torch.argmax(pred_reg[:, start:start + num_head_bin], dim=1)
    ry_res_norm = torch.gather(pred_reg[:, start + num_head_bin:start + 2 * num_head_bin], 1,
                               ry_bin.unsqueeze(1)).squeeze(1)
    if get_ry_fine:          # bbox_transform.py:137-141
        angle_per_class = (np.pi / 2) / num_head_bin
        ry = (ry_bin.float() * angle_per_class + angle_per_class / 2) + ry_res_norm * (angle_per_class / 2) - np.pi / 4
    else:                    # :142-148
        angle_per_class = (2 * np.pi) / num_head_bin
        ry = (ry_bin.float() * angle_per_class + ry_res_norm * (angle_per_class / 2)) % (2 * np.pi)
        ry = torch.where(ry > np.pi, ry - 2 * np.pi, ry)   # `ry[ry > pi] -= 2*pi` without the nonzero() host sync
    size_l = start + 2 * num_head_bin
    size_res_norm = pred_reg[:, size_l:size_l + 3]
    hwl = size_res_norm * anchor_size + anchor_size
    roi_center = roi_box3d[:, 0:3]
    shift_ret = torch.cat((pos_x.view(-1, 1), pos_y.view(-1, 1), pos_z.view(-1, 1), hwl, ry.view(-1, 1)), dim=1)
    ret = shift_ret
    if roi_box3d.shape[1] == 7:
        roi_ry = roi_box3d[:, 6]
        ret = box_utils.rotate_pc_along_y_torch(shift_ret.unsqueeze(1), -roi_ry).squeeze(1)
        ret[:, 6] += roi_ry
    ret[:, 0] += roi_center[:, 0]          # `ret[:, [0, 2]] += roi_center[:, [0, 2]]` without host-built index tensors
    ret[:, 2] += roi_center[:, 2]
    return ret


@torch.no_grad()
def postprocess_detections(roi_boxes3d, rcnn_cls, rcnn_reg, rcnn_feat=None, head_cfg: HeadConfig | None = None,
                           mean_size=(1.52563191462, 1.62856739989, 3.88311640418), score_thresh: float = 0.2,
                           nms_thresh: float = 0.1):
    """Detection post-processing of the evaluation loop (tools/eval.py:96-193): decode the RCNN
    regression against its RoIs, sigmoid scores, score threshold (cfg.RCNN.SCORE_THRESH), rotated BEV NMS ordered by
    the raw scores (cfg.RCNN.NMS_THRESH) — per frame, everything on the device (the reference copies each frame's
    suppression mask to the host).  roi_boxes3d (B, M, 7), rcnn_cls (B*M, 1), rcnn_reg (B*M, C), rcnn_feat (B*M, F, 1)
    -> list of B dicts {boxes3d (k, 7), scores (k,), raw_scores (k,), feat (k, F) | None, roi_index (k,)}."""
    from .iou3d import iou3d_utils
    cfg = head_cfg or HeadConfig()
    B, M = roi_boxes3d.shape[:2]
    anchor = torch.tensor(mean_size, dtype=torch.float32, device=roi_boxes3d.device)
    boxes = decode_bbox_target(roi_boxes3d.reshape(-1, 7), rcnn_reg.reshape(B * M, -1), cfg.loc_scope, cfg.loc_bin_size,
                               cfg.num_head_bin, anchor, get_ry_fine=True).view(B, M, 7)
    raw = rcnn_cls.reshape(B, M)
    norm = torch.sigmoid(raw)
    feat = None if rcnn_feat is None else rcnn_feat.reshape(B, M, -1)
    out = []
    for k in range(B):
        sel = torch.nonzero(norm[k] > score_thresh).flatten()
        if sel.numel() == 0:
            out.append({"boxes3d": boxes.new_zeros(0, 7), "scores": raw.new_zeros(0), "raw_scores": raw.new_zeros(0),
                        "feat": None if feat is None else feat.new_zeros(0, feat.shape[2]), "roi_index": sel})
            continue
        b_sel, r_sel = boxes[k, sel], raw[k, sel]
        keep = iou3d_utils.nms_gpu(box_utils.boxes3d_to_bev_torch(b_sel).contiguous(), r_sel.contiguous(), nms_thresh).view(-1)
        out.append({"boxes3d": b_sel[keep], "scores": norm[k, sel][keep], "raw_scores": r_sel[keep],
                    "feat": None if feat is None else feat[k, sel][keep], "roi_index": sel[keep]})
    return out


class ProposalLayer(nn.Module):
    """proposal_layer.py:10-121 (distance-based proposal, axis-aligned or rotated NMS on the device)."""

    def __init__(self, mode="TEST", cfg: RpnConfig | None = None):
        super().__init__()
        self.mode, self.cfg = mode, cfg or RpnConfig()
        self.batched = True       # False: the reference's per-frame / per-bin loop (same results, host syncs)
        self._mean_size = {}

    @torch.no_grad()
    def forward(self, rpn_scores, rpn_reg, xyz):
        cfg = self.cfg
        B = xyz.shape[0]
        mean_size = self._mean_size.get(xyz.device)
        if mean_size is None:       # built once per device: a pageable H2D copy per call would also break graph capture
            mean_size = self._mean_size[xyz.device] = torch.tensor(cfg.mean_size, dtype=torch.float32, device=xyz.device)
        proposals = decode_bbox_target(xyz.view(-1, 3), rpn_reg.view(-1, rpn_reg.shape[-1]), cfg.loc_scope,
                                       cfg.loc_bin_size, cfg.num_head_bin, mean_size)
        proposals[:, 1] += proposals[:, 3] / 2
        proposals = proposals.view(B, -1, 7)
        _, sorted_idxs = torch.sort(rpn_scores, dim=1, descending=True)
        if self.batched:
            return self._forward_batched(rpn_scores.contiguous(), proposals.contiguous(), sorted_idxs.contiguous())
        ret_bbox3d = rpn_scores.new_zeros(B, cfg.post_nms_top_n, 7)
        ret_scores = rpn_scores.new_zeros(B, cfg.post_nms_top_n)
        for k in range(B):
            s, p = self.distance_based_proposal(rpn_scores[k], proposals[k], sorted_idxs[k])
            ret_bbox3d[k, :p.size(0)] = p
            ret_scores[k, :p.size(0)] = s
        return ret_bbox3d, ret_scores

    def _forward_batched(self, scores, proposals, order):
        """Whole batch in three kernels, no host synchronisation (csrc/proposal.cu)."""
        cfg = self.cfg
        B, N = scores.shape
        L = _lib.lib()
        st = _lib.stream_and_device(scores)
        ws_bytes = L.jmb_proposal_workspace_bytes(B, N, cfg.pre_nms_top_n, cfg.post_nms_top_n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=scores.device)
        ret_bbox3d = torch.empty(B, cfg.post_nms_top_n, 7, dtype=torch.float32, device=scores.device)
        ret_scores = torch.empty(B, cfg.post_nms_top_n, dtype=torch.float32, device=scores.device)
        _lib.check(L.jmb_proposal_layer(B, N, proposals.data_ptr(), scores.data_ptr(), order.data_ptr(),
                                        cfg.pre_nms_top_n, cfg.post_nms_top_n, float(cfg.nms_thresh),
                                        int(cfg.nms_type == "rotate"), ret_bbox3d.data_ptr(), ret_scores.data_ptr(),
                                        ws.data_ptr(), ws_bytes, st), "proposal_layer")
        return ret_bbox3d, ret_scores

    def distance_based_proposal(self, scores, proposals, order):
        cfg = self.cfg
        nms_range_list = [0, 40.0, 80.0]
        pre_top_n_list = [0, int(cfg.pre_nms_top_n * 0.7), cfg.pre_nms_top_n - int(cfg.pre_nms_top_n * 0.7)]
        post_top_n_list = [0, int(cfg.post_nms_top_n * 0.7), cfg.post_nms_top_n - int(cfg.post_nms_top_n * 0.7)]
        scores_ordered, proposals_ordered = scores[order], proposals[order]
        dist = proposals_ordered[:, 2]
        first_mask = (dist > nms_range_list[0]) & (dist <= nms_range_list[1])
        out_s, out_p = [], []
        for i in range(1, len(nms_range_list)):
            dist_mask = (dist > nms_range_list[i - 1]) & (dist <= nms_range_list[i])
            if dist_mask.sum() != 0:
                cur_scores = scores_ordered[dist_mask][:pre_top_n_list[i]]
                cur_proposals = proposals_ordered[dist_mask][:pre_top_n_list[i]]
            else:
                if i == 1:
                    continue
                cur_scores = scores_ordered[first_mask][pre_top_n_list[i - 1]:][:pre_top_n_list[i]]
                cur_proposals = proposals_ordered[first_mask][pre_top_n_list[i - 1]:][:pre_top_n_list[i]]
            boxes_bev = box_utils.boxes3d_to_bev_torch(cur_proposals).contiguous()
            # cur_scores are already sorted (a masked slice of a sorted list); the reference re-sorts them
            # (iou3d_utils.py:82), which is the identity for distinct scores.
            sorder = cur_scores.sort(0, descending=True)[1]
            keep, num = iou3d_cuda.nms_device(boxes_bev[sorder].contiguous(), cfg.nms_thresh,
                                              cfg.nms_type == "rotate", max_keep=post_top_n_list[i])
            keep_idx = sorder[keep[:int(num.item())]]
            out_s.append(cur_scores[keep_idx])
            out_p.append(cur_proposals[keep_idx])
        return torch.cat(out_s, dim=0), torch.cat(out_p, dim=0)


class PointRCNN(nn.Module):
    """point_rcnn.py:9-72 (eval path)."""

    def __init__(self, num_classes=2, use_xyz=True, mode="TEST", rpn_cfg: RpnConfig | None = None,
                 head_cfg: HeadConfig | None = None):
        super().__init__()
        self.mode = mode
        self.rpn = RPN(use_xyz=use_xyz, mode=mode, cfg=rpn_cfg)
        self.rcnn_net = RCNN(num_classes=num_classes, input_channels=128, use_xyz=use_xyz, mode=mode, cfg=head_cfg)

    @torch.no_grad()
    def geometry(self, pts_input: torch.Tensor) -> GeometryPlan:
        """Coordinate-only stage of a batch (see PointNet2MSG.geometry): pts_input (B, N, 3+)."""
        return self.rpn.backbone_net.geometry(pts_input[..., 0:3])

    @torch.no_grad()
    def forward(self, input_data, image_maps=None, rois=None, geometry=None):
        """input_data: pts_input (B,N,3), img (B,3,384,1280), pts_xy (B,N,2).  `rois` (B,M,7) may be injected
        to bypass the proposal layer; `geometry` = self.geometry(pts_input) computed earlier."""
        output = {}
        rpn_output = self.rpn(input_data, image_maps=image_maps, geometry=geometry)
        output.update(rpn_output)
        backbone_xyz, backbone_features = rpn_output["backbone_xyz"], rpn_output["backbone_features"]
        rpn_scores_raw = rpn_output["rpn_cls"][:, :, 0]
        seg_mask = (torch.sigmoid(rpn_scores_raw) > self.rpn.cfg.score_thresh).float()
        pts_depth = torch.norm(backbone_xyz, p=2, dim=2)
        if rois is None:
            rois, roi_scores_raw = self.rpn.proposal_layer(rpn_scores_raw, rpn_output["rpn_reg"], backbone_xyz)
            output["roi_scores_raw"] = roi_scores_raw
        output["rois"] = rois
        output["seg_result"] = seg_mask
        rcnn_input = {"rpn_xyz": backbone_xyz, "rpn_features": backbone_features.permute(0, 2, 1),
                      "seg_mask": seg_mask, "roi_boxes3d": rois, "pts_depth": pts_depth}
        output.update(self.rcnn_net(rcnn_input))
        return output

    @torch.no_grad()
    def pair_affinity(self, rcnn_feat, rois_per_frame):
        """Link / start-end scores of the disjoint frame pairs (0, 1), (2, 3), ... of a batch (an odd last frame is
        left out): rcnn_feat (B*M, 512, 1) -> one (link (M, M), start (M,), end (M,), logits (M, M)) per pair."""
        f = rcnn_feat.view(-1, rois_per_frame, rcnn_feat.shape[1])
        npairs = f.shape[0] // 2
        link, start, end, logits = affinity_batched(self.rcnn_net, f[0:2 * npairs:2], f[1:2 * npairs:2])
        return [(link[i], start[i], end[i], logits[i]) for i in range(npairs)]
