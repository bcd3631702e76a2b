"""Region-proposal fusion head: the per-proposal RCNN network and the link / start-end heads.

Mirror of the reference `RCNN` module (jmodt/detection/modeling/rcnn.py:11-134 construction, :158-202 and
:288-289 eval forward) with the same sub-module names and state_dict keys, so a reference checkpoint's
`rcnn_net.*` tensors load key for key.  forward() is the sm_100a path only: RoI pooling + canonical
transform (one kernel), FPS / ball-query kernels, and every 1x1-conv layer on tcgen05 tensor cores
(jmodt_b200/csrc/tc_gemm.cu) with the grouping fused into the first layer's operand staging and the
max-pool fused into the last layer's epilogue.  There is no torch/cuDNN fallback.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import torch
import torch.nn as nn

from . import _lib, runtime, tc
from .pointnet2 import pointnet2_utils as pu
from .pointnet2 import pytorch_utils as pt_utils
from .pointnet2.pointnet2_modules import PointnetSAModule
from .roipool3d import roipool3d_utils


@dataclass
class HeadConfig:
    """The cfg.RCNN / cfg.REID values the head depends on (reference jmodt/config.py:100-168)."""
    use_intensity: bool = False          # RCNN.USE_INTENSITY
    use_mask: bool = True                # RCNN.USE_MASK
    use_depth: bool = True               # RCNN.USE_DEPTH
    pool_extra_width: float = 0.2        # RCNN.POOL_EXTRA_WIDTH
    num_points: int = 512                # RCNN.NUM_POINTS
    xyz_up_layer: List[int] = field(default_factory=lambda: [128, 128])
    sa_npoints: List[int] = field(default_factory=lambda: [128, 32, -1])
    sa_radius: List[float] = field(default_factory=lambda: [0.2, 0.4, 100])
    sa_nsample: List[int] = field(default_factory=lambda: [64, 64, 64])
    sa_mlps: List[List[int]] = field(default_factory=lambda: [[128, 128, 128], [128, 128, 256], [256, 256, 512]])
    cls_fc: List[int] = field(default_factory=lambda: [512, 512])
    reg_fc: List[int] = field(default_factory=lambda: [512, 512])
    link_fc: List[int] = field(default_factory=lambda: [512, 512])
    se_fc: List[int] = field(default_factory=lambda: [512, 512])
    loc_scope: float = 1.5
    loc_bin_size: float = 0.5
    num_head_bin: int = 9
    dp_ratio: float = 0.0
    use_bn: bool = False

    @property
    def reg_channel(self) -> int:       # rcnn.py:72-76 with LOC_Y_BY_BIN=False
        per_loc_bin_num = int(self.loc_scope / self.loc_bin_size) * 2
        return per_loc_bin_num * 4 + self.num_head_bin * 2 + 3 + 1


def _fc_stack(c_in: int, hidden: List[int], c_out: int, bn: bool, dp: float) -> nn.Sequential:
    """Conv1d stack with a Dropout inserted at index 1, exactly as rcnn.py:45-53 builds cls/reg/link/se."""
    layers, pre = [], c_in
    for h in hidden:
        layers.append(pt_utils.Conv1d(pre, h, bn=bn))
        pre = h
    layers.append(pt_utils.Conv1d(pre, c_out, activation=None))
    if dp >= 0:
        layers.insert(1, nn.Dropout(dp))
    return nn.Sequential(*layers)


def _pack_stack(seq) -> List[tc.PackedLayer]:
    """Fold every conv(+BN)(+ReLU) block of a SharedMLP / Conv1d stack into a PackedLayer."""
    packed = []
    for blk in seq:
        if isinstance(blk, nn.Dropout):
            assert blk.p == 0.0 or not blk.training
            continue
        bn = blk.bn.bn if hasattr(blk, "bn") else None
        w, b = tc.fold_conv_bn(blk.conv, bn)
        packed.append(tc.PackedLayer(w, b, relu=hasattr(blk, "activation")))
    return packed


def run_stack(packed: List[tc.PackedLayer], x: torch.Tensor, pool: int = 0, point_major_out: bool = False) -> torch.Tensor:
    """x (G, K, N) through the layers; the last one max-pools over `pool` columns, or (point_major_out) writes
    (G, N, M) directly — the `.transpose(1, 2).contiguous()` of the reference heads (rpn.py:76-77) for free."""
    fold = len(packed) >= 2 and packed[-1].M == 1 and not pool      # C -> 1 tail: folded into the previous layer's epilogue
    for i, layer in enumerate(packed):
        last = i == len(packed) - 1
        if fold and i == len(packed) - 2:
            x = tc.mlp_layer_dot(layer, packed[-1], x)
            break
        pm = point_major_out and last and layer.M > 1
        x = tc.mlp_layer(layer, x, pool=pool if last else 0, point_major_out=pm)
    if point_major_out and packed[-1].M == 1:
        x = x.transpose(1, 2)            # (G, 1, N) and (G, N, 1) are the same memory
    return x


class RCNN(nn.Module):
    def __init__(self, num_classes: int = 2, input_channels: int = 128, use_xyz: bool = True, mode: str = "TEST",
                 cfg: HeadConfig | None = None):
        super().__init__()
        self.cfg = cfg = cfg or HeadConfig()
        self.mode = mode
        self.SA_modules = nn.ModuleList()      # registered first, as in rcnn.py:16 (state_dict key order)
        self.rcnn_input_channel = 3 + int(cfg.use_intensity) + int(cfg.use_mask) + int(cfg.use_depth)
        self.xyz_up_layer = pt_utils.SharedMLP([self.rcnn_input_channel] + cfg.xyz_up_layer, bn=cfg.use_bn)
        c_out = cfg.xyz_up_layer[-1]
        self.merge_down_layer = pt_utils.SharedMLP([c_out * 2, c_out], bn=cfg.use_bn)

        channel_in = input_channels
        for k in range(len(cfg.sa_npoints)):
            mlps = [channel_in] + list(cfg.sa_mlps[k])
            npoint = cfg.sa_npoints[k] if cfg.sa_npoints[k] != -1 else None
            self.SA_modules.append(PointnetSAModule(npoint=npoint, radius=cfg.sa_radius[k], nsample=cfg.sa_nsample[k],
                                                    mlp=mlps, use_xyz=use_xyz, bn=cfg.use_bn))
            channel_in = mlps[-1]
        cls_channel = 1 if num_classes == 2 else num_classes
        self.cls_layer = _fc_stack(channel_in, cfg.cls_fc, cls_channel, cfg.use_bn, cfg.dp_ratio)
        self.reg_layer = _fc_stack(channel_in, cfg.reg_fc, cfg.reg_channel, cfg.use_bn, cfg.dp_ratio)
        self.link_layer = _fc_stack(channel_in, cfg.link_fc, 1, cfg.use_bn, cfg.dp_ratio)
        self.se_layer = _fc_stack(channel_in, cfg.se_fc, 1, cfg.use_bn, cfg.dp_ratio)
        self.init_weights()
        self._packed = None
        self.fuse_chain = True      # run qualifying SA layers as ONE kernel (csrc/sa_fused.cu)
        self.fuse_input = True      # RoI pooling writes the head layout and xyz_up + merge_down run as ONE kernel

    def init_weights(self):  # rcnn.py:116-134, weight_init='xavier'
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv1d)):
                nn.init.xavier_normal_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        nn.init.normal_(self.reg_layer[-1].conv.weight, mean=0, std=0.001)

    def train(self, mode: bool = True):
        self._packed = None
        return super().train(mode)

    # ---- packing ----------------------------------------------------------------------------
    def pack(self):
        """(Re)build the tensor-core weight images from the current parameters (call after loading weights)."""
        if self.training:
            raise RuntimeError("jmodt_b200.head.RCNN is inference only: call .eval() first")
        P = {
            "xyz_up": _pack_stack(self.xyz_up_layer), "merge_down": _pack_stack(self.merge_down_layer),
            "sa": [_pack_stack(sa.mlps[0]) for sa in self.SA_modules],
            "cls": _pack_stack(self.cls_layer), "reg": _pack_stack(self.reg_layer),
            "link": _pack_stack(self.link_layer), "se": _pack_stack(self.se_layer),
        }
        up = P["xyz_up"]
        if len(up) == 2 and up[0].K <= 8 and up[0]._w32 is not None:
            w8 = torch.zeros(up[0].M, 8, dtype=torch.float32, device=up[0].bias.device)
            w8[:, : up[0].K] = up[0]._w32
            P["xyz_up_w8"] = tc.PackedLayer(w8, up[0].bias[: up[0].M], up[0].relu)
        return P

    @property
    def packed(self):
        """Packed weight images of the current parameters (re-packed after load_state_dict / .to() / in-place updates)."""
        return tc.packed_for(self, self.pack)

    # ---- forward ----------------------------------------------------------------------------
    @torch.no_grad()
    def _input_fusable(self, n_extra: int, n_feat: int) -> bool:
        """The single-kernel input stage (csrc/sa_fused.cu, ROWS mode) covers the reference configuration:
        xyz_up_layer [3 + extras <= 8, 128, 128], 128 RPN channels, merge_down_layer [256, 128], SA0 on the fused path."""
        P = self.packed
        sa0, pk0 = self.SA_modules[0], P["sa"][0]
        return (self.fuse_input and self.fuse_chain and "xyz_up_w8" in P and n_feat == 128 and 3 + n_extra <= 8
                and [(l.M, l.K) for l in P["xyz_up"][1:]] == [(128, 128)] and [(l.M, l.K) for l in P["merge_down"]] == [(128, 256)]
                and all(l.relu for l in P["xyz_up"] + P["merge_down"]) and sa0.npoint is not None
                and tc.sa_fused_supported(pk0, pk0[0].K - 3, sa0.npoint, sa0.groupers[0].nsample))

    @torch.no_grad()
    def pool_rois(self, input_data):
        """ProposalTargetLayer.forward eval branch (proposal_target_layer.py:17-34, 99-115).  Returns the pooled points
        either as the reference's pts_input (G, S, 3 + extras + C) or — when the fused input stage applies — in the
        head layout (G, S, round_up(3 + extras + C, 8)) = [C channels | x, y, z, extras | zeros], which
        forward_points recognises by its row length."""
        cfg = self.cfg
        extra = [input_data["seg_mask"].unsqueeze(2)]
        if cfg.use_depth:
            extra.append((input_data["pts_depth"] / 70.0 - 0.5).unsqueeze(2))
        pts_feature = pack_point_features(extra, input_data["rpn_features"])
        if self._input_fusable(len(extra), input_data["rpn_features"].shape[2]):
            pooled, empty = roipool3d_utils.roipool3d_gpu_canonical_head(
                input_data["rpn_xyz"], pts_feature, input_data["roi_boxes3d"], cfg.pool_extra_width, len(extra),
                sampled_pt_num=cfg.num_points)
            return pooled.view(-1, pooled.shape[2], pooled.shape[3]), empty
        pooled, empty = roipool3d_utils.roipool3d_gpu_canonical(input_data["rpn_xyz"], pts_feature,
                                                                input_data["roi_boxes3d"], cfg.pool_extra_width,
                                                                sampled_pt_num=cfg.num_points)
        return pooled.view(-1, pooled.shape[2], pooled.shape[3]), empty

    @torch.no_grad()
    def forward_points(self, pts_input: torch.Tensor):
        """rcnn.py:172-202 on pts_input (G, 512, 3 + extra + C): returns rcnn_cls (G,1), rcnn_reg (G,46), rcnn_feat (G,512,1)."""
        P = self.packed
        cin = self.rcnn_input_channel
        sa_list = list(zip(self.SA_modules, P["sa"]))
        chain_ok = [self.fuse_chain and sa.npoint is not None and
                    tc.sa_fused_supported(pk, pk[0].K - 3, sa.npoint, sa.groupers[0].nsample) for sa, pk in sa_list]
        head_pitch = (cin + 128 + 7) // 8 * 8
        if pts_input.shape[-1] == head_pitch and head_pitch != cin + 128:
            # head layout from pool_rois: [128 channels | x, y, z, extras | 0...]; one kernel for the whole input stage
            xyz = pts_input[..., 128:131].contiguous()
            # point-major rows when the first set-abstraction level is the fused kernel (its first-layer GEMM reads rows:
            # tc.mlp_rows), else the reference's channel-first layout
            fused_in_pm = bool(chain_ok and chain_ok[0])
            h = tc.rcnn_input_fused(P["xyz_up_w8"], P["xyz_up"][1], P["merge_down"][0], pts_input.contiguous(),
                                    channel_first=not fused_in_pm)                    # (G, 512, 128) | (G, 128, 512)
        else:
            fused_in_pm = False
            xyz = pts_input[..., 0:3].contiguous()
            xyz_input = pts_input[..., 0:cin].transpose(1, 2).contiguous()               # (G, 5, 512)
            rpn_feature = pts_input[..., cin:].transpose(1, 2)                            # (G, 128, 512)
            c_up = P["xyz_up"][-1].M
            both = torch.empty((xyz.shape[0], c_up + rpn_feature.shape[1], xyz.shape[1]), dtype=torch.float32,
                               device=xyz.device)                                      # cat((xyz_feature, rpn_feature))
            both[:, c_up:].copy_(rpn_feature)
            h = xyz_input
            for i, layer in enumerate(P["xyz_up"]):
                h = tc.mlp_layer(layer, h, out=both[:, :c_up] if i == len(P["xyz_up"]) - 1 else None)
            md = P["merge_down"]
            h = both
            for i, layer in enumerate(md):
                h = tc.mlp_layer(layer, h)
        l_xyz, l_feat, pm = xyz, h, fused_in_pm                                           # pm: l_feat is point-major
        for k, (sa, packed) in enumerate(sa_list):
            grouper = sa.groupers[0]
            if sa.npoint is not None:
                fidx = pu.farthest_point_sample(l_xyz, sa.npoint)
                new_xyz = pu.gather_operation(l_xyz.transpose(1, 2).contiguous(), fidx).transpose(1, 2).contiguous()
                idx = pu.ball_query(grouper.radius, grouper.nsample, l_xyz, new_xyz)
                if chain_ok[k]:
                    # channel-first in, channel-first out: the layer's first-layer GEMM over the points (tc.sa_fused)
                    # reads the layout the previous stage writes, no transposes in between
                    l_feat = tc.sa_fused(packed, l_xyz, l_feat, idx, new_xyz, feats_point_major=pm)
                    l_xyz, pm = new_xyz, False
                    continue
                if pm:
                    l_feat, pm = l_feat.transpose(1, 2).contiguous(), False
                h = tc.grouped_first_layer(packed[0], l_xyz, l_feat, idx, new_xyz, grouper.nsample)
                pool = grouper.nsample
            else:                                                                         # GroupAll
                if pm:
                    l_feat, pm = l_feat.transpose(1, 2).contiguous(), False
                new_xyz = None
                h = tc.grouped_first_layer(packed[0], l_xyz, l_feat, None, None, 0)
                pool = l_xyz.shape[1]
            for i, layer in enumerate(packed[1:]):
                h = tc.mlp_layer(layer, h, pool=pool if i == len(packed) - 2 else 0)
            l_xyz, l_feat = new_xyz, h
        # heads: one column per proposal
        feat_t = l_feat.squeeze(-1).t().contiguous().unsqueeze(0)                         # (1, 512, G)
        rcnn_cls, rcnn_reg = runtime.parallel(lambda: run_stack(P["cls"], feat_t)[0].t().contiguous(),   # (G, 1)
                                              lambda: run_stack(P["reg"], feat_t)[0].t().contiguous())   # (G, 46)
        return rcnn_cls, rcnn_reg, l_feat

    @torch.no_grad()
    def forward(self, input_data):
        """Eval path of rcnn.py:158-202,288-289.  input_data: rpn_xyz (B,N,3), rpn_features (B,N,C),
        seg_mask (B,N), pts_depth (B,N), roi_boxes3d (B,M,7)."""
        pts_input, empty = self.pool_rois(input_data)
        rcnn_cls, rcnn_reg, feat = self.forward_points(pts_input)
        return {"rcnn_cls": rcnn_cls, "rcnn_reg": rcnn_reg, "rcnn_feat": feat, "pooled_empty_flag": empty}


def pack_point_features(extra, rpn_features: torch.Tensor) -> torch.Tensor:
    """torch.cat(extra + [rpn_features], dim=2) (proposal_target_layer.py:17-34).  rpn_features (B, N, C) is normally the
    permuted view of the channel-first backbone output (point_rcnn.py:47): then the concat is one tiled transpose
    (csrc/pair_corr.cu: jmb_pack_point_features) instead of a strided gather; otherwise plain torch.cat."""
    B, N, C = rpn_features.shape
    cf = rpn_features.permute(0, 2, 1)
    if not (rpn_features.is_cuda and rpn_features.dtype == torch.float32 and cf.is_contiguous() and len(extra) <= 2
            and C <= 256 and all(e.shape == (B, N, 1) and e.dtype == torch.float32 for e in extra)):
        return torch.cat(list(extra) + [rpn_features], dim=2)
    ex = [e.reshape(B, N).contiguous() for e in extra]
    out = torch.empty((B, N, len(ex) + C), dtype=torch.float32, device=rpn_features.device)
    st = _lib.stream_and_device(rpn_features)
    _lib.check(_lib.lib().jmb_pack_point_features(B, C, N, len(ex), cf.data_ptr(), ex[0].data_ptr() if ex else None,
                                                  ex[1].data_ptr() if len(ex) > 1 else None, out.data_ptr(), st),
               "pack_point_features")
    return out


def pair_corr(pt: torch.Tensor, dt: torch.Tensor, want_cor: bool = True, want_mean_p: bool = True,
              want_mean_d: bool = True):
    """|p_i - d_j| features of G frame pairs and their two means in one kernel (csrc/pair_corr.cu):
    pt (G, K, P), dt (G, K, D) channel-first -> cor (G, K, P*D), mean over i (G, K, D), mean over j (G, K, P);
    an output that is not wanted is not computed (None)."""
    G, K, P = pt.shape
    D = dt.shape[2]
    assert dt.shape[:2] == (G, K) and pt.is_contiguous() and dt.is_contiguous() and pt.dtype == dt.dtype == torch.float32
    new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=pt.device)
    cor = new(G, K, P * D) if want_cor else None
    mean_p = new(G, K, D) if want_mean_p else None
    mean_d = new(G, K, P) if want_mean_d else None
    st = _lib.stream_and_device(pt)
    _lib.check(_lib.lib().jmb_pair_corr(G, K, P, D, pt.data_ptr(), dt.data_ptr(), _lib.ptr(cor), _lib.ptr(mean_p),
                                        _lib.ptr(mean_d), st), "pair_corr")
    return cor, mean_p, mean_d


def dual_softmax(logits: torch.Tensor) -> torch.Tensor:
    """(softmax over successors + softmax over predecessors) / 2 of link logits (..., P, D) — tracker.py:87-89.  One
    definition for the single-GPU and the sharded path, so that equal logits give bit-equal link scores."""
    col = torch.softmax(logits.transpose(-1, -2).contiguous(), dim=-1).transpose(-1, -2)   # softmax over predecessors
    return (torch.softmax(logits, dim=-1) + col) / 2


def _stacks(link_model, se_model):
    """Packed layer stacks of the link / start-end heads (any pt_utils.Conv1d nn.Sequential, e.g. the reference's
    `rcnn_net.link_layer` / `se_layer` handed to the tracker at tools/eval.py:333-336); cached on the modules."""
    return (tc.packed_for(link_model, lambda: _pack_stack(link_model)),
            tc.packed_for(se_model, lambda: _pack_stack(se_model)))


@torch.no_grad()
def affinity_scores_batched(link_model, se_model, pred_features: torch.Tensor, det_features: torch.Tensor):
    """Link and start / end scores of G frame pairs at once (reference tracker.py:81-112 per pair; the training
    twin is rcnn.py:239-258): pred_features (G, P, 512), det_features (G, D, 512) ->
    link (G, P, D) = (softmax over successors + softmax over predecessors) / 2, start (G, D), end (G, P) (after the
    sigmoid), raw link logits (G, P, D).  One launch per layer for all pairs."""
    G, P, _ = pred_features.shape
    D = det_features.shape[1]
    link_stack, se_stack = _stacks(link_model, se_model)
    pt = pred_features.transpose(1, 2).contiguous()                                       # (G, 512, P)
    dt = det_features.transpose(1, 2).contiguous()                                        # (G, 512, D)
    cor, mean_p, mean_d = pair_corr(pt, dt)          # (G, 512, P*D), mean over predecessors (G, 512, D), over successors (G, 512, P)

    def link_branch():
        logits = run_stack(link_stack, cor).view(G, P, D)
        return dual_softmax(logits), logits

    (link, logits), start, end = runtime.parallel(
        link_branch,
        lambda: torch.sigmoid(run_stack(se_stack, mean_p)).view(G, D),
        lambda: torch.sigmoid(run_stack(se_stack, mean_d)).view(G, P))
    return link, start, end, logits


def affinity_batched(rcnn: RCNN, pred_features: torch.Tensor, det_features: torch.Tensor):
    """`affinity_scores_batched` with the heads of an RCNN module."""
    return affinity_scores_batched(rcnn.link_layer, rcnn.se_layer, pred_features, det_features)


def affinity(rcnn: RCNN, pred_features: torch.Tensor, det_features: torch.Tensor):
    """One frame pair: pred_features (P, 512), det_features (D, 512) -> link (P, D), start (D,), end (P,), logits (P, D)."""
    link, start, end, logits = affinity_scores_batched(rcnn.link_layer, rcnn.se_layer, pred_features.unsqueeze(0),
                                                       det_features.unsqueeze(0))
    return link[0], start[0], end[0], logits[0]
