"""Plain-PyTorch fp32 restatement of the reference's per-proposal network and affinity scoring —
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  It composes the reference's own layer definitions
(1x1 convs + ReLU, max-pool) exactly as the reference forward does; grouping goes through explicit torch
indexing.  Run with TF32 disabled (or on CPU)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def shared_mlp(seq, x):
    """SharedMLP / Conv1d stack forward (pytorch_utils.py:6-33): the modules ARE torch modules."""
    return seq(x)


def query_and_group(xyz, new_xyz, features, idx):
    """pointnet2_utils.py:241-264 with torch indexing: xyz (B,N,3), new_xyz (B,m,3), features (B,C,N), idx (B,m,ns)."""
    B, m, ns = idx.shape
    flat = idx.long().view(B, 1, m * ns)
    gx = torch.gather(xyz.transpose(1, 2), 2, flat.expand(-1, 3, -1)).view(B, 3, m, ns)
    gx = gx - new_xyz.transpose(1, 2).unsqueeze(-1)
    gf = torch.gather(features, 2, flat.expand(-1, features.shape[1], -1)).view(B, -1, m, ns)
    return torch.cat([gx, gf], dim=1)


def rcnn_forward_points(rcnn, pts_input, fps_fn, ball_fn):
    """rcnn.py:172-202.  fps_fn(xyz, npoint) -> idx, ball_fn(radius, nsample, xyz, new_xyz) -> idx supply the
    (already parity-pinned) index ops so this function tests the dense arithmetic only."""
    cin = rcnn.rcnn_input_channel
    xyz = pts_input[..., 0:3].contiguous()
    xyz_input = pts_input[..., 0:cin].transpose(1, 2).contiguous().unsqueeze(3)
    xyz_feature = rcnn.xyz_up_layer(xyz_input)
    rpn_feature = pts_input[..., cin:].transpose(1, 2).contiguous().unsqueeze(3)
    merged = rcnn.merge_down_layer(torch.cat((xyz_feature, rpn_feature), dim=1))
    l_xyz, l_feat = xyz, merged.squeeze(3)
    for sa in rcnn.SA_modules:
        if sa.npoint is not None:
            fidx = fps_fn(l_xyz, sa.npoint)
            new_xyz = torch.gather(l_xyz, 1, fidx.long().unsqueeze(-1).expand(-1, -1, 3))
            idx = ball_fn(sa.groupers[0].radius, sa.groupers[0].nsample, l_xyz, new_xyz)
            grouped = query_and_group(l_xyz, new_xyz, l_feat, idx)
        else:
            new_xyz = None
            grouped = torch.cat([l_xyz.transpose(1, 2).unsqueeze(2), l_feat.unsqueeze(2)], dim=1)
        h = sa.mlps[0](grouped)
        h = F.max_pool2d(h, kernel_size=[1, h.size(3)]).squeeze(-1)
        l_xyz, l_feat = new_xyz, h
    rcnn_cls = rcnn.cls_layer(l_feat).squeeze(-1)
    rcnn_reg = rcnn.reg_layer(l_feat).squeeze(-1)
    return rcnn_cls, rcnn_reg, l_feat


def affinity(link_layer, se_layer, pred_features, det_features):
    """tracker.py:81-112 verbatim in behaviour: repeat-based pair tensor, dual softmax, start/end through se_layer."""
    num_pred, num_det = pred_features.shape[0], det_features.shape[0]
    cor_feat = torch.abs(pred_features.unsqueeze(1).repeat(1, num_det, 1)
                         - det_features.unsqueeze(0).repeat(num_pred, 1, 1))
    logits = link_layer(cor_feat.view(num_pred * num_det, -1, 1)).view(num_pred, num_det)
    link = (torch.softmax(logits, dim=1) + torch.softmax(logits, dim=0)) / 2
    start = torch.sigmoid(se_layer(cor_feat.mean(dim=0).unsqueeze(-1))).flatten()
    end = torch.sigmoid(se_layer(cor_feat.mean(dim=1).unsqueeze(-1))).flatten()
    return link, start, end, logits
