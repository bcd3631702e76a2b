"""Plain-PyTorch fp32 restatement of the reference's per-proposal network and affinity scoring —
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  It composes the reference's own layer definitions
(1x1 convs + ReLU, max-pool) exactly as the reference forward does; grouping goes through explicit torch
indexing.  Run with TF32 disabled (or on CPU)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def shared_mlp(seq, x):
    """SharedMLP / Conv1d stack forward (pytorch_utils.py:6-33) through the LEAF torch modules (nn.Conv1d/2d,
    BatchNorm, ReLU, Dropout) in registration order — i.e. exactly nn.Sequential.forward of the reference classes,
    spelled out so that a container whose own forward has been replaced (the package under test routes eval-mode
    blocks to its tensor-core kernel) cannot put the code under test into its own oracle."""
    if isinstance(seq, torch.nn.Sequential):
        for child in seq:
            x = shared_mlp(child, x)
        return x
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
    xyz_feature = shared_mlp(rcnn.xyz_up_layer, xyz_input)
    rpn_feature = pts_input[..., cin:].transpose(1, 2).contiguous().unsqueeze(3)
    merged = shared_mlp(rcnn.merge_down_layer, torch.cat((xyz_feature, rpn_feature), dim=1))
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
        h = shared_mlp(sa.mlps[0], grouped)
        h = F.max_pool2d(h, kernel_size=[1, h.size(3)]).squeeze(-1)
        l_xyz, l_feat = new_xyz, h
    rcnn_cls = shared_mlp(rcnn.cls_layer, l_feat).squeeze(-1)
    rcnn_reg = shared_mlp(rcnn.reg_layer, l_feat).squeeze(-1)
    return rcnn_cls, rcnn_reg, l_feat


def affinity(link_layer, se_layer, pred_features, det_features):
    """tracker.py:81-112 verbatim in behaviour: repeat-based pair tensor, dual softmax, start/end through se_layer."""
    num_pred, num_det = pred_features.shape[0], det_features.shape[0]
    cor_feat = torch.abs(pred_features.unsqueeze(1).repeat(1, num_det, 1)
                         - det_features.unsqueeze(0).repeat(num_pred, 1, 1))
    logits = shared_mlp(link_layer, cor_feat.view(num_pred * num_det, -1, 1)).view(num_pred, num_det)
    link = (torch.softmax(logits, dim=1) + torch.softmax(logits, dim=0)) / 2
    start = torch.sigmoid(shared_mlp(se_layer, cor_feat.mean(dim=0).unsqueeze(-1))).flatten()
    end = torch.sigmoid(shared_mlp(se_layer, cor_feat.mean(dim=1).unsqueeze(-1))).flatten()
    return link, start, end, logits


# --------------------------------------------------------------------------------------------------
# RPN backbone on the CPU: reference forward (backbone.py:159-198, pointnet2_modules.py:20-63,135-164)
# composed from the C oracle's index ops (oracle/cref.py) and torch-CPU layers.  `net` is a module tree
# with the reference's sub-module names (the reference's own PointNet2MSG or this repo's mirror, on CPU).
# --------------------------------------------------------------------------------------------------
def _np(t):
    return t.detach().cpu().numpy()


def sa_msg_forward(sa, xyz, features, cref):
    """pointnet2_modules.py:20-63 for one PointnetSAModuleMSG; xyz (B,N,3) torch CPU."""
    import numpy as np
    idx = torch.from_numpy(cref.fps(_np(xyz), sa.npoint))
    new_xyz = torch.gather(xyz, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3))
    outs = []
    for grouper, mlp in zip(sa.groupers, sa.mlps):
        bidx = torch.from_numpy(cref.ball_query(grouper.radius, grouper.nsample, _np(xyz), _np(new_xyz)))
        B, m, ns = bidx.shape
        flat = bidx.long().view(B, 1, m * ns)
        gx = torch.gather(xyz.transpose(1, 2), 2, flat.expand(-1, 3, -1)).view(B, 3, m, ns)
        gx = gx - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            gf = torch.gather(features, 2, flat.expand(-1, features.shape[1], -1)).view(B, -1, m, ns)
            gx = torch.cat([gx, gf], dim=1)
        h = shared_mlp(mlp, gx)
        outs.append(F.max_pool2d(h, kernel_size=[1, h.size(3)]).squeeze(-1))
    return new_xyz, torch.cat(outs, dim=1), idx


def fp_forward(fp, unknown, known, unknow_feats, known_feats, cref):
    """pointnet2_modules.py:135-164"""
    d2, idx = cref.three_nn(_np(unknown), _np(known))
    dist = torch.sqrt(torch.from_numpy(d2))
    dist_recip = 1.0 / (dist + 1e-8)
    weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
    interp = torch.from_numpy(cref.three_interpolate(_np(known_feats), idx, _np(weight)))
    new_features = interp if unknow_feats is None else torch.cat([interp, unknow_feats], dim=1)
    return shared_mlp(fp.mlp, new_features.unsqueeze(-1)).squeeze(-1)


def ia_fusion_forward(fusion, point_features, img_features):
    """backbone.py:45-76 with plain torch ops on the module's parameters."""
    ia = fusion.IA_Layer
    batch = img_features.size(0)
    img_f = img_features.transpose(1, 2).contiguous().view(-1, ia.ic)
    pt_f = point_features.transpose(1, 2).contiguous().view(-1, ia.pc)
    att = torch.sigmoid(ia.fc3(torch.tanh(ia.fc1(img_f) + ia.fc2(pt_f)))).squeeze(1).view(batch, 1, -1)
    c, bn = ia.conv1[0], ia.conv1[1]
    img_new = F.relu(F.batch_norm(F.conv1d(img_features, c.weight, c.bias), bn.running_mean, bn.running_var,
                                  bn.weight, bn.bias, False, 0.0, bn.eps))
    fused = torch.cat([point_features, img_new * att], dim=1)
    b1 = fusion.bn1
    return F.relu(F.batch_norm(F.conv1d(fused, fusion.conv1.weight, fusion.conv1.bias), b1.running_mean,
                               b1.running_var, b1.weight, b1.bias, False, 0.0, b1.eps))


def grid_gather(feature_map, xy):
    """backbone.py:79-89"""
    return F.grid_sample(feature_map.float(), xy.unsqueeze(1), align_corners=True).squeeze(2)


def backbone_forward(net, xyz, xy, image_maps, cref):
    """backbone.py:159-198 given the image stack's outputs (maps per level, fused map)."""
    l_xyz, l_features, l_xy = [xyz], [None], [xy]
    for i, sa in enumerate(net.SA_modules):
        li_xyz, li_features, li_index = sa_msg_forward(sa, l_xyz[i], l_features[i], cref)
        li_xy = torch.gather(l_xy[i], 1, li_index.long().unsqueeze(-1).repeat(1, 1, 2))
        li_features = ia_fusion_forward(net.Fusion_Conv[i], li_features, grid_gather(image_maps[0][i], li_xy))
        l_xy.append(li_xy)
        l_xyz.append(li_xyz)
        l_features.append(li_features)
    for i in range(-1, -(len(net.FP_modules) + 1), -1):
        l_features[i - 1] = fp_forward(net.FP_modules[i], l_xyz[i - 1], l_xyz[i], l_features[i - 1], l_features[i], cref)
    l_features[0] = ia_fusion_forward(net.final_fusion_img_point, l_features[0], grid_gather(image_maps[1], xy))
    return l_xyz[0], l_features[0]


def basic_block(blk, x):
    """BasicBlock.forward (backbone.py:15-30): conv3x3 -> BatchNorm2d (eval statistics) -> ReLU -> conv3x3 with stride 2,
    spelled out with functional ops on the block's parameters so that a block whose own forward is routed to the
    tensor-core convolution cannot end up in its own oracle.  x (B, Cin, H, W) -> (B, Cout, H/2, W/2)."""
    c1, bn, c2 = blk.conv1, blk.bn1, blk.conv2
    y = F.conv2d(x, c1.weight, c1.bias, stride=c1.stride, padding=c1.padding)
    y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, training=False, eps=bn.eps)
    return F.conv2d(F.relu(y), c2.weight, c2.bias, stride=c2.stride, padding=c2.padding)


def image_stack(net, image):
    """The image branch of PointNet2MSG.forward (backbone.py:170,187-193): the four Img_Block maps and the fused
    full-resolution map relu(image_fusion_bn(image_fusion_conv(cat_i DeConv_i(img_i)))).  Returns (maps, fused)."""
    maps, x = [], image
    for blk in net.Img_Block:
        x = basic_block(blk, x)
        maps.append(x)
    de = torch.cat([F.conv_transpose2d(m, dc.weight, dc.bias, stride=dc.stride) for dc, m in zip(net.DeConv, maps)], dim=1)
    bn = net.image_fusion_bn
    fused = F.conv2d(de, net.image_fusion_conv.weight, net.image_fusion_conv.bias)
    fused = F.relu(F.batch_norm(fused, bn.running_mean, bn.running_var, bn.weight, bn.bias, training=False, eps=bn.eps))
    return maps, fused


def image_decoder_at_points(net, maps, xy):
    """backbone.py:187-194 on given Img_Block maps: the fused map sampled at the projected points (B, 32, N)."""
    de = torch.cat([F.conv_transpose2d(m, dc.weight, dc.bias, stride=dc.stride) for dc, m in zip(net.DeConv, maps)], dim=1)
    bn = net.image_fusion_bn
    fused = F.conv2d(de, net.image_fusion_conv.weight, net.image_fusion_conv.bias)
    fused = F.relu(F.batch_norm(fused, bn.running_mean, bn.running_var, bn.weight, bn.bias, training=False, eps=bn.eps))
    return grid_gather(fused, xy)
