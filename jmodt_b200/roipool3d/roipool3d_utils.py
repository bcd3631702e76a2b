"""Python API of the reference's `jmodt/ops/roipool3d/roipool3d_utils.py` (GPU entry point),
backed by the fused sm_100a kernel in jmodt_b200/csrc/roipool3d.cu."""
from __future__ import annotations

import torch

from . import roipool3d_cuda
from .. import box_utils


def roipool3d_gpu(pts, pts_feature, boxes3d, pool_extra_width, sampled_pt_num=512):
    """reference roipool3d_utils.py:8-29
    :param pts: (B, N, 3)
    :param pts_feature: (B, N, C)
    :param boxes3d: (B, M, 7)
    :return: pooled_features (B, M, sampled_pt_num, 3 + C), pooled_empty_flag (B, M) int32
    """
    batch_size, boxes_num, feature_len = pts.shape[0], boxes3d.shape[1], pts_feature.shape[2]
    pooled_boxes3d = box_utils.enlarge_box3d(boxes3d.view(-1, 7), pool_extra_width).view(batch_size, -1, 7)
    # the kernel writes every element (zeros for empty boxes): no .zero_() pass
    pooled_features = torch.empty((batch_size, boxes_num, sampled_pt_num, 3 + feature_len),
                                  dtype=torch.float32, device=pts.device)
    pooled_empty_flag = torch.empty((batch_size, boxes_num), dtype=torch.int32, device=pts.device)
    roipool3d_cuda.forward(pts.contiguous(), pooled_boxes3d.contiguous(), pts_feature.contiguous(),
                           pooled_features, pooled_empty_flag)
    return pooled_features, pooled_empty_flag


def roipool3d_gpu_canonical(pts, pts_feature, boxes3d, pool_extra_width, sampled_pt_num=512):
    """The whole eval branch of ProposalTargetLayer.forward (proposal_target_layer.py:99-115) in one
    kernel: enlarge, pool, centre on the roi, rotate by ry.  Returns (pts_input-shaped pooled tensor
    (B, M, S, 3 + C), pooled_empty_flag)."""
    batch_size, boxes_num, feature_len = pts.shape[0], boxes3d.shape[1], pts_feature.shape[2]
    pooled_features = torch.empty((batch_size, boxes_num, sampled_pt_num, 3 + feature_len),
                                  dtype=torch.float32, device=pts.device)
    pooled_empty_flag = torch.empty((batch_size, boxes_num), dtype=torch.int32, device=pts.device)
    roipool3d_cuda.forward_canonical(pts.contiguous(), boxes3d.contiguous(), pts_feature.contiguous(),
                                     pool_extra_width, pooled_features, pooled_empty_flag)
    return pooled_features, pooled_empty_flag


def roipool3d_gpu_canonical_head(pts, pts_feature, boxes3d, pool_extra_width, lead, sampled_pt_num=512):
    """roipool3d_gpu_canonical in the head layout: returns (B, M, S, P) with P = round_up(3 + C, 8) and rows
    [pts_feature[lead:] | canonical x, y, z | pts_feature[:lead] | zeros], plus pooled_empty_flag."""
    batch_size, boxes_num, feature_len = pts.shape[0], boxes3d.shape[1], pts_feature.shape[2]
    pitch = (3 + feature_len + 7) // 8 * 8
    pooled_features = torch.empty((batch_size, boxes_num, sampled_pt_num, pitch), dtype=torch.float32,
                                  device=pts.device)
    pooled_empty_flag = torch.empty((batch_size, boxes_num), dtype=torch.int32, device=pts.device)
    roipool3d_cuda.forward_canonical_head(pts.contiguous(), boxes3d.contiguous(), pts_feature.contiguous(),
                                          pool_extra_width, lead, pooled_features, pooled_empty_flag)
    return pooled_features, pooled_empty_flag
