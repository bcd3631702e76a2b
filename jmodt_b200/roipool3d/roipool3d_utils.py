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


# ---- host-side helpers of the reference module (roipool3d_utils.py:32-109): CPU tensors / numpy arrays in, as the
# reference's dataset code uses them.  They run on the host by definition and are not a fallback of the functions above.
def pts_in_boxes3d_cpu(pts, boxes3d):
    """reference roipool3d_utils.py:32-50: pts (N, 3), boxes3d (M, 7) CPU tensors -> list of M boolean masks (N,)."""
    if pts.is_cuda:
        raise NotImplementedError
    pts = pts.float().contiguous()
    boxes3d = boxes3d.float().contiguous()
    pts_flag = torch.empty((boxes3d.size(0), pts.size(0)), dtype=torch.long)
    roipool3d_cuda.pts_in_boxes3d_cpu(pts_flag, pts, boxes3d)
    return [pts_flag[k] > 0 for k in range(boxes3d.shape[0])]


def roipool_pc_cpu(pts, pts_feature, boxes3d, sampled_pt_num):
    """reference roipool3d_utils.py:53-70: pts (N, 3), pts_feature (N, C), boxes3d (M, 7) ->
    pooled_pts (M, S, 3), pooled_features (M, S, C), pooled_empty_flag (M,) int64, all on the CPU."""
    pts = pts.cpu().float().contiguous()
    pts_feature = pts_feature.cpu().float().contiguous()
    boxes3d = boxes3d.cpu().float().contiguous()
    assert pts.shape[0] == pts_feature.shape[0] and pts.shape[1] == 3, '%s %s' % (pts.shape, pts_feature.shape)
    pooled_pts = torch.zeros((boxes3d.shape[0], sampled_pt_num, 3), dtype=torch.float32)
    pooled_features = torch.zeros((boxes3d.shape[0], sampled_pt_num, pts_feature.shape[1]), dtype=torch.float32)
    pooled_empty_flag = torch.zeros(boxes3d.shape[0], dtype=torch.long)
    roipool3d_cuda.roipool3d_cpu(pts, boxes3d, pts_feature, pooled_pts, pooled_features, pooled_empty_flag)
    return pooled_pts, pooled_features, pooled_empty_flag


def roipool3d_cpu(boxes3d, pts, pts_feature, pts_extra_input, pool_extra_width, sampled_pt_num=512,
                  canonical_transform=True):
    """reference roipool3d_utils.py:73-109 (numpy in / numpy out): enlarge the boxes, pool, split the extra inputs from the
    features, and optionally move the pooled points to each box's canonical frame (kitti_utils.rotate_pc_along_y)."""
    import numpy as np
    pooled_boxes3d = box_utils.enlarge_box3d(boxes3d, pool_extra_width)
    pts_feature_all = np.concatenate((pts_extra_input, pts_feature), axis=1)
    pooled_pts, pooled_features, pooled_empty_flag = roipool_pc_cpu(
        torch.from_numpy(pts), torch.from_numpy(pts_feature_all), torch.from_numpy(pooled_boxes3d), sampled_pt_num)
    n_extra = pts_extra_input.shape[1]
    sampled_pts_input = torch.cat((pooled_pts, pooled_features[:, :, 0:n_extra]), dim=2).numpy()
    sampled_pts_feature = pooled_features[:, :, n_extra:].numpy()
    if canonical_transform:
        roi_ry = boxes3d[:, 6] % (2 * np.pi)
        sampled_pts_input[:, :, 0:3] = sampled_pts_input[:, :, 0:3] - boxes3d[:, np.newaxis, 0:3]
        for k in range(sampled_pts_input.shape[0]):     # kitti_utils.py:33-43
            c, s = np.cos(roi_ry[k]), np.sin(roi_ry[k])
            rot = np.array([[c, -s], [s, c]])
            sampled_pts_input[k][:, [0, 2]] = np.dot(sampled_pts_input[k][:, [0, 2]], rot.T)
        return sampled_pts_input, sampled_pts_feature
    return sampled_pts_input, sampled_pts_feature, pooled_empty_flag.numpy()
