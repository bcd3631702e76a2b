"""CPU restatement of the tracker's geometric association terms — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows the reference line by line with plain torch ops on CPU tensors (the reference allocates with
torch.cuda.FloatTensor, kitti_utils.py:115-116, so it cannot run here unmodified):
    boxes3d_to_corners3d      jmodt/utils/kitti_utils.py:107-133
    boxes_dist                jmodt/tracking/data_association.py:10-28
    link_matrix               jmodt/tracking/data_association.py:40-45 (with a caller-supplied IoU matrix)
Parity is tolerance-based (fp32 geometry); pinned by construction against the reference's formulas — the reference
holds no golden vectors for this path.
"""
from __future__ import annotations

import torch


def boxes3d_to_corners3d(boxes3d: torch.Tensor) -> torch.Tensor:
    n = boxes3d.shape[0]
    h, w, l, ry = boxes3d[:, 3:4], boxes3d[:, 4:5], boxes3d[:, 5:6], boxes3d[:, 6:7]
    centers = boxes3d[:, 0:3]
    zeros, ones = torch.zeros(n, 1), torch.ones(n, 1)
    x_corners = torch.cat([l / 2., l / 2., -l / 2., -l / 2., l / 2., l / 2., -l / 2., -l / 2.], dim=1)
    y_corners = torch.cat([zeros, zeros, zeros, zeros, -h, -h, -h, -h], dim=1)
    z_corners = torch.cat([w / 2., -w / 2., -w / 2., w / 2., w / 2., -w / 2., -w / 2., w / 2.], dim=1)
    corners = torch.cat((x_corners.unsqueeze(1), y_corners.unsqueeze(1), z_corners.unsqueeze(1)), dim=1)     # (N, 3, 8)
    cosa, sina = torch.cos(ry), torch.sin(ry)
    R = torch.cat((torch.cat([cosa, zeros, sina], dim=1).unsqueeze(1), torch.cat([zeros, ones, zeros], dim=1).unsqueeze(1),
                   torch.cat([-sina, zeros, cosa], dim=1).unsqueeze(1)), dim=1)                                # (N, 3, 3)
    rotated = torch.matmul(R, corners) + centers.unsqueeze(2).expand(-1, -1, 8)
    return rotated.permute(0, 2, 1)


def boxes_dist(boxes_a: torch.Tensor, boxes_b: torch.Tensor) -> torch.Tensor:
    m, n = len(boxes_a), len(boxes_b)
    ca, cb = boxes3d_to_corners3d(boxes_a), boxes3d_to_corners3d(boxes_b)
    center = torch.linalg.norm(boxes_a[:, :3].unsqueeze(1).repeat(1, n, 1) - boxes_b[:, :3].unsqueeze(0).repeat(m, 1, 1),
                               ord=2, dim=-1)
    corner, _ = torch.max(torch.linalg.norm(ca.view(m, 1, 8, 1, 3).repeat(1, n, 1, 8, 1)
                                            - cb.view(1, n, 1, 8, 3).repeat(m, 1, 8, 1, 1), ord=2, dim=-1).view(m, n, 64), dim=-1)
    return 1. - center / corner


def link_matrix(link_score, iou_matrix, pred_boxes, det_boxes, w_app, w_iou, w_dis):
    return link_score * w_app + iou_matrix * w_iou + boxes_dist(pred_boxes, det_boxes) * w_dis
