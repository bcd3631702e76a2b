"""The few pure-torch box helpers the hot path needs from the reference's
`jmodt/utils/kitti_utils.py` (same names and semantics; file:line cited per function)."""
from __future__ import annotations

import numpy as np
import torch


def enlarge_box3d(boxes3d, extra_width):
    """kitti_utils.py:152-162 — boxes3d (N, 7) [x, y, z, h, w, l, ry]"""
    large = boxes3d.copy() if isinstance(boxes3d, np.ndarray) else boxes3d.clone()
    large[:, 3:6] += extra_width * 2
    large[:, 1] += extra_width
    return large


def boxes3d_to_bev_torch(boxes3d):
    """kitti_utils.py:136-149 — (N, 7) -> (N, 5) [x1, y1, x2, y2, ry]"""
    cu, cv = boxes3d[:, 0], boxes3d[:, 2]
    half_l, half_w = boxes3d[:, 5] / 2, boxes3d[:, 4] / 2
    return torch.stack((cu - half_l, cv - half_w, cu + half_l, cv + half_w, boxes3d[:, 6]), dim=1)


def rotate_pc_along_y_torch(pc, rot_angle):
    """kitti_utils.py:46-64 — pc (N, P, 3 + C), rot_angle (N); rotates x/z in place."""
    cosa = torch.cos(rot_angle).view(-1, 1, 1)
    sina = torch.sin(rot_angle).view(-1, 1, 1)
    x, z = pc[:, :, 0:1].clone(), pc[:, :, 2:3].clone()
    pc[:, :, 0:1] = x * cosa - z * sina
    pc[:, :, 2:3] = x * sina + z * cosa
    return pc
