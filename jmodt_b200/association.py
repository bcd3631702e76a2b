"""Tracker association inputs on the device: the geometric terms of the reference's link matrix.

Mirror of `boxes_dist_gpu` (jmodt/tracking/data_association.py:10-28) and of the weighted sum the MIP solver is fed
(`ortools_solve`, data_association.py:30-45): same function name and argument meaning, one kernel instead of the
(m, n, 8, 8, 3) corner tensor.  The solver itself (ortools / scipy, sequential host code) is out of scope.
"""
from __future__ import annotations

import torch

from . import _lib
from .iou3d.iou3d_utils import boxes_iou3d_gpu


def _check(boxes_a: torch.Tensor, boxes_b: torch.Tensor):
    if not (boxes_a.is_cuda and boxes_b.is_cuda):
        raise _lib.JmodtB200Error("boxes must be CUDA tensors (there is no CPU path)")
    assert boxes_a.dim() == 2 and boxes_a.shape[1] == 7 and boxes_b.dim() == 2 and boxes_b.shape[1] == 7
    return boxes_a.float().contiguous(), boxes_b.float().contiguous()


def boxes_dist_gpu(boxes_a: torch.Tensor, boxes_b: torch.Tensor) -> torch.Tensor:
    """data_association.py:10-28 — boxes (M, 7), (N, 7) [x, y, z, h, w, l, ry] -> (M, N):
    1 - |centre_a - centre_b| / max over the 64 corner pairs of |corner_a - corner_b|."""
    a, b = _check(boxes_a, boxes_b)
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    st = _lib.stream_and_device(a)
    _lib.check(_lib.lib().jmb_boxes_dist(a.shape[0], a.data_ptr(), b.shape[0], b.data_ptr(), out.data_ptr(), None, None,
                                         0.0, 0.0, 0.0, None, st), "boxes_dist")
    return out


def link_matrix(link_score: torch.Tensor, pred_boxes: torch.Tensor, det_boxes: torch.Tensor, w_app: float, w_iou: float,
                w_dis: float) -> torch.Tensor:
    """data_association.py:40-45: link_score * w_app + boxes_iou3d_gpu(pred, det) * w_iou + boxes_dist_gpu(pred, det)
    * w_dis, as two kernels (3-D IoU, then distance + weighted sum) with no intermediate corner tensor.
    link_score (P, D), pred_boxes (P, 7), det_boxes (D, 7) -> (P, D) on the device."""
    a, b = _check(pred_boxes, det_boxes)
    link = link_score.float().contiguous()
    assert tuple(link.shape) == (a.shape[0], b.shape[0])
    iou = boxes_iou3d_gpu(a, b).contiguous()
    out = torch.empty_like(link)
    st = _lib.stream_and_device(a)
    _lib.check(_lib.lib().jmb_boxes_dist(a.shape[0], a.data_ptr(), b.shape[0], b.data_ptr(), None, link.data_ptr(),
                                         iou.data_ptr(), float(w_app), float(w_iou), float(w_dis), out.data_ptr(), st),
               "boxes_dist")
    return out
