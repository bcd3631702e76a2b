"""Python API of the reference's `jmodt/ops/iou3d/iou3d_utils.py`, backed by jmodt_b200/csrc/iou3d.cu."""
from __future__ import annotations

import torch

from . import iou3d_cuda


def boxes_iou_bev(boxes_a, boxes_b):
    """reference iou3d_utils.py:7-19 — (M,5),(N,5) -> (M,N)"""
    ans_iou = torch.empty((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    iou3d_cuda.boxes_iou_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans_iou)
    return ans_iou


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """reference iou3d_utils.py:22-54 — (N,7),(M,7) [x,y,z,h,w,l,ry] -> (N,M); BEV conversion, rotated
    overlap, height overlap and the union are fused into one kernel."""
    ans = torch.empty((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    iou3d_cuda.boxes_iou3d_gpu(boxes_a.float().contiguous(), boxes_b.float().contiguous(), ans)
    return ans


def _nms(boxes, scores, thresh, rotated):
    order = scores.sort(0, descending=True)[1]
    boxes = boxes[order].contiguous()
    keep, num = iou3d_cuda.nms_device(boxes, thresh, rotated)
    # the API returns a variable-length tensor, so the count has to reach the host once
    return order[keep[:int(num.item())]].contiguous()


def nms_gpu(boxes, scores, thresh):
    """reference iou3d_utils.py:57-71 — rotated NMS; returns kept original indices (CUDA LongTensor)."""
    return _nms(boxes, scores, thresh, True)


def nms_normal_gpu(boxes, scores, thresh):
    """reference iou3d_utils.py:74-88 — axis-aligned NMS (ry ignored)."""
    return _nms(boxes, scores, thresh, False)
