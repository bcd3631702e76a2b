"""Same call surface as the reference pybind module `iou3d_cuda`
(jmodt/ops/iou3d/src/iou3d.cpp:170-175), plus device-resident NMS variants."""
from __future__ import annotations

import torch

from .. import _lib


def _chk(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise _lib.JmodtB200Error("tensor must be a CUDA tensor")  # iou3d.cpp:7
        if not t.is_contiguous():
            raise _lib.JmodtB200Error("tensor must be contiguous")  # iou3d.cpp:8


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    _chk(boxes_a, boxes_b, ans_overlap)
    st = _lib.stream_and_device(boxes_a)
    _lib.check(_lib.lib().jmb_boxes_overlap_bev(boxes_a.size(0), boxes_a.data_ptr(), boxes_b.size(0),
                                                boxes_b.data_ptr(), ans_overlap.data_ptr(), st),
               "boxes_overlap_bev")
    return 1


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    _chk(boxes_a, boxes_b, ans_iou)
    st = _lib.stream_and_device(boxes_a)
    _lib.check(_lib.lib().jmb_boxes_iou_bev(boxes_a.size(0), boxes_a.data_ptr(), boxes_b.size(0),
                                            boxes_b.data_ptr(), ans_iou.data_ptr(), st), "boxes_iou_bev")
    return 1


def boxes_iou3d_gpu(boxes_a, boxes_b, ans_iou):
    """(N,7) x (M,7) 3-D IoU in one kernel (iou3d_utils.py:22-54)."""
    _chk(boxes_a, boxes_b, ans_iou)
    st = _lib.stream_and_device(boxes_a)
    _lib.check(_lib.lib().jmb_boxes_iou3d(boxes_a.size(0), boxes_a.data_ptr(), boxes_b.size(0),
                                          boxes_b.data_ptr(), ans_iou.data_ptr(), st), "boxes_iou3d")
    return 1


def nms_device(boxes, thresh, rotated, max_keep=0):
    """boxes (n,5) sorted by score -> (keep int64 (n,) on device, num_keep int32 (1,) on device).
    Nothing is copied to the host."""
    _chk(boxes)
    n = boxes.size(0)
    st = _lib.stream_and_device(boxes)
    L = _lib.lib()
    ws_bytes = L.jmb_nms_workspace_bytes(n)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=boxes.device)
    keep = torch.empty(max(n, 1), dtype=torch.int64, device=boxes.device)
    num = torch.empty(1, dtype=torch.int32, device=boxes.device)
    fn = L.jmb_nms if rotated else L.jmb_nms_normal
    _lib.check(fn(n, boxes.data_ptr(), float(thresh), keep.data_ptr(), num.data_ptr(), int(max_keep),
                  ws.data_ptr(), ws_bytes, st), "nms")
    return keep, num


def _nms_host_keep(boxes, keep, thresh, rotated):
    """Reference contract (iou3d.cpp:73-118): `keep` is a CPU LongTensor filled in place, the
    return value is the number kept."""
    if keep.is_cuda or not keep.is_contiguous():
        raise _lib.JmodtB200Error("keep must be a contiguous CPU LongTensor")
    k_dev, n_dev = nms_device(boxes, thresh, rotated)
    num = int(n_dev.item())
    keep[:num] = k_dev[:num].cpu()
    return num


def nms_gpu(boxes, keep, nms_overlap_thresh):
    return _nms_host_keep(boxes, keep, nms_overlap_thresh, True)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    return _nms_host_keep(boxes, keep, nms_overlap_thresh, False)
