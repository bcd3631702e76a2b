"""numpy front-end of oracle/libjmodt_oracle.so (the C restatement in jmodt_oracle.c).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  Each function mirrors the argument
meaning of the reference wrapper it restates (cited in jmodt_oracle.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libjmodt_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "jmodt_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libjmodt_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_box_overlap.restype = C.c_float
        _lib.orc_iou_bev.restype = C.c_float
        _lib.orc_iou_normal.restype = C.c_float
        for n in ("orc_cuda_sinf", "orc_cuda_cosf"):
            getattr(_lib, n).restype = C.c_float
            getattr(_lib, n).argtypes = [C.c_float]
        _lib.orc_cuda_atan2f.restype = C.c_float
        _lib.orc_cuda_atan2f.argtypes = [C.c_float, C.c_float]
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, new_xyz = _f(xyz), _f(new_xyz)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = np.zeros((b, m, nsample), np.int32)
    lib().orc_ball_query(b, n, m, C.c_float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx))
    return idx


def group_points(points, idx):
    points, idx = _f(points), _i(idx)
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    out = np.empty((b, c, npoints, nsample), np.float32)
    lib().orc_group_points(b, c, n, npoints, nsample, _p(points), _p(idx), _p(out))
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, idx = _f(grad_out), _i(idx)
    b, c, npoints, nsample = grad_out.shape
    g = np.zeros((b, c, n), np.float32)
    lib().orc_group_points_grad(b, c, n, npoints, nsample, _p(grad_out), _p(idx), _p(g))
    return g


def gather_points(points, idx):
    points, idx = _f(points), _i(idx)
    b, c, n = points.shape
    m = idx.shape[1]
    out = np.empty((b, c, m), np.float32)
    lib().orc_gather_points(b, c, n, m, _p(points), _p(idx), _p(out))
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, idx = _f(grad_out), _i(idx)
    b, c, m = grad_out.shape
    g = np.zeros((b, c, n), np.float32)
    lib().orc_gather_points_grad(b, c, n, m, _p(grad_out), _p(idx), _p(g))
    return g


def opt_n_threads(n):
    return lib().orc_opt_n_threads(int(n))


def fps(xyz, npoint, return_temp=False):
    xyz = _f(xyz)
    b, n, _ = xyz.shape
    temp = np.full((b, n), 1e10, np.float32)
    idx = np.empty((b, npoint), np.int32)
    lib().orc_fps(b, n, npoint, _p(xyz), _p(temp), _p(idx))
    return (idx, temp) if return_temp else idx


def three_nn(unknown, known):
    """Returns (dist2, idx) — dist2 is the SQUARED distance the kernel writes; the reference
    Python wrapper applies sqrt afterwards (pointnet2_utils.py:98)."""
    unknown, known = _f(unknown), _f(known)
    b, n, _ = unknown.shape
    m = known.shape[1]
    d2 = np.empty((b, n, 3), np.float32)
    idx = np.empty((b, n, 3), np.int32)
    lib().orc_three_nn(b, n, m, _p(unknown), _p(known), _p(d2), _p(idx))
    return d2, idx


def three_interpolate(points, idx, weight):
    points, idx, weight = _f(points), _i(idx), _f(weight)
    b, c, m = points.shape
    n = idx.shape[1]
    out = np.empty((b, c, n), np.float32)
    lib().orc_three_interpolate(b, c, m, n, _p(points), _p(idx), _p(weight), _p(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _f(grad_out), _i(idx), _f(weight)
    b, c, n = grad_out.shape
    g = np.zeros((b, c, m), np.float32)
    lib().orc_three_interpolate_grad(b, c, n, m, _p(grad_out), _p(idx), _p(weight), _p(g))
    return g


def enlarge_box3d(boxes3d, extra_width):
    """kitti_utils.py:152-162 (fp32, in the reference's operation order)."""
    out = _f(boxes3d).copy()
    ew = np.float32(extra_width)
    out[..., 3:6] += np.float32(ew * np.float32(2))
    out[..., 1] += ew
    return out


def roipool3d(pts, pts_feature, boxes3d_enlarged, sampled_pt_num=512):
    """K10-K12 on already-enlarged boxes.  pts (B,N,3), pts_feature (B,N,C), boxes (B,M,7)."""
    pts, pts_feature, boxes = _f(pts), _f(pts_feature), _f(boxes3d_enlarged)
    b, n, _ = pts.shape
    m = boxes.shape[1]
    c = pts_feature.shape[2]
    pooled = np.zeros((b, m, sampled_pt_num, 3 + c), np.float32)
    empty = np.zeros((b, m), np.int32)
    lib().orc_roipool3d(b, n, m, c, sampled_pt_num, _p(pts), _p(boxes), _p(pts_feature),
                        _p(pooled), _p(empty))
    return pooled, empty


def pts_in_boxes3d(pts, boxes3d):
    pts, boxes = _f(pts), _f(boxes3d)
    flags = np.empty((boxes.shape[0], pts.shape[0]), np.int32)
    lib().orc_pts_in_boxes3d(pts.shape[0], boxes.shape[0], _p(pts), _p(boxes), _p(flags))
    return flags


def boxes_overlap_bev(a, b):
    a, b = _f(a), _f(b)
    out = np.empty((a.shape[0], b.shape[0]), np.float32)
    lib().orc_boxes_overlap_bev(a.shape[0], _p(a), b.shape[0], _p(b), _p(out))
    return out


def boxes_iou_bev(a, b):
    a, b = _f(a), _f(b)
    out = np.empty((a.shape[0], b.shape[0]), np.float32)
    lib().orc_boxes_iou_bev(a.shape[0], _p(a), b.shape[0], _p(b), _p(out))
    return out


def _nms(fn, boxes, scores, thresh):
    boxes = _f(boxes)
    scores = np.asarray(scores, np.float32)
    # the reference sorts with torch.sort(descending) (iou3d_utils.py:65); ties are kept in
    # the order the caller's sort produced, so callers that need tie parity pass `order`.
    order = np.argsort(-scores, kind="stable")
    sb = np.ascontiguousarray(boxes[order])
    keep = np.empty(sb.shape[0], np.int64)
    k = fn(sb.shape[0], _p(sb), C.c_float(thresh), _p(keep))
    return order[keep[:k]]


def nms(boxes, scores, thresh):
    return _nms(lib().orc_nms, boxes, scores, thresh)


def nms_normal(boxes, scores, thresh):
    return _nms(lib().orc_nms_normal, boxes, scores, thresh)


def nms_sorted(boxes_sorted, thresh, rotated):
    """NMS on boxes ALREADY sorted by score; returns kept positions (what iou3d.cpp:73-166 returns)."""
    sb = _f(boxes_sorted)
    keep = np.empty(sb.shape[0], np.int64)
    fn = lib().orc_nms if rotated else lib().orc_nms_normal
    k = fn(sb.shape[0], _p(sb), C.c_float(thresh), _p(keep))
    return keep[:k].copy()


def boxes3d_to_bev(boxes3d):
    """kitti_utils.py:136-149"""
    b = _f(boxes3d)
    out = np.empty((b.shape[0], 5), np.float32)
    half_l, half_w = b[:, 5] / np.float32(2), b[:, 4] / np.float32(2)
    out[:, 0], out[:, 1] = b[:, 0] - half_l, b[:, 2] - half_w
    out[:, 2], out[:, 3] = b[:, 0] + half_l, b[:, 2] + half_w
    out[:, 4] = b[:, 6]
    return out


def boxes_iou3d(boxes_a, boxes_b):
    """iou3d_utils.py:22-54 in fp32 numpy (same operation order)."""
    a, b = _f(boxes_a), _f(boxes_b)
    ov = boxes_overlap_bev(boxes3d_to_bev(a), boxes3d_to_bev(b))
    a_min, a_max = (a[:, 1] - a[:, 3]).reshape(-1, 1), a[:, 1].reshape(-1, 1)
    b_min, b_max = (b[:, 1] - b[:, 3]).reshape(1, -1), b[:, 1].reshape(1, -1)
    oh = np.clip(np.minimum(a_max, b_max) - np.maximum(a_min, b_min), 0, None).astype(np.float32)
    o3 = ov * oh
    va = (a[:, 3] * a[:, 4] * a[:, 5]).reshape(-1, 1)
    vb = (b[:, 3] * b[:, 4] * b[:, 5]).reshape(1, -1)
    return o3 / np.clip(va + vb - o3, np.float32(1e-7), None)
