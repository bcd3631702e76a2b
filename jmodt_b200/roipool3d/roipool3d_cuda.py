"""Same call surface as the GPU entry points of the reference pybind module `roipool3d_cuda`
(jmodt/ops/roipool3d/src/roipool3d.cpp:198-203), including its two HOST helpers `pts_in_boxes3d_cpu` /
`roipool3d_cpu` (roipool3d.cpp:97-195), which the reference's dataset code calls on CPU tensors.  Those two run on the
host by definition; no device op of this package ever routes through them (there is no CPU fallback).
"""
from __future__ import annotations

from .. import _lib


def _chk(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise _lib.JmodtB200Error("tensor must be a CUDAtensor")  # roipool3d.cpp:5
        if not t.is_contiguous():
            raise _lib.JmodtB200Error("tensor must be contiguous")  # roipool3d.cpp:6


def forward(xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag):
    """roipool3d_gpu (roipool3d.cpp:48-79): xyz (B,N,3), boxes3d (B,M,7) enlarged,
    pts_feature (B,N,C) -> pooled_features (B,M,S,3+C), pooled_empty_flag (B,M) int32."""
    _chk(xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag)
    st = _lib.stream_and_device(xyz)
    _lib.check(_lib.lib().jmb_roipool3d(xyz.size(0), xyz.size(1), boxes3d.size(1), pts_feature.size(2),
                                        pooled_features.size(2), xyz.data_ptr(), boxes3d.data_ptr(),
                                        pts_feature.data_ptr(), pooled_features.data_ptr(),
                                        pooled_empty_flag.data_ptr(), st), "roipool3d")
    return 1


forward_slow = forward  # roipool3d.cpp:18-44 computes the same result with a slower kernel


def _host(*tensors):
    for t in tensors:
        if t.is_cuda or not t.is_contiguous():
            raise _lib.JmodtB200Error("host helper: tensors must be contiguous CPU tensors")


def pts_in_boxes3d_cpu(pts_flag, pts, boxes3d):
    """roipool3d.cpp:97-125: pts_flag (M, N) int64 <- 1 where point j lies in box i; pts (N, 3), boxes3d (M, 7) float32."""
    _host(pts_flag, pts, boxes3d)
    rc = _lib.lib().jmb_pts_in_boxes3d_host(pts.size(0), boxes3d.size(0), pts.data_ptr(), boxes3d.data_ptr(), pts_flag.data_ptr())
    if rc:
        raise _lib.JmodtB200Error("pts_in_boxes3d_cpu: bad arguments")
    return 1


def roipool3d_cpu(pts, boxes3d, pts_feature, pooled_pts, pooled_features, pooled_empty_flag):
    """roipool3d.cpp:127-195 on CPU tensors: pts (N, 3), boxes3d (M, 7), pts_feature (N, C) ->
    pooled_pts (M, S, 3), pooled_features (M, S, C), pooled_empty_flag (M) int64."""
    _host(pts, boxes3d, pts_feature, pooled_pts, pooled_features, pooled_empty_flag)
    rc = _lib.lib().jmb_roipool3d_host(pts.size(0), boxes3d.size(0), pts_feature.size(1), pooled_pts.size(1), pts.data_ptr(),
                                       boxes3d.data_ptr(), pts_feature.data_ptr(), pooled_pts.data_ptr(),
                                       pooled_features.data_ptr(), pooled_empty_flag.data_ptr())
    if rc:
        raise _lib.JmodtB200Error("roipool3d_cpu: bad arguments")
    return 1


def forward_canonical(xyz, boxes3d, pts_feature, pool_extra_width, pooled_features, pooled_empty_flag):
    """Fused enlarge + pool + canonical transform (proposal_target_layer.py:99-112); boxes3d are raw rois."""
    _chk(xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag)
    st = _lib.stream_and_device(xyz)
    _lib.check(_lib.lib().jmb_roipool3d_canonical(
        xyz.size(0), xyz.size(1), boxes3d.size(1), pts_feature.size(2), pooled_features.size(2),
        float(pool_extra_width), xyz.data_ptr(), boxes3d.data_ptr(), pts_feature.data_ptr(),
        pooled_features.data_ptr(), pooled_empty_flag.data_ptr(), st), "roipool3d_canonical")
    return 1


def forward_canonical_head(xyz, boxes3d, pts_feature, pool_extra_width, lead, pooled_features, pooled_empty_flag):
    """forward_canonical writing rows as [features lead.. | x, y, z | features 0..lead-1 | 0...] with a pitch of
    round_up(3 + C, 8) floats: the layout the fused input stage of the per-proposal network reads (tc.rcnn_input_fused)."""
    _chk(xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag)
    st = _lib.stream_and_device(xyz)
    _lib.check(_lib.lib().jmb_roipool3d_canonical_head(
        xyz.size(0), xyz.size(1), boxes3d.size(1), pts_feature.size(2), pooled_features.size(2),
        float(pool_extra_width), int(lead), xyz.data_ptr(), boxes3d.data_ptr(), pts_feature.data_ptr(),
        pooled_features.data_ptr(), pooled_empty_flag.data_ptr(), st), "roipool3d_canonical_head")
    return 1
