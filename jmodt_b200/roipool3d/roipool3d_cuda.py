"""Same call surface as the GPU entry points of the reference pybind module `roipool3d_cuda`
(jmodt/ops/roipool3d/src/roipool3d.cpp:198-203).  The reference's two CPU helpers
(`pts_in_boxes3d_cpu`, `roipool3d_cpu`) are host code outside the GPU hot path and are not
provided: jmodt_b200 has no CPU path by design.
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
