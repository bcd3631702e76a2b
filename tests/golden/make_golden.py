#!/usr/bin/env python
"""Generates tests/golden/ref_cuda_ops.npz: outputs of the UNMODIFIED reference CUDA kernels
(oracle/_ref, built by oracle/build_ref.py) on seeded inputs.  Needs a GPU:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/ref_cuda_ops.npz'

then copy the file into tests/golden/.  The CPU test test_oracle_matches_golden_vectors_from_reference_cuda
checks the oracle against it, which pins the oracle to the reference's own results.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import _load_ref, clustered_cloud  # noqa: E402
from jmodt_b200 import synth  # noqa: E402
from oracle import cref  # noqa: E402


def main(out_path):
    ref = _load_ref()
    assert ref is not None, "oracle/_ref missing"
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(2024)
    g = {}
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    b, n, m = 1, 2048, 512
    xyz = np.round(clustered_cloud(rng, b, n, dup_frac=0.2) * 8) / 8      # coarse grid -> many ties
    g["xyz"] = xyz
    fps_idx = torch.empty(b, m, dtype=torch.int32, device=dev)
    temp = torch.full((b, n), 1e10, device=dev)
    ref.pointnet2_cuda.farthest_point_sampling_wrapper(b, n, m, T(xyz), temp, fps_idx)
    g["fps_idx"] = fps_idx.cpu().numpy()
    xr = clustered_cloud(rng, 2, 777, dup_frac=0.3)
    g["xyz_ragged"] = xr
    fr = torch.empty(2, 100, dtype=torch.int32, device=dev)
    ref.pointnet2_cuda.farthest_point_sampling_wrapper(2, 777, 100, T(xr), torch.full((2, 777), 1e10, device=dev), fr)
    g["fps_idx_ragged"] = fr.cpu().numpy()

    new_xyz = np.take_along_axis(xyz, g["fps_idx"].astype(np.int64)[..., None].repeat(3, -1), 1)
    g["new_xyz"] = new_xyz
    for r, ns, key in [(0.5, 32, "ball_idx_r05_n32"), (0.1, 16, "ball_idx_r01_n16")]:
        idx = torch.zeros(b, m, ns, dtype=torch.int32, device=dev)
        ref.pointnet2_cuda.ball_query_wrapper(b, n, m, r, ns, T(new_xyz), T(xyz), idx)
        g[key] = idx.cpu().numpy()

    d2 = torch.empty(b, n, 3, device=dev); ii = torch.empty(b, n, 3, dtype=torch.int32, device=dev)
    ref.pointnet2_cuda.three_nn_wrapper(b, n, m, T(xyz), T(new_xyz), d2, ii)
    g["nn_dist2"], g["nn_idx"] = d2.cpu().numpy(), ii.cpu().numpy()
    feats = rng.normal(size=(b, 16, m)).astype(np.float32)
    w = rng.uniform(size=(b, n, 3)).astype(np.float32)
    out = torch.empty(b, 16, n, device=dev)
    ref.pointnet2_cuda.three_interpolate_wrapper(b, 16, m, n, T(feats), ii, T(w), out)
    g["interp_feats"], g["interp_w"], g["interp_out"] = feats, w, out.cpu().numpy()

    f = synth.make_frame(5, n_points=4096, n_rois=24, with_image=False, empty_rois=3)
    enlarged = cref.enlarge_box3d(f["rois"], 0.2)
    pts = synth.nudge_off_box_faces(f["pts"], enlarged)[None]
    rf = rng.normal(size=(1, 4096, 6)).astype(np.float32)
    pooled = torch.zeros(1, 24, 64, 9, device=dev); empty = torch.zeros(1, 24, dtype=torch.int32, device=dev)
    ref.roipool3d_cuda.forward(T(pts), T(enlarged[None]), T(rf), pooled, empty)
    g["roi_pts"], g["roi_feats"], g["roi_boxes_enlarged"] = pts, rf, enlarged[None]
    g["roi_pooled"], g["roi_empty"] = pooled.cpu().numpy(), empty.cpu().numpy()

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_ops_gpu import _nms_case, _overlapping_pairs
    a7, b7 = _overlapping_pairs(rng, 96)
    a5, b5 = cref.boxes3d_to_bev(a7), cref.boxes3d_to_bev(b7)
    ov = torch.zeros(96, 96, device=dev); io = torch.zeros(96, 96, device=dev)
    ref.iou3d_cuda.boxes_overlap_bev_gpu(T(a5), T(b5), ov)
    ref.iou3d_cuda.boxes_iou_bev_gpu(T(a5), T(b5), io)
    g["bev_a"], g["bev_b"], g["bev_overlap"], g["bev_iou"] = a5, b5, ov.cpu().numpy(), io.cpu().numpy()

    boxes7, scores = _nms_case(rng, 1500)
    order = np.argsort(-scores, kind="stable")
    sb = cref.boxes3d_to_bev(boxes7)[order]
    keep = torch.zeros(1500, dtype=torch.int64)
    k = ref.iou3d_cuda.nms_normal_gpu(T(sb), keep, 0.8)
    g["nms_boxes_sorted"], g["nms_normal_keep"] = sb, keep[:k].numpy().copy()
    keep = torch.zeros(600, dtype=torch.int64)
    k = ref.iou3d_cuda.nms_gpu(T(sb[:600]), keep, 0.3)
    g["nms_rot_keep"] = keep[:k].numpy().copy()
    torch.cuda.synchronize()
    np.savez_compressed(out_path, **g)
    print("wrote", out_path, {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "ref_cuda_ops.npz"))
