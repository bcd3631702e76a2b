#!/usr/bin/env python
"""Kernel-for-kernel: this package's operators vs the UNMODIFIED reference CUDA extensions (oracle/_ref/*.so, built for
sm_100) on the same B200, same tensors, at the shapes one 8-frame step of BASELINE config 2/3 uses.

    gpurun -- python profiles/ref_cuda_timing.py gpurun_out/ref_cuda_timing.json

Timing: CUDA events around each call (both libraries' kernels are enqueued on / synchronise with the legacy default
stream, so torch's current stream = default stream is used for both), 5 warm-ups, median of 30; L2 is NOT flushed (same
for both sides).  The reference's NMS / roipool wrappers include their cudaMalloc / blocking D2H / host sweep — that is
what a caller of the reference pays (iou3d.cpp:73-166, roipool3d_kernel.cu:209-237).
"""
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(fn, warm=5, reps=30):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return statistics.median(ms) * 1e3       # microseconds


def main(out_path):
    from conftest import _load_ref
    from jmodt_b200 import box_utils, synth
    from jmodt_b200.iou3d import iou3d_cuda, iou3d_utils
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    from jmodt_b200.roipool3d import roipool3d_utils as ru
    ref = _load_ref()
    assert ref is not None, "oracle/_ref not available"
    dev = torch.device("cuda:0")
    B = 8
    batch = synth.make_batch(0, B, with_image=False)
    xyz = torch.from_numpy(batch["pts"]).to(dev)
    rois = torch.from_numpy(batch["rois"]).to(dev)
    rows = []

    def add(name, shape, ours, theirs, equal=None):
        a, b = timed(ours), timed(theirs)
        rows.append({"op": name, "shape": shape, "jmodt_b200_us": round(a, 1), "reference_cuda_us": round(b, 1),
                     "speedup": round(b / a, 2), "outputs_equal": equal})
        print(f"{name:28s} {shape:44s} ours {a:10.1f} us   reference {b:10.1f} us   x{b / a:6.2f}  equal={equal}", flush=True)

    # ---- FPS
    lv = [xyz]
    for n_in, m in ((16384, 4096), (4096, 1024), (1024, 256), (256, 64)):
        cur = lv[-1]
        idx_r = torch.zeros(B, m, dtype=torch.int32, device=dev)
        temp = torch.full((B, n_in), 1e10, device=dev)

        def theirs(cur=cur, m=m, idx_r=idx_r, temp=temp, n_in=n_in):
            temp.fill_(1e10)
            ref.pointnet2_cuda.farthest_point_sampling_wrapper(B, n_in, m, cur, temp, idx_r)
        idx = pu.farthest_point_sample(cur, m)
        theirs()
        add("furthest_point_sample", f"B={B} N={n_in} -> {m}", lambda cur=cur, m=m: pu.farthest_point_sample(cur, m), theirs,
            bool(torch.equal(idx, idx_r)))
        lv.append(pu.gather_operation(cur.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous())
    # ---- ball query (RPN levels 0, 1; RCNN SA0)
    for li, (r, ns) in ((0, (0.1, 16)), (0, (0.5, 32)), (1, (0.5, 16)), (1, (1.0, 32))):
        src, ctr = lv[li], lv[li + 1]
        n, m = src.shape[1], ctr.shape[1]
        out_r = torch.zeros(B, m, ns, dtype=torch.int32, device=dev)
        theirs = lambda: ref.pointnet2_cuda.ball_query_wrapper(B, n, m, r, ns, ctr, src, out_r)
        got = pu.ball_query(r, ns, src, ctr)
        out_r.zero_(); theirs()
        add("ball_query", f"B={B} N={n} m={m} r={r} ns={ns}", lambda: pu.ball_query(r, ns, src, ctr), theirs,
            bool(torch.equal(got, out_r)))
    # both radii of RPN level 0 in one call (what PointnetSAModuleMSG issues): cell list here, two scans in the reference
    src, ctr = lv[0], lv[1]
    ra, na, rb, nb = 0.1, 16, 0.5, 32
    oa = torch.zeros(B, 4096, na, dtype=torch.int32, device=dev)
    ob = torch.zeros(B, 4096, nb, dtype=torch.int32, device=dev)

    def both_ref():
        ref.pointnet2_cuda.ball_query_wrapper(B, 16384, 4096, ra, na, ctr, src, oa)
        ref.pointnet2_cuda.ball_query_wrapper(B, 16384, 4096, rb, nb, ctr, src, ob)
    ga, gb = pu.ball_query_msg2(ra, na, rb, nb, src, ctr)
    both_ref()
    add("ball_query x2 radii", f"B={B} N=16384 m=4096 r=0.1/0.5 ns=16/32 (cell list)",
        lambda: pu.ball_query_msg2(ra, na, rb, nb, src, ctr), both_ref, bool(torch.equal(ga, oa) and torch.equal(gb, ob)))
    G = 1024
    pooled_xyz = torch.rand(G, 512, 3, device=dev)
    fidx = pu.farthest_point_sample(pooled_xyz, 128)
    ctr = pu.gather_operation(pooled_xyz.transpose(1, 2).contiguous(), fidx).transpose(1, 2).contiguous()
    idx_r = torch.zeros(G, 128, dtype=torch.int32, device=dev)
    temp = torch.full((G, 512), 1e10, device=dev)

    def fps_ref():
        temp.fill_(1e10)
        ref.pointnet2_cuda.farthest_point_sampling_wrapper(G, 512, 128, pooled_xyz, temp, idx_r)
    fps_ref()
    add("furthest_point_sample", "B=1024 N=512 -> 128 (RCNN SA0)", lambda: pu.farthest_point_sample(pooled_xyz, 128), fps_ref,
        bool(torch.equal(fidx, idx_r)))
    bq_r = torch.zeros(G, 128, 64, dtype=torch.int32, device=dev)
    bq = pu.ball_query(0.2, 64, pooled_xyz, ctr)
    theirs = lambda: ref.pointnet2_cuda.ball_query_wrapper(G, 512, 128, 0.2, 64, ctr, pooled_xyz, bq_r)
    theirs()
    add("ball_query", "B=1024 N=512 m=128 r=0.2 ns=64 (RCNN SA0)", lambda: pu.ball_query(0.2, 64, pooled_xyz, ctr), theirs,
        bool(torch.equal(bq, bq_r)))
    # ---- group_points: the grouped tensor of RCNN SA0 the fused kernel never materialises (1024 x 128 x 128 x 64 floats = 4.3 GB)
    feats = torch.randn(G, 128, 512, device=dev)
    out_r = torch.empty(G, 128, 128, 64, device=dev)
    out_o = torch.empty_like(out_r)
    from jmodt_b200.pointnet2 import pointnet2_cuda as ours_cuda
    add("group_points", "B=1024 C=128 N=512 m=128 ns=64 (4.3 GB out)",
        lambda: ours_cuda.group_points_wrapper(G, 128, 512, 128, 64, feats, bq, out_o),
        lambda: ref.pointnet2_cuda.group_points_wrapper(G, 128, 512, 128, 64, feats, bq, out_r), bool(torch.equal(out_o, out_r)))
    del out_r, out_o, feats
    # ---- three_nn / three_interpolate (FP level 0: 16384 unknown, 4096 known)
    unknown, known = lv[0], lv[1]
    d2r = torch.empty(B, 16384, 3, device=dev)
    ir = torch.empty(B, 16384, 3, dtype=torch.int32, device=dev)
    dist, idx3 = pu.three_nn(unknown, known)
    theirs = lambda: ref.pointnet2_cuda.three_nn_wrapper(B, 16384, 4096, unknown, known, d2r, ir)
    theirs()
    add("three_nn", f"B={B} n=16384 m=4096", lambda: pu.three_nn(unknown, known), theirs, bool(torch.equal(idx3, ir)))
    w = torch.rand(B, 16384, 3, device=dev)
    w = w / w.sum(2, keepdim=True)
    f = torch.randn(B, 256, 4096, device=dev)
    outr = torch.empty(B, 256, 16384, device=dev)
    got = pu.three_interpolate(f, idx3, w)
    theirs = lambda: ref.pointnet2_cuda.three_interpolate_wrapper(B, 256, 4096, 16384, f, ir, w, outr)
    theirs()
    add("three_interpolate", f"B={B} C=256 m=4096 n=16384", lambda: pu.three_interpolate(f, idx3, w), theirs,
        bool(torch.equal(got, outr)))
    # ---- roipool3d (8 frames x 128 boxes x 512 points x 133 floats)
    pf = torch.randn(B, 16384, 130, device=dev)
    enlarged = box_utils.enlarge_box3d(rois.view(-1, 7), 0.2).view(B, -1, 7).contiguous()
    pr = torch.zeros(B, 128, 512, 133, device=dev)
    er = torch.zeros(B, 128, dtype=torch.int32, device=dev)
    got, emp = ru.roipool3d_gpu(xyz, pf, rois, 0.2, sampled_pt_num=512)

    def theirs():
        ref.roipool3d_cuda.forward(xyz, enlarged, pf, pr, er)
    theirs()
    add("roipool3d", f"B={B} N=16384 M=128 C=130 S=512", lambda: ru.roipool3d_gpu(xyz, pf, rois, 0.2, sampled_pt_num=512), theirs,
        bool(torch.equal(emp, er)))
    # ---- iou3d / NMS
    a5 = box_utils.boxes3d_to_bev_torch(rois[0]).contiguous()
    b5 = box_utils.boxes3d_to_bev_torch(rois[1]).contiguous()
    ans_r = torch.zeros(128, 128, device=dev)
    got = iou3d_utils.boxes_iou_bev(a5, b5)
    theirs = lambda: ref.iou3d_cuda.boxes_iou_bev_gpu(a5, b5, ans_r)
    theirs()
    add("boxes_iou_bev", "128 x 128", lambda: iou3d_utils.boxes_iou_bev(a5, b5), theirs, bool(torch.equal(got, ans_r)))
    rng = np.random.default_rng(0)
    for n, rotated in ((6300, False), (2700, False), (6300, True), (128, True)):
        base = batch["rois"][0][rng.integers(0, 128, n)] + rng.normal(0, 0.3, (n, 7)).astype(np.float32) * np.array([1, 1, 1, .1, .1, .1, 1], np.float32)
        bev = box_utils.boxes3d_to_bev_torch(torch.from_numpy(base.astype(np.float32)).to(dev)).contiguous()
        thresh = 0.1 if n == 128 else 0.8
        keep_cpu = torch.zeros(n, dtype=torch.long)
        fn = ref.iou3d_cuda.nms_gpu if rotated else ref.iou3d_cuda.nms_normal_gpu
        num_r = fn(bev, keep_cpu, thresh)
        keep, num = iou3d_cuda.nms_device(bev, thresh, rotated)
        eq = int(num.item()) == num_r and bool(torch.equal(keep[:num_r].cpu(), keep_cpu[:num_r]))
        add("nms_rotated" if rotated else "nms_normal", f"n={n} thresh={thresh} (kept {num_r})",
            lambda: iou3d_cuda.nms_device(bev, thresh, rotated), lambda: fn(bev, keep_cpu, thresh), eq)
    with open(out_path, "w") as fh:
        json.dump({"gpu": torch.cuda.get_device_name(0), "frames": B, "rows": rows,
                   "how": "CUDA events on the default stream, 5 warm-ups, median of 30, no L2 flush; reference = unmodified "
                          "jmodt/ops/*/src built for sm_100 (oracle/build_ref.py)"}, fh, indent=1)
    print("wrote", out_path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_cuda_timing.json")
