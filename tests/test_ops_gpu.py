"""Parity of the sm_100a kernels (called through the reference-shaped Python API, which goes
through the C ABI) against the CPU oracle and — when oracle/_ref is present — against the
reference's own CUDA kernels on the same device.  Integer / index results must be bit-exact;
copied floats must be bit-exact; tolerances for derived floats are stated at the assert."""
import numpy as np
import pytest
import torch

from conftest import clustered_cloud

pytestmark = pytest.mark.gpu


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


# ---------------------------------------------------------------- ball query + grouping
@pytest.mark.parametrize("b,n,m,radius,nsample", [
    (1, 4096, 1024, 0.5, 32),      # BASELINE config 1
    (1, 4096, 4096, 0.5, 32),      # config 1, all-points variant
    (2, 16384, 4096, 0.1, 16),     # RPN level 0, sparse radius (rows rarely fill)
    (2, 16384, 4096, 0.5, 32),
    (3, 1000, 37, 1.0, 5),         # ragged sizes
    (64, 512, 128, 0.2, 64),       # RCNN SA0 shape (B*M batches)
    (1, 33, 7, 100.0, 64),         # nsample > n
    (1, 5, 3, 1e-6, 4),            # nothing but the centre itself / empty rows
])
def test_ball_query_bit_exact(cuda, cref, b, n, m, radius, nsample):
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    rng = np.random.default_rng(n * 7 + m)
    xyz = clustered_cloud(rng, b, n)
    sel = np.stack([rng.choice(n, m, replace=False) for _ in range(b)])
    new_xyz = np.take_along_axis(xyz, sel[..., None].repeat(3, -1), 1)
    if m >= 3:
        new_xyz[:, -1] += 1000.0  # a centre with no neighbour at all -> row of zeros
    want = cref.ball_query(radius, nsample, xyz, new_xyz)
    got = pu.ball_query(radius, nsample, T(xyz, cuda), T(new_xyz, cuda))
    assert got.dtype == torch.int32 and tuple(got.shape) == (b, m, nsample)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_ball_query_vs_reference_cuda(cuda, ref_ext):
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    rng = np.random.default_rng(5)
    for (b, n, m, r, ns) in [(2, 16384, 4096, 0.1, 16), (2, 16384, 4096, 0.5, 32), (16, 512, 128, 0.2, 64)]:
        xyz = T(clustered_cloud(rng, b, n), cuda)
        new_xyz = xyz[:, :m].contiguous()
        ref = torch.zeros(b, m, ns, dtype=torch.int32, device=cuda)
        ref_ext.pointnet2_cuda.ball_query_wrapper(b, n, m, r, ns, new_xyz, xyz, ref)
        torch.cuda.synchronize()
        got = pu.ball_query(r, ns, xyz, new_xyz)
        assert torch.equal(got, ref)


@pytest.mark.parametrize("b,c,n,npoint,nsample", [(2, 3, 4096, 1024, 32), (1, 96, 4096, 1024, 16), (3, 7, 100, 13, 5)])
def test_group_and_gather_bit_exact(cuda, cref, b, c, n, npoint, nsample):
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    rng = np.random.default_rng(c)
    feats = rng.normal(size=(b, c, n)).astype(np.float32)
    idx = rng.integers(0, n, (b, npoint, nsample)).astype(np.int32)
    got = pu.grouping_operation(T(feats, cuda), T(idx, cuda))
    np.testing.assert_array_equal(got.cpu().numpy(), cref.group_points(feats, idx))
    idx1 = rng.integers(0, n, (b, npoint)).astype(np.int32)
    got = pu.gather_operation(T(feats, cuda), T(idx1, cuda))
    np.testing.assert_array_equal(got.cpu().numpy(), cref.gather_points(feats, idx1))


def test_query_and_group_matches_composition(cuda, cref):
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    rng = np.random.default_rng(11)
    b, n, m, c = 2, 4096, 1024, 5
    xyz = clustered_cloud(rng, b, n)
    new_xyz = xyz[:, :m].copy()
    feats = rng.normal(size=(b, c, n)).astype(np.float32)
    out = pu.QueryAndGroup(0.5, 32)(T(xyz, cuda), T(new_xyz, cuda), T(feats, cuda)).cpu().numpy()
    idx = cref.ball_query(0.5, 32, xyz, new_xyz)
    gx = cref.group_points(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx) - new_xyz.transpose(0, 2, 1)[..., None]
    np.testing.assert_array_equal(out[:, :3], gx)
    np.testing.assert_array_equal(out[:, 3:], cref.group_points(feats, idx))


# ---------------------------------------------------------------- farthest point sampling
@pytest.mark.parametrize("b,n,m", [
    (2, 16384, 4096), (2, 4096, 1024), (3, 1024, 256), (3, 256, 64),   # RPN levels
    (40, 512, 128), (40, 128, 32),                                     # RCNN levels
    (2, 1000, 300), (2, 37, 37), (1, 3, 2), (1, 1, 1), (2, 5000, 64),  # ragged / tiny
])
def test_fps_bit_exact(cuda, cref, b, n, m):
    from jmodt_b200.pointnet2 import pointnet2_cuda as pc
    rng = np.random.default_rng(n + m)
    xyz = clustered_cloud(rng, b, n, dup_frac=0.2)  # many exact ties
    want_idx, want_temp = cref.fps(xyz, m, return_temp=True)
    x = T(xyz, cuda)
    idx = torch.empty(b, m, dtype=torch.int32, device=cuda)
    temp = torch.empty(b, n, dtype=torch.float32, device=cuda)
    pc.farthest_point_sampling_wrapper(b, n, m, x, temp, idx)
    np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
    np.testing.assert_array_equal(temp.cpu().numpy(), want_temp)


def test_fps_all_identical_points(cuda, cref):
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    xyz = np.ones((2, 2048, 3), np.float32)
    got = pu.farthest_point_sample(T(xyz, cuda), 64).cpu().numpy()
    np.testing.assert_array_equal(got, cref.fps(xyz, 64))


def test_fps_large_cloud_fallback(cuda, cref):
    from jmodt_b200.pointnet2 import pointnet2_cuda as pc
    rng = np.random.default_rng(3)
    b, n, m = 1, 20000, 48
    xyz = clustered_cloud(rng, b, n)
    idx = torch.empty(b, m, dtype=torch.int32, device=cuda)
    temp = torch.empty(b, n, dtype=torch.float32, device=cuda)
    pc.farthest_point_sampling_wrapper(b, n, m, T(xyz, cuda), temp, idx)
    np.testing.assert_array_equal(idx.cpu().numpy(), cref.fps(xyz, m))


def test_fps_vs_reference_cuda(cuda, ref_ext):
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    rng = np.random.default_rng(9)
    for (b, n, m) in [(2, 16384, 4096), (4, 4096, 1024), (8, 512, 128), (8, 128, 32), (2, 777, 100)]:
        xyz = T(clustered_cloud(rng, b, n, dup_frac=0.2), cuda)
        ref = torch.empty(b, m, dtype=torch.int32, device=cuda)
        temp = torch.full((b, n), 1e10, dtype=torch.float32, device=cuda)
        ref_ext.pointnet2_cuda.farthest_point_sampling_wrapper(b, n, m, xyz, temp, ref)
        torch.cuda.synchronize()
        assert torch.equal(pu.farthest_point_sample(xyz, m), ref)


# ---------------------------------------------------------------- three_nn / three_interpolate
@pytest.mark.parametrize("b,n,m", [(2, 16384, 4096), (2, 4096, 1024), (2, 1024, 256), (2, 256, 64),
                                   (1, 100, 2), (1, 10, 1), (3, 999, 77)])
def test_three_nn_bit_exact(cuda, cref, b, n, m):
    from jmodt_b200.pointnet2 import pointnet2_cuda as pc
    rng = np.random.default_rng(n - m)
    unknown = clustered_cloud(rng, b, n, dup_frac=0.2)
    known = unknown[:, rng.choice(n, m, replace=(m > n))].copy()  # includes exact matches + duplicates
    want_d2, want_idx = cref.three_nn(unknown, known)
    d2 = torch.empty(b, n, 3, device=cuda)
    idx = torch.empty(b, n, 3, dtype=torch.int32, device=cuda)
    pc.three_nn_wrapper(b, n, m, T(unknown, cuda), T(known, cuda), d2, idx)
    np.testing.assert_array_equal(idx.cpu().numpy(), want_idx)
    np.testing.assert_array_equal(d2.cpu().numpy(), want_d2)


def test_three_nn_and_interpolate_vs_reference_cuda(cuda, ref_ext):
    from jmodt_b200.pointnet2 import pointnet2_cuda as pc
    rng = np.random.default_rng(2)
    for (b, n, m, c) in [(2, 16384, 4096, 64), (2, 256, 64, 512), (1, 50, 5, 3)]:
        unknown = T(clustered_cloud(rng, b, n, dup_frac=0.2), cuda)
        known = unknown[:, :m].contiguous()
        d2r = torch.empty(b, n, 3, device=cuda); ir = torch.empty(b, n, 3, dtype=torch.int32, device=cuda)
        ref_ext.pointnet2_cuda.three_nn_wrapper(b, n, m, unknown, known, d2r, ir)
        d2 = torch.empty_like(d2r); ii = torch.empty_like(ir)
        pc.three_nn_wrapper(b, n, m, unknown, known, d2, ii)
        torch.cuda.synchronize()
        assert torch.equal(ii, ir) and torch.equal(d2, d2r)
        feats = torch.randn(b, c, m, device=cuda)
        w = torch.rand(b, n, 3, device=cuda)
        outr = torch.empty(b, c, n, device=cuda); out = torch.empty_like(outr)
        ref_ext.pointnet2_cuda.three_interpolate_wrapper(b, c, m, n, feats, ir, w, outr)
        pc.three_interpolate_wrapper(b, c, m, n, feats, ir, w, out)
        torch.cuda.synchronize()
        assert torch.equal(out, outr)


def test_three_interpolate_bit_exact(cuda, cref):
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    rng = np.random.default_rng(4)
    b, c, m, n = 2, 37, 64, 300
    feats = rng.normal(size=(b, c, m)).astype(np.float32)
    idx = rng.integers(0, m, (b, n, 3)).astype(np.int32)
    w = rng.uniform(size=(b, n, 3)).astype(np.float32)
    got = pu.three_interpolate(T(feats, cuda), T(idx, cuda), T(w, cuda))
    np.testing.assert_array_equal(got.cpu().numpy(), cref.three_interpolate(feats, idx, w))


def test_backward_ops_are_deterministic_and_match_oracle_and_reference_cuda(cuda, cref, ref_ext):
    """The three backward ops with a FIXED summation order (ascending position per target = the order of the oracle's
    sequential loops): bit-equal to the oracle, bit-identical run to run, and within fp32 reordering (1e-5) of the
    reference's atomicAdd kernels at a real level shape."""
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    rng = np.random.default_rng(18)
    b, c, n, npoint, ns = 2, 16, 4096, 1024, 32
    feats = torch.from_numpy(rng.normal(size=(b, c, n)).astype(np.float32)).to(cuda).requires_grad_(True)
    idx = np.minimum(rng.integers(0, n, (b, npoint, ns)), rng.integers(0, n, (b, npoint, ns))).astype(np.int32)   # skewed: long runs
    g = rng.normal(size=(b, c, npoint, ns)).astype(np.float32)
    grads = []
    for _ in range(2):
        feats.grad = None
        pu.grouping_operation(feats, T(idx, cuda)).backward(T(g, cuda))
        grads.append(feats.grad.clone())
    assert torch.equal(grads[0], grads[1])
    np.testing.assert_array_equal(grads[0].cpu().numpy(), cref.group_points_grad(g, idx, n))
    ref_grad = torch.zeros(b, c, n, device=cuda)
    ref_ext.pointnet2_cuda.group_points_grad_wrapper(b, c, n, npoint, ns, T(g, cuda), T(idx, cuda), ref_grad)
    np.testing.assert_allclose(grads[0].cpu().numpy(), ref_grad.cpu().numpy(), rtol=1e-5, atol=2e-5)
    feats.grad = None
    idx1 = rng.integers(0, n, (b, npoint)).astype(np.int32)
    g1 = rng.normal(size=(b, c, npoint)).astype(np.float32)
    pu.gather_operation(feats, T(idx1, cuda)).backward(T(g1, cuda))
    np.testing.assert_array_equal(feats.grad.cpu().numpy(), cref.gather_points_grad(g1, idx1, n))
    ref_grad.zero_()
    ref_ext.pointnet2_cuda.gather_points_grad_wrapper(b, c, n, npoint, T(g1, cuda), T(idx1, cuda), ref_grad)
    np.testing.assert_allclose(feats.grad.cpu().numpy(), ref_grad.cpu().numpy(), rtol=1e-5, atol=2e-5)
    feats.grad = None
    nu = 3000
    idx3 = rng.integers(0, n, (b, nu, 3)).astype(np.int32)
    w = rng.uniform(size=(b, nu, 3)).astype(np.float32)
    g3 = rng.normal(size=(b, c, nu)).astype(np.float32)
    pu.three_interpolate(feats, T(idx3, cuda), T(w, cuda)).backward(T(g3, cuda))
    np.testing.assert_array_equal(feats.grad.cpu().numpy(), cref.three_interpolate_grad(g3, idx3, w, n))
    ref_grad.zero_()
    ref_ext.pointnet2_cuda.three_interpolate_grad_wrapper(b, c, nu, n, T(g3, cuda), T(idx3, cuda), T(w, cuda), ref_grad)
    np.testing.assert_allclose(feats.grad.cpu().numpy(), ref_grad.cpu().numpy(), rtol=1e-5, atol=2e-5)


def test_backward_ops_match_oracle(cuda, cref):
    """Small shapes incl. the legacy atomic entry points of the pybind surface."""
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    rng = np.random.default_rng(8)
    b, c, n, npoint, ns = 2, 6, 200, 50, 8
    feats = torch.from_numpy(rng.normal(size=(b, c, n)).astype(np.float32)).to(cuda).requires_grad_(True)
    idx = rng.integers(0, n, (b, npoint, ns)).astype(np.int32)
    g = rng.normal(size=(b, c, npoint, ns)).astype(np.float32)
    pu.grouping_operation(feats, T(idx, cuda)).backward(T(g, cuda))
    np.testing.assert_allclose(feats.grad.cpu().numpy(), cref.group_points_grad(g, idx, n), rtol=1e-5, atol=1e-5)
    feats.grad = None
    idx1 = rng.integers(0, n, (b, npoint)).astype(np.int32)
    g1 = rng.normal(size=(b, c, npoint)).astype(np.float32)
    pu.gather_operation(feats, T(idx1, cuda)).backward(T(g1, cuda))
    np.testing.assert_allclose(feats.grad.cpu().numpy(), cref.gather_points_grad(g1, idx1, n), rtol=1e-5, atol=1e-5)
    feats.grad = None
    m = n
    idx3 = rng.integers(0, m, (b, 70, 3)).astype(np.int32)
    w = rng.uniform(size=(b, 70, 3)).astype(np.float32)
    g3 = rng.normal(size=(b, c, 70)).astype(np.float32)
    pu.three_interpolate(feats, T(idx3, cuda), T(w, cuda)).backward(T(g3, cuda))
    np.testing.assert_allclose(feats.grad.cpu().numpy(), cref.three_interpolate_grad(g3, idx3, w, m),
                               rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------- roipool3d
def _roi_case(seed, b, n, m, c):
    from jmodt_b200 import synth
    batch = synth.make_batch(seed, b, n_points=n, n_rois=m, with_image=False, empty_rois=4 if m >= 8 else 1)
    rng = np.random.default_rng(seed)
    feats = rng.normal(size=(b, n, c)).astype(np.float32)
    return batch["pts"], feats, batch["rois"]


@pytest.mark.parametrize("b,n,m,c,s", [(1, 16384, 128, 130, 512), (2, 16384, 128, 130, 512), (2, 4096, 20, 7, 64),
                                       (1, 300, 3, 1, 16)])
def test_roipool3d_bit_exact(cuda, cref, b, n, m, c, s):
    from jmodt_b200 import synth
    from jmodt_b200.roipool3d import roipool3d_utils as ru
    pts, feats, rois = _roi_case(100 + n, b, n, m, c)
    enlarged = cref.enlarge_box3d(rois, 0.2)
    for bi in range(b):
        pts[bi] = synth.nudge_off_box_faces(pts[bi], enlarged[bi])
    want, want_empty = cref.roipool3d(pts, feats, enlarged, s)
    got, got_empty = ru.roipool3d_gpu(T(pts, cuda), T(feats, cuda), T(rois, cuda), 0.2, sampled_pt_num=s)
    assert got_empty.dtype == torch.int32
    np.testing.assert_array_equal(got_empty.cpu().numpy(), want_empty)
    assert want_empty.sum() >= 1 and (want_empty == 0).sum() >= 1   # the case covers empty and non-empty boxes
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_roipool3d_vs_reference_cuda(cuda, ref_ext):
    from jmodt_b200.roipool3d import roipool3d_utils as ru
    from jmodt_b200 import box_utils
    pts, feats, rois = _roi_case(7, 2, 16384, 128, 130)
    p, f, r = T(pts, cuda), T(feats, cuda), T(rois, cuda)
    enlarged = box_utils.enlarge_box3d(r.view(-1, 7), 0.2).view(2, -1, 7).contiguous()
    ref = torch.zeros(2, 128, 512, 133, device=cuda)
    ref_empty = torch.zeros(2, 128, dtype=torch.int32, device=cuda)
    ref_ext.roipool3d_cuda.forward(p, enlarged, f, ref, ref_empty)
    torch.cuda.synchronize()
    got, got_empty = ru.roipool3d_gpu(p, f, r, 0.2)
    assert torch.equal(got_empty, ref_empty)
    assert torch.equal(got, ref)


def test_roipool3d_canonical_matches_reference_eval_branch(cuda):
    """proposal_target_layer.py:99-115 restated with torch ops; the rotation goes through a matmul in
    the reference, so xyz is compared to 1e-5 absolute (coordinates are O(10 m)); features exactly."""
    from jmodt_b200.roipool3d import roipool3d_utils as ru
    from jmodt_b200 import box_utils
    pts, feats, rois = _roi_case(21, 2, 16384, 128, 130)
    p, f, r = T(pts, cuda), T(feats, cuda), T(rois, cuda)
    pooled, empty = ru.roipool3d_gpu(p, f, r, 0.2)
    pooled[:, :, :, 0:3] -= r[:, :, 0:3].unsqueeze(2)
    for k in range(2):
        pooled[k, :, :, 0:3] = box_utils.rotate_pc_along_y_torch(pooled[k, :, :, 0:3], r[k, :, 6])
    got, got_empty = ru.roipool3d_gpu_canonical(p, f, r, 0.2)
    assert torch.equal(got_empty, empty)
    assert torch.equal(got[..., 3:], pooled[..., 3:])
    np.testing.assert_allclose(got[..., :3].cpu().numpy(), pooled[..., :3].cpu().numpy(), atol=1e-5, rtol=0)


# ---------------------------------------------------------------- iou3d / nms
def _boxes7(rng, n, spread=30.0):
    ctr = rng.uniform(-spread, spread, (n, 3)); ctr[:, 1] = rng.normal(1.6, 0.2, n)
    size = np.array([1.5, 1.6, 3.9]) * rng.uniform(0.7, 1.3, (n, 3))
    ry = rng.uniform(-np.pi, np.pi, (n, 1))
    return np.concatenate([ctr, size, ry], 1).astype(np.float32)


def _overlapping_pairs(rng, n):
    a = _boxes7(rng, n)
    b = a.copy()
    b[:, [0, 2]] += rng.normal(0, 0.8, (n, 2)); b[:, 6] += rng.normal(0, 0.3, n); b[:, 1] += rng.normal(0, 0.2, n)
    b[: n // 8] = a[: n // 8]                       # identical boxes
    b[n // 8: n // 4, 6] = a[n // 8: n // 4, 6]     # parallel edges
    return a, b.astype(np.float32)


def test_bev_overlap_and_iou_match_oracle(cuda, cref):
    """Same arithmetic as the reference binaries, so the oracle is expected to agree bit for bit; the
    assert allows 2e-6 absolute on areas O(1..10 m^2) for libdevice-version drift in sinf/cosf/atan2f."""
    from jmodt_b200.iou3d import iou3d_cuda, iou3d_utils
    from jmodt_b200 import box_utils
    rng = np.random.default_rng(31)
    a7, b7 = _overlapping_pairs(rng, 128)
    a5, b5 = cref.boxes3d_to_bev(a7), cref.boxes3d_to_bev(b7)
    ov = torch.empty(128, 128, device=cuda)
    iou3d_cuda.boxes_overlap_bev_gpu(T(a5, cuda), T(b5, cuda), ov)
    want = cref.boxes_overlap_bev(a5, b5)
    got = ov.cpu().numpy()
    np.testing.assert_allclose(got, want, atol=2e-6, rtol=0)
    assert (got == want).mean() > 0.999
    iou = iou3d_utils.boxes_iou_bev(T(a5, cuda), T(b5, cuda)).cpu().numpy()
    np.testing.assert_allclose(iou, cref.boxes_iou_bev(a5, b5), atol=2e-6, rtol=0)
    iou3 = iou3d_utils.boxes_iou3d_gpu(T(a7, cuda), T(b7, cuda)).cpu().numpy()
    np.testing.assert_allclose(iou3, cref.boxes_iou3d(a7, b7), atol=2e-6, rtol=0)
    bev = box_utils.boxes3d_to_bev_torch(T(a7, cuda)).cpu().numpy()
    np.testing.assert_array_equal(bev, a5)


def test_bev_overlap_vs_reference_cuda(cuda, ref_ext):
    from jmodt_b200.iou3d import iou3d_cuda
    rng = np.random.default_rng(32)
    a7, b7 = _overlapping_pairs(rng, 256)
    from oracle import cref
    a5, b5 = T(cref.boxes3d_to_bev(a7), cuda), T(cref.boxes3d_to_bev(b7), cuda)
    ref = torch.zeros(256, 256, device=cuda); got = torch.empty_like(ref)
    ref_ext.iou3d_cuda.boxes_overlap_bev_gpu(a5, b5, ref)
    iou3d_cuda.boxes_overlap_bev_gpu(a5, b5, got)
    torch.cuda.synchronize()
    assert torch.equal(got, ref)
    ref_ext.iou3d_cuda.boxes_iou_bev_gpu(a5, b5, ref)
    iou3d_cuda.boxes_iou_bev_gpu(a5, b5, got)
    torch.cuda.synchronize()
    assert torch.equal(got, ref)


def _nms_case(rng, n):
    base = _boxes7(rng, max(n // 6, 1), spread=25.0)
    boxes = base[rng.integers(0, base.shape[0], n)].copy()
    boxes[:, [0, 2]] += rng.normal(0, 0.5, (n, 2)); boxes[:, 6] += rng.normal(0, 0.2, n)
    scores = rng.uniform(size=n).astype(np.float32)
    return boxes.astype(np.float32), scores


@pytest.mark.parametrize("n,thresh,rotated", [(6300, 0.8, False), (2700, 0.8, False), (128, 0.1, True),
                                              (1000, 0.5, True), (65, 0.3, False), (1, 0.5, True), (0, 0.5, False)])
def test_nms_keep_list_bit_exact(cuda, cref, n, thresh, rotated):
    from jmodt_b200.iou3d import iou3d_utils
    rng = np.random.default_rng(n + 1)
    boxes7, scores = _nms_case(rng, n) if n else (np.zeros((0, 7), np.float32), np.zeros(0, np.float32))
    bev = cref.boxes3d_to_bev(boxes7)
    s = T(scores, cuda)
    order = s.sort(0, descending=True)[1].cpu().numpy()      # same sort call as the reference wrapper
    want = order[cref.nms_sorted(bev[order], thresh, rotated)] if n else np.zeros(0, np.int64)
    fn = iou3d_utils.nms_gpu if rotated else iou3d_utils.nms_normal_gpu
    got = fn(T(bev, cuda), s, thresh)
    assert got.dtype == torch.int64 and got.is_cuda
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    if n > 10:
        assert 0 < len(want) < n


def test_nms_vs_reference_cuda_and_max_keep(cuda, ref_ext):
    from jmodt_b200.iou3d import iou3d_cuda
    from oracle import cref
    rng = np.random.default_rng(77)
    for n, thresh, rotated in [(6300, 0.8, False), (900, 0.4, True)]:
        boxes7, scores = _nms_case(rng, n)
        order = np.argsort(-scores, kind="stable")
        bev = T(cref.boxes3d_to_bev(boxes7)[order], cuda)
        keep_ref = torch.zeros(n, dtype=torch.int64)
        fn = ref_ext.iou3d_cuda.nms_gpu if rotated else ref_ext.iou3d_cuda.nms_normal_gpu
        num = fn(bev, keep_ref, thresh)
        keep, num_dev = iou3d_cuda.nms_device(bev, thresh, rotated)
        assert int(num_dev.item()) == num
        assert torch.equal(keep[:num].cpu(), keep_ref[:num])
        keep2, num2 = iou3d_cuda.nms_device(bev, thresh, rotated, max_keep=50)
        assert int(num2.item()) == min(50, num)
        assert torch.equal(keep2[:min(50, num)].cpu(), keep_ref[:min(50, num)])
        host_keep = torch.zeros(n, dtype=torch.int64)
        legacy = iou3d_cuda.nms_gpu if rotated else iou3d_cuda.nms_normal_gpu
        assert legacy(bev, host_keep, thresh) == num and torch.equal(host_keep[:num], keep_ref[:num])


# ---------------------------------------------------------------- error behaviour
def test_errors_are_raised_not_fatal(cuda):
    from jmodt_b200 import _lib
    from jmodt_b200.pointnet2 import pointnet2_cuda as pc
    x = torch.zeros(1, 8, 3, device=cuda)
    with pytest.raises(_lib.JmodtB200Error):
        pc.ball_query_wrapper(1, 8, 8, 0.5, 4, x.cpu(), x, torch.zeros(1, 8, 4, dtype=torch.int32, device=cuda))
    with pytest.raises(_lib.JmodtB200Error):
        pc.ball_query_wrapper(1, 8, 8, 0.5, 4, x.transpose(1, 2), x, torch.zeros(1, 8, 4, dtype=torch.int32, device=cuda))
    assert _lib.lib().jmb_ball_query(-1, 8, 8, 0.5, 4, None, None, None, None) == -1
    assert b"negative" in _lib.lib().jmb_last_error()
    assert _lib.lib().jmb_nms(10, x.data_ptr(), 0.5, x.data_ptr(), x.data_ptr(), 0, None, 0, None) == -3


def test_ball_query_msg2_equals_two_queries(cuda, cref):
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    rng = np.random.default_rng(77)
    for (b, n, m, ra, na, rb, nb) in [(2, 16384, 4096, 0.1, 16, 0.5, 32), (3, 1000, 37, 1.0, 5, 2.0, 9), (8, 256, 64, 2.0, 16, 4.0, 32)]:
        xyz = clustered_cloud(rng, b, n)
        new_xyz = xyz[:, :m].copy()
        new_xyz[:, -1] += 1000.0
        ia, ib = pu.ball_query_msg2(ra, na, rb, nb, T(xyz, cuda), T(new_xyz, cuda))
        np.testing.assert_array_equal(ia.cpu().numpy(), cref.ball_query(ra, na, xyz, new_xyz))
        np.testing.assert_array_equal(ib.cpu().numpy(), cref.ball_query(rb, nb, xyz, new_xyz))


def test_ball_query_cell_list_equals_oracle_and_reference_cuda(cuda, cref, ref_ext):
    """Large clouds go through the hashed cell list (csrc/ball_query.cu: ball_query_grid_kernel): identical rows to the
    oracle / the reference kernel on (i) a KITTI-shaped synthetic frame at the RPN level-0 shape, (ii) a cloud so dense that
    balls hold more than the 256 buffered hits (ordered-scan fallback), (iii) negative coordinates and far-away centres
    without any hit."""
    from jmodt_b200 import synth
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    rng = np.random.default_rng(5)
    batch = synth.make_batch(31, 2, with_image=False)["pts"]
    cases = [(batch, 4096, 0.1, 16, 0.5, 32),
             ((rng.normal(0, 0.35, (1, 8192, 3))).astype(np.float32), 512, 0.3, 16, 0.6, 32),
             ((rng.uniform(-30, 30, (2, 9000, 3))).astype(np.float32), 300, 1.0, 8, 2.5, 64)]
    for xyz, m, ra, na, rb, nb in cases:
        b, n, _ = xyz.shape
        new_xyz = np.ascontiguousarray(xyz[:, rng.permutation(n)[:m]])
        new_xyz[:, -1] += 500.0                                # a centre with no neighbour at all
        x, c = T(xyz, cuda), T(new_xyz, cuda)
        ia, ib = pu.ball_query_msg2(ra, na, rb, nb, x, c)
        np.testing.assert_array_equal(ia.cpu().numpy(), cref.ball_query(ra, na, xyz, new_xyz))
        np.testing.assert_array_equal(ib.cpu().numpy(), cref.ball_query(rb, nb, xyz, new_xyz))
        for r, ns, got in ((ra, na, ia), (rb, nb, ib)):
            ref = torch.zeros(b, m, ns, dtype=torch.int32, device=cuda)
            ref_ext.pointnet2_cuda.ball_query_wrapper(b, n, m, r, ns, c, x, ref)
            assert torch.equal(got, ref)
    dense_hits = (np.linalg.norm(cases[1][0][0][None] - cases[1][0][0][:64, None], axis=2) < 0.6).sum(1)
    assert dense_hits.max() > 256                              # the fallback path was exercised
