"""Same call surface as the reference pybind module `pointnet2_cuda`
(jmodt/ops/pointnet2/src/pointnet2_api.cpp:10-24): caller-allocated outputs, sizes passed
redundantly as ints.  Each wrapper forwards raw pointers to the C ABI of libjmodt_b200.so
on torch's CURRENT stream (the reference launches on the legacy default stream).
"""
from __future__ import annotations

from .. import _lib


def _chk(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise _lib.JmodtB200Error("tensor must be a CUDA tensor")  # ball_query.cpp:10 CHECK_CUDA
        if not t.is_contiguous():
            raise _lib.JmodtB200Error("tensor must be contiguous")  # ball_query.cpp:11 CHECK_CONTIGUOUS


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    _chk(new_xyz, xyz, idx)
    st = _lib.stream_and_device(xyz)
    _lib.check(_lib.lib().jmb_ball_query(b, n, m, radius, nsample, new_xyz.data_ptr(), xyz.data_ptr(),
                                         idx.data_ptr(), st), "ball_query")
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    _chk(points, idx, out)
    st = _lib.stream_and_device(points)
    _lib.check(_lib.lib().jmb_group_points(b, c, n, npoints, nsample, points.data_ptr(), idx.data_ptr(),
                                           out.data_ptr(), st), "group_points")
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    _chk(grad_out, idx, grad_points)
    st = _lib.stream_and_device(grad_out)
    _lib.check(_lib.lib().jmb_group_points_grad(b, c, n, npoints, nsample, grad_out.data_ptr(),
                                                idx.data_ptr(), grad_points.data_ptr(), st),
               "group_points_grad")
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    _chk(points, idx, out)
    st = _lib.stream_and_device(points)
    _lib.check(_lib.lib().jmb_gather_points(b, c, n, npoints, points.data_ptr(), idx.data_ptr(),
                                            out.data_ptr(), st), "gather_points")
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    _chk(grad_out, idx, grad_points)
    st = _lib.stream_and_device(grad_out)
    _lib.check(_lib.lib().jmb_gather_points_grad(b, c, n, npoints, grad_out.data_ptr(), idx.data_ptr(),
                                                 grad_points.data_ptr(), st), "gather_points_grad")
    return 1


def farthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    _chk(points, idx)
    st = _lib.stream_and_device(points)
    _lib.check(_lib.lib().jmb_furthest_point_sampling(b, n, m, points.data_ptr(), _lib.ptr(temp),
                                                      idx.data_ptr(), st), "furthest_point_sampling")
    return 1


def wait_indices(idx, k0, k1, timeout_ms=2000):
    """Stream gate (no reference counterpart): returns in stream order once idx[:, k0:k1] >= 0 everywhere."""
    _chk(idx)
    st = _lib.stream_and_device(idx)
    _lib.check(_lib.lib().jmb_wait_indices(idx.data_ptr(), idx.shape[0], idx.stride(0), k0, k1, timeout_ms, st),
               "wait_indices")
    return 1


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    _chk(unknown, known, dist2, idx)
    st = _lib.stream_and_device(unknown)
    _lib.check(_lib.lib().jmb_three_nn(b, n, m, unknown.data_ptr(), known.data_ptr(), dist2.data_ptr(),
                                       idx.data_ptr(), st), "three_nn")


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    _chk(points, idx, weight, out)
    st = _lib.stream_and_device(points)
    _lib.check(_lib.lib().jmb_three_interpolate(b, c, m, n, points.data_ptr(), idx.data_ptr(),
                                                weight.data_ptr(), out.data_ptr(), st), "three_interpolate")


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    _chk(grad_out, idx, weight, grad_points)
    st = _lib.stream_and_device(grad_out)
    _lib.check(_lib.lib().jmb_three_interpolate_grad(b, c, n, m, grad_out.data_ptr(), idx.data_ptr(),
                                                     weight.data_ptr(), grad_points.data_ptr(), st),
               "three_interpolate_grad")
