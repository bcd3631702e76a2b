"""Tracker association inputs: boxes_dist_gpu / link_matrix on the sm_100a kernel vs the CPU
restatement of the reference (oracle/tracking_ref.py).  fp32 geometry: tolerance 1e-5 absolute on values in [-1, 1]."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _boxes(n, seed):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(n, 3, generator=g) * torch.tensor([80.0, 4.0, 70.0]) + torch.tensor([-40.0, -1.0, 0.0])
    hwl = torch.tensor([1.5256, 1.6286, 3.8831]) * (0.8 + 0.4 * torch.rand(n, 3, generator=g))
    ry = (torch.rand(n, 1, generator=g) * 2 - 1) * np.pi
    return torch.cat([xyz, hwl, ry], dim=1)


@pytest.mark.parametrize("P,D", [(128, 128), (37, 50), (1, 3), (200, 1)])
def test_boxes_dist_and_link_matrix_match_the_reference_restatement(cuda, P, D):
    from jmodt_b200 import association
    from jmodt_b200.iou3d.iou3d_utils import boxes_iou3d_gpu
    from oracle import tracking_ref
    a, b = _boxes(P, 1), _boxes(D, 2)
    b[: min(P, D)] = a[: min(P, D)] + 0.05 * torch.randn(min(P, D), 7, generator=torch.Generator().manual_seed(3))
    got = association.boxes_dist_gpu(a.to(cuda), b.to(cuda))
    want = tracking_ref.boxes_dist(a, b)
    assert got.shape == (P, D)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), atol=1e-5, rtol=0)
    assert float(got.max()) <= 1.0 + 1e-6
    link = torch.rand(P, D, generator=torch.Generator().manual_seed(4))
    w_app, w_iou, w_dis = 0.5, 0.3, 0.2
    score = association.link_matrix(link.to(cuda), a.to(cuda), b.to(cuda), w_app, w_iou, w_dis)
    iou = boxes_iou3d_gpu(a.to(cuda), b.to(cuda)).cpu()          # pinned bit-exactly elsewhere (tests/test_ops_gpu.py)
    want_score = tracking_ref.link_matrix(link, iou, a, b, w_app, w_iou, w_dis)
    np.testing.assert_allclose(score.cpu().numpy(), want_score.numpy(), atol=1e-5, rtol=0)


def test_identical_boxes_have_distance_score_one(cuda):
    from jmodt_b200 import association
    a = _boxes(16, 7).to(cuda)
    d = association.boxes_dist_gpu(a, a)
    assert torch.allclose(torch.diagonal(d), torch.ones(16, device=cuda))
