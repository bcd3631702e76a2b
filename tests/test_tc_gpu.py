"""tcgen05 MLP layer vs a plain PyTorch fp32 reference (TF32 disabled).  Tolerance: the 3-term bf16 split
has <= 2^-16 relative error per product, so we require 2e-5 of the output scale per layer (north_star: 1e-4)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _check(got, want, tol=2e-5):
    scale = want.abs().max().item() + 1e-12
    err = (got - want).abs().max().item() / scale
    assert err < tol, f"normwise error {err:.3e} >= {tol}"


@pytest.mark.parametrize("G,K,N,M,relu", [(1, 32, 128, 128, False), (2, 64, 256, 128, True), (3, 131, 8192, 128, True),
                                          (2, 128, 300, 256, True), (1, 5, 512, 128, True), (4, 512, 100, 46, False),
                                          (1, 512, 16384, 1, False)])
def test_dense_layer(cuda, G, K, N, M, relu):
    from jmodt_b200 import tc
    g = torch.Generator(device="cpu").manual_seed(K * 1000 + N)
    w = torch.randn(M, K, generator=g) / K ** 0.5
    b = torch.randn(M, generator=g)
    x = torch.randn(G, K, N, generator=g)
    layer = tc.PackedLayer(w.to(cuda), b.to(cuda), relu)
    got = tc.mlp_layer(layer, x.to(cuda).contiguous())
    want = torch.einsum("mk,gkn->gmn", w.double(), x.double()) + b.double()[None, :, None]
    if relu:
        want = want.clamp_min(0)
    _check(got.cpu().double(), want)


def test_maxpool_epilogue_and_grouped_prologue(cuda):
    from jmodt_b200 import tc
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    g = torch.Generator(device="cpu").manual_seed(7)
    G, n_pts, C, npoint, ns, M = 6, 512, 128, 128, 64, 128
    xyz = (torch.rand(G, n_pts, 3, generator=g) * 2).to(cuda)
    feats = torch.randn(G, C, n_pts, generator=g).to(cuda)
    w = (torch.randn(M, 3 + C, generator=g) / 11).to(cuda)
    b = torch.randn(M, generator=g).to(cuda)
    fidx = pu.farthest_point_sample(xyz, npoint)
    centres = pu.gather_operation(xyz.transpose(1, 2).contiguous(), fidx).transpose(1, 2).contiguous()
    idx = pu.ball_query(0.4, ns, xyz, centres)
    grouped = pu.QueryAndGroup(0.4, ns)(xyz, centres, feats)                     # (G, 3+C, npoint, ns)
    want = torch.einsum("mk,gkps->gmps", w.double(), grouped.double()) + b.double()[None, :, None, None]
    want = want.clamp_min(0)
    layer = tc.PackedLayer(w, b, True)
    got = tc.grouped_first_layer(layer, xyz, feats, idx, centres, ns)
    _check(got.view(G, M, npoint, ns).double(), want)
    got_pool = tc.grouped_first_layer(layer, xyz, feats, idx, centres, ns, pool=ns)
    _check(got_pool.double(), want.max(dim=3)[0])
    dense_pool = tc.mlp_layer(layer, grouped.view(G, 3 + C, npoint * ns).contiguous(), pool=ns)
    _check(dense_pool.double(), want.max(dim=3)[0])
    # GroupAll (no idx, no centring)
    ga = tc.grouped_first_layer(layer, xyz, feats, None, None, 0, pool=128)
    want_ga = torch.einsum("mk,gkn->gmn", w.double(), torch.cat([xyz.transpose(1, 2), feats], 1).double()) + b.double()[None, :, None]
    _check(ga.double(), want_ga.clamp_min(0).view(G, M, n_pts // 128, 128).max(dim=3)[0])
