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


@pytest.mark.parametrize("G,K,N,M", [(5, 64, 32, 128), (1024, 70, 32, 200), (7, 259, 64, 256), (9, 40, 8, 64), (3, 96, 16, 128),
                                     (2, 33, 130, 128), (3, 64, 1026, 96), (2, 48, 12, 16)])
def test_narrow_groups_packed_into_one_tile_and_unaligned_rows(cuda, G, K, N, M):
    """N < 128 dividing 128 packs 128/N groups per tile (bulk-copy paths), N % 4 != 0 takes the register / direct-store
    paths; both against fp64, for the dense, pooled and point-major epilogues and for a channel-slice output."""
    from jmodt_b200 import tc
    g = torch.Generator(device="cpu").manual_seed(G * 7 + K * 1000 + N)
    w = torch.randn(M, K, generator=g) / K ** 0.5
    b = torch.randn(M, generator=g)
    x = torch.randn(G, K, N, generator=g)
    layer = tc.PackedLayer(w.to(cuda), b.to(cuda), True)
    xd = x.to(cuda).contiguous()
    want = (torch.einsum("mk,gkn->gmn", w.double(), x.double()) + b.double()[None, :, None]).clamp_min(0)
    _check(tc.mlp_layer(layer, xd).cpu().double(), want)
    _check(tc.mlp_layer(layer, xd, point_major_out=True).cpu().double(), want.transpose(1, 2))
    wide = torch.full((G, M + 24, N), -7.0, device=cuda)
    tc.mlp_layer(layer, xd, out=wide[:, 8:8 + M])
    _check(wide[:, 8:8 + M].cpu().double(), want)
    assert (wide[:, :8] == -7).all() and (wide[:, 8 + M:] == -7).all()
    for pool in (8, 32, 64, 128):
        if N % pool == 0:
            _check(tc.mlp_layer(layer, xd, pool=pool).cpu().double(), want.view(G, M, N // pool, pool).max(dim=3)[0])


def test_group_all_gather_with_packed_groups(cuda):
    """GroupAll prologue (no idx, no centring) over 32 points per group: four groups share a tile."""
    from jmodt_b200 import tc
    g = torch.Generator(device="cpu").manual_seed(11)
    G, n_pts, C, M = 37, 32, 256, 256
    xyz = torch.rand(G, n_pts, 3, generator=g).to(cuda)
    feats = torch.randn(G, C, n_pts, generator=g).to(cuda)
    w = (torch.randn(M, 3 + C, generator=g) / 16).to(cuda)
    b = torch.randn(M, generator=g).to(cuda)
    layer = tc.PackedLayer(w, b, True)
    want = (torch.einsum("mk,gkn->gmn", w.double(), torch.cat([xyz.transpose(1, 2), feats], 1).double())
            + b.double()[None, :, None]).clamp_min(0)
    _check(tc.grouped_first_layer(layer, xyz, feats, None, None, 0).double(), want)
    _check(tc.grouped_first_layer(layer, xyz, feats, None, None, 0, pool=32).double(), want.max(dim=2, keepdim=True)[0])


def test_many_launches_reuse_the_scheduler_slots(cuda):
    """More launches than scheduler slots, alternating shapes: every launch must see a re-armed tile counter."""
    from jmodt_b200 import tc
    g = torch.Generator(device="cpu").manual_seed(3)
    w = torch.randn(128, 64, generator=g) / 8
    layer = tc.PackedLayer(w.to(cuda), None, False)
    xs = [torch.randn(2, 64, n, generator=g).to(cuda) for n in (128, 4096, 640)]
    wants = [torch.einsum("mk,gkn->gmn", w.double(), x.cpu().double()) for x in xs]
    for i in range(200):
        got = tc.mlp_layer(layer, xs[i % 3])
        if i % 17 == 0 or i > 190:
            _check(got.cpu().double(), wants[i % 3])


@pytest.mark.parametrize("C,npoint,ns,C3,C1,C2", [(128, 128, 64, 128, 128, 128), (128, 32, 64, 256, 128, 128),
                                                  (64, 64, 16, 128, 128, 128), (128, 16, 8, 256, 128, 128),
                                                  (0, 128, 16, 32, 16, 16), (0, 64, 32, 64, 32, 32),       # RPN level 0
                                                  (96, 64, 16, 128, 64, 64), (96, 32, 32, 128, 64, 96),    # RPN level 1
                                                  (8, 16, 8, 200, 24, 40)])
def test_sa_fused_single_kernel_matches_layerwise_and_torch(cuda, C, npoint, ns, C3, C1, C2):
    """The one-kernel set-abstraction layer vs (a) the layer-by-layer tcgen05 path and (b) torch fp32."""
    from jmodt_b200 import tc
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    g = torch.Generator(device="cpu").manual_seed(C + npoint)
    G, n_pts = 37, 512
    xyz = (torch.rand(G, n_pts, 3, generator=g) * 2).to(cuda)
    feats = torch.randn(G, C, n_pts, generator=g).to(cuda) if C else None
    dims = [3 + C, C1, C2, C3]
    layers, ws = [], []
    for i in range(3):
        w = (torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5).to(cuda)
        b = (torch.randn(dims[i + 1], generator=g) * 0.2).to(cuda)
        ws.append((w, b))
        layers.append(tc.PackedLayer(w, b, True))
    fidx = pu.farthest_point_sample(xyz, npoint)
    centres = pu.gather_operation(xyz.transpose(1, 2).contiguous(), fidx).transpose(1, 2).contiguous()
    idx = pu.ball_query(0.4, ns, xyz, centres)
    assert tc.sa_fused_supported(layers, C, npoint, ns)
    got = tc.sa_fused(layers, xyz, feats, idx, centres)
    h = tc.grouped_first_layer(layers[0], xyz, feats, idx, centres, ns)
    h = tc.mlp_layer(layers[1], h)
    layerwise = tc.mlp_layer(layers[2], h, pool=ns)
    assert got.shape == layerwise.shape == (G, C3, npoint)
    _check(got.double(), layerwise.double(), tol=2e-5)        # same products; the K order of the fp32 accumulation differs (xyz last)
    x = pu.QueryAndGroup(0.4, ns)(xyz, centres, feats).double()
    for w, b in ws:
        x = (torch.einsum("mk,gkps->gmps", w.double(), x) + b.double()[None, :, None, None]).clamp_min(0)
    _check(got.double(), x.max(dim=3)[0], tol=5e-5)


@pytest.mark.parametrize("G,K,N,M,relu", [(4, 512, 16384, 512, True), (8, 128, 4096, 128, True), (1, 512, 1024, 256, True),
                                          (3, 96, 100, 46, False), (8, 64, 32, 128, True), (2, 256, 37, 200, True)])
def test_single_output_tail_folded_into_the_previous_epilogue(cuda, G, K, N, M, relu):
    """tc.mlp_layer_dot: layer (K -> M, bias, ReLU) followed by a M -> 1 layer in one launch (the C -> 1 tails of the cls /
    link / start-end heads) against fp64; ragged N, several M tiles, narrow groups packed into one tile; bit-identical
    when the columns are split differently (the sharded affinity relies on it)."""
    from jmodt_b200 import tc
    g = torch.Generator(device="cpu").manual_seed(K * 7 + N)
    w1, b1 = torch.randn(M, K, generator=g) / K ** 0.5, torch.randn(M, generator=g)
    w2, b2 = torch.randn(1, M, generator=g) / M ** 0.5, torch.randn(1, generator=g)
    x = torch.randn(G, K, N, generator=g)
    l1 = tc.PackedLayer(w1.to(cuda), b1.to(cuda), relu)
    l2 = tc.PackedLayer(w2.to(cuda), b2.to(cuda), False)
    xd = x.to(cuda).contiguous()
    got = tc.mlp_layer_dot(l1, l2, xd)
    h = torch.einsum("mk,gkn->gmn", w1.double(), x.double()) + b1.double()[None, :, None]
    if relu:
        h = h.clamp_min(0)
    want = torch.einsum("om,gmn->gon", w2.double(), h) + b2.double()
    assert got.shape == (G, 1, N)
    _check(got.cpu().double(), want)
    assert torch.equal(tc.mlp_layer_dot(l1, l2, xd), got)
    if N % 8 == 0 and N >= 256:       # the same columns as one group of half the width twice over
        half = tc.mlp_layer_dot(l1, l2, xd[:, :, : N // 2].contiguous())
        assert torch.equal(half, got[:, :, : N // 2])


@pytest.mark.parametrize("G,N,C,M,relu", [(4, 512, 128, 128, False), (3, 100, 32, 46, True), (2, 1000, 256, 200, True),
                                          (1, 128, 64, 128, True)])
def test_point_major_rows_layer(cuda, G, N, C, M, relu):
    """tc.mlp_rows: x (G, N, C) point-major -> (G, N, M) (one-tap convolution mode of the layer kernel) against fp64."""
    from jmodt_b200 import tc
    g = torch.Generator(device="cpu").manual_seed(C * 11 + N)
    w, b = torch.randn(M, C, generator=g) / C ** 0.5, torch.randn(M, generator=g)
    x = torch.randn(G, N, C, generator=g)
    layer = tc.PackedLayer(w.to(cuda), b.to(cuda), relu)
    got = tc.mlp_rows(layer, x.to(cuda).contiguous())
    want = torch.einsum("mk,gnk->gnm", w.double(), x.double()) + b.double()
    if relu:
        want = want.clamp_min(0)
    assert got.shape == (G, N, M)
    _check(got.cpu().double(), want)
    # the same numbers as the channel-first launch with a point-major epilogue
    ref = tc.mlp_layer(layer, x.to(cuda).transpose(1, 2).contiguous(), point_major_out=True)
    assert torch.equal(got, ref)
