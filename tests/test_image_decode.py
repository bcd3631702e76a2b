"""Image decoder of the LI-Fusion backbone (reference backbone.py:187-196) evaluated at the sampled pixels only
(csrc/image_decode.cu) against the reference's dense formulation: ConvTranspose2d x4 -> cat -> 1x1 conv -> BatchNorm ->
ReLU -> grid_sample, run by torch on the CPU in fp32."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def _net(seed=0):
    from jmodt_b200.detector import PointNet2MSG, RpnConfig
    from jmodt_b200.synth import fill_deterministic
    torch.manual_seed(seed)
    net = fill_deterministic(PointNet2MSG(input_channels=0, cfg=RpnConfig())).eval()
    with torch.no_grad():      # non-trivial BatchNorm statistics and decoder biases
        bn = net.image_fusion_bn
        bn.running_mean.copy_(torch.linspace(-0.2, 0.3, 32))
        bn.running_var.copy_(torch.linspace(0.5, 1.5, 32))
        for dc in net.DeConv:
            dc.bias.copy_(torch.linspace(-0.1, 0.1, 16))
    return net


def _maps(B, H, W, seed=1):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, c, H >> (l + 1), W >> (l + 1), generator=g) for l, c in enumerate((64, 128, 256, 512))]


def _xy(B, N, seed=2):
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand(B, N, 2, generator=g) * 2.2 - 1.1            # some points fall outside the image
    xy[:, 0] = torch.tensor([-1.0, -1.0])                        # exact corners and the centre
    xy[:, 1] = torch.tensor([1.0, 1.0])
    xy[:, 2] = torch.tensor([0.0, 0.0])
    xy[:, 3] = torch.tensor([1.0, -1.0])
    xy[:, 4] = torch.tensor([-3.0, 5.0])                         # far outside: no tap
    return xy


def _dense(net, maps, xy):
    """The reference's dense formulation, from the oracle (functional ops on the module's parameters)."""
    from oracle import modules_ref
    with torch.no_grad():
        de = torch.cat([dc(m) for dc, m in zip(net.DeConv, maps)], dim=1)
        fused = F.relu(net.image_fusion_bn(net.image_fusion_conv(de)))
        want = modules_ref.image_decoder_at_points(net, maps, xy)
        assert torch.allclose(want, F.grid_sample(fused, xy.unsqueeze(1), mode="bilinear", padding_mode="zeros",
                                                  align_corners=True).squeeze(2), rtol=1e-5, atol=1e-6)
        return want, fused


def test_decoder_pack_reproduces_dense_pixels_on_cpu():
    """The packed phase weights + folded 1x1 / BatchNorm / biases, applied per pixel in torch, equal the dense map."""
    net = _net()
    maps = _maps(1, 32, 64)
    _, fused = _dense(net, maps, torch.zeros(1, 1, 2))
    wexp, w1, b1 = net._decoder_pack()
    assert wexp.shape == (256, 30, 2, 16, 16) and wexp.dtype == torch.int32 and w1.shape == (32, 64) and b1.shape == (32,)
    wexp = wexp.view(torch.bfloat16).float().sum(dim=2)          # hi + lo planes of bf16 pairs -> (256, 30, 16, 32)
    rng = np.random.default_rng(0)
    for _ in range(40):
        y, x = int(rng.integers(0, 32)), int(rng.integers(0, 64))
        ph = (y % 16) * 16 + (x % 16)
        d, chunk = [], 0
        for l, m in enumerate(maps):
            src = m[0, :, y >> (l + 1), x >> (l + 1)]
            acc = torch.zeros(16)
            for k in range(0, src.numel(), 32):
                acc += wexp[ph, chunk] @ src[k:k + 32]
                chunk += 1
            d.append(acc)
        got = torch.relu(w1 @ torch.cat(d) + b1)
        assert torch.allclose(got, fused[0, :, y, x], rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,H,W", [(2, 1000, 96, 320), (1, 16384, 384, 1280), (3, 7, 32, 64)])
def test_decode_gather_matches_dense_reference(cuda, B, N, H, W):
    net = _net()
    maps, xy = _maps(B, H, W), _xy(B, N)
    want, _ = _dense(net, maps, xy)
    net = net.to(cuda)
    for fmt in (torch.contiguous_format, torch.channels_last):
        got = net.decode_gather([m.to(cuda).contiguous(memory_format=fmt) for m in maps], xy.to(cuda), image_hw=(H, W))
        torch.cuda.synchronize()
        assert got.shape == (B, 32, N)
        err = (got.cpu() - want).abs().max().item()
        assert err <= 1e-4 * max(1.0, want.abs().max().item()), err
    assert torch.count_nonzero(got[:, :, 4]) == 0          # the far-outside point samples nothing


@pytest.mark.gpu
def test_decode_gather_is_deterministic(cuda):
    net = _net().to(cuda)
    maps, xy = [m.to(cuda) for m in _maps(2, 96, 320)], _xy(2, 4096).to(cuda)
    a = net.decode_gather(maps, xy, image_hw=(96, 320))
    b = net.decode_gather(maps, xy, image_hw=(96, 320))
    assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("C,H,W,N", [(64, 192, 640, 4096), (512, 24, 80, 64), (6, 5, 7, 33)])
def test_channels_last_gather_equals_channel_first_bitwise(cuda, C, H, W, N):
    from jmodt_b200.detector import feature_gather
    g = torch.Generator().manual_seed(5)
    fm = torch.randn(2, C, H, W, generator=g).to(cuda)
    xy = _xy(2, N).to(cuda)
    want = feature_gather(fm, xy)
    got = feature_gather(fm.contiguous(memory_format=torch.channels_last), xy)
    assert torch.equal(got, want)
    ref = F.grid_sample(fm, xy.unsqueeze(1), mode="bilinear", padding_mode="zeros", align_corners=True).squeeze(2)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_backbone_with_sparse_decoder_matches_dense_maps(cuda):
    """PointNet2MSG.forward fed (channels-last maps, None) — the decoder evaluated at the points — against the same
    network fed the reference's dense (maps, fused)."""
    from jmodt_b200.synth import make_batch
    net = _net().to(cuda)
    b = make_batch(80, 1)
    xyz, xy, img = (torch.from_numpy(b[k]).to(cuda) for k in ("pts", "pts_xy", "img"))
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False         # the dense side in plain fp32
    try:
        with torch.no_grad():
            dense = net.image_features(img)
            sparse = net.image_features(img, dense=False)
            assert sparse[1] is None and all(m.is_contiguous(memory_format=torch.channels_last) for m in sparse[0])
            want_xyz, want = net(xyz, None, xy, image_maps=dense)
            got_xyz, got = net(xyz, None, xy, image_maps=(dense[0], None))      # same maps: the decoder alone differs
            cl_xyz, cl = net(xyz, None, xy, image_maps=sparse)                  # channels-last maps end to end
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert torch.equal(got_xyz, want_xyz) and torch.equal(cl_xyz, want_xyz)
    scale = want.abs().max().item()
    assert (got - want).abs().max().item() <= 1e-4 * scale
    assert (cl - want).abs().max().item() <= 1e-4 * scale


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout,B,H,W", [(3, 64, 2, 32, 64), (64, 128, 1, 24, 40), (128, 128, 2, 17, 23),
                                             (256, 512, 1, 12, 20), (3, 64, 1, 384, 1280)])
def test_basic_block_on_tensor_cores_matches_torch_fp32(cuda, cin, cout, B, H, W):
    """BasicBlock (backbone.py:15-30: conv3x3 -> BN -> ReLU -> conv3x3 stride 2) through the tcgen05 implicit-GEMM
    convolution against torch on the CPU in fp32; odd sizes exercise the zero padding and the ragged last tile."""
    from jmodt_b200.detector import BasicBlock
    from jmodt_b200.synth import fill_deterministic
    torch.manual_seed(1)
    blk = fill_deterministic(BasicBlock(cin, cout)).eval()
    with torch.no_grad():
        blk.bn1.running_mean.copy_(torch.linspace(-0.3, 0.2, cout))
        blk.bn1.running_var.copy_(torch.linspace(0.6, 1.7, cout))
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, cin, H, W, generator=g)
    from oracle import modules_ref
    with torch.no_grad():
        want = modules_ref.basic_block(blk, x)
        assert torch.allclose(want, blk(x), rtol=1e-5, atol=1e-6)         # the CPU module forward is the same composition
    blk = blk.to(cuda)
    with torch.no_grad():
        got = blk(x.to(cuda))
        got_cl = blk(x.to(cuda).contiguous(memory_format=torch.channels_last))
    assert got.shape == want.shape and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got, got_cl)
    scale = want.abs().max().item()
    assert (got.cpu() - want).abs().max().item() <= 1e-4 * scale
