"""Host-side logic that needs no GPU: weight folding / packing (the contract between tc.py and the tensor-core
kernels), the sync-free box decode against the reference golden vectors, frame sharding helpers."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def test_fold_conv_bn_equals_eval_mode_conv_bn():
    """tc.fold_conv_bn (eval-mode BatchNorm folded into the 1x1 conv, pytorch_utils.py:36-102) vs torch."""
    from jmodt_b200 import tc
    torch.manual_seed(0)
    conv = torch.nn.Conv1d(37, 50, 1, bias=True)
    bn = torch.nn.BatchNorm1d(50)
    with torch.no_grad():
        bn.running_mean.normal_(0, 1); bn.running_var.uniform_(0.5, 2.0)
        bn.weight.normal_(1, 0.2); bn.bias.normal_(0, 0.3)
    bn.eval()
    x = torch.randn(3, 37, 11)
    w, b = tc.fold_conv_bn(conv, bn)
    got = torch.einsum("mk,bkn->bmn", w, x) + b[None, :, None]
    with torch.no_grad():
        want = bn(conv(x))
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)
    w2, b2 = tc.fold_conv_bn(torch.nn.Conv1d(8, 4, 1, bias=False))
    assert w2.shape == (4, 8) and not b2.any()


def test_packed_layer_image_layout_and_hi_lo_split():
    """PackedLayer stores W as bf16 hi/lo chunk images in the UMMA no-swizzle K-major core-matrix layout
    (DESIGN.md §3): element (m, k) of a 128-row x 32-column chunk sits at byte offset
    (k/8)*2048 + (m/8)*128 + (m%8)*16 + (k%8)*2 of its 8 KB image, hi image first, then lo; hi + lo ~= w to 2^-16."""
    from jmodt_b200 import tc
    torch.manual_seed(1)
    M, K = 200, 70
    w = torch.randn(M, K)
    layer = tc.PackedLayer(w, torch.arange(M, dtype=torch.float32), relu=True)
    Mt, Kc = 2, 3
    assert layer.M == M and layer.K == K and layer.relu
    assert tuple(layer.wpack.shape) == (Mt, Kc, 2, 4, 16, 8, 8) and layer.wpack.dtype == torch.bfloat16
    assert layer.bias.shape == (Mt * 128,) and torch.equal(layer.bias[:M], torch.arange(M, dtype=torch.float32))
    assert not layer.bias[M:].any()
    flat = layer.wpack.reshape(Mt, Kc, 2, -1)                 # 4096 bf16 elements per image
    hi = torch.zeros(Mt * 128, Kc * 32)
    lo = torch.zeros(Mt * 128, Kc * 32)
    for m in range(Mt * 128):
        for k in range(0, Kc * 32, 7):                          # sample columns (incl. the zero padding)
            mt, mr, kc, kr = m // 128, m % 128, k // 32, k % 32
            off = ((kr // 8) * 2048 + (mr // 8) * 128 + (mr % 8) * 16 + (kr % 8) * 2) // 2
            hi[m, k] = flat[mt, kc, 0, off].float()
            lo[m, k] = flat[mt, kc, 1, off].float()
    wp = torch.zeros(Mt * 128, Kc * 32)
    wp[:M, :K] = w
    cols = list(range(0, Kc * 32, 7))
    assert torch.equal(hi[:, cols], wp[:, cols].to(torch.bfloat16).float())
    err = (hi[:, cols] + lo[:, cols] - wp[:, cols]).abs().max().item()
    assert err <= 2.0 ** -16 * wp.abs().max().item()
    # the fused set-abstraction kernel wants the input columns as [channels, xyz]
    perm = layer.repacked_xyz_last()
    flat_p = perm.wpack.reshape(Mt, Kc, 2, -1)
    k_new, k_old = 0, 3                                          # first channel column moved to the front
    off = ((k_new % 32 // 8) * 2048 + (5 // 8) * 128 + (5 % 8) * 16 + (k_new % 8) * 2) // 2
    assert flat_p[0, 0, 0, off].float() == w[5, k_old].to(torch.bfloat16).float()


def test_decode_bbox_target_matches_the_reference_golden_on_cpu():
    """The sync-free rewrite of decode_bbox_target (torch.where instead of masked assignment, column adds instead
    of list indexing) against outputs of the reference function (tests/golden/make_golden_modules.py)."""
    from jmodt_b200.detector import RpnConfig, decode_bbox_target
    G = np.load(os.path.join(HERE, "golden", "ref_modules.npz"))
    cfg = RpnConfig()
    out = decode_bbox_target(torch.from_numpy(G["dec_xyz"]), torch.from_numpy(G["dec_reg"]), cfg.loc_scope,
                             cfg.loc_bin_size, cfg.num_head_bin, torch.tensor(cfg.mean_size, dtype=torch.float32))
    np.testing.assert_allclose(out.numpy(), G["dec_out"], atol=2e-5, rtol=1e-5)
    assert float(out[:, 6].max()) <= np.pi + 1e-6 and float(out[:, 6].min()) >= -np.pi - 1e-6


def test_rcnn_decode_with_rotated_rois_and_fine_heading_matches_the_reference_golden():
    """decode_bbox_target as the evaluation loop calls it on the RCNN head (tools/eval.py:109-116): 7-column RoIs,
    get_ry_fine=True — against outputs of the reference function on the same inputs."""
    from jmodt_b200.detector import decode_bbox_target
    from jmodt_b200.head import HeadConfig
    G = np.load(os.path.join(HERE, "golden", "ref_modules.npz"))
    cfg = HeadConfig()
    out = decode_bbox_target(torch.from_numpy(G["dec7_rois"]).clone(), torch.from_numpy(G["dec7_reg"]), cfg.loc_scope,
                             cfg.loc_bin_size, cfg.num_head_bin,
                             torch.tensor((1.52563191462, 1.62856739989, 3.88311640418), dtype=torch.float32),
                             get_ry_fine=True)
    np.testing.assert_allclose(out.numpy(), G["dec7_out"], atol=2e-5, rtol=1e-5)


def test_sa_fused_supported_shapes():
    from jmodt_b200 import tc
    mk = lambda dims: [tc.PackedLayer(torch.zeros(dims[i + 1], dims[i]), None, True) for i in range(3)]
    assert tc.sa_fused_supported(mk([131, 128, 128, 128]), 128, 128, 64)          # RCNN SA0
    assert tc.sa_fused_supported(mk([131, 128, 128, 256]), 128, 32, 64)           # RCNN SA1
    assert tc.sa_fused_supported(mk([3, 16, 16, 32]), 0, 4096, 16)                # RPN level 0
    assert tc.sa_fused_supported(mk([99, 64, 96, 128]), 96, 1024, 32)             # RPN level 1
    assert not tc.sa_fused_supported(mk([259, 128, 196, 256]), 256, 256, 16)      # RPN level 2: too wide
    assert not tc.sa_fused_supported(mk([131, 128, 128, 128]), 128, 3, 8)         # npoint * nsample not a multiple of 128
    assert not tc.sa_fused_supported(mk([131, 128, 128, 128])[:2], 128, 128, 64)


def test_tracking_oracle_corners_and_distance_properties():
    """oracle/tracking_ref.py (restatement of kitti_utils.py:107-133 and data_association.py:10-28): corners against an
    independent numpy construction; the distance score is 1 on the diagonal, symmetric, and < 1 elsewhere."""
    from oracle import tracking_ref
    g = torch.Generator().manual_seed(0)
    n = 9
    boxes = torch.cat([torch.rand(n, 3, generator=g) * 20, 1 + torch.rand(n, 3, generator=g) * 3,
                       (torch.rand(n, 1, generator=g) * 2 - 1) * np.pi], dim=1)
    corners = tracking_ref.boxes3d_to_corners3d(boxes).numpy()
    for i in range(n):
        x, y, z, h, w, l, ry = boxes[i].double().numpy()
        xs = np.array([l, l, -l, -l, l, l, -l, -l]) / 2
        ys = np.array([0, 0, 0, 0, -h, -h, -h, -h])
        zs = np.array([w, -w, -w, w, w, -w, -w, w]) / 2
        want = np.stack([np.cos(ry) * xs + np.sin(ry) * zs + x, ys + y, -np.sin(ry) * xs + np.cos(ry) * zs + z], axis=1)
        np.testing.assert_allclose(corners[i], want, atol=1e-5)
    d = tracking_ref.boxes_dist(boxes, boxes)
    assert torch.allclose(torch.diagonal(d), torch.ones(n)) and torch.allclose(d, d.t(), atol=1e-6)
    assert float((d - torch.eye(n)).max()) < 1.0
