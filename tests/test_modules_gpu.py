"""This package's sm_100a path against outputs of the REAL reference modules (tests/golden/ref_modules.npz,
generated on CPU by tests/golden/make_golden_modules.py from /root/reference with name-hashed weights), plus
fused-vs-unfused checks of the set-abstraction / feature-propagation modules.  fp32 tolerance from north_star:
1e-4 relative (normwise)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

G = np.load(os.path.join(ROOT, "tests", "golden", "ref_modules.npz"))


def _rel(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.abs(got - want).max() / (np.abs(want).max() + 1e-12)


def test_state_dict_keys_equal_the_reference():
    """Not a GPU test: the drop-in contract on parameter names."""
    from jmodt_b200.detector import PointNet2MSG
    from jmodt_b200.head import RCNN
    assert list(PointNet2MSG(input_channels=0).state_dict().keys()) == json.loads(str(G["keys_backbone"]))
    assert list(RCNN().state_dict().keys()) == json.loads(str(G["keys_rcnn"]))


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


@pytest.mark.gpu
def test_attention_fusion_vs_reference(cuda):
    from jmodt_b200.detector import AttentionFusion
    from jmodt_b200.synth import fill_deterministic
    af = fill_deterministic(AttentionFusion(64, 96, 96)).to(cuda).eval()
    out = af(torch.from_numpy(G["af_point"]).to(cuda), torch.from_numpy(G["af_img"]).to(cuda))
    assert _rel(out.cpu().numpy(), G["af_out"]) < 1e-4


@pytest.mark.gpu
def test_feature_gather_vs_reference(cuda):
    from jmodt_b200.detector import feature_gather
    out = feature_gather(torch.from_numpy(G["fg_map"]).to(cuda), torch.from_numpy(G["fg_xy"]).to(cuda))
    np.testing.assert_allclose(out.cpu().numpy(), G["fg_out"], atol=2e-6, rtol=1e-5)
    assert (G["fg_out"] == 0).any() and (G["fg_out"] != 0).any()      # covers out-of-image taps


@pytest.mark.gpu
def test_decode_bbox_target_vs_reference(cuda):
    from jmodt_b200.detector import RpnConfig, decode_bbox_target
    cfg = RpnConfig()
    out = decode_bbox_target(torch.from_numpy(G["dec_xyz"]).to(cuda), torch.from_numpy(G["dec_reg"]).to(cuda),
                             cfg.loc_scope, cfg.loc_bin_size, cfg.num_head_bin,
                             torch.tensor(cfg.mean_size, dtype=torch.float32, device=cuda))
    np.testing.assert_allclose(out.cpu().numpy(), G["dec_out"], atol=2e-5, rtol=1e-5)


@pytest.mark.gpu
def test_rcnn_dense_layers_and_affinity_vs_reference(cuda):
    from jmodt_b200.head import RCNN, affinity, run_stack
    from jmodt_b200.synth import fill_deterministic
    rcnn = fill_deterministic(RCNN()).to(cuda).eval()
    P = rcnn.pack()
    pts = torch.from_numpy(G["rcnn_pts_input"]).to(cuda)
    xyz_feature = run_stack(P["xyz_up"], pts[..., 0:5].transpose(1, 2).contiguous())
    merged = run_stack(P["merge_down"], torch.cat((xyz_feature, pts[..., 5:].transpose(1, 2)), 1).contiguous())
    assert _rel(merged.cpu().numpy(), G["rcnn_merged"]) < 1e-4
    feat_t = torch.from_numpy(G["rcnn_feat"]).to(cuda).squeeze(-1).t().contiguous().unsqueeze(0)
    assert _rel(run_stack(P["cls"], feat_t)[0].t().cpu().numpy(), G["rcnn_cls"]) < 1e-4
    assert _rel(run_stack(P["reg"], feat_t)[0].t().cpu().numpy(), G["rcnn_reg"]) < 1e-4
    link, start, end, logits = affinity(rcnn, torch.from_numpy(G["aff_pred"]).to(cuda),
                                        torch.from_numpy(G["aff_det"]).to(cuda))
    assert _rel(logits.cpu().numpy(), G["aff_logits"]) < 1e-4
    assert _rel(link.cpu().numpy(), G["aff_link"]) < 1e-4
    assert _rel(start.cpu().numpy(), G["aff_start"]) < 1e-4 and _rel(end.cpu().numpy(), G["aff_end"]) < 1e-4


@pytest.mark.gpu
def test_fp_shared_mlp_with_bn_vs_reference(cuda):
    from jmodt_b200.detector import PointNet2MSG
    from jmodt_b200.pointnet2.pointnet2_modules import pack_shared_mlp
    from jmodt_b200.synth import fill_deterministic
    from jmodt_b200 import tc
    fp = fill_deterministic(PointNet2MSG(input_channels=0).FP_modules[0]).to(cuda).eval()
    h = torch.from_numpy(G["fp_in"]).to(cuda)
    for layer in pack_shared_mlp(fp.mlp):
        h = tc.mlp_layer(layer, h)
    assert _rel(h.cpu().numpy(), G["fp_out"]) < 1e-4


@pytest.mark.gpu
def test_sa_msg_and_fp_fused_equal_unfused_composition(cuda):
    """Fused inference path (grouping in the first layer's operand staging, max-pool in the last epilogue) vs the
    reference composition ball_query -> group -> SharedMLP(cuDNN fp32) -> max_pool of the same module."""
    from jmodt_b200.pointnet2.pointnet2_modules import PointnetFPModule, PointnetSAModuleMSG
    from jmodt_b200.synth import fill_deterministic, make_batch
    pts = torch.from_numpy(make_batch(11, 2, with_image=False)["pts"]).to(cuda)
    feats = torch.randn(2, 96, 16384, device=cuda)
    sa = fill_deterministic(PointnetSAModuleMSG(npoint=1024, radii=[0.5, 1.0], nsamples=[16, 32],
                                                mlps=[[96, 64, 64, 128], [96, 64, 96, 128]], bn=True)).to(cuda).eval()
    with torch.no_grad():
        new_xyz, fused, idx = sa(pts, feats)
        sa.fused = False
        new_xyz2, unfused, idx2 = sa(pts, feats)
    assert torch.equal(idx, idx2) and torch.equal(new_xyz, new_xyz2) and fused.shape == (2, 256, 1024)
    assert _rel(fused.cpu().numpy(), unfused.cpu().numpy()) < 1e-4
    sa0 = fill_deterministic(PointnetSAModuleMSG(npoint=4096, radii=[0.1, 0.5], nsamples=[16, 32],
                                                 mlps=[[0, 16, 16, 32], [0, 32, 32, 64]], bn=True)).to(cuda).eval()
    with torch.no_grad():
        _, f0, _ = sa0(pts, None)
        sa0.fused = False
        _, u0, _ = sa0(pts, None)
    assert _rel(f0.cpu().numpy(), u0.cpu().numpy()) < 1e-4
    fp = fill_deterministic(PointnetFPModule(mlp=[256 + 96, 128, 128])).to(cuda).eval()
    with torch.no_grad():
        a = fp(pts, new_xyz, feats, fused)
        fp.fused = False
        b = fp(pts, new_xyz, feats, fused)
    assert a.shape == (2, 128, 16384) and _rel(a.cpu().numpy(), b.cpu().numpy()) < 1e-4


@pytest.mark.gpu
def test_point_rcnn_end_to_end_runs_and_is_consistent(cuda):
    """Whole pipeline on 2 synthetic frames: shapes of the reference output dict (point_rcnn.py:23-72), the proposal
    layer's zero-padding contract, and RCNN outputs equal to running the head on the same rois again."""
    from jmodt_b200.detector import PointRCNN, RpnConfig
    from jmodt_b200.synth import fill_deterministic, make_batch
    cfg = RpnConfig(post_nms_top_n=128)
    model = fill_deterministic(PointRCNN(rpn_cfg=cfg)).to(cuda).eval()
    b = make_batch(40, 2)
    inp = {"pts_input": torch.from_numpy(b["pts"]).to(cuda), "img": torch.from_numpy(b["img"]).to(cuda),
           "pts_xy": torch.from_numpy(b["pts_xy"]).to(cuda)}
    out = model(inp)
    assert out["rpn_cls"].shape == (2, 16384, 1) and out["rpn_reg"].shape == (2, 16384, 76)
    assert out["backbone_features"].shape == (2, 128, 16384) and out["rois"].shape == (2, 128, 7)
    assert out["rcnn_cls"].shape == (256, 1) and out["rcnn_reg"].shape == (256, 46) and out["rcnn_feat"].shape == (256, 512, 1)
    for k in ("rpn_cls", "rpn_reg", "rcnn_cls", "rcnn_reg", "rcnn_feat"):
        assert torch.isfinite(out[k]).all(), k
    out2 = model(inp, rois=out["rois"])
    assert torch.equal(out2["rcnn_feat"], out["rcnn_feat"])
    res = model.pair_affinity(out["rcnn_feat"], 128)
    assert len(res) == 1 and res[0][0].shape == (128, 128) and res[0][1].shape == (128,)
    assert torch.allclose(res[0][0].sum(), torch.tensor(128.0, device=cuda), atol=1e-2)   # (rowsoftmax+colsoftmax)/2


@pytest.mark.gpu
def test_rpn_backbone_vs_cpu_oracle(cuda, cref):
    """Whole RPN point path (4x SA-MSG + LI-Fusion + 4x FP + final fusion) on the sm_100a kernels against the CPU
    restatement (C-oracle index ops + torch-CPU layers) on a reduced cloud (2048 points, scaled-down sampling)."""
    from jmodt_b200.detector import PointNet2MSG, RpnConfig
    from jmodt_b200.synth import fill_deterministic, make_batch
    from oracle import modules_ref
    cfg = RpnConfig(sa_npoints=[512, 128, 32, 16])
    net = fill_deterministic(PointNet2MSG(input_channels=0, cfg=cfg)).eval()
    b = make_batch(60, 1, n_points=2048, n_rois=16, empty_rois=0)
    xyz, xy, img = torch.from_numpy(b["pts"]), torch.from_numpy(b["pts_xy"]), torch.from_numpy(b["img"])[:, :, :96, :320].contiguous()
    with torch.no_grad():
        maps_cpu = net.image_features(img)
        want_xyz, want = modules_ref.backbone_forward(net, xyz, xy, maps_cpu, cref)
    net_gpu = net.to(cuda)
    maps_gpu = ([m.to(cuda) for m in maps_cpu[0]], maps_cpu[1].to(cuda))     # same image maps on both sides
    got_xyz, got = net_gpu(xyz.to(cuda), None, xy.to(cuda), image_maps=maps_gpu)
    assert torch.equal(got_xyz.cpu(), want_xyz) and got.shape == (1, 128, 2048)
    assert _rel(got.cpu().numpy(), want.numpy()) < 1e-4


@pytest.mark.gpu
def test_batched_proposal_layer_equals_reference_loop(cuda):
    """csrc/proposal.cu (3 kernels, no host sync, quota-bounded greedy NMS) vs the reference's per-frame / per-bin
    procedure (proposal_layer.py:36-121) executed with this package's single-set mask + sweep NMS: identical boxes and
    scores — including heavily overlapping proposals (deep scans, quota not reached) and rotated NMS."""
    from jmodt_b200.detector import ProposalLayer, RpnConfig
    from jmodt_b200.synth import make_batch
    g = torch.Generator().manual_seed(9)
    for case, post in [("both bins", 128), ("far bin empty", 100), ("near bin empty", 64), ("dense overlaps", 128),
                       ("dense overlaps rotated", 128), ("everything suppressed", 512)]:
        cfg = RpnConfig(post_nms_top_n=post)
        if case.startswith("dense"):
            cfg.nms_thresh = 0.3
            cfg.nms_type = "rotate" if "rotated" in case else "normal"
        if case == "everything suppressed":
            cfg.nms_thresh = 0.01
        B, N = 3, 16384
        xyz = torch.from_numpy(make_batch(70, B, with_image=False)["pts"]).to(cuda)
        if case == "far bin empty":
            xyz[..., 2] = xyz[..., 2].clamp(max=35.0)
        if case == "near bin empty":
            xyz[..., 2] = xyz[..., 2] * 0.3 + 45.0
        reg = (torch.randn(B, N, 76, generator=g) * (0.02 if ("dense" in case or "suppressed" in case) else 0.5)).to(cuda)
        scores = torch.randn(B, N, generator=g).to(cuda)
        layer = ProposalLayer(cfg=cfg)
        boxes_b, scores_b = layer(scores, reg, xyz)
        layer.batched = False
        boxes_l, scores_l = layer(scores, reg, xyz)
        assert boxes_b.shape == (B, post, 7)
        assert torch.equal(scores_b, scores_l), case
        assert torch.equal(boxes_b, boxes_l), case
        assert (scores_b != 0).any()


@pytest.mark.gpu
def test_geometry_overlap_and_chunked_level0_are_bit_identical_to_serial(cuda):
    """The side-stream coordinate chain and the gated, chunked level-0 set abstraction (detector.PointNet2MSG._geometry)
    reorder launches only: features and coordinates must equal the single-stream run bit for bit, repeatedly (the index
    buffer is recycled between calls, which is what the `filled` event protects)."""
    from jmodt_b200.detector import PointNet2MSG, RpnConfig
    from jmodt_b200.synth import fill_deterministic, make_batch
    net = fill_deterministic(PointNet2MSG(input_channels=0, cfg=RpnConfig())).to(cuda).eval()
    b = make_batch(70, 2)
    xyz, xy, img = (torch.from_numpy(b[k]).to(cuda) for k in ("pts", "pts_xy", "img"))
    maps = net.image_features(img)
    net.overlap_geometry = False
    want_xyz, want = net(xyz, None, xy, image_maps=maps)
    net.overlap_geometry = True
    for chunks in (4, 1, 8, 4):
        net.l0_chunks = chunks
        got_xyz, got = net(xyz, None, xy, image_maps=maps)
        torch.cuda.synchronize()
        assert torch.equal(got_xyz, want_xyz) and torch.equal(got, want), chunks


@pytest.mark.gpu
def test_wait_indices_gate_orders_a_consumer_behind_a_running_sampler(cuda):
    """jmb_wait_indices: a copy queued behind the gate on stream B sees the complete prefix that FPS on stream A has
    produced, for every prefix length."""
    from jmodt_b200.pointnet2 import pointnet2_cuda, pointnet2_utils
    from jmodt_b200.synth import make_batch
    xyz = torch.from_numpy(make_batch(80, 4)["pts"]).to(cuda)
    want = pointnet2_utils.farthest_point_sample(xyz, 4096)
    a, bstream = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for k1 in (64, 1024, 4096):
        with torch.cuda.stream(a):
            idx = torch.full((4, 4096), -1, dtype=torch.int32, device=cuda)
            filled = torch.cuda.Event()
            filled.record(a)
            pointnet2_cuda.farthest_point_sampling_wrapper(4, 16384, 4096, xyz, None, idx)
        with torch.cuda.stream(bstream):
            bstream.wait_event(filled)
            pointnet2_cuda.wait_indices(idx, 0, k1)
            seen = idx[:, :k1].clone()
        torch.cuda.synchronize()
        assert torch.equal(seen, want[:, :k1]), k1


@pytest.mark.gpu
def test_fp_module_interpolating_after_the_first_layer_is_equivalent(cuda):
    """Without skip features relu(W . interp(f) + b) = relu(interp(W . f + b)) (the weights sum to one): the fused FP
    path that runs the first layer on the known points must match the reference order to fp32 rounding."""
    from jmodt_b200.pointnet2.pointnet2_modules import PointnetFPModule
    from jmodt_b200.synth import fill_deterministic
    g = torch.Generator().manual_seed(21)
    fp = fill_deterministic(PointnetFPModule(mlp=[256, 128, 128])).to(cuda).eval()
    unknown = torch.rand(2, 4096, 3, generator=g).to(cuda)
    known = unknown[:, :1024].contiguous()
    feats = torch.randn(2, 256, 1024, generator=g).to(cuda)
    with torch.no_grad():
        fp.interp_after_first_layer = True
        a = fp(unknown, known, None, feats)
        fp.interp_after_first_layer = False
        b = fp(unknown, known, None, feats)
        fp.fused = False
        want = fp(unknown, known, None, feats)          # torch layers (cuDNN fp32), reference op order
    assert a.shape == b.shape == want.shape == (2, 128, 4096)
    assert _rel(a.cpu().numpy(), want.cpu().numpy()) < 1e-4 and _rel(b.cpu().numpy(), want.cpu().numpy()) < 1e-4
    assert _rel(a.cpu().numpy(), b.cpu().numpy()) < 2e-5
    # with skip features: relu(interp(W_a . f) + W_b . skip + b)
    fp2 = fill_deterministic(PointnetFPModule(mlp=[256 + 96, 256, 256])).to(cuda).eval()
    skip = torch.randn(2, 96, 4096, generator=g).to(cuda)
    with torch.no_grad():
        a2 = fp2(unknown, known, skip, feats)
        fp2.interp_after_first_layer = False
        b2 = fp2(unknown, known, skip, feats)
        fp2.fused = False
        want2 = fp2(unknown, known, skip, feats)
    assert a2.shape == want2.shape == (2, 256, 4096)
    assert _rel(a2.cpu().numpy(), want2.cpu().numpy()) < 1e-4 and _rel(b2.cpu().numpy(), want2.cpu().numpy()) < 1e-4
    assert _rel(a2.cpu().numpy(), b2.cpu().numpy()) < 2e-5


@pytest.mark.gpu
def test_wide_sa_level_first_layer_applied_before_the_gather(cuda):
    """RPN level 2 / 3 shapes (widths beyond the one-kernel path): the first SharedMLP layer runs over the POINTS and
    csrc/interpolate.cu:sa_first_layer_kernel finishes it per grouped neighbour — vs the reference composition
    ball_query -> group -> SharedMLP (cuDNN fp32) -> max_pool of the same module."""
    from jmodt_b200.pointnet2.pointnet2_modules import PointnetSAModuleMSG
    from jmodt_b200.synth import fill_deterministic, make_batch
    pts = torch.from_numpy(make_batch(21, 2, with_image=False)["pts"][:, :1024]).to(cuda).contiguous()
    for c_in, npoint, mlps in ((256, 256, [[256, 128, 196, 256], [256, 128, 196, 256]]),
                               (512, 64, [[512, 256, 256, 512], [512, 256, 384, 512]])):
        feats = torch.randn(2, c_in, 1024, device=cuda)
        sa = fill_deterministic(PointnetSAModuleMSG(npoint=npoint, radii=[1.0, 2.0], nsamples=[16, 32], mlps=mlps,
                                                    bn=True)).to(cuda).eval()
        with torch.no_grad():
            new_xyz, fused, idx = sa(pts, feats)
            sa.fused = False
            new_xyz2, unfused, idx2 = sa(pts, feats)
        assert torch.equal(idx, idx2) and fused.shape == unfused.shape
        assert _rel(fused.cpu().numpy(), unfused.cpu().numpy()) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("ic,pc,B,N", [(64, 96, 8, 4096), (128, 256, 8, 4100), (32, 128, 1, 32768), (32, 40, 13, 2731)])
def test_ia_attention_kernel_matches_the_reference_layer(cuda, ic, pc, B, N):
    """IALayer (backbone.py:33-58) through jmb_ia_attention (thread per point, all reduced channels in registers) against
    the three Linear layers + tanh + sigmoid in torch on the CPU; ragged N and rc = 10 ... 64."""
    from jmodt_b200.detector import IALayer
    from jmodt_b200.synth import fill_deterministic
    torch.manual_seed(3)
    layer = fill_deterministic(IALayer([ic, pc])).eval()
    g = torch.Generator().manual_seed(4)
    img, pts = torch.randn(B, ic, N, generator=g), torch.randn(B, pc, N, generator=g)
    with torch.no_grad():
        ri = layer.fc1(img.transpose(1, 2).reshape(-1, ic))
        rp = layer.fc2(pts.transpose(1, 2).reshape(-1, pc))
        att = torch.sigmoid(layer.fc3(torch.tanh(ri + rp))).view(B, 1, N)
        want = layer.conv1(img) * att
    layer = layer.to(cuda)
    assert B * N >= 32768 and pc // 4 <= 64         # the SIMT kernel's gate in IALayer.forward
    got = layer(img.to(cuda), pts.to(cuda)).cpu()
    assert (got - want).abs().max().item() <= 1e-4 * want.abs().max().item()      # fp32 logits within 1e-4 of the scale
