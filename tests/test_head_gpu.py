"""Fusion head (per-proposal RCNN network, link / start-end affinity) on tcgen05 vs the torch fp32 restatement
of the reference forward (oracle/modules_ref.py, TF32 disabled).  north_star tolerance: logits within 1e-4
relative (normwise: max abs error / max abs reference)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _rel(got, want):
    return (got.double() - want.double()).abs().max().item() / (want.double().abs().max().item() + 1e-12)


def _frame_inputs(cuda, B, seed):
    from jmodt_b200 import synth
    batch = synth.make_batch(seed, B, with_image=False)
    g = torch.Generator().manual_seed(seed)
    pts = torch.from_numpy(batch["pts"]).to(cuda)
    return {"rpn_xyz": pts, "rpn_features": torch.randn(B, 16384, 128, generator=g).to(cuda),
            "seg_mask": (torch.rand(B, 16384, generator=g) > 0.5).float().to(cuda),
            "pts_depth": torch.norm(pts, p=2, dim=2), "roi_boxes3d": torch.from_numpy(batch["rois"]).to(cuda)}


def test_rcnn_head_matches_torch_reference(cuda):
    from jmodt_b200.head import RCNN
    from jmodt_b200.pointnet2 import pointnet2_utils as pu
    from oracle import modules_ref
    torch.manual_seed(0)
    rcnn = RCNN().to(cuda).eval()
    with torch.no_grad():
        for p in rcnn.parameters():            # biases are zero-initialised in the reference; make them matter
            if p.dim() == 1:
                p.normal_(0, 0.1)
    rcnn.pack()
    inp = _frame_inputs(cuda, 1, 3)
    rois = inp["roi_boxes3d"]
    inp["roi_boxes3d"] = torch.cat((rois[:, :20], rois[:, -4:]), dim=1).contiguous()     # the last rois are empty
    rcnn.fuse_input = False
    pts_input, empty = rcnn.pool_rois(inp)               # the reference's pts_input layout
    assert pts_input.shape == (24, 512, 133)
    assert 0 < int(empty.sum()) < 24
    rcnn.fuse_input = True
    pts_head, empty_h = rcnn.pool_rois(inp)              # head layout: [128 channels | xyz, mask, depth | 0 0 0]
    assert pts_head.shape == (24, 512, 136) and torch.equal(empty, empty_h)
    assert torch.equal(pts_head[..., :128], pts_input[..., 5:]) and torch.equal(pts_head[..., 128:133], pts_input[..., :5])
    assert not pts_head[..., 133:].any()
    with torch.no_grad():
        wcls, wreg, wfeat = modules_ref.rcnn_forward_points(rcnn, pts_input, pu.farthest_point_sample, pu.ball_query)
    for x in (pts_head, pts_input):                      # single-kernel input stage, and the layer-by-layer one
        cls, reg, feat = rcnn.forward_points(x)
        assert cls.shape == (24, 1) and reg.shape == (24, 46) and feat.shape == (24, 512, 1)
        assert _rel(feat, wfeat) < 1e-4, _rel(feat, wfeat)
        assert _rel(cls, wcls) < 1e-4, _rel(cls, wcls)
        assert _rel(reg, wreg) < 1e-4, _rel(reg, wreg)
    cls, reg, feat = rcnn.forward_points(pts_head)
    out = rcnn(inp)
    assert torch.equal(out["rcnn_feat"], feat) and out["pooled_empty_flag"].shape == (1, 24)


def test_rcnn_input_stage_single_kernel_vs_layers(cuda):
    """xyz_up_layer + cat + merge_down_layer (rcnn.py:172-186) as ONE kernel vs torch fp32 on random rows."""
    from jmodt_b200 import tc
    from jmodt_b200.head import RCNN
    torch.manual_seed(2)
    rcnn = RCNN().to(cuda).eval()
    with torch.no_grad():
        for p in rcnn.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.1)
    P = rcnn.pack()
    g = torch.Generator().manual_seed(11)
    rows = torch.zeros(5, 512, 136)
    rows[..., :133] = torch.randn(5, 512, 133, generator=g)
    rows = rows.to(cuda)
    got = tc.rcnn_input_fused(P["xyz_up_w8"], P["xyz_up"][1], P["merge_down"][0], rows)
    from jmodt_b200.pointnet2 import pytorch_utils as pt_utils
    with torch.no_grad(), pt_utils.torch_layers():                                       # torch (cuDNN fp32) forward
        xyz_in = rows[..., 128:133].transpose(1, 2).unsqueeze(3)                         # (G, 5, 512, 1)
        up = rcnn.xyz_up_layer(xyz_in)
        want = rcnn.merge_down_layer(torch.cat((up, rows[..., :128].transpose(1, 2).unsqueeze(3)), dim=1))
    want = want.squeeze(3).transpose(1, 2)
    assert got.shape == (5, 512, 128)
    assert _rel(got, want) < 1e-4, _rel(got, want)


def test_affinity_matches_torch_reference(cuda):
    from jmodt_b200.head import RCNN, affinity
    from oracle import modules_ref
    torch.manual_seed(1)
    rcnn = RCNN().to(cuda).eval()
    with torch.no_grad():
        for p in rcnn.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.1)
    rcnn.pack()
    g = torch.Generator().manual_seed(5)
    for P, D in [(128, 128), (37, 50), (1, 3)]:
        pf = torch.randn(P, 512, generator=g).abs().to(cuda)     # rcnn_feat is post-ReLU / max-pool: non-negative
        df = (pf[torch.randint(0, P, (D,), generator=g)] + 0.1 * torch.randn(D, 512, generator=g).to(cuda)).abs()
        link, start, end, logits = affinity(rcnn, pf, df)
        with torch.no_grad():
            wl, ws, we, wlog = modules_ref.affinity(rcnn.link_layer, rcnn.se_layer, pf, df)
        assert link.shape == (P, D) and start.shape == (D,) and end.shape == (P,)
        assert _rel(logits, wlog) < 1e-4, _rel(logits, wlog)
        assert _rel(link, wl) < 1e-4 and _rel(start, ws) < 1e-4 and _rel(end, we) < 1e-4


def test_batched_affinity_and_pair_corr_match_torch_reference(cuda):
    """affinity_batched (one pair_corr launch + one launch per layer for all frame pairs) vs the torch fp32
    restatement pair by pair; the pair-correlation kernel alone vs torch (|p - d| exact, means to fp32 rounding)."""
    from jmodt_b200.head import RCNN, affinity_batched, pair_corr
    from oracle import modules_ref
    torch.manual_seed(4)
    rcnn = RCNN().to(cuda).eval()
    with torch.no_grad():
        for p in rcnn.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.1)
    rcnn.pack()
    g = torch.Generator().manual_seed(8)
    for G, P, D in [(4, 128, 128), (3, 37, 50), (1, 1, 3), (2, 200, 9)]:
        pf = torch.randn(G, P, 512, generator=g).abs().to(cuda)
        df = torch.randn(G, D, 512, generator=g).abs().to(cuda)
        pt, dt = pf.transpose(1, 2).contiguous(), df.transpose(1, 2).contiguous()
        cor, mean_p, mean_d = pair_corr(pt, dt)
        want = (pt.unsqueeze(3) - dt.unsqueeze(2)).abs()                                   # (G, 512, P, D)
        assert torch.equal(cor.view(G, 512, P, D), want)
        assert torch.allclose(mean_p, want.mean(dim=2), rtol=1e-5, atol=1e-6)
        assert torch.allclose(mean_d, want.mean(dim=3), rtol=1e-5, atol=1e-6)
        link, start, end, logits = affinity_batched(rcnn, pf, df)
        assert link.shape == (G, P, D) and start.shape == (G, D) and end.shape == (G, P)
        for k in range(G):
            with torch.no_grad():
                wl, ws, we, wlog = modules_ref.affinity(rcnn.link_layer, rcnn.se_layer, pf[k], df[k])
            assert _rel(logits[k], wlog) < 1e-4, (G, P, D, _rel(logits[k], wlog))
            assert _rel(link[k], wl) < 1e-4 and _rel(start[k], ws) < 1e-4 and _rel(end[k], we) < 1e-4


def test_pack_point_features_equals_torch_cat(cuda):
    """torch.cat((mask, depth, features.permute(0, 2, 1)), dim=2) as one tiled transpose: bit-identical."""
    from jmodt_b200.head import pack_point_features
    g = torch.Generator().manual_seed(3)
    for B, C, N, E in [(2, 128, 16384, 2), (1, 96, 1000, 2), (3, 7, 33, 1), (1, 128, 64, 0)]:
        feat = torch.randn(B, C, N, generator=g).to(cuda)
        extra = [torch.randn(B, N, 1, generator=g).to(cuda) for _ in range(E)]
        rf = feat.permute(0, 2, 1)                       # what point_rcnn.py:47 hands to the pooling stage
        got = pack_point_features(extra, rf)
        assert got.is_contiguous() and torch.equal(got, torch.cat(extra + [rf], dim=2))
        assert torch.equal(pack_point_features(extra, rf.contiguous()), torch.cat(extra + [rf], dim=2))   # fallback path


def test_state_dict_keys_match_reference_layout(cuda):
    from jmodt_b200.head import RCNN
    keys = set(RCNN().state_dict())
    for k in ["xyz_up_layer.layer0.conv.weight", "xyz_up_layer.layer1.conv.bias", "merge_down_layer.layer0.conv.weight",
              "SA_modules.0.mlps.0.layer0.conv.weight", "SA_modules.2.mlps.0.layer2.conv.bias",
              "cls_layer.0.conv.weight", "cls_layer.2.conv.weight", "cls_layer.3.conv.bias",
              "reg_layer.3.conv.weight", "link_layer.0.conv.weight", "link_layer.3.conv.weight", "se_layer.2.conv.bias"]:
        assert k in keys, k
    assert RCNN().reg_layer[-1].conv.weight.shape == (46, 512, 1)
