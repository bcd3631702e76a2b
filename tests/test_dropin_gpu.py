"""The drop-in, end to end: the UNMODIFIED reference python (`jmodt.detection.modeling.point_rcnn.PointRCNN`, its RPN,
backbone, proposal / proposal-target layers, RCNN) imported over `jmodt_b200.dropin.install()`, i.e. running on this
package's operators, modules and tcgen05 layers — compared with this package's own detector and with the outputs of
the reference over ITS OWN kernels (tests/golden/ref_gpu.npz).

The reference python reaches the GPU box as the git-ignored mirror oracle/_ref/py (oracle/refpy.py)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import cases  # noqa: E402

_PATH = os.path.join(HERE, "golden", "ref_gpu.npz")
G = np.load(_PATH) if os.path.exists(_PATH) else None


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


@pytest.fixture(scope="module")
def models(cuda):
    from oracle import refpy
    if not refpy.available():
        pytest.skip("reference python not available (oracle/_ref/py)")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    refpy.import_reference(ops="dropin")
    refpy.set_eval_cfg(post_nms_top_n=128)
    import jmodt.ops.pointnet2.pytorch_utils as ref_pt
    import jmodt_b200.pointnet2.pytorch_utils as our_pt
    assert ref_pt is our_pt                                   # the reference imports THIS package's modules
    from jmodt.detection.modeling.point_rcnn import PointRCNN as RefPointRCNN
    from jmodt_b200.detector import PointRCNN, RpnConfig
    from jmodt_b200.synth import fill_deterministic
    torch.manual_seed(0)
    ref = fill_deterministic(RefPointRCNN(num_classes=2, use_xyz=True, mode="TEST")).to(cuda).eval()
    ours = fill_deterministic(PointRCNN(rpn_cfg=RpnConfig(post_nms_top_n=128))).to(cuda).eval()
    assert list(ref.state_dict().keys()) == list(ours.state_dict().keys())
    pts, pts_xy, img = cases.detector_inputs(1)
    inp = {"pts_input": torch.from_numpy(pts).to(cuda), "pts_xy": torch.from_numpy(pts_xy).to(cuda),
           "img": torch.from_numpy(img).to(cuda)}
    return ref, ours, inp


@pytest.mark.gpu
def test_reference_point_rcnn_over_the_dropin(cuda, models):
    from jmodt_b200 import _lib
    ref, ours, inp = models
    n0 = _lib.launch_count
    with torch.no_grad():
        out_ref = ref(inp)
    assert _lib.launch_count - n0 > 50          # the reference forward ran on this library's kernels
    out = ours(inp)
    for k in ("rpn_cls", "rpn_reg", "backbone_features"):
        assert _rel(out_ref[k].cpu().numpy(), out[k].cpu().numpy()) < 1e-4, k
    assert out_ref["rois"].shape == out["rois"].shape == (1, 128, 7)
    same = (torch.abs(out_ref["rois"] - out["rois"]).amax(-1) < 1e-3).float().mean().item()
    assert same > 0.95, same
    # stage-wise: the reference's RCNN module (rcnn.py:158-202,288-289) and this package's, on the same RoIs
    seg = (torch.sigmoid(out["rpn_cls"][:, :, 0]) > 0.2).float()
    rc_in = {"rpn_xyz": out["backbone_xyz"], "rpn_features": out["backbone_features"].permute(0, 2, 1),
             "seg_mask": seg, "roi_boxes3d": out["rois"], "pts_depth": torch.norm(out["backbone_xyz"], p=2, dim=2)}
    with torch.no_grad():
        r_ref = ref.rcnn_net(dict(rc_in))
    r_our = ours.rcnn_net(dict(rc_in))
    for k in ("rcnn_cls", "rcnn_reg", "rcnn_feat"):
        assert r_ref[k].shape == r_our[k].shape
        assert _rel(r_ref[k].cpu().numpy(), r_our[k].cpu().numpy()) < 1e-4, k
    if G is not None:      # ... and vs the reference over its own CUDA extensions, on the reference's RoIs
        rc_in["roi_boxes3d"] = torch.from_numpy(G["det_rois"]).to(cuda)
        with torch.no_grad():
            r = ref.rcnn_net(dict(rc_in))
        assert _rel(r["rcnn_feat"].cpu().numpy(), G["det_rcnn_feat"]) < 2e-4
        assert _rel(r["rcnn_reg"].cpu().numpy(), G["det_rcnn_reg"]) < 2e-4


@pytest.mark.gpu
def test_tracker_affinity_calls_reach_the_tensor_cores(cuda, models):
    """tracker.py:81-112 executed verbatim on the reference RCNN's link / start-end heads (tools/eval.py:333-336 hands
    exactly these modules to Tracker): every model call is one tcgen05 launch per layer, and the scores equal the
    fused `jmodt_b200.tracking.affinity_scores` path and the torch fp32 restatement."""
    from jmodt_b200 import tc
    from jmodt_b200.tracking import affinity_scores
    from oracle import modules_ref
    ref, ours, _ = models
    link_model, se_model = ref.rcnn_net.link_layer, ref.rcnn_net.se_layer
    g = torch.Generator().manual_seed(8)
    pred_features = torch.rand(37, 512, generator=g).to(cuda)
    det_features = torch.rand(45, 512, generator=g).to(cuda)
    num_pred, num_det = 37, 45
    tc.profiler.reset()
    tc.profiler.enabled = True
    with torch.no_grad():
        cor_feat = torch.abs(pred_features.unsqueeze(1).repeat(1, num_det, 1)
                             - det_features.unsqueeze(0).repeat(num_pred, 1, 1))
        link_scores = link_model(cor_feat.view(num_pred * num_det, -1, 1)).view(num_pred, num_det)
        logits = link_scores
        link_score_pred = torch.softmax(link_scores, dim=1)
        link_score_det = torch.softmax(link_scores, dim=0)
        link_scores = (link_score_pred + link_score_det) / 2
        start_scores = torch.sigmoid(se_model(cor_feat.mean(dim=0).unsqueeze(-1))).flatten()
        end_scores = torch.sigmoid(se_model(cor_feat.mean(dim=1).unsqueeze(-1))).flatten()
    tc.profiler.enabled = False
    launches = [r for r in tc.profiler.records if r.kind == "tc_gemm_kernel"]
    assert len(launches) == 9                      # 3 layers x (link, start, end)
    assert any("N=1665" in r.desc for r in launches)      # the 37 x 45 pairs went in as ONE 1665-column problem
    tc.profiler.reset()
    link2, start2, end2 = affinity_scores(link_model, se_model, pred_features, det_features)
    np.testing.assert_allclose(link_scores.cpu().numpy(), link2.cpu().numpy(), atol=1e-6)
    np.testing.assert_allclose(start_scores.cpu().numpy(), start2.cpu().numpy(), atol=1e-6)
    np.testing.assert_allclose(end_scores.cpu().numpy(), end2.cpu().numpy(), atol=1e-6)
    with torch.no_grad():
        wl, ws, we, wlog = modules_ref.affinity(link_model, se_model, pred_features, det_features)
    assert _rel(logits.cpu().numpy(), wlog.cpu().numpy()) < 1e-4
    np.testing.assert_allclose(link_scores.cpu().numpy(), wl.cpu().numpy(), atol=1e-5)


@pytest.mark.gpu
def test_packed_weights_follow_load_state_dict_and_device_moves(cuda):
    """a weight image cached by the first forward must not survive load_state_dict / in-place updates."""
    from jmodt_b200.pointnet2 import pytorch_utils as pt_utils
    torch.manual_seed(3)
    mlp = pt_utils.SharedMLP([16, 32, 8], bn=True).to(cuda).eval()
    x = torch.randn(2, 16, 64, 4, device=cuda)
    with torch.no_grad():
        y0 = mlp(x)
        with pt_utils.torch_layers():
            assert _rel(y0.cpu().numpy(), mlp(x).cpu().numpy()) < 1e-4
        sd = {k: (v * 1.5 + 0.1 if v.is_floating_point() else v) for k, v in mlp.state_dict().items()}
        mlp.load_state_dict(sd)
        y1 = mlp(x)
        with pt_utils.torch_layers():
            want = mlp(x)
    assert _rel(y1.cpu().numpy(), want.cpu().numpy()) < 1e-4 and _rel(y1.cpu().numpy(), y0.cpu().numpy()) > 1e-2
