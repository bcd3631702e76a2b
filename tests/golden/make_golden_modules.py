#!/usr/bin/env python
"""Generates tests/golden/ref_modules.npz by IMPORTING THE REFERENCE (from /root/reference) in the CPU container
and running its own nn.Modules / functions on seeded inputs with name-hashed weights
(jmodt_b200.synth.fill_deterministic).  The reference cannot travel to the GPU box, so its outputs are committed
as fixtures; tests/test_modules_gpu.py checks this package's sm_100a path against them.

    python tests/golden/make_golden_modules.py        (needs /root/reference; an easydict shim is created on the fly)
"""
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("JMODT_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

SHIM = '''
class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in {**(d or {}), **kw}.items():
            setattr(self, k, v)
    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)
    __setitem__ = __setattr__
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)
'''


def import_reference():
    shim_dir = tempfile.mkdtemp()
    open(os.path.join(shim_dir, "easydict.py"), "w").write(SHIM)
    sys.path.insert(0, shim_dir)
    sys.path.insert(0, REF)
    # the reference's python wrappers import their pybind modules at import time; give them the compiled
    # reference extensions (oracle/_ref) — they are not called on the CPU.
    for pkg, name in [("jmodt.ops.pointnet2", "pointnet2_cuda"), ("jmodt.ops.roipool3d", "roipool3d_cuda"),
                      ("jmodt.ops.iou3d", "iou3d_cuda")]:
        so = os.path.join(ROOT, "oracle", "_ref", name + ".so")
        if os.path.exists(so):
            spec = importlib.util.spec_from_file_location(name, so)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        else:
            mod = types.ModuleType(name)
        sys.modules[pkg + "." + name] = mod


def main(out_path):
    import_reference()
    from jmodt.config import cfg
    from jmodt.detection.modeling import backbone as rb
    from jmodt.detection.modeling.rcnn import RCNN as RefRCNN
    from jmodt.utils import bbox_transform as rbt

    from jmodt_b200.synth import fill_deterministic
    torch.manual_seed(0)
    g = {}
    gen = torch.Generator().manual_seed(77)

    # ---- state_dict key lists (drop-in contract)
    ref_backbone = rb.PointNet2MSG(input_channels=0).eval()
    ref_rcnn = RefRCNN(num_classes=2, input_channels=128, mode="TEST").eval()
    g["keys_backbone"] = np.array(json.dumps(list(ref_backbone.state_dict().keys())))
    g["keys_rcnn"] = np.array(json.dumps(list(ref_rcnn.state_dict().keys())))

    # ---- AttentionFusion / IALayer (backbone.py:33-76)
    af = fill_deterministic(rb.AttentionFusion(64, 96, 96)).eval()
    pf = torch.randn(2, 96, 200, generator=gen)
    im = torch.randn(2, 64, 200, generator=gen)
    with torch.no_grad():
        g["af_point"], g["af_img"], g["af_out"] = pf.numpy(), im.numpy(), af(pf, im).numpy()

    # ---- feature_gather (backbone.py:79-89)
    fmap = torch.randn(2, 8, 12, 20, generator=gen)
    xy = torch.rand(2, 50, 2, generator=gen) * 2.3 - 1.15
    g["fg_map"], g["fg_xy"], g["fg_out"] = fmap.numpy(), xy.numpy(), rb.feature_gather(fmap, xy).numpy()

    # ---- decode_bbox_target as ProposalLayer.forward calls it (proposal_layer.py:24-32)
    xyz = torch.rand(300, 3, generator=gen) * torch.tensor([80.0, 4.0, 70.0]) - torch.tensor([40.0, 1.0, 0.0])
    reg = torch.randn(300, 76, generator=gen)
    mean_size = torch.from_numpy(cfg.CLS_MEAN_SIZE[0])
    orig = torch.Tensor.get_device
    torch.Tensor.get_device = lambda self: "cpu"      # decode_bbox_target does anchor_size.to(x.get_device())
    try:
        dec = rbt.decode_bbox_target(xyz, reg, anchor_size=mean_size, loc_scope=cfg.RPN.LOC_SCOPE,
                                     loc_bin_size=cfg.RPN.LOC_BIN_SIZE, num_head_bin=cfg.RPN.NUM_HEAD_BIN,
                                     get_xz_fine=cfg.RPN.LOC_XZ_FINE, get_y_by_bin=False, get_ry_fine=False)
    finally:
        torch.Tensor.get_device = orig
    g["dec_xyz"], g["dec_reg"], g["dec_out"] = xyz.numpy(), reg.numpy(), dec.numpy()

    # ---- RCNN dense pieces + affinity (rcnn.py:178-196, tracker.py:81-112)
    fill_deterministic(ref_rcnn)
    pts_in = torch.randn(3, 64, 133, generator=gen)
    with torch.no_grad():
        xyz_input = pts_in[..., 0:5].transpose(1, 2).contiguous().unsqueeze(3)
        xyz_feature = ref_rcnn.xyz_up_layer(xyz_input)
        rpn_feature = pts_in[..., 5:].transpose(1, 2).contiguous().unsqueeze(3)
        merged = ref_rcnn.merge_down_layer(torch.cat((xyz_feature, rpn_feature), dim=1)).squeeze(3)
        feat = torch.randn(6, 512, 1, generator=gen).abs()
        g["rcnn_pts_input"], g["rcnn_merged"] = pts_in.numpy(), merged.numpy()
        g["rcnn_feat"], g["rcnn_cls"], g["rcnn_reg"] = feat.numpy(), ref_rcnn.cls_layer(feat).squeeze(-1).numpy(), \
            ref_rcnn.reg_layer(feat).squeeze(-1).numpy()
        pred, det = torch.randn(9, 512, generator=gen).abs(), torch.randn(7, 512, generator=gen).abs()
        num_pred, num_det = 9, 7
        cor_feat = torch.abs(pred.unsqueeze(1).repeat(1, num_det, 1) - det.unsqueeze(0).repeat(num_pred, 1, 1))
        link_scores = ref_rcnn.link_layer(cor_feat.view(num_pred * num_det, -1, 1)).view(num_pred, num_det)
        g["aff_logits"] = link_scores.numpy()
        g["aff_link"] = ((torch.softmax(link_scores, dim=1) + torch.softmax(link_scores, dim=0)) / 2).numpy()
        g["aff_start"] = torch.sigmoid(ref_rcnn.se_layer(cor_feat.mean(dim=0).unsqueeze(-1))).numpy().flatten()
        g["aff_end"] = torch.sigmoid(ref_rcnn.se_layer(cor_feat.mean(dim=1).unsqueeze(-1))).numpy().flatten()
        g["aff_pred"], g["aff_det"] = pred.numpy(), det.numpy()

    # ---- SharedMLP with eval-mode BN + FP-module dense part (pytorch_utils.py:6-33, pointnet2_modules.py:155-164)
    fp = fill_deterministic(ref_backbone.FP_modules[0]).eval()          # mlp [256, 128, 128], bn=True
    x = torch.randn(2, 256, 300, generator=gen)
    with torch.no_grad():
        g["fp_in"], g["fp_out"] = x.numpy(), fp.mlp(x.unsqueeze(-1)).squeeze(-1).numpy()

    # ---- decode_bbox_target as the evaluation loop calls it on the RCNN head (tools/eval.py:109-116): 7-column RoIs
    # (rotated back by the RoI heading) and the fine heading bins.  Drawn LAST so the arrays above keep their values.
    rois7 = torch.cat([torch.rand(200, 3, generator=gen) * torch.tensor([80.0, 4.0, 70.0]) - torch.tensor([40.0, 1.0, 0.0]),
                       torch.tensor([1.5256, 1.6286, 3.8831]) * (0.8 + 0.4 * torch.rand(200, 3, generator=gen)),
                       (torch.rand(200, 1, generator=gen) * 2 - 1) * np.pi], dim=1)
    reg46 = torch.randn(200, 46, generator=gen)
    torch.Tensor.get_device = lambda self: "cpu"
    try:
        dec7 = rbt.decode_bbox_target(rois7.clone(), reg46, anchor_size=mean_size, loc_scope=cfg.RCNN.LOC_SCOPE,
                                      loc_bin_size=cfg.RCNN.LOC_BIN_SIZE, num_head_bin=cfg.RCNN.NUM_HEAD_BIN,
                                      get_xz_fine=True, get_y_by_bin=cfg.RCNN.LOC_Y_BY_BIN,
                                      loc_y_scope=cfg.RCNN.LOC_Y_SCOPE, loc_y_bin_size=cfg.RCNN.LOC_Y_BIN_SIZE,
                                      get_ry_fine=True)
    finally:
        torch.Tensor.get_device = orig
    g["dec7_rois"], g["dec7_reg"], g["dec7_out"] = rois7.numpy(), reg46.numpy(), dec7.numpy()

    np.savez_compressed(out_path, **g)
    print("wrote", out_path, sorted(g))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "ref_modules.npz"))
