#!/usr/bin/env python
"""Generates tests/golden/ref_gpu.npz ON A B200 by running the UNMODIFIED reference — its Python
(`jmodt/detection/layers/proposal_layer.py`, `jmodt/tracking/data_association.py`,
`jmodt/detection/modeling/point_rcnn.py`, ...) over ITS OWN compiled CUDA extensions (oracle/_ref/*.so) — on the
seeded inputs of tests/golden/cases.py.  These parts of the reference need a GPU (`ProposalLayer.__init__` calls
`.cuda()`, `kitti_utils.boxes3d_to_corners3d_torch` allocates `torch.cuda.FloatTensor`s, NMS and IoU exist only as CUDA
kernels), so unlike make_golden_modules.py this generator cannot run in the CPU container:

    gpurun -- python tests/golden/make_golden_gpu.py gpurun_out/ref_gpu.npz      # then copy it to tests/golden/

The reference python travels to the box as the mirror oracle/refpy.stage() writes under oracle/_ref/py (git-ignored).
Nothing of this package's kernels runs here: `import_reference(ops="reference")` binds the reference wrappers to the
reference extensions.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main(out_path):
    from oracle import refpy
    refpy.import_reference(ops="reference")
    import cases
    from jmodt.config import cfg
    from jmodt_b200.synth import fill_deterministic
    torch.backends.cuda.matmul.allow_tf32 = False       # the reference on its fp32 path (SURVEY §7 hard parts)
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    g = {}

    # ---- ProposalLayer.forward (proposal_layer.py:16-121) ---------------------------------------------------------
    from jmodt.detection.layers.proposal_layer import ProposalLayer
    for name, (frames, n, zr, nms_type, post) in cases.PROPOSAL_CASES.items():
        cfg.RPN.NMS_TYPE = nms_type
        cfg.TEST.RPN_POST_NMS_TOP_N = post
        scores, reg, xyz = cases.proposal_inputs(name)
        layer = ProposalLayer(mode="TEST")
        with torch.no_grad():
            boxes, sc = layer(torch.from_numpy(scores).to(dev), torch.from_numpy(reg).to(dev), torch.from_numpy(xyz).to(dev))
        g[f"prop_{name}_boxes"] = boxes.cpu().numpy()
        g[f"prop_{name}_scores"] = sc.cpu().numpy()
        nz = int((boxes.abs().sum(-1) > 0).sum())
        print(f"proposal case {name}: {nz} of {boxes.shape[0] * boxes.shape[1]} rows filled")
    cfg.RPN.NMS_TYPE = "normal"

    # ---- tracker association inputs (data_association.py:10-45) ---------------------------------------------------
    from jmodt.ops.iou3d.iou3d_utils import boxes_iou3d_gpu
    from jmodt.tracking.data_association import boxes_dist_gpu
    pred, det, link = (torch.from_numpy(a).to(dev) for a in cases.association_inputs())
    with torch.no_grad():
        iou = boxes_iou3d_gpu(pred, det)
        dis = boxes_dist_gpu(pred, det)
        link_matrix = link * cases.W_APP + iou * cases.W_IOU + dis * cases.W_DIS      # data_association.py:42-44
    g["assoc_iou"], g["assoc_dist"], g["assoc_link_matrix"] = iou.cpu().numpy(), dis.cpu().numpy(), link_matrix.cpu().numpy()

    # ---- the whole detector at the BASELINE config-3 frame shape (point_rcnn.py:23-72) ----------------------------
    refpy.set_eval_cfg(post_nms_top_n=128)
    from jmodt.detection.modeling.point_rcnn import PointRCNN
    torch.manual_seed(0)
    model = fill_deterministic(PointRCNN(num_classes=2, use_xyz=True, mode="TEST")).to(dev).eval()
    pts, pts_xy, img = cases.detector_inputs(1)
    inp = {"pts_input": torch.from_numpy(pts).to(dev), "pts_xy": torch.from_numpy(pts_xy).to(dev),
           "img": torch.from_numpy(img).to(dev)}
    with torch.no_grad():
        out = model(inp)
    S = cases.STRIDE
    g["det_rpn_cls"] = out["rpn_cls"].cpu().numpy()                                  # (1, 16384, 1)
    g["det_rpn_reg"] = out["rpn_reg"][:, ::S].cpu().numpy()                          # (1, 1024, 76)
    g["det_backbone_features"] = out["backbone_features"][:, :, ::S].cpu().numpy()   # (1, 128, 1024)
    g["det_rois"] = out["rois"].cpu().numpy()                                        # (1, 128, 7)
    g["det_roi_scores_raw"] = out["roi_scores_raw"].cpu().numpy()
    g["det_seg_result"] = out["seg_result"].cpu().numpy().astype(np.uint8)
    g["det_rcnn_cls"] = out["rcnn_cls"].cpu().numpy()                                # (128, 1)
    g["det_rcnn_reg"] = out["rcnn_reg"].cpu().numpy()                                # (128, 46)
    g["det_rcnn_feat"] = out["rcnn_feat"].cpu().numpy() if "rcnn_feat" in out else np.zeros(0, np.float32)
    for k in ("det_rpn_cls", "det_rois", "det_rcnn_cls", "det_rcnn_reg", "det_rcnn_feat"):
        print(k, g[k].shape, float(np.abs(g[k]).mean()))
    # affinity of the frame with itself shifted (tracker.py:81-112 executed on the reference modules)
    if g["det_rcnn_feat"].size:
        feat = out["rcnn_feat"].reshape(128, -1)
        pf, df = feat[:64], feat[64:]
        with torch.no_grad():
            cor = torch.abs(pf.unsqueeze(1).repeat(1, 64, 1) - df.unsqueeze(0).repeat(64, 1, 1))
            logits = model.rcnn_net.link_layer(cor.view(64 * 64, -1, 1)).view(64, 64)
            start = torch.sigmoid(model.rcnn_net.se_layer(cor.mean(dim=0).unsqueeze(-1))).flatten()
            end = torch.sigmoid(model.rcnn_net.se_layer(cor.mean(dim=1).unsqueeze(-1))).flatten()
        g["det_aff_logits"], g["det_aff_start"], g["det_aff_end"] = (t.cpu().numpy() for t in (logits, start, end))
    np.savez_compressed(out_path, **g)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_gpu.npz"))
