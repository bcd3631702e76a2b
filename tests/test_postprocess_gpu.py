"""Detection post-processing (tools/eval.py:96-193) on the device vs the same procedure with the C
oracle's rotated NMS on the host: identical kept RoI indices, boxes and scores."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_postprocess_detections_matches_host_procedure(cuda, cref):
    from jmodt_b200 import box_utils, synth
    from jmodt_b200.detector import decode_bbox_target, postprocess_detections
    from jmodt_b200.head import HeadConfig
    cfg = HeadConfig()
    B, M = 3, 128
    g = torch.Generator().manual_seed(5)
    rois = torch.from_numpy(synth.make_batch(11, B, with_image=False)["rois"])             # (B, 128, 7)
    rois[1, 64:] = rois[1, :64] + 0.05 * torch.randn(64, 7, generator=g)                     # heavy overlaps in frame 1
    reg = torch.randn(B * M, cfg.reg_channel, generator=g) * 0.1
    cls = torch.randn(B * M, 1, generator=g) * 2
    cls[2 * M:] = -10.0                                                                      # frame 2: nothing passes
    feat = torch.randn(B * M, 512, 1, generator=g)
    got = postprocess_detections(rois.to(cuda), cls.to(cuda), reg.to(cuda), feat.to(cuda), head_cfg=cfg)
    assert len(got) == B and got[2]["boxes3d"].shape == (0, 7) and got[2]["feat"].shape == (0, 512)
    anchor = torch.tensor((1.52563191462, 1.62856739989, 3.88311640418), dtype=torch.float32)
    boxes = decode_bbox_target(rois.to(cuda).view(-1, 7), reg.to(cuda), cfg.loc_scope, cfg.loc_bin_size, cfg.num_head_bin,
                               anchor.to(cuda), get_ry_fine=True).view(B, M, 7).cpu()
    raw = cls.view(B, M)
    for k in range(2):
        sel = torch.nonzero(torch.sigmoid(raw[k]) > 0.2).flatten()
        order = raw[k, sel].sort(0, descending=True)[1]
        bev = box_utils.boxes3d_to_bev_torch(boxes[k, sel]).numpy()
        keep = order.numpy()[cref.nms_sorted(np.ascontiguousarray(bev[order.numpy()]), 0.1, True)]
        want_idx = sel.numpy()[keep]
        assert np.array_equal(got[k]["roi_index"].cpu().numpy(), want_idx), k
        assert torch.equal(got[k]["boxes3d"].cpu(), boxes[k, want_idx])
        assert torch.allclose(got[k]["scores"].cpu(), torch.sigmoid(raw[k, want_idx]))
        assert torch.equal(got[k]["feat"].cpu(), feat.view(B, M, 512)[k, want_idx])
    assert 0 < len(got[1]["roi_index"]) < int((torch.sigmoid(raw[1]) > 0.2).sum())          # overlaps were suppressed
