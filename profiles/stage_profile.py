"""Runs ONE stage of the e2e step between cudaProfilerStart/Stop so `ncu --profile-from-start off` captures just it.
usage: python profiles/stage_profile.py {step|rpn|rcnn|affinity|proposal} [frames]"""
import sys, torch
sys.path.insert(0, '.')
import bench
stage = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device('cuda:0')
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
host_np = bench.make_inputs_e2e(0, B)
suite = bench.FusionE2E(dev, B, host_np)
host = {k: torch.from_numpy(v) for k, v in host_np.items() if k != 'img'}
d = bench.to_device_e2e(host, dev, torch, non_blocking=False)
m = suite.model
for _ in range(2): out = suite.step(d)
inp = {"pts_input": d["pts"], "pts_xy": d["pts_xy"]}
rpn = m.rpn(inp, image_maps=suite.image_maps)
scores = rpn["rpn_cls"][:, :, 0]
rc_in = {"rpn_xyz": rpn["backbone_xyz"], "rpn_features": rpn["backbone_features"].permute(0, 2, 1),
         "seg_mask": (torch.sigmoid(scores) > 0.2).float(), "roi_boxes3d": d["rois"], "pts_depth": torch.norm(rpn["backbone_xyz"], p=2, dim=2)}
pts_input, _ = m.rcnn_net.pool_rois(rc_in)
feat = m.rcnn_net.forward_points(pts_input)[2]
torch.cuda.synchronize()
torch.cuda.profiler.start()
if stage == 'step': suite.step(d)
elif stage == 'rpn': m.rpn(inp, image_maps=suite.image_maps)
elif stage == 'rcnn': m.rcnn_net.forward_points(pts_input)
elif stage == 'affinity': m.pair_affinity(feat, 128)
elif stage == 'proposal': m.rpn.proposal_layer(scores, rpn["rpn_reg"], rpn["backbone_xyz"])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
