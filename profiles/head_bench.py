import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from jmodt_b200.head import RCNN, affinity
from jmodt_b200 import synth
dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
rcnn = RCNN().to(dev).eval(); rcnn.pack()
batch = synth.make_batch(0, B, with_image=False)
pts = torch.from_numpy(batch['pts']).to(dev)
inp = {"rpn_xyz": pts, "rpn_features": torch.randn(B, 16384, 128, device=dev), "seg_mask": (torch.rand(B,16384,device=dev)>0.5).float(),
       "pts_depth": torch.norm(pts, p=2, dim=2), "roi_boxes3d": torch.from_numpy(batch['rois']).to(dev)}
def step():
    out = rcnn(inp)
    f = out['rcnn_feat'].view(B, 128, 512)
    res = [affinity(rcnn, f[i], f[i+1]) for i in range(0, B-1, 2)]
    return out, res
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/5
print(f"B={B} head step {ms:.2f} ms -> {B*128/ms*1e3:.0f} proposals/s ; RCNN flops {147.4e9*B/ms/1e9:.1f} TFLOP/s(fp32-equivalent)")
# phase breakdown
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
pts_input, _ = rcnn.pool_rois(inp)
print("pool_rois", t(lambda: rcnn.pool_rois(inp)))
print("forward_points", t(lambda: rcnn.forward_points(pts_input)))
f = rcnn.forward_points(pts_input)[2].view(B,128,512)
print("affinity x1", t(lambda: affinity(rcnn, f[0], f[1])))
