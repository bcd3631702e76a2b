"""One RCNN-SA0-shaped sa_fused launch with JMB_SA_DEBUG=1: the kernel prints clock64 stamps of CTA 0
(issuer of part 0, epilogue warp 0) to stderr.  usage: JMB_SA_DEBUG=1 python profiles/sa_timeline.py"""
import sys, torch
sys.path.insert(0, '.')
from jmodt_b200 import tc
from jmodt_b200.pointnet2 import pointnet2_utils as pu
cuda = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
G, n_pts, C, npoint, ns = 1024, 512, 128, 128, 64
xyz = (torch.rand(G, n_pts, 3, generator=g)).to(cuda)
feats = torch.randn(G, n_pts, C, generator=g).to(cuda)
dims = [3 + C, 128, 128, 128]
layers = [tc.PackedLayer((torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5).to(cuda),
                         torch.zeros(dims[i + 1]).to(cuda), True) for i in range(3)]
fidx = pu.farthest_point_sample(xyz, npoint)
centres = pu.gather_operation(xyz.transpose(1, 2).contiguous(), fidx).transpose(1, 2).contiguous()
idx = pu.ball_query(0.2, ns, xyz, centres)
for _ in range(2):
    tc.sa_fused(layers, xyz, feats, idx, centres, feats_point_major=True)
torch.cuda.synchronize()
