"""One big dense tc_gemm launch (the affinity link layer shape) with JMB_TC_DEBUG=1: CTA 0 prints clock64 stamps.
usage: JMB_TC_DEBUG=1 python profiles/tc_timeline.py [M K G N]"""
import sys, torch
sys.path.insert(0, '.')
from jmodt_b200 import tc
M, K, G, N = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (512, 512, 4, 16384)
cuda = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
layer = tc.PackedLayer((torch.randn(M, K, generator=g) / K ** 0.5).to(cuda), torch.zeros(M).to(cuda), True)
x = torch.randn(G, K, N, generator=g).to(cuda)
for _ in range(2):
    tc.mlp_layer(layer, x)
torch.cuda.synchronize()
