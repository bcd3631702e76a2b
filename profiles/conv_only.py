"""Runs one BasicBlock (3x3 conv + BN + ReLU + 3x3 conv stride 2) on the tcgen05 implicit-GEMM path (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmodt_b200.detector import BasicBlock  # noqa: E402
from jmodt_b200.synth import fill_deterministic  # noqa: E402

dev = torch.device("cuda", 0)
cin, cout, H, W = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (128, 256, 96, 320)
blk = fill_deterministic(BasicBlock(cin, cout)).to(dev).eval()
x = torch.randn(8, cin, H, W, device=dev).contiguous(memory_format=torch.channels_last)
with torch.no_grad():
    for _ in range(3):
        y = blk(x)
torch.cuda.synchronize()
print(y.shape)
