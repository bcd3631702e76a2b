"""Runs the sparse image decoder a few times at the config-3 shape (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmodt_b200.detector import PointNet2MSG, RpnConfig  # noqa: E402
from jmodt_b200.synth import fill_deterministic, make_batch  # noqa: E402

dev = torch.device("cuda", 0)
net = fill_deterministic(PointNet2MSG(input_channels=0, cfg=RpnConfig())).to(dev).eval()
b = make_batch(0, 8)
xy = torch.from_numpy(b["pts_xy"]).to(dev)
g = torch.Generator(device="cuda").manual_seed(0)
maps = [torch.randn(8, c, 384 >> (l + 1), 1280 >> (l + 1), device=dev, generator=g).contiguous(memory_format=torch.channels_last)
        for l, c in enumerate((64, 128, 256, 512))]
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    net.decode_gather(maps, xy)
torch.cuda.synchronize()
