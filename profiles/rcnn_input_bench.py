"""rcnn_input_fused (xyz_up_layer x2 + merge_down_layer of the per-proposal network) alone at the config-3 shape
(1024 proposals x 512 points): time per call, both output layouts; JMB_SA_DEBUG=1 prints the in-kernel timeline."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmodt_b200 import tc  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
mk = lambda m, k, relu=True: tc.PackedLayer((torch.randn(m, k, generator=g) / k ** 0.5).to(dev), (torch.randn(m, generator=g) * 0.1).to(dev), relu)
w1, w2, w3 = mk(128, 8), mk(128, 128), mk(128, 256)
rows = torch.randn(1024, 512, 136, generator=g).to(dev)
dbg = os.environ.get("JMB_SA_DEBUG") == "1"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for cf in (True, False):
    run = lambda: tc.rcnn_input_fused(w1, w2, w3, rows, channel_first=cf)
    if dbg:
        print(f"--- timeline channel_first={cf}", file=sys.stderr, flush=True)
        run(); torch.cuda.synchronize()
        continue
    for _ in range(3):
        run()
    ms = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    print(f"rcnn_input_fused channel_first={cf}: {ms[len(ms) // 2] * 1e3:.1f} us (min {ms[0] * 1e3:.1f})", flush=True)
