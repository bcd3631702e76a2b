#!/usr/bin/env python
"""sa_fused_kernel alone at the two RCNN set-abstraction shapes of BASELINE config 3 (8 frames x 128 proposals):
time per call = the first-layer GEMM over the points (tc_gemm_kernel) + sa_fused_kernel (CUDA events, L2 flushed),
algorithmic TFLOP/s of the reference layer, and — with JMB_SA_DEBUG=1 — the in-kernel clock64
timeline of CTA 0's issuer thread and epilogue warp 0 on stderr.

    gpurun -- python profiles/sa_bench.py            # add JMB_SA_DEBUG=1 for the timeline (one launch per shape)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmodt_b200 import tc  # noqa: E402
from jmodt_b200.pointnet2 import pointnet2_utils as pu  # noqa: E402

dev = torch.device("cuda:0")
dbg = os.environ.get("JMB_SA_DEBUG") == "1"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator().manual_seed(0)
for (G, n_pts, C, npoint, ns, widths, r) in ((1024, 512, 128, 128, 64, (128, 128, 128), 0.2),
                                             (1024, 128, 128, 32, 64, (128, 128, 256), 0.4),
                                             (8, 16384, 0, 4096, 32, (32, 32, 64), 0.5),
                                             (8, 4096, 96, 1024, 32, (64, 96, 128), 1.0)):
    xyz = (torch.rand(G, n_pts, 3, generator=g) * (1.0 if n_pts <= 512 else 20.0)).to(dev)
    feats = torch.randn(G, C, n_pts, generator=g).to(dev) if C else None     # channel-first, as the previous stage writes it
    dims = [3 + C, *widths]
    layers = [tc.PackedLayer((torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5).to(dev),
                             (torch.randn(dims[i + 1], generator=g) * 0.1).to(dev), True) for i in range(3)]
    fidx = pu.farthest_point_sample(xyz, npoint)
    ctr = pu.gather_operation(xyz.transpose(1, 2).contiguous(), fidx).transpose(1, 2).contiguous()
    idx = pu.ball_query(r, ns, xyz, ctr)
    run = lambda: tc.sa_fused(layers, xyz, feats, idx, ctr)     # first-layer GEMM over the points + sa_fused_kernel
    if dbg:
        print(f"--- timeline G={G} n_pts={n_pts} C={C} npoint={npoint} ns={ns} widths={widths}", file=sys.stderr, flush=True)
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
    flops = 2.0 * G * npoint * ns * sum(dims[i] * dims[i + 1] for i in range(3))
    med = ms[len(ms) // 2]
    print(f"sa_fused G={G} n_pts={n_pts} C={C} npoint={npoint} ns={ns} widths={widths}: {med * 1e3:8.1f} us  "
          f"{flops / med / 1e9:7.1f} TFLOP/s algorithmic  (min {ms[0] * 1e3:.1f} us)", flush=True)
