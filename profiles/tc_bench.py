#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 MLP-layer kernel on the layer shapes of the region-proposal fusion path.

usage (GPU box): python profiles/tc_bench.py [--reps 20] > gpurun_out/tc_bench.txt
Each line: measured us per launch (CUDA events around `reps` back-to-back launches), algorithmic TFLOP/s, GB/s of
compulsory traffic (x read once + y written once), and the HBM / tensor lower bounds for comparison.
"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from jmodt_b200 import tc  # noqa: E402

# (kind, M, K, G, N, extra)   kind: dense | pool | pm (point-major out) | grouped (ns, n_pts)
SHAPES = [
    ("dense", 16, 16, 8, 16384 * 4, 0), ("pool", 32, 16, 8, 16384 * 4, 16), ("dense", 32, 32, 8, 32768 * 4, 0),
    ("pool", 64, 32, 8, 32768 * 4, 32),
    ("dense", 24, 64, 8, 4096, 0), ("dense", 1, 24, 8, 4096, 0), ("dense", 96, 192, 8, 4096, 0),
    ("dense", 64, 64, 8, 16384, 0), ("pool", 128, 64, 8, 16384, 16), ("dense", 96, 64, 8, 32768, 0),
    ("pool", 128, 96, 8, 32768, 32),
    ("dense", 64, 256, 8, 1024, 0), ("dense", 256, 512, 8, 1024, 0), ("dense", 196, 128, 8, 8192, 0),
    ("pool", 256, 196, 8, 8192, 32),
    ("dense", 128, 512, 8, 256, 0), ("dense", 512, 1024, 8, 256, 0), ("dense", 384, 256, 8, 2048, 0),
    ("pool", 512, 384, 8, 2048, 32),
    ("dense", 256, 1024, 8, 64, 0), ("dense", 1024, 2048, 8, 64, 0), ("dense", 512, 1536, 8, 256, 0),
    ("dense", 512, 768, 8, 1024, 0), ("dense", 256, 608, 8, 4096, 0), ("dense", 128, 256, 8, 16384, 0),
    ("dense", 128, 128, 8, 16384, 0), ("dense", 1, 128, 8, 16384, 0), ("dense", 76, 128, 8, 16384, 0),
    ("dense", 128, 5, 1024, 512, 0), ("dense", 128, 128, 1024, 512, 0), ("pm", 128, 256, 1024, 512, 0),
    ("dense", 256, 256, 1024, 32, 0), ("pool", 512, 256, 1024, 32, 32),
    ("dense", 512, 512, 1, 1024, 0), ("dense", 512, 512, 4, 16384, 0), ("dense", 1, 512, 4, 16384, 0),
    ("dense", 512, 512, 4, 128, 0),
    ("pm", 128, 128, 1024, 512, 0), ("pm", 128, 128, 1024, 128, 0), ("pm", 64, 96, 8, 4096, 0),     # first-layer GEMMs over the points (sa_fused Z)
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peaks = {"hbm": 6.5e12, "tensor": 1.4e15}
    total = 0.0
    for kind, M, K, G, N, extra in SHAPES:
        g = torch.Generator(device="cpu").manual_seed(1)
        layer = tc.PackedLayer((torch.randn(M, K, generator=g) / K ** 0.5).to(dev), torch.randn(M, generator=g).to(dev), True)
        x = torch.randn(G, K, N, device=dev)
        fn = {"dense": lambda: tc.mlp_layer(layer, x), "pool": lambda: tc.mlp_layer(layer, x, pool=extra),
              "pm": lambda: tc.mlp_layer(layer, x, point_major_out=True)}[kind]
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.reps
        flops = 2.0 * M * K * G * N
        out_cols = N // extra if kind == "pool" else N
        bytes_ = 4.0 * G * (K * N + M * out_cols)
        lb_hbm, lb_tc = bytes_ / peaks["hbm"] * 1e6, 3 * flops / peaks["tensor"] * 1e6
        total += us
        print(f"{us:8.1f} us  {flops / us / 1e6:7.1f} TF/s  {bytes_ / us / 1e3:7.0f} GB/s   bounds: hbm {lb_hbm:6.1f} us, tensor {lb_tc:6.1f} us   "
              f"{kind:6s} M={M} K={K} G={G} N={N} {extra or ''}")
        del x
    print(f"total {total:.1f} us")


if __name__ == "__main__":
    main()
