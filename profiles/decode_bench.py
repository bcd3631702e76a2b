"""Times the sparse image decoder (csrc/image_decode.cu) at the config-3 shape (8 frames x 16 384 points, 384 x 1280 image)
next to the dense cuDNN formulation the reference runs, and the cuDNN 3x3 convolution stack in both layouts."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmodt_b200.detector import PointNet2MSG, RpnConfig  # noqa: E402
from jmodt_b200.pointnet2 import pytorch_utils as pt_utils  # noqa: E402
from jmodt_b200.synth import fill_deterministic, make_batch  # noqa: E402


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def main():
    dev = torch.device("cuda", 0)
    B = 8
    net = fill_deterministic(PointNet2MSG(input_channels=0, cfg=RpnConfig())).to(dev).eval()
    b = make_batch(0, B)
    xy, img = (torch.from_numpy(b[k]).to(dev) for k in ("pts_xy", "img"))
    res = {}
    with torch.no_grad():
        maps_cl, _ = net.image_features(img, dense=False)
        with pt_utils.torch_layers():
            maps, _ = net.image_features(img, dense=True)[0], None
        res["decode_gather_sparse_us"] = timed(lambda: net.decode_gather(maps_cl, xy))

        def dense():
            de = torch.cat([dc(m) for dc, m in zip(net.DeConv, maps)], dim=1)
            fused = torch.relu(net.image_fusion_bn(net.image_fusion_conv(de)))
            return torch.nn.functional.grid_sample(fused, xy.unsqueeze(1), align_corners=True).squeeze(2)
        res["decode_dense_cudnn_tf32_us"] = timed(dense, n=3, warm=1)
        def own_stack():
            y = img
            for blk in net.Img_Block:
                y = blk(y)
            return y
        res["conv_stack_tcgen05_fp32_grade_us"] = timed(own_stack, n=5, warm=2)
        y = img
        for i, blk in enumerate(net.Img_Block):
            res["conv_block%d_tcgen05_us" % (i + 1)] = timed(lambda: blk(y), n=5, warm=1)
            y = blk(y)
        ctx = pt_utils.torch_layers()
        ctx.__enter__()
        for name, tf32, cl in (("nchw_tf32", True, False), ("nhwc_tf32", True, True), ("nchw_fp32", False, False),
                               ("nhwc_fp32", False, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            x = img.contiguous(memory_format=torch.channels_last) if cl else img

            def stack():
                y = x
                for blk in net.Img_Block:
                    y = blk(y)
                return y
            res["conv_stack_cudnn_%s_us" % name] = timed(stack, n=3, warm=2)
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cudnn.benchmark = True
        x = img.contiguous(memory_format=torch.channels_last)

        def stack():
            y = x
            for blk in net.Img_Block:
                y = blk(y)
            return y
        res["conv_stack_cudnn_nhwc_tf32_autotuned_us"] = timed(stack, n=3, warm=3)
        xb = x.bfloat16()
        netb = fill_deterministic(PointNet2MSG(input_channels=0, cfg=RpnConfig())).to(dev).eval().bfloat16()

        def stackb():
            y = xb
            for blk in netb.Img_Block:
                y = blk(y)
            return y
        res["conv_stack_cudnn_nhwc_bf16_autotuned_us"] = timed(stackb, n=3, warm=3)
        ctx.__exit__(None, None, None)
    res["conv_stack_gflop"] = 92.3 * B
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/decode_bench.json", "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
