"""Host side of the tcgen05 MLP layer (jmodt_b200/csrc/tc_gemm.cu): weight packing and launch helpers.

Weights are split once into bf16 hi/lo and rearranged into the exact shared-memory image the kernel's
A-operand descriptors expect (K-major, no-swizzle core matrices of 8 rows x 16 bytes), so a pipeline
stage is filled with ONE contiguous 16 KB bulk copy.
"""
from __future__ import annotations

import itertools

import torch

from . import _lib
from .runtime import EventLog

BM, BK = 128, 32


class Profiler(EventLog):
    """Optional per-launch accounting for bench.py: algorithmic FLOPs (2*M*K*columns, unpadded) and a CUDA-event
    pair per launch on the launching stream (runtime.EventLog; graph-resident when a capture is in progress)."""

    def launch(self, flops, fn, kind="tc_gemm_kernel", desc=""):
        return self.timed(kind, fn, flops=flops, kind=kind, desc=desc)

    def table(self):
        """Per-launch list of one step: (desc, kind, flops, mean ms).  Graph mode: one record per launch site;
        eager mode: the records of the first recorded step."""
        return [(r.desc, r.kind, r.flops, sum(r.ms) / max(1, len(r.ms))) for r in self.records]


profiler = Profiler()


class PackedLayer:
    """bf16 hi/lo chunk images of one layer's weight (+ fp32 bias, zero padded to a multiple of 128)."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor | None, relu: bool, xyz_last: bool = False):
        w = weight.detach().reshape(weight.shape[0], -1).float()
        if xyz_last:   # input columns [dx, dy, dz, channels...] -> [channels..., dx, dy, dz] (csrc/sa_fused.cu)
            w = torch.cat((w[:, 3:], w[:, :3]), dim=1)
        self.M, self.K = w.shape
        Mt, Kc = -(-self.M // BM), -(-self.K // BK)
        wp = torch.zeros(Mt * BM, Kc * BK, dtype=torch.float32, device=w.device)
        wp[: self.M, : self.K] = w
        hi = wp.to(torch.bfloat16)
        lo = (wp - hi.float()).to(torch.bfloat16)

        def image(t):  # (Mt*128, Kc*32) -> [Mt][Kc][k8=4][m8=16][mr=8][kr=8]
            return t.view(Mt, 16, 8, Kc, 4, 8).permute(0, 3, 4, 1, 2, 5)

        self.wpack = torch.stack((image(hi), image(lo)), dim=2).contiguous()  # [Mt][Kc][2][4][16][8][8]
        self._w32, self._xyz_last_pack = (None if xyz_last else weight.detach().reshape(weight.shape[0], -1).float()), None
        b = torch.zeros(Mt * BM, dtype=torch.float32, device=w.device)
        if bias is not None:
            b[: self.M] = bias.detach().float()
        self.bias = b
        self.relu = bool(relu)

    def repacked_xyz_last(self) -> "PackedLayer":
        """This layer with its input columns reordered to [channels, xyz] (cached)."""
        if self._xyz_last_pack is None:
            self._xyz_last_pack = PackedLayer(self._w32, self.bias[: self.M], self.relu, xyz_last=True)
        return self._xyz_last_pack


def weights_token(module) -> tuple:
    """Identity of a module's current weights: (storage address, in-place version counter) of every parameter and
    buffer.  `load_state_dict` and optimizer steps bump the version counter; `.to()` / `.cuda()` / `.float()`
    replace the storage."""
    return tuple((t.data_ptr(), t._version) for t in itertools.chain(module.parameters(), module.buffers()))


def packed_for(module, build, attr: str = "_packed"):
    """The packed weight images cached on `module` under `attr`, rebuilt by `build()` whenever the module's
    parameters or buffers have changed since they were packed (the reference loads checkpoints with strict=False,
    train_utils.py:31-47, after the model may already have run once — a stale image would go unnoticed)."""
    d = module.__dict__
    tok = weights_token(module)
    if d.get(attr) is None or d.get("_pack_token") != tok:
        d[attr] = None
        d[attr] = build()
        d["_pack_token"] = tok
    return d[attr]


def fold_conv_bn(conv, bn=None):
    """Eval-mode BatchNorm folded into the preceding 1x1 conv: returns (weight (Cout, Cin), bias (Cout))."""
    w = conv.weight.detach().reshape(conv.weight.shape[0], -1).float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    if bn is not None:
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        w = w * scale[:, None]
        b = (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
    return w, b


def pack_conv3x3(conv, bn=None, relu: bool = False) -> PackedLayer:
    """3x3 Conv2d (+ eval-mode BatchNorm folded in) as the (Cout, 9 * Cpad) matrix [cout][3 dy + dx][c] that
    csrc/tc_gemm.cu's convolution mode multiplies; Cpad = 4 for the RGB input, the channel count otherwise."""
    w = conv.weight.detach().float()                                   # (Cout, Cin, 3, 3)
    cout, cin = w.shape[0], w.shape[1]
    assert tuple(w.shape[2:]) == (3, 3) and conv.padding == (1, 1) and conv.stride[0] == conv.stride[1] and conv.stride[0] in (1, 2)
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(cout, device=w.device)
    if bn is not None:
        scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        w = w * scale[:, None, None, None]
        b = (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
    cpad = 4 if cin <= 4 else cin
    assert cpad & (cpad - 1) == 0 and (cpad == 4 or cpad % 32 == 0), "conv3x3: channels must be <= 4 or a power of two >= 32"
    wk = torch.zeros(cout, 3, 3, cpad, dtype=torch.float32, device=w.device)
    wk[..., :cin] = w.permute(0, 2, 3, 1)
    layer = PackedLayer(wk.reshape(cout, 9 * cpad), b, relu)
    layer.cpad, layer.stride = cpad, conv.stride[0]
    return layer


def conv3x3(layer: PackedLayer, x: torch.Tensor) -> torch.Tensor:
    """x (B, H, W, Cpad) channels-last fp32 -> (B, OH, OW, Cout) on the tcgen05 kernel (implicit GEMM, 9 taps x channels)."""
    assert x.dim() == 4 and x.is_contiguous() and x.dtype == torch.float32 and x.shape[3] == layer.cpad
    B, H, W, C = x.shape
    s = layer.stride
    OH, OW = (H - 1) // s + 1, (W - 1) // s + 1
    y = torch.empty((B, OH, OW, layer.M), dtype=torch.float32, device=x.device)
    st = _lib.stream_and_device(x)
    flops = 2.0 * B * OH * OW * layer.M * 9 * C
    desc = f"conv3x3 Cout={layer.M} C={C} B={B} {H}x{W} stride={s}"
    profiler.launch(flops, lambda: _lib.check(
        _lib.lib().jmb_tc_conv3x3(layer.wpack.data_ptr(), layer.bias.data_ptr(), layer.M, C, B, H, W, s, x.data_ptr(),
                                  int(layer.relu), y.data_ptr(), st), "tc_conv3x3"), desc=desc)
    return y


def mlp_layer(layer: PackedLayer, x: torch.Tensor, *, out: torch.Tensor | None = None,
              pool: int = 0, point_major_out: bool = False) -> torch.Tensor:
    """Dense layer: x (G, K, N) channel-first fp32 -> (G, M, N), or (G, M, N / pool) with pool > 0, or the
    point-major (G, N, M) with point_major_out (the layout the fused set-abstraction kernel gathers from).
    `out` may be a channel slice `buf[:, c0:c0+M]` of a wider contiguous (G, C, N) tensor (replaces torch.cat)."""
    assert x.dim() == 3 and x.is_contiguous() and x.dtype == torch.float32
    G, K, N = x.shape
    assert K == layer.K
    if point_major_out:
        assert not pool and out is None
        y = torch.empty((G, N, layer.M), dtype=torch.float32, device=x.device)
        st = _lib.stream_and_device(x)
        profiler.launch(2.0 * layer.M * K * G * N, lambda: _lib.check(
            _lib.lib().jmb_tc_mlp_layer(layer.wpack.data_ptr(), layer.bias.data_ptr(), layer.M, K, G, N, 0,
                                        x.data_ptr(), K * N, N, None, None, None, 0, 0, 2, 0, int(layer.relu),
                                        y.data_ptr(), 0, st), "tc_mlp_layer"),
            desc=f"dense M={layer.M} K={K} G={G} N={N} point-major-out")
        return y
    shape = (G, layer.M, N // pool) if pool else (G, layer.M, N)
    y = out if out is not None else torch.empty(shape, dtype=torch.float32, device=x.device)
    assert tuple(y.shape) == shape and y.stride(2) == 1 and y.stride(1) == shape[2]
    y_gs = y.stride(0) if G > 1 else 0
    st = _lib.stream_and_device(x)
    profiler.launch(2.0 * layer.M * K * G * N, lambda: _lib.check(
        _lib.lib().jmb_tc_mlp_layer(layer.wpack.data_ptr(), layer.bias.data_ptr(), layer.M, K, G, N, 0,
                                    x.data_ptr(), K * N, N, None, None, None, 0, 0,
                                    1 if pool else 0, pool, int(layer.relu), y.data_ptr(), y_gs, st), "tc_mlp_layer"),
        desc=f"dense M={layer.M} K={K} G={G} N={N} pool={pool}")
    return y


def mlp_rows(layer: PackedLayer, x: torch.Tensor) -> torch.Tensor:
    """Dense layer over point-major rows: x (G, N, C) -> (G, N, M); C a power of two >= 32 (the one-tap convolution mode of
    csrc/tc_gemm.cu: a row's channels are one contiguous run)."""
    assert x.dim() == 3 and x.is_contiguous() and x.dtype == torch.float32
    G, N, C = x.shape
    assert C == layer.K and C >= BK and C & (C - 1) == 0
    y = torch.empty((G, N, layer.M), dtype=torch.float32, device=x.device)
    st = _lib.stream_and_device(x)
    profiler.launch(2.0 * layer.M * C * G * N, lambda: _lib.check(
        _lib.lib().jmb_tc_mlp_rows(layer.wpack.data_ptr(), layer.bias.data_ptr(), layer.M, C, G, N, x.data_ptr(),
                                   int(layer.relu), y.data_ptr(), st), "tc_mlp_rows"),
        desc=f"rows M={layer.M} K={C} G={G} N={N} point-major")
    return y


def mlp_layer_dot(layer: PackedLayer, last: PackedLayer, x: torch.Tensor) -> torch.Tensor:
    """last(layer(x)) in ONE launch when `last` has a single output channel: x (G, K, N) -> (G, 1, N).  The epilogue of
    `layer` multiplies its activated rows by `last`'s weights and reduces them per 32-row block in fp32 (deterministic
    order); the 4 * ceil(M / 128) partial rows and `last`'s bias are added here.  Saves the (G, M, N) activation and a
    launch whose 128-row tile would hold one useful row."""
    assert x.dim() == 3 and x.is_contiguous() and x.dtype == torch.float32
    assert last.M == 1 and last.K == layer.M and last._w32 is not None
    G, K, N = x.shape
    assert K == layer.K
    Mt = -(-layer.M // BM)
    if getattr(last, "_dot_vec", None) is None:
        v = torch.zeros(Mt * BM, dtype=torch.float32, device=last._w32.device)
        v[: layer.M] = last._w32.reshape(-1)
        last._dot_vec = v
    partial = torch.empty((G, 4 * Mt, N), dtype=torch.float32, device=x.device)
    st = _lib.stream_and_device(x)
    profiler.launch(2.0 * (layer.M + 1) * K * G * N, lambda: _lib.check(
        _lib.lib().jmb_tc_mlp_layer_dot(layer.wpack.data_ptr(), layer.bias.data_ptr(), layer.M, K, G, N, x.data_ptr(),
                                        K * N, N, int(layer.relu), last._dot_vec.data_ptr(), partial.data_ptr(), st),
        "tc_mlp_layer_dot"), desc=f"dense M={layer.M} K={K} G={G} N={N} dot")
    if getattr(last, "_dot_bias", None) is None:
        last._dot_bias = float(last.bias[0].item())      # host copy, made once (not during a graph capture)
    y = torch.empty((G, 1, N), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().jmb_tc_dot_finish(G, 4 * Mt, N, partial.data_ptr(), last._dot_bias, int(last.relu), y.data_ptr(), st),
               "tc_dot_finish")
    return y


def grouped_first_layer(layer: PackedLayer, xyz: torch.Tensor, feats: torch.Tensor | None,
                        idx: torch.Tensor | None, centres: torch.Tensor | None, nsample: int,
                        *, pool: int = 0) -> torch.Tensor:
    """First SharedMLP layer fused with QueryAndGroup / GroupAll (pointnet2_utils.py:231-290).
    xyz (G, n_pts, 3); feats (G, C, n_pts) or None; idx (G, npoint, nsample) int32 or None (GroupAll);
    centres (G, npoint, 3) or None.  Returns (G, M, npoint * nsample) (or pooled)."""
    G, n_pts, _ = xyz.shape
    C = 0 if feats is None else feats.shape[1]
    assert layer.K == 3 + C
    N = idx.shape[1] * idx.shape[2] if idx is not None else n_pts
    shape = (G, layer.M, N // pool) if pool else (G, layer.M, N)
    y = torch.empty(shape, dtype=torch.float32, device=xyz.device)
    st = _lib.stream_and_device(xyz)
    fx = feats if feats is not None else xyz  # never dereferenced for rows >= 3 when C == 0
    profiler.launch(2.0 * layer.M * layer.K * G * N, lambda: _lib.check(
        _lib.lib().jmb_tc_mlp_layer(layer.wpack.data_ptr(), layer.bias.data_ptr(), layer.M, layer.K, G, N, 1,
                                    fx.data_ptr(), C * n_pts, n_pts, _lib.ptr(idx), xyz.data_ptr(),
                                    _lib.ptr(centres), nsample if centres is not None else 0, n_pts,
                                    1 if pool else 0, pool, int(layer.relu), y.data_ptr(), 0, st),
        "tc_mlp_layer(grouped)"), desc=f"grouped M={layer.M} K={layer.K} G={G} N={N} ns={nsample} pool={pool}")
    return y


def hoisted_first_layer(layer: PackedLayer, xyz: torch.Tensor, feats: torch.Tensor, idx: torch.Tensor,
                        centres: torch.Tensor) -> torch.Tensor:
    """First SharedMLP layer of a set-abstraction scale, applied BEFORE the gather (it is linear up to its ReLU and
    grouping only selects columns): Z = W1[:, 3:] . feats + b1 over the n_pts POINTS (tc_gemm_kernel), then
    csrc/interpolate.cu:sa_first_layer_kernel emits relu(Z[idx] + W1[:, :3] . (xyz[idx] - centre)) for every grouped
    neighbour.  xyz (G, n_pts, 3), feats (G, C, n_pts), idx (G, npoint, nsample), centres (G, npoint, 3) ->
    (G, C1, npoint * nsample), what `grouped_first_layer` computes with a gathered GEMM over all npoint * nsample columns."""
    G, n_pts, _ = xyz.shape
    npoint, nsample = idx.shape[1], idx.shape[2]
    cache = getattr(layer, "_hoist", None)
    if cache is None:
        w1 = layer._w32
        cache = layer._hoist = (PackedLayer(w1[:, 3:].contiguous(), layer.bias[: layer.M], relu=False),
                                torch.cat((w1[:, :3], layer.bias[: layer.M, None]), dim=1).contiguous())
    w1f, w1x = cache
    z = mlp_layer(w1f, feats.contiguous(), point_major_out=True)               # (G, n_pts, C1)
    out = torch.empty((G, layer.M, npoint * nsample), dtype=torch.float32, device=xyz.device)
    st = _lib.stream_and_device(xyz)
    _lib.check(_lib.lib().jmb_sa_first_layer(z.data_ptr(), w1x.data_ptr(), layer.M, G, npoint, nsample, n_pts,
                                             idx.data_ptr(), xyz.data_ptr(), centres.contiguous().data_ptr(),
                                             out.data_ptr(), st), "sa_first_layer")
    return out


def sa_fused_supported(layers, n_feat_channels: int, npoint: int, nsample: int) -> bool:
    """Shapes the single-kernel set-abstraction path handles (see csrc/sa_fused.cu): three ReLU layers of widths
    C1, C2 <= 128 (C1 a multiple of 8) and C3 <= 256 (narrower layers run zero-padded to the 128-row MMA tile); any
    number of input feature channels (the first layer is applied to the points before the gather)."""
    return (len(layers) == 3 and layers[0].M <= 128 and layers[0].M % 8 == 0 and layers[1].M <= 128
            and layers[1].K == layers[0].M and layers[2].K == layers[1].M and layers[2].M <= 256
            and all(l.relu for l in layers) and all(l._w32 is not None for l in layers)
            and n_feat_channels >= 0 and layers[0].K == n_feat_channels + 3
            and nsample in (8, 16, 32, 64) and (npoint * nsample) % 128 == 0)


def _sa_fused_packs(layers):
    """What sa_fused_kernel consumes of the three layers (cached on the first):
    W1f = W1[:, 3:] with b1 as a linear PackedLayer (applied to the POINTS before the gather; None without input features),
    w1x (C1, 4) fp32 rows [W1[k, 0:3], b1[k]] (finished by the gather threads; b1 is used only when there is no Z), W2 (128 x 128) and W3 (128|256 x 128)
    zero-padded images."""
    l1, l2, l3 = layers
    cache = getattr(l1, "_sa_packs", None)
    if cache is None:
        def padded(layer, K):
            w = torch.zeros(layer.M, K, dtype=torch.float32, device=layer.bias.device)
            w[:, : layer.K] = layer._w32
            return PackedLayer(w, layer.bias[: layer.M], layer.relu)
        w1 = l1._w32
        w1f = PackedLayer(w1[:, 3:].contiguous(), l1.bias[: l1.M], relu=False) if l1.K > 3 else None      # Z carries b1
        w1x = torch.cat((w1[:, :3], l1.bias[: l1.M, None]), dim=1).contiguous().cpu()     # host table: goes into the launch parameters
        cache = l1._sa_packs = (w1f, w1x, l2 if l2.K == 128 else padded(l2, 128), l3 if l3.K == 128 else padded(l3, 128))
    return cache


def sa_fused(layers, xyz: torch.Tensor, feats: torch.Tensor | None, idx: torch.Tensor, centres: torch.Tensor,
             feats_point_major: bool = False, out_point_major: bool = False) -> torch.Tensor:
    """Whole set-abstraction layer: xyz (G, n_pts, 3), feats (G, C, n_pts) channel-first ((G, n_pts, C) with
    feats_point_major: transposed here), or None (coordinates only), idx (G, npoint, nsample) int32, centres
    (G, npoint, 3) -> (G, C3, npoint) or point-major (G, npoint, C3).  Two launches: Z = W1[:, 3:] . feats over the n_pts
    points (tc_gemm_kernel, point-major rows), then sa_fused_kernel (gather + rest of layer 1 + layers 2, 3 + max-pool)."""
    G, n_pts, _ = xyz.shape
    C = 0 if feats is None else (feats.shape[2] if feats_point_major else feats.shape[1])
    npoint, nsample = idx.shape[1], idx.shape[2]
    assert layers[0].K == 3 + C and idx.is_contiguous() and xyz.is_contiguous()
    w1f, w1x, l2, l3 = _sa_fused_packs(layers)
    C1, C2, C3 = layers[0].M, layers[1].M, layers[2].M
    z = None
    if feats is not None:
        if feats_point_major and C >= BK and C & (C - 1) == 0:
            z = mlp_rows(w1f, feats.contiguous())                             # point-major in, point-major out
        else:
            if feats_point_major:
                feats = feats.transpose(1, 2)
            z = mlp_layer(w1f, feats.contiguous(), point_major_out=True)      # (G, n_pts, C1)
    oshape = (G, npoint, C3) if out_point_major else (G, C3, npoint)
    out = torch.empty(oshape, dtype=torch.float32, device=xyz.device)
    st = _lib.stream_and_device(xyz)
    cols = float(G) * npoint * nsample
    # algorithmic = the reference layer (three 1x1 convs over the grouped tensor); the kernel itself executes layers 2
    # and 3 on the tensor cores and the coordinate part of layer 1 on the CUDA cores (its feature part is the Z launch)
    flops = 2.0 * cols * sum(l.M * l.K for l in layers)
    executed = 2.0 * cols * (C1 * 3 + C2 * C1 + C3 * C2)
    profiler.launch(flops, lambda: _lib.check(
        _lib.lib().jmb_sa_fused(_lib.ptr(z), w1x.data_ptr(), l2.wpack.data_ptr(), l2.bias.data_ptr(),
                                l3.wpack.data_ptr(), l3.bias.data_ptr(), C1, C2, C3, G, npoint, nsample, n_pts,
                                idx.data_ptr(), xyz.data_ptr(), centres.contiguous().data_ptr(),
                                out.data_ptr(), int(out_point_major), st), "sa_fused"), kind="sa_fused_kernel",
        desc=f"sa_fused C={C} widths={C1},{C2},{C3} G={G} npoint={npoint} ns={nsample} executed_flops={executed:.4g}")
    return out


def rcnn_input_fused(w1: PackedLayer, w2: PackedLayer, w3: PackedLayer, rows_in: torch.Tensor,
                     channel_first: bool = False) -> torch.Tensor:
    """xyz_up_layer (2 layers) + concat + merge_down_layer of the per-proposal network (rcnn.py:172-186) in one
    kernel: rows_in (..., 136) in the head layout [128 channels | x, y, z, mask, depth | 0, 0, 0] ->
    (..., 128) point-major merged features, or — channel_first, rows_in (G, S, 136) with S a multiple of 128 —
    (G, 128, S), the layout the reference's merge_down_layer produces (rcnn.py:186).  w1 is the first xyz_up layer
    packed as 128 x 8."""
    assert rows_in.is_contiguous() and rows_in.dtype == torch.float32 and rows_in.shape[-1] == 136
    assert (w1.M, w1.K) == (128, 8) and (w2.M, w2.K) == (128, 128) and (w3.M, w3.K) == (128, 256)
    assert w1.relu and w2.relu and w3.relu
    rows = rows_in.numel() // 136
    rpg = 0
    if channel_first:
        assert rows_in.dim() == 3 and rows_in.shape[1] % 128 == 0
        rpg = rows_in.shape[1]
        out = torch.empty((rows_in.shape[0], 128, rpg), dtype=torch.float32, device=rows_in.device)
    else:
        out = torch.empty(rows_in.shape[:-1] + (128,), dtype=torch.float32, device=rows_in.device)
    st = _lib.stream_and_device(rows_in)
    flops = 2.0 * rows * (128 * 5 + 128 * 128 + 128 * 256)
    profiler.launch(flops, lambda: _lib.check(
        _lib.lib().jmb_rcnn_input_fused(w1.wpack.data_ptr(), w1.bias.data_ptr(), w2.wpack.data_ptr(),
                                        w2.bias.data_ptr(), w3.wpack.data_ptr(), w3.bias.data_ptr(), rows, 136,
                                        rows_in.data_ptr(), out.data_ptr(), rpg, st), "rcnn_input_fused"),
        kind="rcnn_input_kernel", desc=f"rcnn_input_fused rows={rows}")
    return out
