"""Stream / CUDA-graph runtime (jmodt_b200/runtime.py): forked branches and the captured step must give exactly
the results of the plain launch order."""
import pytest
import torch


def test_parallel_is_sequential_without_cuda():
    from jmodt_b200 import runtime
    order = []
    outs = runtime.parallel(lambda: order.append(0) or 10, lambda: order.append(1) or 11, lambda: order.append(2) or 12)
    assert outs == [10, 11, 12]
    if not torch.cuda.is_available():
        assert order == [0, 1, 2]


def test_event_log_disabled_is_a_passthrough():
    from jmodt_b200.runtime import EventLog
    log = EventLog()
    assert log.timed("x", lambda: 7) == 7 and log.launches == 1 and log.records == []
    assert log.summary() == {"flops": 0.0, "ms": 0.0, "launches": 0}


@pytest.mark.gpu
def test_branches_on_forked_streams_match_sequential(cuda):
    from jmodt_b200 import runtime, tc
    g = torch.Generator().manual_seed(0)
    layers = [tc.PackedLayer((torch.randn(m, 64, generator=g) / 8).to(cuda), torch.randn(m, generator=g).to(cuda), True)
              for m in (32, 128, 200)]
    x = torch.randn(4, 64, 1000, generator=g).to(cuda)
    fns = [(lambda l=l: tc.mlp_layer(l, x)) for l in layers]
    old = runtime.branch_parallel
    try:
        runtime.branch_parallel = False
        want = runtime.parallel(*fns)
        runtime.branch_parallel = True
        got = runtime.parallel(*fns)
    finally:
        runtime.branch_parallel = old
    torch.cuda.synchronize()
    for a, b in zip(got, want):
        assert torch.equal(a, b)


@pytest.mark.gpu
def test_captured_step_replays_bit_identically(cuda):
    """RoI pooling + per-proposal network + pair affinity captured as one CUDA graph: replays on NEW inputs loaded
    into the static buffers must equal the eager result bit for bit (same kernels, same order)."""
    from jmodt_b200 import synth
    from jmodt_b200.head import RCNN, affinity_batched
    from jmodt_b200.runtime import CapturedPath
    torch.manual_seed(0)
    rcnn = RCNN().to(cuda).eval()
    rcnn.pack()

    def frames(seed):
        batch = synth.make_batch(seed, 2, with_image=False)
        g = torch.Generator().manual_seed(seed)
        pts = torch.from_numpy(batch["pts"])
        return {"rpn_xyz": pts, "rpn_features": torch.randn(2, 16384, 128, generator=g),
                "seg_mask": (torch.rand(2, 16384, generator=g) > 0.5).float(), "pts_depth": torch.norm(pts, p=2, dim=2),
                "roi_boxes3d": torch.from_numpy(batch["rois"])[:, :32].contiguous()}

    def step(d):
        out = rcnn(d)
        f = out["rcnn_feat"].view(2, 32, 512)
        link, start, end, _ = affinity_batched(rcnn, f[0:1], f[1:2])
        return {"cls": out["rcnn_cls"], "reg": out["rcnn_reg"], "link": link, "start": start, "end": end}

    static = {k: v.to(cuda) for k, v in frames(1).items()}
    step(static)                                   # lazy initialisation happens outside the capture
    cap = CapturedPath(step, static, warmup=1)
    assert cap.launches_per_replay > 10
    for seed in (1, 5):
        host = {k: v.pin_memory() for k, v in frames(seed).items()}
        cap.load(host)
        got = {k: v.clone() for k, v in cap.replay().items()}
        want = step({k: v.to(cuda) for k, v in host.items()})
        torch.cuda.synchronize()
        for k in want:
            assert torch.equal(got[k], want[k]), k


@pytest.mark.gpu
def test_geometry_one_step_ahead_equals_the_unpipelined_forward(cuda):
    """runtime.GeometryAhead: the coordinate stage of batch k + 1 computed on a side stream under batch k's feature
    stages (eagerly and as a captured graph) must give, for every batch of a stream of DIFFERENT batches, exactly the
    unpipelined forward's outputs."""
    from jmodt_b200 import synth
    from jmodt_b200.detector import PointRCNN, RpnConfig
    from jmodt_b200.runtime import CapturedPath, GeometryAhead
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    model = synth.fill_deterministic(PointRCNN(rpn_cfg=RpnConfig(post_nms_top_n=64))).to(cuda).eval()
    batches = []
    for seed in (3, 4, 5, 6):
        b = synth.make_batch(seed * 10, 2, with_image=(seed == 3))
        batches.append({k: torch.from_numpy(b[k]).to(cuda) for k in ("pts", "pts_xy", "rois")})
        if seed == 3:
            with torch.no_grad():
                maps, fused = model.rpn.backbone_net.image_features(torch.from_numpy(b["img"]).to(cuda))
    image_maps = ([m.contiguous() for m in maps], fused.contiguous())
    keys = ("rpn_cls", "rpn_reg", "rois", "rcnn_cls", "rcnn_reg", "rcnn_feat")

    def forward(b, plan=None):
        out = model({"pts_input": b["pts"], "pts_xy": b["pts_xy"]}, image_maps=image_maps, geometry=plan)
        return {k: out[k] for k in keys}

    want = [{k: v.clone() for k, v in forward(b).items()} for b in batches]
    # eager pipeline
    ahead = GeometryAhead(model.geometry, batches[0]["pts"])
    for i, b in enumerate(batches):
        nxt = batches[(i + 1) % len(batches)]["pts"]
        got = ahead.step(lambda plan: forward(b, plan), nxt)
        torch.cuda.synchronize()
        for k in keys:
            assert torch.equal(got[k], want[i][k]), (i, k)
    # the same as one captured graph over static buffers
    static = {"pts": batches[0]["pts"].clone(), "pts_xy": batches[0]["pts_xy"].clone(), "next_pts": batches[0]["pts"].clone()}
    ahead2 = GeometryAhead(model.geometry, static["pts"])

    def step(d):
        out = ahead2.step(lambda plan: forward(d, plan), d["next_pts"])
        d["pts"].copy_(d["next_pts"])
        return out
    cap = CapturedPath(step, static, warmup=1)
    # after the warm-up / capture calls the static state is "current = next = batch 0"; stream batches 1, 2, 3 through it
    cap.load({"next_pts": batches[1]["pts"], "pts_xy": batches[0]["pts_xy"]})
    got0 = {k: v.clone() for k, v in cap.replay().items()}
    for k in keys:
        assert torch.equal(got0[k], want[0][k]), ("graph", 0, k)
    for i in (1, 2):
        cap.load({"next_pts": batches[i + 1]["pts"], "pts_xy": batches[i]["pts_xy"]})
        got = cap.replay()
        torch.cuda.synchronize()
        for k in keys:
            assert torch.equal(got[k], want[i][k]), ("graph", i, k)


@pytest.mark.gpu
def test_streamed_path_returns_every_steps_results_in_order(cuda):
    """runtime.StreamedPath: inputs staged on a copy stream, results read back one call later — each step's host results
    equal the serial upload / replay / read-back of the same inputs."""
    from jmodt_b200.runtime import CapturedPath, StreamedPath

    def fn(inp):
        return {"y": inp["x"] * 2.0 + inp["b"].sum(), "z": (inp["x"] > 0.5).float().sum(dim=1)}
    static = {"x": torch.zeros(64, 1000, device=cuda), "b": torch.zeros(10, device=cuda)}
    path = CapturedPath(fn, static, warmup=1)
    g = torch.Generator().manual_seed(0)
    batches = [{"x": torch.rand(64, 1000, generator=g).pin_memory(), "b": torch.rand(10, generator=g).pin_memory()}
               for _ in range(7)]
    sp = StreamedPath(path, batches[0], ("y", "z"))
    got = []
    for bt in batches:
        r = sp.step(bt)
        if r is not None:
            got.append([t.clone() for t in r])
    got.append([t.clone() for t in sp.drain()])
    assert len(got) == len(batches)
    for bt, (y, z) in zip(batches, got):
        assert torch.allclose(y, bt["x"] * 2.0 + bt["b"].sum(), rtol=1e-6, atol=1e-6)      # b.sum(): device vs host order
        assert torch.equal(z, (bt["x"] > 0.5).float().sum(dim=1))
