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
