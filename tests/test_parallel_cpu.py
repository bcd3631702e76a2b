"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: frame sharding covers every frame exactly once, and
the row-sharded affinity with its single all-gather of link-logit tiles equals the unsharded computation
(exact equality of gathered logits).  The dense arithmetic here is the torch restatement of the
reference heads (oracle/modules_ref.py); on GPUs the same functions are fed by the tcgen05 layers."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_frame_shard_partitions():
    from jmodt_b200.parallel import frame_shard
    for n, w in [(1024, 8), (10, 4), (3, 8), (0, 2)]:
        seen = [f for r in range(w) for f in frame_shard(n, r, w)]
        assert seen == list(range(n))
        sizes = [len(frame_shard(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, P, D, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jmodt_b200.parallel import sharded_affinity
    from jmodt_b200.head import RCNN
    from jmodt_b200.synth import fill_deterministic
    from oracle import modules_ref
    torch.manual_seed(0)
    rcnn = fill_deterministic(RCNN()).eval()
    g = torch.Generator().manual_seed(3)
    pred, det = torch.randn(P, 512, generator=g).abs(), torch.randn(D, 512, generator=g).abs()

    def logits_fn(p_rows, d_all):
        cor = (p_rows.unsqueeze(1) - d_all.unsqueeze(0)).abs()
        return rcnn.link_layer(cor.view(-1, 512, 1)).view(p_rows.shape[0], d_all.shape[0])

    def se_fn(x):
        return torch.sigmoid(rcnn.se_layer(x.unsqueeze(-1))).flatten()

    with torch.no_grad():
        link, start, end, logits = sharded_affinity(logits_fn, se_fn, pred, det)
        wl, ws, we, wlog = modules_ref.affinity(rcnn.link_layer, rcnn.se_layer, pred, det)
    ok = (torch.allclose(logits, wlog, atol=1e-5, rtol=1e-5) and torch.allclose(link, wl, atol=1e-6)
          and torch.allclose(start, ws, atol=1e-6) and torch.allclose(end, we, atol=1e-6)
          and logits.shape == (P, D))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("P,D", [(16, 12), (7, 5)])
def test_sharded_affinity_equals_single_process_gloo(P, D):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + P) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, P, D, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
