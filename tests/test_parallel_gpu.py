"""BASELINE config 4 on devices: one P x D frame pair scored by row shards whose link-logit tiles come from the tcgen05
layers, gathered with ONE all-gather — exact equality with the single-GPU result.

* one GPU: the shards of a 4-rank job evaluated one after the other on the same device and concatenated (what the
  all-gather would return) — pins the claim the exchange rests on: a logit tile does not depend on how the rows are
  split;
* >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2|4`): real NCCL ranks through `parallel.sharded_affinity_device`,
  plus the shard-boundary feature exchange and a single process driving two devices through the C ABI.
"""
import os
import sys

import pytest
import torch

from conftest import ROOT


def _features(P, D, seed=3):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(P, 512, generator=g).abs(), torch.randn(D, 512, generator=g).abs()


def _rcnn(dev):
    from jmodt_b200.head import RCNN
    from jmodt_b200.synth import fill_deterministic
    torch.manual_seed(0)
    return fill_deterministic(RCNN()).to(dev).eval()


@pytest.mark.gpu
@pytest.mark.parametrize("P,D,world", [(128, 128, 4), (128, 128, 8), (37, 45, 4), (5, 9, 8)])
def test_row_shards_reproduce_the_unsharded_logits_bit_for_bit(cuda, P, D, world):
    from jmodt_b200.head import _stacks, affinity, pair_corr, run_stack
    from jmodt_b200.parallel import row_shard
    rcnn = _rcnn(cuda)
    pred, det = (t.to(cuda) for t in _features(P, D))
    link, start, end, logits = affinity(rcnn, pred, det)
    link_stack, se_stack = _stacks(rcnn.link_layer, rcnn.se_layer)
    pt, dt = pred.t().contiguous(), det.t().contiguous()
    tiles, ends, starts = [], [], []
    for r in range(world):
        lo, hi = row_shard(P, r, world)
        clo, chi = row_shard(D, r, world)
        if hi > lo:
            cor, _, mean_d = pair_corr(pt[:, lo:hi].contiguous().unsqueeze(0), dt.unsqueeze(0), want_mean_p=False)
            tiles.append(run_stack(link_stack, cor).view(hi - lo, D))
            ends.append(torch.sigmoid(run_stack(se_stack, mean_d)).view(-1))
        if chi > clo:
            _, mean_p, _ = pair_corr(pt.unsqueeze(0), dt[:, clo:chi].contiguous().unsqueeze(0), want_cor=False,
                                     want_mean_d=False)
            starts.append(torch.sigmoid(run_stack(se_stack, mean_p)).view(-1))
    assert torch.equal(torch.cat(tiles), logits)
    assert torch.equal(torch.cat(ends), end) and torch.equal(torch.cat(starts), start)
    from jmodt_b200.head import dual_softmax
    assert torch.equal(dual_softmax(torch.cat(tiles)), link)


def _nccl_worker(rank, world, port, P, D, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from jmodt_b200.head import affinity
    from jmodt_b200.parallel import exchange_boundary_features, sharded_affinity_device
    rcnn = _rcnn(dev)
    pred, det = (t.to(dev) for t in _features(P, D))
    link, start, end, logits = sharded_affinity_device(rcnn.link_layer, rcnn.se_layer, pred, det)
    wl, ws, we, wlog = affinity(rcnn, pred, det)
    checks = {"logits": torch.equal(logits, wlog), "link": torch.equal(link, wl), "start": torch.equal(start, ws),
              "end": torch.equal(end, we), "shape": logits.shape == (P, D)}
    # shard-boundary exchange: rank r receives rank r + 1's first-frame features
    mine = torch.full((8, 512), float(rank), device=dev)
    nb = exchange_boundary_features(mine)
    checks["boundary"] = (nb is None) if rank == world - 1 else bool((nb == rank + 1).all())
    torch.cuda.synchronize()
    bad = [k for k, v in checks.items() if not v]
    q.put((rank, True if not bad else "mismatch: " + ",".join(bad)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("P,D", [(128, 128), (37, 45)])
def test_sharded_affinity_nccl_equals_single_gpu_exactly(cuda, P, D):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() + P) % 2000
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, P, D, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert sorted(res) == [(r, True) for r in range(world)]


@pytest.mark.gpu
def test_one_process_two_devices(cuda):
    """per-device kernel attributes / SM counts / scheduler counters — the 217 KB dynamic-smem
    launches must work on a second device of the same process (the reference's nn.DataParallel mode)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    from jmodt_b200.head import affinity
    outs = []
    for d in range(2):
        dev = torch.device("cuda", d)
        with torch.cuda.device(dev):
            rcnn = _rcnn(dev)
            pred, det = (t.to(dev) for t in _features(64, 64))
            outs.append([t.cpu() for t in affinity(rcnn, pred, det)])
            torch.cuda.synchronize(dev)
    for a, b in zip(*outs):
        assert torch.equal(a, b)
