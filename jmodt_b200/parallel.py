"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed (NCCL on GPUs,
gloo in the CPU tests).

* Detection / per-proposal feature extraction: frames are independent, so ranks own contiguous blocks of frames
  and NO collective is needed (`frame_shard`).
* Affinity of one frame pair sharded over ranks (BASELINE config 4): predecessor rows are split across ranks;
  each rank computes its (P/W, D) tile of link logits and the `end` scores of its rows locally, and the `start`
  scores of its shard of successor columns; the only exchange is ONE all-gather of the logit tiles, required
  because softmax(dim=0) (reference tracker.py:88) spans all rows.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def frame_shard(num_frames: int, rank: int, world: int) -> range:
    """Contiguous block of frames owned by `rank` (the first num_frames % world ranks get one extra frame)."""
    base, extra = divmod(num_frames, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def row_shard(n: int, rank: int, world: int) -> Tuple[int, int]:
    r = frame_shard(n, rank, world)
    return r.start, r.stop


def gather_rows(tile: torch.Tensor, total_rows: int, group=None) -> torch.Tensor:
    """All-gather row tiles of possibly different heights into the full (total_rows, D) matrix."""
    world = dist.get_world_size(group)
    if world == 1:
        return tile
    D = tile.shape[1]
    max_rows = -(-total_rows // world)
    padded = tile.new_zeros(max_rows, D)
    padded[: tile.shape[0]] = tile
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    parts = []
    for r in range(world):
        lo, hi = row_shard(total_rows, r, world)
        parts.append(out[r][: hi - lo])
    return torch.cat(parts, dim=0)


def gather_packed(parts, sizes_fn, group=None, timer=None):
    """ONE all-gather for several ragged per-rank pieces: every rank flattens its pieces into one buffer padded to the
    largest rank's size, the buffers are exchanged with a single `all_gather_into_tensor`, and piece k of rank r is
    returned as a flat view of length sizes_fn(r)[k].  Returns a list (per piece) of lists (per rank)."""
    world = dist.get_world_size(group)
    sizes = [sizes_fn(r) for r in range(world)]
    cap = max(sum(sz) for sz in sizes)
    flat = torch.cat([p.reshape(-1) for p in parts]) if parts else None
    buf = flat.new_zeros(cap)
    buf[: flat.numel()] = flat
    out = flat.new_empty(world * cap)
    if timer is not None:
        timer[0].record()
    dist.all_gather_into_tensor(out, buf, group=group)
    if timer is not None:
        timer[1].record()
    res = [[] for _ in parts]
    for r in range(world):
        off = r * cap
        for k, n in enumerate(sizes[r]):
            res[k].append(out[off: off + n])
            off += n
    return res


@torch.no_grad()
def sharded_affinity_device(link_model, se_model, pred_features: torch.Tensor, det_features: torch.Tensor, group=None,
                            timer=None):
    """BASELINE config 4 on the sm_100a kernels: one P x D frame pair scored by all ranks of `group`.

    Both feature matrices are replicated (256 KB each).  Rank r owns predecessor rows [lo, hi) and successor columns
    [clo, chi): it computes its (p, D) tile of link logits (pair-correlation kernel on its rows -> link stack on
    tcgen05), the `end` scores of its rows, and the `start` scores of its columns over ALL predecessors.  The one
    exchange is a single all-gather that carries the logit tile with the two score slices behind it (16.5 KB per
    rank at 128 x 128 on 4 ranks); it is required because softmax(dim=0) (reference tracker.py:88) spans all rows.
    A column of the layer kernel's output depends on that column's inputs only, so the gathered logits are
    bit-identical to the single-GPU result.  Returns link (P, D), start (D,), end (P,), logits (P, D) on every rank."""
    from .head import _stacks, dual_softmax, pair_corr, run_stack
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    P, D = pred_features.shape[0], det_features.shape[0]
    lo, hi = row_shard(P, rank, world)
    clo, chi = row_shard(D, rank, world)
    link_stack, se_stack = _stacks(link_model, se_model)
    pt = pred_features.t().contiguous()            # (C, P)
    dt = det_features.t().contiguous()             # (C, D)
    dev = pred_features.device
    tile = torch.empty((0, D), dtype=torch.float32, device=dev)
    end_local = torch.empty(0, dtype=torch.float32, device=dev)
    start_local = torch.empty(0, dtype=torch.float32, device=dev)
    if hi > lo:
        cor, _, mean_d = pair_corr(pt[:, lo:hi].contiguous().unsqueeze(0), dt.unsqueeze(0), want_mean_p=False)
        tile = run_stack(link_stack, cor).view(hi - lo, D)
        end_local = torch.sigmoid(run_stack(se_stack, mean_d)).view(-1)
    if chi > clo:
        _, mean_p, _ = pair_corr(pt.unsqueeze(0), dt[:, clo:chi].contiguous().unsqueeze(0), want_cor=False,
                                 want_mean_d=False)
        start_local = torch.sigmoid(run_stack(se_stack, mean_p)).view(-1)
    if world > 1:
        def sizes(r):
            a, b = row_shard(P, r, world)
            c, d = row_shard(D, r, world)
            return ((b - a) * D, b - a, d - c)
        tiles, ends, starts = gather_packed([tile, end_local, start_local], sizes, group, timer)
        logits = torch.cat(tiles).view(P, D)
        end, start = torch.cat(ends), torch.cat(starts)
    else:
        logits, end, start = tile, end_local, start_local
    return dual_softmax(logits), start, end, logits


def exchange_boundary_features(first_frame_features: torch.Tensor, group=None, timer=None):
    """Frame-sharded sequences: the affinity pair that straddles two shards (last frame of rank r, first frame of rank
    r + 1) needs the right neighbour's first-frame RCNN features (M, C).  One all-gather of those (256 KB per rank);
    returns the neighbour's features, or None on the last rank."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    x = first_frame_features.contiguous()
    out = x.new_empty((world,) + tuple(x.shape))
    if timer is not None:
        timer[0].record()
    dist.all_gather_into_tensor(out, x, group=group)
    if timer is not None:
        timer[1].record()
    return out[rank + 1] if rank + 1 < world else None


def sharded_affinity(logits_fn: Callable, se_fn: Callable, pred_features: torch.Tensor, det_features: torch.Tensor,
                     group=None):
    """Link / start / end scores of one frame pair computed by all ranks of `group`.

    logits_fn(pred_rows (p,C), det (D,C)) -> (p, D) raw link logits;
    se_fn(x (n, C)) -> (n,) sigmoid start/end scores of mean |p - d| features.
    Both feature matrices are replicated (they are 256 KB); returns the full (P, D) link scores, start (D,),
    end (P,) on every rank, equal to the single-GPU result (reference tracker.py:81-112).
    """
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    P, D = pred_features.shape[0], det_features.shape[0]
    lo, hi = row_shard(P, rank, world)
    tile = logits_fn(pred_features[lo:hi], det_features) if hi > lo else pred_features.new_zeros(0, D)
    # end scores: mean over successors of |p_i - d_j| for the local rows
    cor_rows = (pred_features[lo:hi].unsqueeze(1) - det_features.unsqueeze(0)).abs()           # (p, D, C)
    end_local = se_fn(cor_rows.mean(dim=1)) if hi > lo else pred_features.new_zeros(0)
    # start scores: this rank's shard of successor columns, over ALL predecessors (no exchange needed)
    clo, chi = row_shard(D, rank, world)
    cor_cols = (pred_features.unsqueeze(1) - det_features[clo:chi].unsqueeze(0)).abs()         # (P, d, C)
    start_local = se_fn(cor_cols.mean(dim=0)) if chi > clo else pred_features.new_zeros(0)
    if world > 1:
        def sizes(r):
            a, b = row_shard(P, r, world)
            c, d = row_shard(D, r, world)
            return ((b - a) * D, b - a, d - c)
        tiles, ends, starts = gather_packed([tile.contiguous(), end_local, start_local], sizes, group)   # the ONE exchange
        logits, end, start = torch.cat(tiles).view(P, D), torch.cat(ends), torch.cat(starts)
    else:
        logits, end, start = tile, end_local, start_local
    link = (torch.softmax(logits, dim=1) + torch.softmax(logits, dim=0)) / 2
    return link, start, end, logits
