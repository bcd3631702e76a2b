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
    logits = gather_rows(tile.contiguous(), P, group) if world > 1 else tile
    link = (torch.softmax(logits, dim=1) + torch.softmax(logits, dim=0)) / 2
    # end scores: mean over successors of |p_i - d_j| for the local rows
    cor_rows = (pred_features[lo:hi].unsqueeze(1) - det_features.unsqueeze(0)).abs()           # (p, D, C)
    end_local = se_fn(cor_rows.mean(dim=1)) if hi > lo else pred_features.new_zeros(0)
    # start scores: this rank's shard of successor columns, over ALL predecessors (no exchange needed)
    clo, chi = row_shard(D, rank, world)
    cor_cols = (pred_features.unsqueeze(1) - det_features[clo:chi].unsqueeze(0)).abs()         # (P, d, C)
    start_local = se_fn(cor_cols.mean(dim=0)) if chi > clo else pred_features.new_zeros(0)
    if world > 1:
        end = gather_rows(end_local.view(-1, 1), P, group).flatten()
        start = gather_rows(start_local.view(-1, 1), D, group).flatten()
    else:
        end, start = end_local, start_local
    return link, start, end, logits
