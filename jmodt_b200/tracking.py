"""Tracker-side entry points of the hot path: the affinity block of `Tracker.update`
(reference jmodt/tracking/tracker.py:81-112) and the association inputs of `ortools_solve`
(jmodt/tracking/data_association.py:31-45).  Track bookkeeping, the Kalman filter and the MIP / Hungarian solvers stay
the reference's host code (out of scope).

Two ways in for the reference code base:

* unchanged `tracker.py`: after `jmodt_b200.dropin.install()` the `link_model` / `se_model` the evaluation script
  hands to `Tracker` (tools/eval.py:333-336: `rcnn_net.link_layer`, `rcnn_net.se_layer`) are this package's
  `pt_utils.Conv1d` stacks, whose eval-mode forward runs on the tcgen05 layer kernel — the three model calls at
  tracker.py:86,106,109 reach the tensor cores as they are;
* `affinity_scores(link_model, se_model, pred_features, det_features)`: the whole block — pair features, both means,
  link stack, dual softmax, start / end stacks — on this package's kernels without materialising the three
  (P, D, 512) temporaries; it replaces tracker.py:81-89 and :105-110 (INTEGRATION.md shows the edit).
"""
from __future__ import annotations

import torch

from .association import boxes_dist_gpu, link_matrix  # noqa: F401  (re-exported: data_association.py:10-45)
from .head import affinity_scores_batched


@torch.no_grad()
def affinity_scores(link_model, se_model, pred_features: torch.Tensor, det_features: torch.Tensor):
    """tracker.py:81-112 for one frame: pred_features (P, C), det_features (D, C) ->
    link_scores (P, D) = (softmax(dim=1) + softmax(dim=0)) / 2 of the link logits,
    start_scores (D,) = sigmoid(se_model(mean over predecessors of |p - d|)),
    end_scores (P,) = sigmoid(se_model(mean over successors)).  The tracker's w_se weighting, the numpy
    concatenations and the solver call stay with the caller."""
    link, start, end, _ = affinity_scores_batched(link_model, se_model, pred_features.unsqueeze(0).contiguous(),
                                                  det_features.unsqueeze(0).contiguous())
    return link[0], start[0], end[0]
