"""Seeded inputs shared by tests/golden/make_golden_gpu.py (which runs the UNMODIFIED reference on them on a B200) and
the tests that compare this package's sm_100a path with the committed outputs (tests/golden/ref_gpu.npz).
Pure numpy, deterministic."""
from __future__ import annotations

import numpy as np

PROPOSAL_CASES = {
    # name: (frames, points, z-range the clouds are squeezed into (None = as generated), nms type, post-NMS top n)
    "both_bins": (2, 16384, None, "normal", 100),
    "both_bins_128": (1, 16384, None, "normal", 128),
    "far_bin_empty": (2, 16384, (1.0, 30.0), "normal", 100),     # proposal_layer.py:93-103: back-fill from the first bin's
                                                                  # candidates beyond its 6300 pre-NMS quota
    "far_bin_empty_short": (1, 4096, (1.0, 30.0), "normal", 100),  # ... which do not exist here: the reference runs its NMS on
                                                                  # zero boxes (and prints "CUDA Error!"); 70 rows + padding
    "near_bin_empty": (1, 4096, (48.0, 70.0), "normal", 100),    # :94-95: `continue` for the first bin
    "rotated_nms": (1, 4096, None, "rotate", 100),                # cfg.RPN.NMS_TYPE = 'rotate'
    "few_points": (2, 64, None, "normal", 100),                   # fewer candidates than the quotas -> zero padding
}


def proposal_inputs(name: str):
    """rpn_scores (B, N), rpn_reg (B, N, 76), xyz (B, N, 3) float32 for one proposal-layer case."""
    from jmodt_b200 import synth
    frames, n, zr, _, _ = PROPOSAL_CASES[name]
    seed = sum(ord(c) for c in name)
    rng = np.random.RandomState(seed)
    xyz = np.stack([synth.make_frame(900 + seed + k, n_points=n, with_image=False)["pts"] for k in range(frames)])
    if zr is not None:
        z = xyz[..., 2]
        xyz[..., 2] = zr[0] + (z - z.min()) / (z.max() - z.min()) * (zr[1] - zr[0])
    # regression channels (rpn.py:32-37 / bbox_transform.py:27-260): 12 x-bins, 12 z-bins, 12+12 residuals, y, 12 heading
    # bins, 12 heading residuals, 3 sizes.  Small residuals keep most decoded boxes near their points, so the clusters
    # of the synthetic cloud produce heavily overlapping candidates and the NMS has work to do.
    reg = (rng.standard_normal((frames, n, 76)) * 0.6).astype(np.float32)
    reg[..., 73:76] *= 0.15
    scores = (rng.standard_normal((frames, n)) * 2.0).astype(np.float32)
    return scores, reg, xyz.astype(np.float32)


def association_inputs():
    """pred_boxes (P, 7), det_boxes (D, 7) [x, y, z, h, w, l, ry] and link_score (P, D) as Tracker.update hands them to
    ortools_solve (tracker.py:113-123): detections of two consecutive synthetic frames."""
    from jmodt_b200 import synth
    rng = np.random.RandomState(5)
    a = synth.make_frame(40, n_points=1024, with_image=False, empty_rois=0)["rois"][:48]
    b = a[rng.permutation(48)[:40]] + rng.normal(0, 0.3, (40, 7)).astype(np.float32) * np.array([1, .1, 1, .05, .05, .05, .2], np.float32)
    b[-6:, [0, 2]] += rng.uniform(5, 30, (6, 2)).astype(np.float32)          # a few new objects far from every track
    link = rng.uniform(0, 1, (48, 40)).astype(np.float32)
    return a.astype(np.float32), b.astype(np.float32), link


W_APP, W_IOU, W_DIS = 0.4, 0.35, 0.25          # any weights: the sum is linear in them


def detector_inputs(frames: int = 1):
    """One full-size frame batch for the detector goldens: pts (B, 16384, 3), pts_xy (B, 16384, 2), img (B, 3, 384, 1280)."""
    from jmodt_b200 import synth
    b = synth.make_batch(700, frames, with_image=True)
    return b["pts"], b["pts_xy"], b["img"]


STRIDE = 16      # the detector goldens keep every 16th point of the per-point outputs (file size)
