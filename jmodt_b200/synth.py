"""Deterministic synthetic KITTI-shaped frames for tests and bench.py.

There is no KITTI data on the box; this generator mimics the input contract of the
reference loader (jmodt/detection/datasets/kitti_dataset.py): 16 384 points per frame in
rectified-camera coordinates inside cfg.PC_AREA_SCOPE (config.py:34-36), a 375x1242 image
normalised with the ImageNet statistics and zero-padded to 384x1280 (kitti_dataset.py:13,
40-41,100-106), point->pixel coordinates normalised to [-1, 1] by the padded size (:254-255),
and 128 proposals per frame around car-sized clusters (CLS_MEAN_SIZE, config.py:38).
Everything is numpy on the host; callers move the arrays where they need them.
"""
from __future__ import annotations

import numpy as np

N_POINTS = 16384
N_ROIS = 128
IMG_H, IMG_W = 375, 1242
PAD_H, PAD_W = 384, 1280
MEAN_SIZE = np.array([1.52563191462, 1.62856739989, 3.88311640418], dtype=np.float32)  # h, w, l
P2 = np.array([[721.5377, 0.0, 609.5593, 44.85728],
               [0.0, 721.5377, 172.854, 0.2163791],
               [0.0, 0.0, 1.0, 0.002745884]], dtype=np.float64)


def make_frame(frame_id: int, n_points: int = N_POINTS, n_rois: int = N_ROIS, with_image: bool = True,
               empty_rois: int = 4):
    """Returns a dict of numpy arrays: pts (N,3) f32, intensity (N,), pts_xy (N,2) f32 in [-1,1],
    rois (M,7) f32 [x, y_bottom, z, h, w, l, ry], img (3,384,1280) f32 (optional)."""
    rng = np.random.default_rng(1234 + frame_id)
    n_ground = int(n_points * 0.60)
    n_obj = int(n_points * 0.35)
    n_dup = n_points - n_ground - n_obj

    # ground: y ~ 1.65 m below the camera, depth density ~ 1/z
    zg = np.exp(rng.uniform(np.log(2.0), np.log(70.4), n_ground))
    xg = rng.uniform(-1.0, 1.0, n_ground) * np.minimum(40.0, zg * 0.8)
    yg = 1.65 + rng.normal(0, 0.05, n_ground)
    ground = np.stack([xg, yg, zg], 1)

    # car-sized clusters, surface-sampled
    ctr_z = np.exp(rng.uniform(np.log(5.0), np.log(65.0), n_rois))
    ctr_x = rng.uniform(-1.0, 1.0, n_rois) * np.minimum(35.0, ctr_z * 0.6)
    size = MEAN_SIZE[None, :] * rng.uniform(0.9, 1.1, (n_rois, 3))  # h, w, l
    ry = rng.uniform(-np.pi, np.pi, n_rois)
    y_bottom = 1.65 + rng.normal(0, 0.05, n_rois)
    which = rng.integers(0, n_rois, n_obj)
    u = rng.uniform(-0.5, 0.5, (n_obj, 3))
    face = rng.integers(0, 3, n_obj)
    u[np.arange(n_obj), face] = np.sign(u[np.arange(n_obj), face] + 1e-9) * 0.5  # push onto a face
    lx = u[:, 0] * size[which, 2]  # along length
    lz = u[:, 1] * size[which, 1]  # along width
    ly = u[:, 2] * size[which, 0]  # height, centred
    c, s = np.cos(ry[which]), np.sin(ry[which])
    ox = ctr_x[which] + lx * c + lz * s
    oz = ctr_z[which] - lx * s + lz * c
    oy = y_bottom[which] - size[which, 0] / 2 + ly
    objs = np.stack([ox, oy, oz], 1)

    pts = np.concatenate([ground, objs], 0)
    pts[:, 0] = np.clip(pts[:, 0], -40, 40)
    pts[:, 1] = np.clip(pts[:, 1], -1, 3)
    pts[:, 2] = np.clip(pts[:, 2], 0.5, 70.4)
    # exact duplicates of earlier points: the reference loader pads clouds this way
    # (kitti_dataset.py:243-247); they exercise the FPS / three_nn tie rules
    dup = pts[rng.integers(0, pts.shape[0], n_dup)]
    pts = np.concatenate([pts, dup], 0)
    rng.shuffle(pts, axis=0)
    pts = pts.astype(np.float32)
    intensity = rng.uniform(0, 1, n_points).astype(np.float32)

    # projection to the (padded) image plane, kitti_dataset.py:254-255 / calibration.py:60-69
    hom = np.concatenate([pts.astype(np.float64), np.ones((n_points, 1))], 1)
    uvw = hom @ P2.T
    uu, vv = uvw[:, 0] / uvw[:, 2], uvw[:, 1] / uvw[:, 2]
    pts_xy = np.stack([uu / (PAD_W - 1.0) * 2.0 - 1.0, vv / (PAD_H - 1.0) * 2.0 - 1.0], 1).astype(np.float32)

    rois = np.stack([ctr_x + rng.normal(0, 0.2, n_rois), y_bottom + rng.normal(0, 0.05, n_rois),
                     ctr_z + rng.normal(0, 0.2, n_rois), size[:, 0], size[:, 1], size[:, 2],
                     ry + rng.normal(0, 0.05, n_rois)], 1).astype(np.float32)
    if empty_rois:
        rois[-empty_rois:, 0] += 200.0  # far outside the cloud -> pooled_empty_flag = 1

    out = {"pts": pts, "intensity": intensity, "pts_xy": pts_xy, "rois": rois}
    if with_image:
        img = rng.integers(0, 256, (IMG_H, IMG_W, 3)).astype(np.float32) / 255.0
        img = (img - np.array([0.485, 0.456, 0.406], np.float32)) / np.array([0.229, 0.224, 0.225], np.float32)
        pad = np.zeros((PAD_H, PAD_W, 3), np.float32)
        pad[:IMG_H, :IMG_W] = img
        out["img"] = np.ascontiguousarray(pad.transpose(2, 0, 1))
    return out


def make_batch(first_frame: int, batch: int, **kw):
    """Stacks `batch` consecutive frames: pts (B,N,3), pts_xy (B,N,2), rois (B,M,7) [, img (B,3,384,1280)]."""
    frames = [make_frame(first_frame + i, **kw) for i in range(batch)]
    return {k: np.stack([f[k] for f in frames], 0) for k in frames[0]}


def nudge_off_box_faces(pts, boxes_enlarged, margin=1e-4):
    """Moves points that lie within `margin` of a face of any box far away (y += 50) so that the
    in-box predicate does not depend on the last ulp of sinf/cosf.  pts (N,3), boxes (M,7)."""
    p = pts.astype(np.float64)
    bad = np.zeros(p.shape[0], bool)
    for bx in boxes_enlarged.astype(np.float64):
        cx, by, cz, h, w, l, ry = bx
        cy = by - h / 2
        dx, dz = p[:, 0] - cx, p[:, 2] - cz
        xr = dx * np.cos(ry) - dz * np.sin(ry)
        zr = dx * np.sin(ry) + dz * np.cos(ry)
        near = (np.abs(np.abs(xr) - l / 2) < margin) | (np.abs(np.abs(zr) - w / 2) < margin) | \
               (np.abs(np.abs(p[:, 1] - cy) - h / 2) < margin) | (np.abs(np.abs(dx) - 10) < margin) | \
               (np.abs(np.abs(dz) - 10) < margin)
        bad |= near
    out = pts.copy()
    out[bad, 1] += 50.0
    return out


def fill_deterministic(module, seed: int = 0):
    """Fills every parameter / buffer of a torch module with values that depend only on the tensor's NAME and
    shape (not on construction order), so the reference modules and this package's mirrors can be given
    identical weights without shipping checkpoints.  Used by tests/golden/make_golden_modules.py and the tests."""
    import zlib

    import torch
    with torch.no_grad():
        for name, t in module.state_dict().items():
            if name.endswith("num_batches_tracked"):
                continue
            g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + seed) & 0x7FFFFFFF)
            if name.endswith("running_var") or (t.dim() == 1 and name.endswith("weight")):
                v = torch.rand(t.shape, generator=g) + 0.5
            elif t.dim() == 1:
                v = torch.randn(t.shape, generator=g) * 0.1
            else:
                fan_in = int(np.prod(t.shape[1:]))
                v = torch.randn(t.shape, generator=g) / float(np.sqrt(max(fan_in, 1)))
            t.copy_(v.to(t.dtype))
    return module
