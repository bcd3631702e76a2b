// Host-side helpers of the reference's roipool3d module (jmodt/ops/roipool3d/src/roipool3d.cpp:82-195:
// `pts_in_boxes3d_cpu`, `roipool3d_cpu`).  The reference's dataset code calls them on CPU tensors (ground-truth
// augmentation); they are part of the operator API this library mirrors, NOT a fallback of the GPU path — nothing in
// jmodt_b200 routes device work here.  Plain pointers in host memory, same results as the reference functions
// (tests/test_oracle_cpu.py compares with the reference's own extension, oracle/_ref).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/jmodt_b200.h"

namespace {

// Point inside a box rotated about y?  Arithmetic as the reference evaluates it: the `/ 2.0` terms are double
// divisions, cos / sin are the double functions rounded to float, the products and sums are float.
inline int point_in_box(const float *pt, const float *box) {
    const float cx = box[0], bottom_y = box[1], cz = box[2], h = box[3], w = box[4], l = box[5], ry = box[6];
    const float cy = (float)((double)bottom_y - (double)h / 2.0);
    const float dx = pt[0] - cx, dz = pt[2] - cz;
    if ((double)fabsf(pt[1] - cy) > (double)h / 2.0 || fabsf(dx) > 10.0f || fabsf(dz) > 10.0f) return 0;
    const float c = (float)cos((double)ry), s = (float)sin((double)ry);
    const float xr = dx * c + dz * (-s);
    const float zr = dx * s + dz * c;
    const double hl = (double)l / 2.0, hw = (double)w / 2.0;
    return ((double)xr >= -hl) & ((double)xr <= hl) & ((double)zr >= -hw) & ((double)zr <= hw);
}

}  // namespace

extern "C" int jmb_pts_in_boxes3d_host(int n_pts, int n_boxes, const float *pts, const float *boxes3d, int64_t *flags) {
    if (n_pts < 0 || n_boxes < 0 || ((n_pts > 0 && n_boxes > 0) && (!pts || !boxes3d || !flags))) return JMB_ERR_INVALID_ARG;
    for (int b = 0; b < n_boxes; ++b)
        for (int j = 0; j < n_pts; ++j) flags[(size_t)b * n_pts + j] = point_in_box(pts + 3 * (size_t)j, boxes3d + 7 * (size_t)b);
    return JMB_OK;
}

extern "C" int jmb_roipool3d_host(int n_pts, int n_boxes, int feat_len, int sampled, const float *pts,
                                  const float *boxes3d, const float *pts_feature, float *pooled_pts,
                                  float *pooled_features, int64_t *empty_flag) {
    if (n_pts < 0 || n_boxes < 0 || feat_len < 0 || sampled < 0) return JMB_ERR_INVALID_ARG;
    if (n_boxes == 0) return JMB_OK;
    if (!boxes3d || !empty_flag || (n_pts > 0 && !pts) || (sampled > 0 && (!pooled_pts || (feat_len > 0 && !pooled_features))))
        return JMB_ERR_INVALID_ARG;
    for (int b = 0; b < n_boxes; ++b) {
        float *op = pooled_pts + (size_t)b * sampled * 3;
        float *of = pooled_features + (size_t)b * sampled * feat_len;
        int taken = 0;
        for (int j = 0; j < n_pts && taken < sampled; ++j) {       // first `sampled` points inside, in index order
            if (!point_in_box(pts + 3 * (size_t)j, boxes3d + 7 * (size_t)b)) continue;
            memcpy(op + 3 * (size_t)taken, pts + 3 * (size_t)j, 3 * sizeof(float));
            if (feat_len) memcpy(of + (size_t)taken * feat_len, pts_feature + (size_t)j * feat_len, feat_len * sizeof(float));
            ++taken;
        }
        empty_flag[b] = taken == 0;
        for (int k = taken; taken > 0 && k < sampled; ++k) {        // wrap-around padding with the points already taken
            memcpy(op + 3 * (size_t)k, op + 3 * (size_t)(k % taken), 3 * sizeof(float));
            if (feat_len) memcpy(of + (size_t)k * feat_len, of + (size_t)(k % taken) * feat_len, feat_len * sizeof(float));
        }
    }
    return JMB_OK;
}
