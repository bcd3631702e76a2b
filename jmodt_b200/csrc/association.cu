// Tracker association inputs for sm_100a.
//
// Replaces boxes_dist_gpu (reference jmodt/tracking/data_association.py:10-28), which materialises the corners of both
// box sets (kitti_utils.boxes3d_to_corners3d_torch, kitti_utils.py:107-133: nine cat / matmul / permute launches),
// repeats them into an (m, n, 8, 8, 3) tensor (98 KB per box pair) and reduces it with two norms and a max, and the
// weighted sum of data_association.py:42-45 (link * w_app + iou * w_iou + dist * w_dis: three more passes).
// Here one thread owns one box pair: both boxes' eight corners are rebuilt in registers, the 64 corner distances
// and the centre distance are reduced on the fly, and — when the link scores and the 3-D IoU matrix are passed — the
// association score is written in the same pass.
#include "common.cuh"

namespace jmb {

struct Corners { float x[8], y[8], z[8]; };

// kitti_utils.py:107-133: corners in the box frame, rotated about y by ry (R = [[c,0,s],[0,1,0],[-s,0,c]]), translated
__device__ __forceinline__ void box_corners(const float *b, Corners &c) {
    const float h = b[3], w = b[4], l = b[5];
    const float ca = cosf(b[6]), sa = sinf(b[6]);
    const float hl = l * 0.5f, hw = w * 0.5f;
    const float xs[8] = {hl, hl, -hl, -hl, hl, hl, -hl, -hl};
    const float ys[8] = {0.f, 0.f, 0.f, 0.f, -h, -h, -h, -h};
    const float zs[8] = {hw, -hw, -hw, hw, hw, -hw, -hw, hw};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c.x[i] = __fmaf_rn(sa, zs[i], __fmul_rn(ca, xs[i])) + b[0];
        c.y[i] = ys[i] + b[1];
        c.z[i] = __fmaf_rn(ca, zs[i], __fmul_rn(-sa, xs[i])) + b[2];
    }
}

__global__ void __launch_bounds__(128)
boxes_dist_kernel(int na, int nb, const float *__restrict__ boxes_a, const float *__restrict__ boxes_b,
                  float *__restrict__ dist, const float *__restrict__ link, const float *__restrict__ iou,
                  float w_app, float w_iou, float w_dis, float *__restrict__ score) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= nb) return;
    float a[7], b[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) { a[k] = __ldg(boxes_a + (size_t)i * 7 + k); b[k] = __ldg(boxes_b + (size_t)j * 7 + k); }
    Corners ca, cb;
    box_corners(a, ca);
    box_corners(b, cb);
    float far2 = 0.f;
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float dx = ca.x[p] - cb.x[q], dy = ca.y[p] - cb.y[q], dz = ca.z[p] - cb.z[q];
            far2 = fmaxf(far2, dx * dx + dy * dy + dz * dz);
        }
    const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    const float d = 1.f - sqrtf(dx * dx + dy * dy + dz * dz) / sqrtf(far2);
    const size_t o = (size_t)i * nb + j;
    if (dist) dist[o] = d;
    if (score) score[o] = __ldg(link + o) * w_app + __ldg(iou + o) * w_iou + d * w_dis;
}

}  // namespace jmb

extern "C" int jmb_boxes_dist(int na, const float *boxes_a, int nb, const float *boxes_b, float *dist,
                              const float *link, const float *iou, float w_app, float w_iou, float w_dis,
                              float *score, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(na >= 0 && nb >= 0, "boxes_dist: negative size");
    if (na == 0 || nb == 0) return JMB_OK;
    JMB_REQUIRE(boxes_a && boxes_b && (dist || score), "boxes_dist: null pointer");
    JMB_REQUIRE(!score || (link && iou), "boxes_dist: the association score needs the link and IoU matrices");
    JMB_REQUIRE(na <= 65535, "boxes_dist: too many rows");
    boxes_dist_kernel<<<dim3(div_up(nb, 128), na), 128, 0, (cudaStream_t)stream>>>(na, nb, boxes_a, boxes_b, dist, link, iou,
                                                                              w_app, w_iou, w_dis, score);
    return check_launch("boxes_dist");
}
