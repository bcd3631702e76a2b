// Bilinear sampling of image feature maps at projected LiDAR points (LI-Fusion), sm_100a.
//
// Replaces feature_gather (reference jmodt/detection/modeling/backbone.py:79-89):
//     F.grid_sample(feature_map (B,C,H,W), xy (B,1,N,2), mode='bilinear', padding_mode='zeros',
//                   align_corners=True).squeeze(2)                      -> (B, C, N)
// One thread owns one point: the four tap offsets and weights are computed once and reused for every
// channel; a warp writes 32 consecutive points of a channel row (coalesced), the tap reads are gathers
// that mostly hit L2 (the level-1..4 maps are 31 / 16 / 8 / 4 MB per frame).
#include "common.cuh"

namespace jmb {

__global__ void __launch_bounds__(128)
feature_gather_kernel(int c, int h, int w, int n, const float *__restrict__ fmap,
                      const float *__restrict__ xy, float *__restrict__ out) {
    const int b = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float gx = __ldg(xy + ((size_t)b * n + p) * 2), gy = __ldg(xy + ((size_t)b * n + p) * 2 + 1);
    // align_corners=True un-normalisation: ((coord + 1) / 2) * (size - 1)
    const float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)(w - 1));
    const float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)(h - 1));
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = (fx + 1.f) - ix, wy0 = (fy + 1.f) - iy;
    const float w_nw = wx0 * wy0, w_ne = wx1 * wy0, w_sw = wx0 * wy1, w_se = wx1 * wy1;
    const bool vx0 = x0 >= 0 && x0 < w, vx1 = x1 >= 0 && x1 < w, vy0 = y0 >= 0 && y0 < h, vy1 = y1 >= 0 && y1 < h;
    const int o_nw = y0 * w + x0, o_ne = y0 * w + x1, o_sw = y1 * w + x0, o_se = y1 * w + x1;
    const size_t plane = (size_t)h * w;
    const float *src = fmap + (size_t)b * c * plane;
    float *dst = out + (size_t)b * c * n + p;
    for (int ci = 0; ci < c; ++ci, src += plane, dst += n) {
        float acc = 0.f;
        if (vx0 && vy0) acc += __ldg(src + o_nw) * w_nw;
        if (vx1 && vy0) acc += __ldg(src + o_ne) * w_ne;
        if (vx0 && vy1) acc += __ldg(src + o_sw) * w_sw;
        if (vx1 && vy1) acc += __ldg(src + o_se) * w_se;
        *dst = acc;
    }
}

}  // namespace jmb

extern "C" int jmb_feature_gather(int b, int c, int h, int w, int n, const float *fmap, const float *xy,
                                  float *out, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && c >= 0 && h > 0 && w > 0 && n >= 0, "feature_gather: bad sizes");
    if (b == 0 || c == 0 || n == 0) return JMB_OK;
    JMB_REQUIRE(fmap && xy && out, "feature_gather: null pointer");
    JMB_REQUIRE(b <= 65535, "feature_gather: batch too large");
    dim3 grid(div_up(n, 128), b);
    feature_gather_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(c, h, w, n, fmap, xy, out);
    return check_launch("feature_gather");
}
