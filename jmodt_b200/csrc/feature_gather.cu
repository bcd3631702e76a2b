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

constexpr int FG_CH = 8;  // channels per thread: 4*FG_CH independent tap loads in flight

__global__ void __launch_bounds__(128)
feature_gather_kernel(int c, int h, int w, int n, const float *__restrict__ fmap,
                      const float *__restrict__ xy, float *__restrict__ out) {
    const int b = blockIdx.z;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float gx = __ldg(xy + ((size_t)b * n + p) * 2), gy = __ldg(xy + ((size_t)b * n + p) * 2 + 1);
    // align_corners=True un-normalisation: ((coord + 1) / 2) * (size - 1)
    const float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)(w - 1));
    const float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)(h - 1));
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = (fx + 1.f) - ix, wy0 = (fy + 1.f) - iy;
    // out-of-image taps contribute zero (padding_mode='zeros'): zero their weight and clamp their address
    const bool vx0 = x0 >= 0 && x0 < w, vx1 = x1 >= 0 && x1 < w, vy0 = y0 >= 0 && y0 < h, vy1 = y1 >= 0 && y1 < h;
    const float w_nw = (vx0 && vy0) ? wx0 * wy0 : 0.f, w_ne = (vx1 && vy0) ? wx1 * wy0 : 0.f;
    const float w_sw = (vx0 && vy1) ? wx0 * wy1 : 0.f, w_se = (vx1 && vy1) ? wx1 * wy1 : 0.f;
    const int cx0 = min(max(x0, 0), w - 1), cx1 = min(max(x1, 0), w - 1);
    const int cy0 = min(max(y0, 0), h - 1), cy1 = min(max(y1, 0), h - 1);
    const int o_nw = cy0 * w + cx0, o_ne = cy0 * w + cx1, o_sw = cy1 * w + cx0, o_se = cy1 * w + cx1;
    const size_t plane = (size_t)h * w;
    const int c0 = blockIdx.y * FG_CH;
    const float *src = fmap + ((size_t)b * c + c0) * plane;
    float *dst = out + ((size_t)b * c + c0) * n + p;
    float v[FG_CH][4];
#pragma unroll
    for (int i = 0; i < FG_CH; ++i) {
        if (c0 + i < c) {
            const float *s2 = src + (size_t)i * plane;
            v[i][0] = __ldg(s2 + o_nw); v[i][1] = __ldg(s2 + o_ne); v[i][2] = __ldg(s2 + o_sw); v[i][3] = __ldg(s2 + o_se);
        }
    }
#pragma unroll
    for (int i = 0; i < FG_CH; ++i) {
        if (c0 + i < c) {
            float acc = 0.f;      // same accumulation order as before: nw, ne, sw, se
            acc += v[i][0] * w_nw; acc += v[i][1] * w_ne; acc += v[i][2] * w_sw; acc += v[i][3] * w_se;
            dst[(size_t)i * n] = acc;
        }
    }
}

}  // namespace jmb

extern "C" int jmb_feature_gather(int b, int c, int h, int w, int n, const float *fmap, const float *xy,
                                  float *out, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && c >= 0 && h > 0 && w > 0 && n >= 0, "feature_gather: bad sizes");
    if (b == 0 || c == 0 || n == 0) return JMB_OK;
    JMB_REQUIRE(fmap && xy && out, "feature_gather: null pointer");
    JMB_REQUIRE(b <= 65535 && div_up(c, FG_CH) <= 65535, "feature_gather: batch / channel count too large");
    dim3 grid(div_up(n, 128), div_up(c, FG_CH), b);
    feature_gather_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(c, h, w, n, fmap, xy, out);
    return check_launch("feature_gather");
}
