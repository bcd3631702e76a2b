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

// ---- LI-Fusion attention weight (reference jmodt/detection/modeling/backbone.py:33-58, IALayer) --------------------
//     att[b, n] = sigmoid( fc3( tanh( fc1(img[b, :, n]) + fc2(pt[b, :, n]) ) ) )
// Three tiny Linear layers (ic -> rc, pc -> rc, rc -> 1 with rc = pc / 4), two elementwise ops and a sigmoid: six
// launches whose GEMMs are far below one tensor-core tile wave (rc = 24 ... 256 rows, 64 ... 16 384 points per frame).
// One SIMT kernel in plain fp32: a CTA owns 32 points and ALL rc rows, walks K = ic + pc in 32-deep shared-memory tiles of
// [W1 | W2] and [img ; pt], and finishes with tanh, the fc3 dot product (a reduction over the CTA's own rows) and the
// sigmoid.  w12 (rc, ic + pc) row-major, b12 (rc) = b1 + b2, w3 (rc), b3 scalar; img (B, ic, N), pt (B, pc, N) channel-first.
namespace jmb {

constexpr int IA_TN = 32, IA_KT = 32, IA_THREADS = 256, IA_MAXR = 32;     // up to 8 * 32 = 256 rows

template <int ROWS>      // rows per thread: rc <= 8 * ROWS
__global__ void __launch_bounds__(IA_THREADS)
ia_attention_kernel(int ic, int pc, int rc, int N, const float *__restrict__ img, const float *__restrict__ pt,
                    const float *__restrict__ w12, const float *__restrict__ b12, const float *__restrict__ w3,
                    float b3, float *__restrict__ att) {
    extern __shared__ float ia_smem[];
    const int K = ic + pc;
    float *sw = ia_smem;                        // [rc][IA_KT + 1]
    float *sx = ia_smem + rc * (IA_KT + 1);     // [IA_KT][IA_TN]
    float *sred = sx + IA_KT * IA_TN;           // [8][IA_TN]
    const int b = blockIdx.y, n0 = blockIdx.x * IA_TN;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // column, row group (rows ty, ty + 8, ...)
    float acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
    const float *imgb = img + (size_t)b * ic * N, *ptb = pt + (size_t)b * pc * N;
    for (int k0 = 0; k0 < K; k0 += IA_KT) {
        for (int e = threadIdx.x; e < rc * IA_KT; e += IA_THREADS) {
            const int r = e / IA_KT, kk = e - r * IA_KT, k = k0 + kk;
            sw[r * (IA_KT + 1) + kk] = k < K ? __ldg(w12 + (size_t)r * K + k) : 0.f;
        }
        for (int e = threadIdx.x; e < IA_KT * IA_TN; e += IA_THREADS) {
            const int kk = e / IA_TN, c = e - kk * IA_TN, k = k0 + kk, n = n0 + c;
            float v = 0.f;
            if (k < K && n < N) v = k < ic ? __ldg(imgb + (size_t)k * N + n) : __ldg(ptb + (size_t)(k - ic) * N + n);
            sx[kk * IA_TN + c] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < IA_KT; ++kk) {
            const float x = sx[kk * IA_TN + tx];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const int row = ty + 8 * r;       // rows >= rc read zero-filled... guard instead: the tile holds rc rows only
                if (row < rc) acc[r] = fmaf(sw[row * (IA_KT + 1) + kk], x, acc[r]);
            }
        }
        __syncthreads();
    }
    float part = 0.f;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int row = ty + 8 * r;
        if (row < rc) part = fmaf(__ldg(w3 + row), tanhf(acc[r] + __ldg(b12 + row)), part);
    }
    sred[ty * IA_TN + tx] = part;
    __syncthreads();
    if (ty == 0 && n0 + tx < N) {
        float s = b3;
#pragma unroll
        for (int g = 0; g < 8; ++g) s += sred[g * IA_TN + tx];
        att[(size_t)b * N + n0 + tx] = 1.f / (1.f + expf(-s));
    }
}

}  // namespace jmb

extern "C" int jmb_ia_attention(int B, int ic, int pc, int rc, int N, const float *img, const float *pt,
                                const float *w12, const float *b12, const float *w3, float b3, float *att, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(B >= 0 && ic > 0 && pc > 0 && rc > 0 && N >= 0, "ia_attention: bad sizes");
    if (B == 0 || N == 0) return JMB_OK;
    JMB_REQUIRE(img && pt && w12 && b12 && w3 && att, "ia_attention: null pointer");
    JMB_REQUIRE(rc <= 8 * IA_MAXR, "ia_attention: %d reduced channels exceed the tile (256)", rc);
    JMB_REQUIRE(B <= 65535, "ia_attention: batch too large");
    const size_t smem = ((size_t)rc * (IA_KT + 1) + IA_KT * IA_TN + 8 * IA_TN) * sizeof(float);
    dim3 grid(div_up(N, IA_TN), B);
    cudaStream_t st = (cudaStream_t)stream;
    const int rows = div_up(rc, 8);
    if (rows <= 4) ia_attention_kernel<4><<<grid, IA_THREADS, smem, st>>>(ic, pc, rc, N, img, pt, w12, b12, w3, b3, att);
    else if (rows <= 8) ia_attention_kernel<8><<<grid, IA_THREADS, smem, st>>>(ic, pc, rc, N, img, pt, w12, b12, w3, b3, att);
    else if (rows <= 16) ia_attention_kernel<16><<<grid, IA_THREADS, smem, st>>>(ic, pc, rc, N, img, pt, w12, b12, w3, b3, att);
    else ia_attention_kernel<32><<<grid, IA_THREADS, smem, st>>>(ic, pc, rc, N, img, pt, w12, b12, w3, b3, att);
    return check_launch("ia_attention");
}
