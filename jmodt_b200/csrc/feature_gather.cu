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
// launches whose GEMMs are far below one tensor-core tile wave.  One SIMT kernel in plain fp32: a thread owns ONE point
// and all rc rows in registers; its K = ic + pc inputs are read straight from the channel-first tensors (a warp reads 32
// consecutive points of a channel row: coalesced, no staging), [W1 | W2] sits transposed in shared memory and is read as
// warp-wide broadcasts; tanh, the fc3 dot product and the sigmoid are thread-local.  No barrier after the weight load.
// The first version (a CTA of 32 points walking K in 32-deep shared-memory tiles with two barriers per tile) took 106 us
// for 131 072 points where the inputs amount to 84 MB.
// w12 (rc, ic + pc) row-major, b12 (rc) = b1 + b2, w3 (rc), b3 scalar; img (B, ic, N), pt (B, pc, N) channel-first.
namespace jmb {

constexpr int IA_THREADS = 128;

template <int RC>      // rows held per thread: rc <= RC
__global__ void __launch_bounds__(IA_THREADS)
ia_attention_kernel(int ic, int pc, int rc, int N, const float *__restrict__ img, const float *__restrict__ pt,
                    const float *__restrict__ w12, const float *__restrict__ b12, const float *__restrict__ w3,
                    float b3, float *__restrict__ att) {
    extern __shared__ __align__(16) float ia_smem[];      // [K][RC] transposed weights, rows >= rc zero
    const int K = ic + pc;
    for (int e = threadIdx.x; e < K * RC; e += IA_THREADS) {
        const int k = e / RC, r = e - k * RC;
        ia_smem[e] = r < rc ? __ldg(w12 + (size_t)r * K + k) : 0.f;
    }
    __syncthreads();
    const int b = blockIdx.y, n = blockIdx.x * IA_THREADS + threadIdx.x;
    if (n >= N) return;
    float acc[RC];
#pragma unroll
    for (int r = 0; r < RC; ++r) acc[r] = r < rc ? __ldg(b12 + r) : 0.f;
    const float *x = img + (size_t)b * ic * N + n;
    const float *w = ia_smem;
    for (int half = 0; half < 2; ++half) {
        const int kn = half == 0 ? ic : pc;
        int k = 0;
        for (; k + 8 <= kn; k += 8) {          // eight independent loads in flight: the loop is bound by their latency
            float xv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) xv[u] = __ldg(x + (size_t)(k + u) * N);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int r4 = 0; r4 < RC; r4 += 4) {
                    const float4 wv = *reinterpret_cast<const float4 *>(w + (k + u) * RC + r4);
                    acc[r4 + 0] = fmaf(wv.x, xv[u], acc[r4 + 0]); acc[r4 + 1] = fmaf(wv.y, xv[u], acc[r4 + 1]);
                    acc[r4 + 2] = fmaf(wv.z, xv[u], acc[r4 + 2]); acc[r4 + 3] = fmaf(wv.w, xv[u], acc[r4 + 3]);
                }
            }
        }
        for (; k < kn; ++k) {
            const float x0 = __ldg(x + (size_t)k * N);
#pragma unroll
            for (int r = 0; r < RC; ++r) acc[r] = fmaf(w[k * RC + r], x0, acc[r]);
        }
        x = pt + (size_t)b * pc * N + n;
        w += (size_t)ic * RC;
    }
    float s = b3;
#pragma unroll
    for (int r = 0; r < RC; ++r)
        if (r < rc) s = fmaf(__ldg(w3 + r), tanhf(acc[r]), s);
    att[(size_t)b * N + n] = 1.f / (1.f + expf(-s));
}

template <int RC>
int launch_ia_attention(int B, int ic, int pc, int rc, int N, const float *img, const float *pt, const float *w12,
                        const float *b12, const float *w3, float b3, float *att, cudaStream_t st) {
    const int smem = (ic + pc) * RC * (int)sizeof(float);
    int dev = 0, sms = 0;
    {
        const int rcode = device_info(&dev, &sms);
        if (rcode != JMB_OK) return rcode;
    }
    if (smem > 48 * 1024)
        JMB_FUNC_ATTR_ONCE(ia_attention_kernel<RC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024, dev);
    dim3 grid(div_up(N, IA_THREADS), B);
    ia_attention_kernel<RC><<<grid, IA_THREADS, smem, st>>>(ic, pc, rc, N, img, pt, w12, b12, w3, b3, att);
    return check_launch("ia_attention");
}

}  // namespace jmb

extern "C" int jmb_ia_attention(int B, int ic, int pc, int rc, int N, const float *img, const float *pt,
                                const float *w12, const float *b12, const float *w3, float b3, float *att,
                                void *stream) {
    using namespace jmb;
    JMB_REQUIRE(B >= 0 && ic > 0 && pc > 0 && rc > 0 && N >= 0, "ia_attention: bad sizes");
    if (B == 0 || N == 0) return JMB_OK;
    JMB_REQUIRE(img && pt && w12 && b12 && w3 && att, "ia_attention: null pointer");
    JMB_REQUIRE(rc <= 64, "ia_attention: at most 64 reduced channels (got %d)", rc);
    JMB_REQUIRE(B <= 65535, "ia_attention: batch too large");
    const int RCp = rc <= 16 ? 16 : rc <= 32 ? 32 : 64;
    JMB_REQUIRE((long long)(ic + pc) * RCp * 4 <= 200 * 1024, "ia_attention: ic + pc = %d too large for %d reduced channels",
                ic + pc, rc);
    cudaStream_t st = (cudaStream_t)stream;
    if (RCp == 16) return launch_ia_attention<16>(B, ic, pc, rc, N, img, pt, w12, b12, w3, b3, att, st);
    if (RCp == 32) return launch_ia_attention<32>(B, ic, pc, rc, N, img, pt, w12, b12, w3, b3, att, st);
    return launch_ia_attention<64>(B, ic, pc, rc, N, img, pt, w12, b12, w3, b3, att, st);
}
