// Pair correlation features of the link / start-end heads, sm_100a.
//
// Replaces the tensor glue of the reference tracker's affinity block (jmodt/tracking/tracker.py:81-112; training
// twin jmodt/detection/modeling/rcnn.py:239-258):
//     cor = |pred.unsqueeze(1).repeat(1, D, 1) - det.unsqueeze(0).repeat(P, 1, 1)|         (P, D, 512) materialised
//     start_in = cor.mean(dim=0), end_in = cor.mean(dim=1)
// = two repeats, a subtraction, an abs and two reductions that each re-read the 33.5 MB tensor.  Here one CTA owns
// one (pair, channel): the channel's P predecessor and D successor values sit in shared memory, the P x D tile of
// |p_i - d_j| is written once (coalesced, channel-first, the layout the link MLP's operand staging reads), and the
// two means are produced from shared memory in the same launch.
#include "common.cuh"

namespace jmb {

__global__ void __launch_bounds__(256)
pair_corr_kernel(int K, int P, int D, const float *__restrict__ pt, const float *__restrict__ dt,
                 float *__restrict__ cor, float *__restrict__ mean_over_p, float *__restrict__ mean_over_d) {
    extern __shared__ float pc_smem[];
    float *sa = pc_smem, *sb = pc_smem + P;
    const int k = blockIdx.x, g = blockIdx.y;
    const float *a = pt + ((size_t)g * K + k) * P;
    const float *b = dt + ((size_t)g * K + k) * D;
    for (int i = threadIdx.x; i < P; i += blockDim.x) sa[i] = __ldg(a + i);
    for (int j = threadIdx.x; j < D; j += blockDim.x) sb[j] = __ldg(b + j);
    __syncthreads();
    if (cor) {
        float *out = cor + ((size_t)g * K + k) * (size_t)P * D;
        const int total = P * D;
        int i = threadIdx.x / D, j = threadIdx.x - i * D;       // element e = i * D + j, advanced incrementally
        const int di = blockDim.x / D, dj = blockDim.x - di * D;
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            out[e] = fabsf(__fsub_rn(sa[i], sb[j]));
            i += di; j += dj;
            if (j >= D) { j -= D; ++i; }
        }
    }
    // means: sequential fp32 sums over the other index, then one division (torch: sum / count)
    if (mean_over_p) {
        for (int j = threadIdx.x; j < D; j += blockDim.x) {
            const float bj = sb[j];
            float s = 0.f;
            for (int i = 0; i < P; ++i) s += fabsf(__fsub_rn(sa[i], bj));
            mean_over_p[((size_t)g * K + k) * D + j] = s / (float)P;
        }
    }
    if (mean_over_d) {
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            const float ai = sa[i];
            float s = 0.f;
            for (int j = 0; j < D; ++j) s += fabsf(__fsub_rn(ai, sb[j]));
            mean_over_d[((size_t)g * K + k) * P + i] = s / (float)D;
        }
    }
}

}  // namespace jmb

extern "C" int jmb_pair_corr(int G, int K, int P, int D, const float *pt, const float *dt, float *cor,
                             float *mean_over_p, float *mean_over_d, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(G >= 0 && K >= 0 && P >= 0 && D >= 0, "pair_corr: negative size");
    if (G == 0 || K == 0 || P == 0 || D == 0) return JMB_OK;
    JMB_REQUIRE(pt && dt, "pair_corr: null pointer");
    JMB_REQUIRE(G <= 65535, "pair_corr: too many pairs");
    JMB_REQUIRE((long long)P * D < (1LL << 31), "pair_corr: pair matrix too large");
    const size_t smem = (size_t)(P + D) * sizeof(float);
    JMB_REQUIRE(smem <= 48 * 1024, "pair_corr: P + D = %d exceeds the shared-memory tile", P + D);
    pair_corr_kernel<<<dim3(K, G), 256, smem, (cudaStream_t)stream>>>(K, P, D, pt, dt, cor, mean_over_p, mean_over_d);
    return check_launch("pair_corr");
}

// ---- point-feature packing for RoI pooling ---------------------------------------------------------------------
// The reference builds the per-point feature vector of the RoI-pooling stage with
//     pts_feature = torch.cat((seg_mask.unsqueeze(2), depth.unsqueeze(2), rpn_features), dim=2)
// (proposal_target_layer.py:17-34) where rpn_features is backbone_features.permute(0, 2, 1) (point_rcnn.py:47): a
// strided gather of the channel-first (B, C, N) tensor, 32 bytes fetched per 4 bytes used.  Here a CTA transposes a
// 32-point x C-channel tile through shared memory: coalesced reads along the points, coalesced row writes.
namespace jmb {

__global__ void __launch_bounds__(256)
pack_point_features_kernel(int C, int N, int E, const float *__restrict__ feat, const float *__restrict__ e0,
                           const float *__restrict__ e1, float *__restrict__ out) {
    extern __shared__ float pk_tile[];                 // [32][C + 1]
    const int b = blockIdx.y, n0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *src = feat + (size_t)b * C * N;
    const int pitch = C + 1;
    const int n = n0 + lane;
    for (int c = warp; c < C; c += 8) pk_tile[lane * pitch + c] = n < N ? __ldg(src + (size_t)c * N + n) : 0.f;
    __syncthreads();
    const int row = E + C;
    for (int r = warp; r < 32; r += 8) {
        const int p = n0 + r;
        if (p >= N) break;
        float *dst = out + ((size_t)b * N + p) * row;
        if (lane == 0 && E > 0) dst[0] = __ldg(e0 + (size_t)b * N + p);
        if (lane == 1 && E > 1) dst[1] = __ldg(e1 + (size_t)b * N + p);
        for (int c = lane; c < C; c += 32) dst[E + c] = pk_tile[r * pitch + c];
    }
}

}  // namespace jmb

extern "C" int jmb_pack_point_features(int B, int C, int N, int E, const float *feat, const float *extra0,
                                       const float *extra1, float *out, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(B >= 0 && C >= 0 && N >= 0 && E >= 0 && E <= 2, "pack_point_features: bad sizes");
    if (B == 0 || N == 0) return JMB_OK;
    JMB_REQUIRE((feat || C == 0) && out && (E < 1 || extra0) && (E < 2 || extra1), "pack_point_features: null pointer");
    JMB_REQUIRE(B <= 65535, "pack_point_features: batch too large");
    const size_t smem = (size_t)32 * (C + 1) * sizeof(float);
    JMB_REQUIRE(smem <= 48 * 1024, "pack_point_features: %d channels exceed the shared-memory tile", C);
    pack_point_features_kernel<<<dim3(div_up(N, 32), B), 256, smem, (cudaStream_t)stream>>>(C, N, E, feat, extra0, extra1, out);
    return check_launch("pack_point_features");
}
