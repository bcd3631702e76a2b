// three_nn / three_interpolate (+grad) and group_points (+grad) for sm_100a.
//
// Replaces three_nn_kernel_fast, three_interpolate_kernel_fast, three_interpolate_grad_kernel_fast
// (reference jmodt/ops/pointnet2/src/interpolate_gpu.cu:9-52, 77-97, 120-142) and
// group_points_kernel_fast / group_points_grad_kernel_fast (group_points_gpu.cu:47-66, 8-25).
//
// three_nn: the reference keeps its running bests in fp64 (:30) purely as storage for fp32
// values; comparing the widened floats is the same as comparing floats, so fp32 registers
// give identical results (1e40 becomes +inf, which is also what the reference stores when
// fewer than three known points exist).  Selection is the three smallest (d2, index) pairs
// in lexicographic order — strict '<' while scanning in ascending index order (:37-48) —
// so the known set can be split between SPLIT lanes and merged exactly.
#include "common.cuh"

namespace jmb {

constexpr int NN_THREADS = 128;
constexpr int NN_TILE = 1024;  // known points per shared-memory tile (16 KB as float4)

struct Top3 {
    float d1, d2, d3;
    int i1, i2, i3;
};

// insert (d,k) with k larger than every index already present (ascending scan)
__device__ __forceinline__ void top3_push(Top3 &t, float d, int k) {
    if (d < t.d3) {
        if (d < t.d1) {
            t.d3 = t.d2; t.i3 = t.i2; t.d2 = t.d1; t.i2 = t.i1; t.d1 = d; t.i1 = k;
        } else if (d < t.d2) {
            t.d3 = t.d2; t.i3 = t.i2; t.d2 = d; t.i2 = k;
        } else {
            t.d3 = d; t.i3 = k;
        }
    }
}

// lexicographic (d, idx) insert for merging partial results of different index ranges
__device__ __forceinline__ bool lex_less(float da, int ia, float db, int ib) {
    return da < db || (da == db && ia < ib);
}
__device__ __forceinline__ void top3_merge_one(Top3 &t, float d, int k) {
    if (lex_less(d, k, t.d3, t.i3)) {
        if (lex_less(d, k, t.d1, t.i1)) {
            t.d3 = t.d2; t.i3 = t.i2; t.d2 = t.d1; t.i2 = t.i1; t.d1 = d; t.i1 = k;
        } else if (lex_less(d, k, t.d2, t.i2)) {
            t.d3 = t.d2; t.i3 = t.i2; t.d2 = d; t.i2 = k;
        } else {
            t.d3 = d; t.i3 = k;
        }
    }
}

// SPLIT consecutive lanes share one unknown point; lane s scans tile entries s, s+SPLIT, ...
template <int SPLIT>
__global__ void __launch_bounds__(NN_THREADS)
three_nn_kernel(int n, int m, const float *__restrict__ unknown, const float *__restrict__ known,
                float *__restrict__ dist2, int *__restrict__ idx) {
    __shared__ float4 s_known[NN_TILE];
    const int b = blockIdx.y;
    constexpr int PTS_PER_CTA = NN_THREADS / SPLIT;
    const int sub = threadIdx.x % SPLIT;
    const int p = blockIdx.x * PTS_PER_CTA + threadIdx.x / SPLIT;
    const bool valid = p < n;
    const float *kn = known + (size_t)b * m * 3;

    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (valid) {
        const float *u = unknown + ((size_t)b * n + p) * 3;
        ux = __ldg(u); uy = __ldg(u + 1); uz = __ldg(u + 2);
    }
    Top3 t;
    t.d1 = t.d2 = t.d3 = __int_as_float(0x7f800000);  // (float)1e40
    t.i1 = t.i2 = t.i3 = 0;
    // the reference leaves index 0 in unused slots; for the lexicographic merge the padding
    // must sort after every real candidate, which +inf already guarantees (real d2 < inf).

    for (int t0 = 0; t0 < m; t0 += NN_TILE) {
        const int tn = min(NN_TILE, m - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < tn; k += NN_THREADS) {
            const float *q = kn + (size_t)(t0 + k) * 3;
            s_known[k] = make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), 0.f);
        }
        __syncthreads();
#pragma unroll 4
        for (int k = sub; k < tn; k += SPLIT) {
            const float4 q = s_known[k];
            const float d = dist2_ref(ux - q.x, uy - q.y, uz - q.z);
            top3_push(t, d, t0 + k);
        }
    }

    if (SPLIT > 1) {
        // butterfly merge across the SPLIT lanes that share this unknown point
#pragma unroll
        for (int off = 1; off < SPLIT; off <<= 1) {
            const float e1 = __shfl_xor_sync(0xffffffffu, t.d1, off);
            const float e2 = __shfl_xor_sync(0xffffffffu, t.d2, off);
            const float e3 = __shfl_xor_sync(0xffffffffu, t.d3, off);
            const int j1 = __shfl_xor_sync(0xffffffffu, t.i1, off);
            const int j2 = __shfl_xor_sync(0xffffffffu, t.i2, off);
            const int j3 = __shfl_xor_sync(0xffffffffu, t.i3, off);
            // +inf padding carries index 0 on both sides; skip it so it cannot displace
            // a real (inf, k) candidate — real distances are finite for finite inputs.
            if (e1 < __int_as_float(0x7f800000)) top3_merge_one(t, e1, j1);
            if (e2 < __int_as_float(0x7f800000)) top3_merge_one(t, e2, j2);
            if (e3 < __int_as_float(0x7f800000)) top3_merge_one(t, e3, j3);
        }
    }
    if (valid && sub == 0) {
        float *dd = dist2 + ((size_t)b * n + p) * 3;
        int *ii = idx + ((size_t)b * n + p) * 3;
        dd[0] = t.d1; dd[1] = t.d2; dd[2] = t.d3;
        ii[0] = t.i1; ii[1] = t.i2; ii[2] = t.i3;
    }
}

constexpr int TI_CH = 8;  // channels handled by one thread (idx / weight reuse)

// out[b,c,p] = fma(w2,f2, fma(w0,f0, fl(w1*f1)))  — order read from the reference SASS
__global__ void __launch_bounds__(256)
three_interpolate_kernel(int c, int m, int n, const float *__restrict__ points,
                         const int *__restrict__ idx, const float *__restrict__ weight,
                         float *__restrict__ out) {
    const int b = blockIdx.z;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int *id = idx + ((size_t)b * n + p) * 3;
    const float *w = weight + ((size_t)b * n + p) * 3;
    const int i0 = __ldg(id), i1 = __ldg(id + 1), i2 = __ldg(id + 2);
    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const int c0 = blockIdx.y * TI_CH;
#pragma unroll
    for (int cc = 0; cc < TI_CH; ++cc) {
        const int ci = c0 + cc;
        if (ci < c) {
            const float *src = points + ((size_t)b * c + ci) * m;
            const float v = __fmaf_rn(w2, __ldg(src + i2),
                                      __fmaf_rn(w0, __ldg(src + i0), __fmul_rn(w1, __ldg(src + i1))));
            out[((size_t)b * c + ci) * n + p] = v;
        }
    }
}

__global__ void __launch_bounds__(256)
three_interpolate_grad_kernel(int c, int n, int m, const float *__restrict__ grad_out,
                              const int *__restrict__ idx, const float *__restrict__ weight,
                              float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int *id = idx + ((size_t)b * n + p) * 3;
    const float *w = weight + ((size_t)b * n + p) * 3;
    const int i0 = __ldg(id), i1 = __ldg(id + 1), i2 = __ldg(id + 2);
    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const int c0 = blockIdx.y * TI_CH;
#pragma unroll
    for (int cc = 0; cc < TI_CH; ++cc) {
        const int ci = c0 + cc;
        if (ci < c) {
            const float g = __ldg(grad_out + ((size_t)b * c + ci) * n + p);
            float *dst = grad_points + ((size_t)b * c + ci) * m;
            atomicAdd(dst + i0, __fmul_rn(g, w0));
            atomicAdd(dst + i1, __fmul_rn(g, w1));
            atomicAdd(dst + i2, __fmul_rn(g, w2));
        }
    }
}

constexpr int GP_CH = 8;

__global__ void __launch_bounds__(256)
group_points_kernel(int c, int n, long long per_batch, const float *__restrict__ points,
                    const int *__restrict__ idx, float *__restrict__ out) {
    const int b = blockIdx.z;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (point, sample) pair
    if (e >= per_batch) return;
    const int src = __ldg(idx + (size_t)b * per_batch + e);
    const int c0 = blockIdx.y * GP_CH;
#pragma unroll
    for (int cc = 0; cc < GP_CH; ++cc) {
        const int ci = c0 + cc;
        if (ci < c)
            out[((size_t)b * c + ci) * per_batch + e] = __ldg(points + ((size_t)b * c + ci) * n + src);
    }
}

__global__ void __launch_bounds__(256)
group_points_grad_kernel(int c, int n, long long per_batch, const float *__restrict__ grad_out,
                         const int *__restrict__ idx, float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= per_batch) return;
    const int dst = __ldg(idx + (size_t)b * per_batch + e);
    const int c0 = blockIdx.y * GP_CH;
#pragma unroll
    for (int cc = 0; cc < GP_CH; ++cc) {
        const int ci = c0 + cc;
        if (ci < c)
            atomicAdd(grad_points + ((size_t)b * c + ci) * n + dst,
                      __ldg(grad_out + ((size_t)b * c + ci) * per_batch + e));
    }
}

}  // namespace jmb

extern "C" int jmb_three_nn(int b, int n, int m, const float *unknown, const float *known,
                            float *dist2, int *idx, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && n >= 0 && m >= 0, "three_nn: negative size");
    if (b == 0 || n == 0) return JMB_OK;
    JMB_REQUIRE(unknown && dist2 && idx && (known || m == 0), "three_nn: null pointer");
    JMB_REQUIRE(b <= 65535, "three_nn: batch %d exceeds grid.y limit", b);
    cudaStream_t st = (cudaStream_t)stream;
    // split the known set across lanes when there are too few unknown points to fill the GPU
    const long long pts = (long long)b * n;
    if (pts >= 148LL * 1024) {
        dim3 grid(div_up(n, NN_THREADS), b);
        three_nn_kernel<1><<<grid, NN_THREADS, 0, st>>>(n, m, unknown, known, dist2, idx);
    } else if (pts >= 148LL * 256) {
        dim3 grid(div_up(n, NN_THREADS / 4), b);
        three_nn_kernel<4><<<grid, NN_THREADS, 0, st>>>(n, m, unknown, known, dist2, idx);
    } else {
        dim3 grid(div_up(n, NN_THREADS / 16), b);
        three_nn_kernel<16><<<grid, NN_THREADS, 0, st>>>(n, m, unknown, known, dist2, idx);
    }
    return check_launch("three_nn");
}

extern "C" int jmb_three_interpolate(int b, int c, int m, int n, const float *points,
                                     const int *idx, const float *weight, float *out, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && c >= 0 && m >= 0 && n >= 0, "three_interpolate: negative size");
    if (b == 0 || c == 0 || n == 0) return JMB_OK;
    JMB_REQUIRE(points && idx && weight && out, "three_interpolate: null pointer");
    JMB_REQUIRE(b <= 65535 && div_up(c, TI_CH) <= 65535, "three_interpolate: grid limit");
    dim3 grid(div_up(n, 256), div_up(c, TI_CH), b);
    three_interpolate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, m, n, points, idx, weight, out);
    return check_launch("three_interpolate");
}

extern "C" int jmb_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                          const int *idx, const float *weight, float *grad_points,
                                          void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && c >= 0 && m >= 0 && n >= 0, "three_interpolate_grad: negative size");
    if (b == 0 || c == 0 || n == 0) return JMB_OK;
    JMB_REQUIRE(grad_out && idx && weight && grad_points, "three_interpolate_grad: null pointer");
    JMB_REQUIRE(b <= 65535 && div_up(c, TI_CH) <= 65535, "three_interpolate_grad: grid limit");
    dim3 grid(div_up(n, 256), div_up(c, TI_CH), b);
    three_interpolate_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, m, grad_out, idx,
                                                                          weight, grad_points);
    return check_launch("three_interpolate_grad");
}

extern "C" int jmb_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                                const int *idx, float *out, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, "group_points: negative size");
    if (b == 0 || c == 0 || npoints == 0 || nsample == 0) return JMB_OK;
    JMB_REQUIRE(points && idx && out, "group_points: null pointer");
    const long long per_batch = (long long)npoints * nsample;
    JMB_REQUIRE(b <= 65535 && div_up(c, GP_CH) <= 65535, "group_points: grid limit");
    dim3 grid((unsigned)div_up_ll(per_batch, 256), div_up(c, GP_CH), b);
    group_points_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, per_batch, points, idx, out);
    return check_launch("group_points");
}

extern "C" int jmb_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                     const float *grad_out, const int *idx, float *grad_points,
                                     void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, "group_points_grad: negative size");
    if (b == 0 || c == 0 || npoints == 0 || nsample == 0) return JMB_OK;
    JMB_REQUIRE(grad_out && idx && grad_points, "group_points_grad: null pointer");
    const long long per_batch = (long long)npoints * nsample;
    JMB_REQUIRE(b <= 65535 && div_up(c, GP_CH) <= 65535, "group_points_grad: grid limit");
    dim3 grid((unsigned)div_up_ll(per_batch, 256), div_up(c, GP_CH), b);
    group_points_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, per_batch, grad_out, idx,
                                                                     grad_points);
    return check_launch("group_points_grad");
}

// ---- first SharedMLP layer of a set-abstraction level, applied BEFORE the gather -------------------------------
// relu(W1 . [xyz_j - c ; f_j] + b1) = relu(Z_j + W1x . (xyz_j - c)),  Z = W1f . F + b1 over the level's n_pts points
// (one small dense GEMM, jmb_tc_mlp_layer with point-major output).  This kernel finishes the layer for every grouped
// neighbour and writes the (G, C1, npoint * nsample) channel-first activation the next dense layer reads — it replaces
// the grouped-gather GEMM over all npoint * nsample columns (K = 3 + C_in, e.g. 259 x 262 144 columns at RPN level 2)
// of the layer-by-layer path (reference pointnet2_utils.py:241-264 + the first Conv2d of the SharedMLP).
// CTA = 32 consecutive columns of one group: the neighbours' Z rows are staged through shared memory with coalesced
// 16-byte loads (a row is contiguous), then each warp emits whole channels as 128-byte stores.
namespace jmb {

__global__ void __launch_bounds__(256)
sa_first_layer_kernel(int C1, int npoint, int nsample, int n_pts, const float *__restrict__ z,
                      const float4 *__restrict__ w1x, const int *__restrict__ idx, const float *__restrict__ xyz,
                      const float *__restrict__ centres, float *__restrict__ out) {
    extern __shared__ __align__(16) float fl_smem[];   // C1 float4 weights, then [32][C1 + 1] rows
    const int pitch = C1 + 1;
    float4 *wx = reinterpret_cast<float4 *>(fl_smem);
    float *rows = fl_smem + 4 * C1;
    const int g = blockIdx.y;
    const int N = npoint * nsample;
    const int n0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < C1; k += blockDim.x) wx[k] = __ldg(w1x + k);
    // stage: warp w loads rows w, w + 8, ...; a row of C1 floats = C1 / 4 16-byte pieces
    const int pieces = C1 >> 2;
    for (int r = warp; r < 32; r += 8) {
        const int n = n0 + r;
        if (n >= N) break;
        const int pi = __ldg(idx + (size_t)g * N + n);
        const float4 *src = reinterpret_cast<const float4 *>(z + ((size_t)g * n_pts + pi) * C1);
        for (int c = lane; c < pieces; c += 32) {
            const float4 v = __ldg(src + c);
            float *dst = rows + r * pitch + c * 4;
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
    }
    // this lane's column: relative coordinates (pointnet2_utils.py:252)
    float dx = 0.f, dy = 0.f, dz = 0.f;
    const int n = n0 + lane;
    if (n < N) {
        const int pi = __ldg(idx + (size_t)g * N + n);
        const float *pt = xyz + ((size_t)g * n_pts + pi) * 3;
        const float *cen = centres + ((size_t)g * npoint + n / nsample) * 3;
        dx = __fsub_rn(__ldg(pt), __ldg(cen));
        dy = __fsub_rn(__ldg(pt + 1), __ldg(cen + 1));
        dz = __fsub_rn(__ldg(pt + 2), __ldg(cen + 2));
    }
    __syncthreads();
    if (n < N) {
        float *o = out + (size_t)g * C1 * N + n;
        for (int k = warp; k < C1; k += 8) {
            const float4 w = wx[k];
            const float base = z ? rows[lane * pitch + k] : w.w;
            o[(size_t)k * N] = fmaxf(fmaf(w.x, dx, fmaf(w.y, dy, fmaf(w.z, dz, base))), 0.f);
        }
    }
}

}  // namespace jmb

extern "C" int jmb_sa_first_layer(const float *z, const float *w1x, int C1, int G, int npoint, int nsample, int n_pts,
                                  const int *idx, const float *xyz, const float *centres, float *out, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(G >= 0 && C1 > 0 && npoint > 0 && nsample > 0 && n_pts > 0, "sa_first_layer: bad sizes");
    if (G == 0) return JMB_OK;
    JMB_REQUIRE(z && w1x && idx && xyz && centres && out, "sa_first_layer: null pointer");
    JMB_REQUIRE(C1 % 4 == 0 && C1 <= 1024, "sa_first_layer: width %d must be a multiple of 4, <= 1024", C1);
    JMB_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15u) == 0 && (reinterpret_cast<uintptr_t>(w1x) & 15u) == 0,
                "sa_first_layer: z and w1x must be 16-byte aligned");
    JMB_REQUIRE(G <= 65535, "sa_first_layer: too many groups");
    const long long N = (long long)npoint * nsample;
    JMB_REQUIRE(N < (1LL << 31), "sa_first_layer: too many columns");
    const size_t smem = ((size_t)32 * (C1 + 1) + 4) * sizeof(float) + (size_t)C1 * sizeof(float4);
    if (smem > 48 * 1024)
        JMB_CUDA(cudaFuncSetAttribute(sa_first_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)G);
    sa_first_layer_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(C1, npoint, nsample, n_pts, z,
                                                                    reinterpret_cast<const float4 *>(w1x), idx, xyz,
                                                                    centres, out);
    return check_launch("sa_first_layer");
}

// ---- deterministic scatter-add for the three backward ops ---------------------------------------------------------
// The reference accumulates its gradients with float atomicAdd (group_points_gpu.cu:8-25, sampling_gpu.cu:46-63,
// interpolate_gpu.cu:120-142): the summation order, and with it the last bits of the result, change from run to run.
// Here the contributions to every target are summed in a FIXED order.  The caller sorts the (batch-local) flat index
// list once (stable sort: positions of equal targets stay ascending) and passes
//     order   (B, Lq)          positions q of the flat index list, sorted by target
//     seg_off (B, n_tgt + 1)   start of every target's run in `order`
// contribution of position q: src[b][c][q / rep] * (weight ? weight[b][q] : 1)
//     group_points_grad        Lq = npoint * nsample, rep = 1
//     gather_points_grad       Lq = npoint,           rep = 1
//     three_interpolate_grad   Lq = 3 n,              rep = 3, weight = interpolation weights
namespace jmb {

__global__ void __launch_bounds__(128)
segmented_scatter_kernel(int C, int L_src, int Lq, int n_tgt, int rep, const float *__restrict__ src,
                         const int *__restrict__ order, const int *__restrict__ seg_off,
                         const float *__restrict__ weight, float *__restrict__ out) {
    const int b = blockIdx.z, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tgt) return;
    const int q0 = __ldg(seg_off + (size_t)b * (n_tgt + 1) + t), q1 = __ldg(seg_off + (size_t)b * (n_tgt + 1) + t + 1);
    const int *ord = order + (size_t)b * Lq;
    const float *w = weight ? weight + (size_t)b * Lq : nullptr;
    for (int c = blockIdx.y; c < C; c += gridDim.y) {
        const float *s = src + ((size_t)b * C + c) * L_src;
        float acc = 0.f;
        for (int q = q0; q < q1; ++q) {
            const int pos = __ldg(ord + q);
            const float v = __ldg(s + pos / rep);
            acc = __fadd_rn(acc, w ? __fmul_rn(v, __ldg(w + pos)) : v);     // product rounded, then added: the reference's atomicAdd(p, g * w)
        }
        out[((size_t)b * C + c) * n_tgt + t] = acc;
    }
}

}  // namespace jmb

extern "C" int jmb_segmented_scatter_add(int B, int C, int L_src, int Lq, int n_tgt, int rep, const float *src,
                                         const int *order, const int *seg_off, const float *weight, float *out,
                                         void *stream) {
    using namespace jmb;
    JMB_REQUIRE(B >= 0 && C >= 0 && L_src >= 0 && Lq >= 0 && n_tgt >= 0 && rep >= 1, "segmented_scatter_add: bad sizes");
    if (B == 0 || C == 0 || n_tgt == 0) return JMB_OK;
    JMB_REQUIRE(src && order && seg_off && out, "segmented_scatter_add: null pointer");
    JMB_REQUIRE(B <= 65535, "segmented_scatter_add: batch too large");
    dim3 grid(div_up(n_tgt, 128), C < 64 ? C : 64, B);
    segmented_scatter_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(C, L_src, Lq, n_tgt, rep, src, order, seg_off, weight, out);
    return check_launch("segmented_scatter_add");
}
