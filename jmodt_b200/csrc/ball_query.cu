// Ball query for sm_100a.
//
// Replaces ball_query_kernel_fast (reference jmodt/ops/pointnet2/src/ball_query_gpu.cu:9-45),
// which runs ONE THREAD per centre scanning all n points serially (<= 16 CTAs busy at the
// RPN level-0 shape).  Here a warp owns CPW centres and the 32 lanes test 32 consecutive
// points per step: ballot + prefix-popcount gives every hit its slot in ascending point
// order, so the "first nsample neighbours in index order" contract is kept exactly, and a
// warp stops as soon as all its centres are full.  Point tiles are staged once per CTA in
// shared memory (AoS, stride-3 reads are bank-conflict free because gcd(3,32)=1) and shared
// by the CTA's 8 warps.
#include "common.cuh"

namespace jmb {

constexpr int BQ_WARPS = 8;
constexpr int BQ_THREADS = BQ_WARPS * 32;
constexpr int BQ_TILE = 2048;  // points per shared-memory tile (24 KB)

template <int CPW>
__global__ void __launch_bounds__(BQ_THREADS)
ball_query_kernel(int n, int m, float radius2, int nsample, const float *__restrict__ new_xyz,
                  const float *__restrict__ xyz, int *__restrict__ idx) {
    __shared__ __align__(16) float s_pts[BQ_TILE * 3];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;
    const int c0 = (blockIdx.x * BQ_WARPS + warp) * CPW;

    const float *pts = xyz + (size_t)b * n * 3;
    float cx[CPW], cy[CPW], cz[CPW];
    int cnt[CPW], first[CPW];
    int open = 0;  // centres of this warp still collecting
#pragma unroll
    for (int i = 0; i < CPW; ++i) {
        const int c = c0 + i;
        first[i] = 0;
        if (c < m) {
            const float *p = new_xyz + ((size_t)b * m + c) * 3;
            cx[i] = __ldg(p); cy[i] = __ldg(p + 1); cz[i] = __ldg(p + 2);
            cnt[i] = 0;
            ++open;
        } else {
            cx[i] = cy[i] = cz[i] = 0.f;
            cnt[i] = nsample;  // nothing to do
        }
    }
    if (nsample <= 0) open = 0;

    for (int t0 = 0; t0 < n; t0 += BQ_TILE) {
        const int tn = min(BQ_TILE, n - t0);
        const float *src = pts + (size_t)t0 * 3;
        const int nf = tn * 3;
        if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
            const float4 *src4 = reinterpret_cast<const float4 *>(src);
            float4 *dst4 = reinterpret_cast<float4 *>(s_pts);
            for (int f = threadIdx.x; f < (nf >> 2); f += BQ_THREADS) dst4[f] = __ldg(src4 + f);
            for (int f = (nf & ~3) + threadIdx.x; f < nf; f += BQ_THREADS) s_pts[f] = __ldg(src + f);
        } else {
            for (int f = threadIdx.x; f < nf; f += BQ_THREADS) s_pts[f] = __ldg(src + f);
        }
        __syncthreads();

        if (open > 0) {
            for (int p0 = 0; p0 < tn; p0 += 32) {
                const int p = p0 + (int)lane;
                const bool valid = p < tn;
                const int pp = valid ? p : 0;
                const float x = s_pts[pp * 3], y = s_pts[pp * 3 + 1], z = s_pts[pp * 3 + 2];
#pragma unroll
                for (int i = 0; i < CPW; ++i) {
                    if (cnt[i] < nsample) {  // warp-uniform
                        const float d2 = dist2_ref(cx[i] - x, cy[i] - y, cz[i] - z);
                        const bool hit = valid && (d2 < radius2);
                        const unsigned mk = __ballot_sync(0xffffffffu, hit);
                        if (mk) {
                            if (cnt[i] == 0) first[i] = t0 + p0 + __ffs(mk) - 1;
                            const int pos = cnt[i] + __popc(mk & lt_mask);
                            if (hit && pos < nsample)
                                idx[((size_t)b * m + c0 + i) * nsample + pos] = t0 + p;
                            cnt[i] += __popc(mk);
                            if (cnt[i] >= nsample) { cnt[i] = nsample; --open; }
                        }
                    }
                }
                if (open == 0) break;
            }
        }
        if (!__syncthreads_or(open > 0)) break;
    }

    // Tail of each row: the reference pre-fills the row with the first hit (:36-40);
    // rows without any hit stay 0 (the caller's zero-init, pointnet2_utils.py:218).
#pragma unroll
    for (int i = 0; i < CPW; ++i) {
        const int c = c0 + i;
        if (c < m) {
            const int have = min(cnt[i], nsample);
            int *row = idx + ((size_t)b * m + c) * nsample;
            for (int l = have + (int)lane; l < nsample; l += 32) row[l] = first[i];
        }
    }
}

}  // namespace jmb

extern "C" int jmb_ball_query(int b, int n, int m, float radius, int nsample,
                              const float *new_xyz, const float *xyz, int *idx, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample >= 0, "ball_query: negative size");
    if (b == 0 || m == 0 || nsample == 0) return JMB_OK;
    JMB_REQUIRE(new_xyz && xyz && idx, "ball_query: null pointer");
    JMB_REQUIRE(b <= 65535, "ball_query: batch %d exceeds grid.y limit", b);
    const float radius2 = radius * radius;  // fp32, as ball_query_gpu.cu:23
    cudaStream_t st = (cudaStream_t)stream;
    const long long centres = (long long)b * m;
    // more centres per warp amortise the shared-memory reads; fewer keep small problems parallel
    if (centres >= 4 * 4096) {
        dim3 grid(div_up(m, BQ_WARPS * 4), b);
        ball_query_kernel<4><<<grid, BQ_THREADS, 0, st>>>(n, m, radius2, nsample, new_xyz, xyz, idx);
    } else if (centres >= 2 * 2048) {
        dim3 grid(div_up(m, BQ_WARPS * 2), b);
        ball_query_kernel<2><<<grid, BQ_THREADS, 0, st>>>(n, m, radius2, nsample, new_xyz, xyz, idx);
    } else {
        dim3 grid(div_up(m, BQ_WARPS), b);
        ball_query_kernel<1><<<grid, BQ_THREADS, 0, st>>>(n, m, radius2, nsample, new_xyz, xyz, idx);
    }
    return check_launch("ball_query");
}
