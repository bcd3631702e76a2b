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

// NR radii (multi-scale grouping queries the same centres with several radii, pointnet2_modules.py:41-42) share
// one scan of the cloud: the distance is computed once per (centre, point) and compared against each radius.
struct BallQueryScales {
    float radius2[2];
    int nsample[2];
    int *idx[2];
};

template <int CPW, int NR>
__global__ void __launch_bounds__(BQ_THREADS)
ball_query_kernel(int n, int m, BallQueryScales sc, const float *__restrict__ new_xyz,
                  const float *__restrict__ xyz) {
    __shared__ __align__(16) float s_pts[BQ_TILE * 3];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;
    const int c0 = (blockIdx.x * BQ_WARPS + warp) * CPW;

    const float *pts = xyz + (size_t)b * n * 3;
    float cx[CPW], cy[CPW], cz[CPW];
    int cnt[CPW][NR], first[CPW][NR];
    int open = 0;  // (centre, radius) rows of this warp still collecting
#pragma unroll
    for (int i = 0; i < CPW; ++i) {
        const int c = c0 + i;
        if (c < m) {
            const float *p = new_xyz + ((size_t)b * m + c) * 3;
            cx[i] = __ldg(p); cy[i] = __ldg(p + 1); cz[i] = __ldg(p + 2);
        } else {
            cx[i] = cy[i] = cz[i] = 0.f;
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            first[i][r] = 0;
            cnt[i][r] = (c < m && sc.nsample[r] > 0) ? 0 : sc.nsample[r];
            if (cnt[i][r] < sc.nsample[r]) ++open;
        }
    }

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
                    bool any_open = false;
#pragma unroll
                    for (int r = 0; r < NR; ++r) any_open |= cnt[i][r] < sc.nsample[r];
                    if (any_open) {  // warp-uniform
                        const float d2 = dist2_ref(cx[i] - x, cy[i] - y, cz[i] - z);
#pragma unroll
                        for (int r = 0; r < NR; ++r) {
                            if (cnt[i][r] < sc.nsample[r]) {
                                const bool hit = valid && (d2 < sc.radius2[r]);
                                const unsigned mk = __ballot_sync(0xffffffffu, hit);
                                if (mk) {
                                    if (cnt[i][r] == 0) first[i][r] = t0 + p0 + __ffs(mk) - 1;
                                    const int pos = cnt[i][r] + __popc(mk & lt_mask);
                                    if (hit && pos < sc.nsample[r])
                                        sc.idx[r][((size_t)b * m + c0 + i) * sc.nsample[r] + pos] = t0 + p;
                                    cnt[i][r] += __popc(mk);
                                    if (cnt[i][r] >= sc.nsample[r]) { cnt[i][r] = sc.nsample[r]; --open; }
                                }
                            }
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
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const int have = min(cnt[i][r], sc.nsample[r]);
                int *row = sc.idx[r] + ((size_t)b * m + c) * sc.nsample[r];
                for (int l = have + (int)lane; l < sc.nsample[r]; l += 32) row[l] = first[i][r];
            }
        }
    }
}

template <int NR>
static int launch_ball_query(int b, int n, int m, const BallQueryScales &sc, const float *new_xyz, const float *xyz,
                             cudaStream_t st) {
    const long long centres = (long long)b * m;
    // more centres per warp amortise the shared-memory reads; fewer keep small problems parallel
    if (centres >= 4 * 4096) {
        dim3 grid(div_up(m, BQ_WARPS * 4), b);
        ball_query_kernel<4, NR><<<grid, BQ_THREADS, 0, st>>>(n, m, sc, new_xyz, xyz);
    } else if (centres >= 2 * 2048) {
        dim3 grid(div_up(m, BQ_WARPS * 2), b);
        ball_query_kernel<2, NR><<<grid, BQ_THREADS, 0, st>>>(n, m, sc, new_xyz, xyz);
    } else {
        dim3 grid(div_up(m, BQ_WARPS), b);
        ball_query_kernel<1, NR><<<grid, BQ_THREADS, 0, st>>>(n, m, sc, new_xyz, xyz);
    }
    return check_launch("ball_query");
}

}  // namespace jmb

extern "C" int jmb_ball_query(int b, int n, int m, float radius, int nsample,
                              const float *new_xyz, const float *xyz, int *idx, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample >= 0, "ball_query: negative size");
    if (b == 0 || m == 0 || nsample == 0) return JMB_OK;
    JMB_REQUIRE(new_xyz && xyz && idx, "ball_query: null pointer");
    JMB_REQUIRE(b <= 65535, "ball_query: batch %d exceeds grid.y limit", b);
    BallQueryScales sc;
    sc.radius2[0] = radius * radius;  // fp32, as ball_query_gpu.cu:23
    sc.nsample[0] = nsample; sc.idx[0] = idx;
    sc.radius2[1] = 0.f; sc.nsample[1] = 0; sc.idx[1] = nullptr;
    return launch_ball_query<1>(b, n, m, sc, new_xyz, xyz, (cudaStream_t)stream);
}

extern "C" int jmb_ball_query_msg2(int b, int n, int m, float radius_a, int nsample_a, float radius_b, int nsample_b,
                                   const float *new_xyz, const float *xyz, int *idx_a, int *idx_b, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample_a > 0 && nsample_b > 0, "ball_query_msg2: bad sizes");
    if (b == 0 || m == 0) return JMB_OK;
    JMB_REQUIRE(new_xyz && xyz && idx_a && idx_b, "ball_query_msg2: null pointer");
    JMB_REQUIRE(b <= 65535, "ball_query_msg2: batch %d exceeds grid.y limit", b);
    BallQueryScales sc;
    sc.radius2[0] = radius_a * radius_a; sc.nsample[0] = nsample_a; sc.idx[0] = idx_a;
    sc.radius2[1] = radius_b * radius_b; sc.nsample[1] = nsample_b; sc.idx[1] = idx_b;
    return launch_ball_query<2>(b, n, m, sc, new_xyz, xyz, (cudaStream_t)stream);
}
