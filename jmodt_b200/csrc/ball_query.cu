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


// ---- cell-list variant for large clouds ---------------------------------------------------------------------------
// The scan above tests every point against every centre: 8 x 4 096 x 16 384 = 537 M distance tests for RPN level 0, 0.86 ms
// per batch, almost all of them misses (the 0.1 m ball rarely fills, so there is no early exit).  Here the points of a
// frame are bucketed once by a hashed uniform grid of cell size >= the largest radius; a centre then only visits the 27
// neighbouring cells.  The CONTRACT is unchanged — the first nsample hits in ascending point index, the row padded with
// the first hit — because the hit set is the same (same d2 arithmetic, every point within the radius lies in one of the
// 27 cells) and the hits are put in index order before they are written (rank sort in shared memory).  A centre whose
// ball holds more than BG_CAP points falls back to the ordered scan for that row (dense balls fill quickly).
constexpr int BG_BUCKETS = 16384;          // per frame (power of two)
constexpr int BG_CAP = 256;                // hits buffered per (centre, radius)
constexpr int BG_WARPS = 8;
constexpr int BG_START_PITCH = (BG_BUCKETS + 4) & ~3;   // ints per frame of the bucket-start table (16-byte multiple)

__device__ __forceinline__ unsigned bg_hash(int cx, int cy, int cz) {
    return ((unsigned)cx * 73856093u ^ (unsigned)cy * 19349663u ^ (unsigned)cz * 83492791u) & (BG_BUCKETS - 1);
}
__device__ __forceinline__ int bg_cell(float v, float inv_cs) { return (int)floorf(v * inv_cs); }

// one CTA per frame: histogram -> exclusive scan -> scatter (order inside a bucket is arbitrary: hits are sorted later)
__global__ void __launch_bounds__(1024)
bq_build_grid_kernel(int n, float inv_cs, const float *__restrict__ xyz, int *__restrict__ start, float4 *__restrict__ list) {
    extern __shared__ int bg_hist[];            // [BG_BUCKETS]
    __shared__ int s_warp_sum[32];
    const int b = blockIdx.x;
    const float *pts = xyz + (size_t)b * n * 3;
    int *st = start + (size_t)b * BG_START_PITCH;
    float4 *ls = list + (size_t)b * n;       // bucket-ordered copies (x, y, z, index): a candidate is ONE 16-byte load
    for (int i = threadIdx.x; i < BG_BUCKETS; i += 1024) bg_hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 1024)
        atomicAdd(&bg_hist[bg_hash(bg_cell(__ldg(pts + 3 * i), inv_cs), bg_cell(__ldg(pts + 3 * i + 1), inv_cs),
                                   bg_cell(__ldg(pts + 3 * i + 2), inv_cs))], 1);
    __syncthreads();
    // exclusive scan: 16 consecutive buckets per thread
    constexpr int PER = BG_BUCKETS / 1024;
    int local[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) { local[j] = bg_hist[threadIdx.x * PER + j]; sum += local[j]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp_sum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        s_warp_sum[lane] = wi - w;
    }
    __syncthreads();
    int run = s_warp_sum[warp] + incl - sum;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        st[threadIdx.x * PER + j] = run;
        bg_hist[threadIdx.x * PER + j] = run;       // becomes the scatter cursor
        run += local[j];
    }
    if (threadIdx.x == 1023) st[BG_BUCKETS] = run;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 1024) {
        const float x = __ldg(pts + 3 * i), y = __ldg(pts + 3 * i + 1), z = __ldg(pts + 3 * i + 2);
        const int pos = atomicAdd(&bg_hist[bg_hash(bg_cell(x, inv_cs), bg_cell(y, inv_cs), bg_cell(z, inv_cs))], 1);
        ls[pos] = make_float4(x, y, z, __int_as_float(i));
    }
}

template <int NR>
__global__ void __launch_bounds__(BG_WARPS * 32)
ball_query_grid_kernel(int n, int m, float inv_cs, BallQueryScales sc, const float *__restrict__ new_xyz,
                       const float *__restrict__ xyz, const int *__restrict__ start, const float4 *__restrict__ list) {
    __shared__ int s_hits[BG_WARPS][NR][BG_CAP];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;
    const int c = blockIdx.x * BG_WARPS + warp;
    if (c >= m) return;                                   // whole warp
    const float *pts = xyz + (size_t)b * n * 3;
    const int *st = start + (size_t)b * BG_START_PITCH;
    const float4 *ls = list + (size_t)b * n;
    const float *cp = new_xyz + ((size_t)b * m + c) * 3;
    const float cx = __ldg(cp), cy = __ldg(cp + 1), cz = __ldg(cp + 2);
    // the 27 neighbouring cells -> distinct buckets (hash collisions among them are visited once)
    const int ix = bg_cell(cx, inv_cs), iy = bg_cell(cy, inv_cs), iz = bg_cell(cz, inv_cs);
    unsigned bucket = 0xffffffffu;
    if (lane < 27) bucket = bg_hash(ix + (int)(lane % 3) - 1, iy + (int)((lane / 3) % 3) - 1, iz + (int)(lane / 9) - 1);
    const unsigned same = __match_any_sync(0xffffffffu, bucket);
    unsigned leaders = __ballot_sync(0xffffffffu, lane < 27 && (unsigned)(__ffs(same) - 1) == lane);
    int cnt[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) cnt[r] = 0;
    while (leaders) {
        const int j = __ffs(leaders) - 1;
        leaders &= leaders - 1;
        const unsigned bk = __shfl_sync(0xffffffffu, bucket, j);
        const int q0 = __ldg(st + bk), q1 = __ldg(st + bk + 1);
        for (int q = q0; q < q1; q += 32) {
            const bool valid = q + (int)lane < q1;
            int k = 0;
            float d2 = 3.0e38f;
            if (valid) {
                const float4 e = __ldg(ls + q + lane);
                k = __float_as_int(e.w);
                d2 = dist2_ref(cx - e.x, cy - e.y, cz - e.z);
            }
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const bool hit = valid && d2 < sc.radius2[r];
                const unsigned mk = __ballot_sync(0xffffffffu, hit);
                const int pos = cnt[r] + __popc(mk & lt_mask);
                if (hit && pos < BG_CAP) s_hits[warp][r][pos] = k;
                cnt[r] += __popc(mk);
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int ns = sc.nsample[r];
        int *row = sc.idx[r] + ((size_t)b * m + c) * ns;
        if (cnt[r] <= BG_CAP) {
            // rank sort: indices are distinct, so the rank of a hit is the number of smaller hits
            const int *h = s_hits[warp][r];
            int first = 0x7fffffff;
            for (int e = lane; e < cnt[r]; e += 32) {
                const int v = h[e];
                int rank = 0;
                for (int t = 0; t < cnt[r]; ++t) rank += h[t] < v;
                if (rank < ns) row[rank] = v;
                first = min(first, v);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
            if (cnt[r] == 0) first = 0;                     // no hit: the row stays 0 (the caller's zero-init in the reference)
            for (int l = min(cnt[r], ns) + (int)lane; l < ns; l += 32) row[l] = first;
        } else {
            // dense ball: ordered scan of the cloud with early exit (the contract of the kernel above, one row)
            int have = 0, first = 0;
            for (int p0 = 0; p0 < n && have < ns; p0 += 32) {
                const int pidx = p0 + (int)lane;
                bool hit = false;
                if (pidx < n) {
                    const float *p = pts + (size_t)pidx * 3;
                    hit = dist2_ref(cx - __ldg(p), cy - __ldg(p + 1), cz - __ldg(p + 2)) < sc.radius2[r];
                }
                const unsigned mk = __ballot_sync(0xffffffffu, hit);
                if (mk) {
                    if (have == 0) first = p0 + __ffs(mk) - 1;
                    const int pos = have + __popc(mk & lt_mask);
                    if (hit && pos < ns) row[pos] = pidx;
                    have += __popc(mk);
                }
            }
            for (int l = min(have, ns) + (int)lane; l < ns; l += 32) row[l] = first;
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

extern "C" size_t jmb_ball_query_grid_workspace_bytes(int b, int n) {
    if (b <= 0 || n <= 0) return 0;
    // per frame: bucket starts (BG_BUCKETS + 1 ints, padded to 16 bytes) + one float4 per point
    return (size_t)b * ((((size_t)jmb::BG_BUCKETS + 4) & ~(size_t)3) * sizeof(int) + (size_t)n * sizeof(float4));
}

// Same contract as jmb_ball_query_msg2 (radius_b / nsample_b / idx_b may be 0 / 0 / NULL for a single radius), through a
// hashed cell list built per call in `workspace` (jmb_ball_query_grid_workspace_bytes).  Worth it for large clouds and radii
// that are small against their extent (RPN level 0: 16 384 points, 0.1 / 0.5 m).
extern "C" int jmb_ball_query_msg2_grid(int b, int n, int m, float radius_a, int nsample_a, float radius_b, int nsample_b,
                                        const float *new_xyz, const float *xyz, int *idx_a, int *idx_b, void *workspace,
                                        size_t workspace_bytes, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample_a > 0 && nsample_b >= 0, "ball_query_grid: bad sizes");
    if (b == 0 || m == 0) return JMB_OK;
    JMB_REQUIRE(n > 0 && new_xyz && xyz && idx_a && (nsample_b == 0 || idx_b), "ball_query_grid: null pointer");
    JMB_REQUIRE(b <= 65535, "ball_query_grid: batch %d exceeds grid.y limit", b);
    const size_t need = jmb_ball_query_grid_workspace_bytes(b, n);
    if (!workspace || workspace_bytes < need) {
        set_error("ball_query_grid: workspace of %zu bytes required", need);
        return JMB_ERR_WORKSPACE;
    }
    const float rmax = fmaxf(radius_a, nsample_b > 0 ? radius_b : 0.f);
    JMB_REQUIRE(rmax > 0.f, "ball_query_grid: radius must be positive");
    const float inv_cs = 1.0f / (rmax * 1.001f);       // cell a little larger than the largest radius: rounding of v * inv_cs
                                                       // can then never put two points within the radius two cells apart
    JMB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, "ball_query_grid: workspace must be 16-byte aligned");
    int *start = static_cast<int *>(workspace);
    float4 *list = reinterpret_cast<float4 *>(start + (size_t)b * BG_START_PITCH);
    cudaStream_t st = (cudaStream_t)stream;
    const int hist_bytes = BG_BUCKETS * (int)sizeof(int);
    {
        int dev = 0, sms = 0;
        const int rc = device_info(&dev, &sms);
        if (rc != JMB_OK) return rc;
        JMB_FUNC_ATTR_ONCE(bq_build_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_bytes, dev);
    }
    bq_build_grid_kernel<<<b, 1024, hist_bytes, st>>>(n, inv_cs, xyz, start, list);
    int rc = check_launch("ball_query_grid(build)");
    if (rc != JMB_OK) return rc;
    BallQueryScales sc;
    sc.radius2[0] = radius_a * radius_a; sc.nsample[0] = nsample_a; sc.idx[0] = idx_a;
    sc.radius2[1] = radius_b * radius_b; sc.nsample[1] = nsample_b; sc.idx[1] = idx_b;
    dim3 grid(div_up(m, BG_WARPS), b);
    if (nsample_b > 0) ball_query_grid_kernel<2><<<grid, BG_WARPS * 32, 0, st>>>(n, m, inv_cs, sc, new_xyz, xyz, start, list);
    else ball_query_grid_kernel<1><<<grid, BG_WARPS * 32, 0, st>>>(n, m, inv_cs, sc, new_xyz, xyz, start, list);
    return check_launch("ball_query_grid");
}
