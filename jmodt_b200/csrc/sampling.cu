// Farthest point sampling + gather for sm_100a.
//
// Replaces farthest_point_sampling_kernel (reference jmodt/ops/pointnet2/src/sampling_gpu.cu:93-209)
// and gather_points_kernel_fast (:8-24).
//
// FPS is a chain of (m-1) dependent argmax steps.  The reference re-reads xyz and temp from
// global memory every step and reduces through a log2(bs)-deep __syncthreads tree.  Here
//   * xyz lives in shared memory for the whole kernel (SoA, conflict-free), the running
//     min-distances live in registers (PPT per thread);
//   * the block argmax is two REDUX steps per warp + ONE __syncthreads per iteration
//     (double-buffered partials), instead of up to 11 barriers.
//
// Tie-breaking must reproduce the reference bit for bit.  In the reference, thread tid
// (of bs = opt_n_threads(n) threads, cuda_utils.h:10-14) scans k = tid, tid+bs, ... with a
// strict '>' (sampling_gpu.cu:135-136), so it keeps its LOWEST k among equal maxima; the
// tree (:86-91,143-203) keeps the LEFT operand on ties, which orders threads by the
// bit-reversed tid.  The winner is therefore the point with the smallest
//       rank(k) = bitrev_{log2 bs}(k mod bs) * S + (k div bs),      S = ceil(n / bs)
// among those with maximal distance.  We lay the points out BY RANK (slot r holds the point
// of rank r; thread t owns slots t, t+T, t+2T, ...), so "strict '>' in slot order, then
// (max distance, min slot)" is exactly the reference's choice, for any thread count T.
#include "common.cuh"

#include <math.h>

namespace jmb {

// inverse of rank(): slot r -> point index k (>= n for the padding slots)
__device__ __forceinline__ int fps_slot_to_point(int r, int bs, int log2bs, int S, int n) {
    const int q = r / S;
    const int jj = r - q * S;
    if (q >= bs) return n;
    const unsigned rev = (log2bs == 0) ? 0u : (__brev((unsigned)q) >> (32 - log2bs));
    const int k = jj * bs + (int)rev;
    return k < n ? k : n;
}

template <int T, int PPT>
__global__ void __launch_bounds__(T)
fps_kernel(int n, int m, int bs, int log2bs, int S, const float *__restrict__ dataset,
           float *__restrict__ temp_out, int *__restrict__ idxs) {
    extern __shared__ __align__(16) float fps_smem[];
    constexpr int SLOTS = T * PPT;
    float *sx = fps_smem, *sy = sx + SLOTS, *sz = sy + SLOTS;
    __shared__ int s_d[2][32];
    __shared__ unsigned s_r[2][32];

    const int t = threadIdx.x;
    const unsigned lane = lane_id();
    const int warp = t >> 5;
    const float *pts = dataset + (size_t)blockIdx.x * n * 3;
    int *out = idxs + (size_t)blockIdx.x * m;

    float tmp[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int r = j * T + t;
        const int k = fps_slot_to_point(r, bs, log2bs, S, n);
        const bool valid = k < n;
        sx[r] = valid ? __ldg(pts + (size_t)k * 3) : 0.f;
        sy[r] = valid ? __ldg(pts + (size_t)k * 3 + 1) : 0.f;
        sz[r] = valid ? __ldg(pts + (size_t)k * 3 + 2) : 0.f;
        tmp[j] = valid ? 1e10f : -1.0f;  // 1e10: pointnet2_utils.py:26; -1 never beats best=-1
    }
    __syncthreads();

    int old_r = 0;  // point 0 has rank 0
    if (t == 0) out[0] = 0;

    for (int it = 1; it < m; ++it) {
        const float x1 = sx[old_r], y1 = sy[old_r], z1 = sz[old_r];
        float best = -1.0f;
        int bj = 0;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const int r = j * T + t;
            const float d = dist2_ref(sx[r] - x1, sy[r] - y1, sz[r] - z1);
            const float d2 = fminf(d, tmp[j]);
            tmp[j] = d2;
            if (d2 > best) { best = d2; bj = j; }
        }
        // (max distance, min slot): non-negative floats order like their int bit patterns,
        // and the -1.0f of an all-padding thread is negative as an int.
        const int db = __float_as_int(best);
        const unsigned rr = (unsigned)(bj * T + t);
        int wm = __reduce_max_sync(0xffffffffu, db);
        unsigned wr = __reduce_min_sync(0xffffffffu, db == wm ? rr : 0xffffffffu);
        if (T > 32) {
            const int buf = it & 1;
            if (lane == 0) { s_d[buf][warp] = wm; s_r[buf][warp] = wr; }
            __syncthreads();
            const int vd = (int)lane < T / 32 ? s_d[buf][lane] : (int)0x80000000;
            const unsigned vr = (int)lane < T / 32 ? s_r[buf][lane] : 0xffffffffu;
            wm = __reduce_max_sync(0xffffffffu, vd);
            wr = __reduce_min_sync(0xffffffffu, vd == wm ? vr : 0xffffffffu);
        }
        old_r = (int)wr;
        if (t == 0) out[it] = fps_slot_to_point(old_r, bs, log2bs, S, n);
    }

    if (temp_out != nullptr) {
        float *tp = temp_out + (size_t)blockIdx.x * n;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const int k = fps_slot_to_point(j * T + t, bs, log2bs, S, n);
            if (k < n) tp[k] = tmp[j];
        }
    }
}

// Fallback for clouds that do not fit the register/shared-memory kernel (bs*S > 16384
// slots): same rank layout, distances kept in the caller's temp buffer.
__global__ void __launch_bounds__(1024)
fps_kernel_large(int n, int m, int bs, int log2bs, int S, const float *__restrict__ dataset,
                 float *__restrict__ temp, int *__restrict__ idxs) {
    constexpr int T = 1024;
    __shared__ int s_d[2][32];
    __shared__ unsigned s_r[2][32];
    const int t = threadIdx.x;
    const unsigned lane = lane_id();
    const int warp = t >> 5;
    const float *pts = dataset + (size_t)blockIdx.x * n * 3;
    float *tp = temp + (size_t)blockIdx.x * n;
    int *out = idxs + (size_t)blockIdx.x * m;
    const int slots = bs * S;
    for (int k = t; k < n; k += T) tp[k] = 1e10f;
    __syncthreads();
    int old = 0;
    if (t == 0) out[0] = 0;
    for (int it = 1; it < m; ++it) {
        const float x1 = __ldg(pts + (size_t)old * 3), y1 = __ldg(pts + (size_t)old * 3 + 1),
                    z1 = __ldg(pts + (size_t)old * 3 + 2);
        float best = -1.0f;
        unsigned br = 0xffffffffu;
        for (int r = t; r < slots; r += T) {
            const int k = fps_slot_to_point(r, bs, log2bs, S, n);
            if (k >= n) continue;
            const float d = dist2_ref(__ldg(pts + (size_t)k * 3) - x1, __ldg(pts + (size_t)k * 3 + 1) - y1,
                                      __ldg(pts + (size_t)k * 3 + 2) - z1);
            const float d2 = fminf(d, tp[k]);
            tp[k] = d2;
            if (d2 > best) { best = d2; br = (unsigned)r; }
        }
        const int db = __float_as_int(best);
        int wm = __reduce_max_sync(0xffffffffu, db);
        unsigned wr = __reduce_min_sync(0xffffffffu, db == wm ? br : 0xffffffffu);
        const int buf = it & 1;
        if (lane == 0) { s_d[buf][warp] = wm; s_r[buf][warp] = wr; }
        __syncthreads();
        const int vd = s_d[buf][lane];
        const unsigned vr = s_r[buf][lane];
        wm = __reduce_max_sync(0xffffffffu, vd);
        wr = __reduce_min_sync(0xffffffffu, vd == wm ? vr : 0xffffffffu);
        old = fps_slot_to_point((int)wr, bs, log2bs, S, n);
        if (t == 0) out[it] = old;
    }
}

template <int T, int PPT>
static int launch_fps(int b, int n, int m, int bs, int log2bs, int S, const float *dataset,
                      float *temp, int *idxs, cudaStream_t st) {
    const size_t smem = (size_t)3 * T * PPT * sizeof(float);
    if (smem > 32 * 1024) {  // static smem (reduction scratch) counts against the 48 KB default too
        cudaError_t e = cudaFuncSetAttribute(fps_kernel<T, PPT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("fps: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return JMB_ERR_CUDA;
        }
    }
    fps_kernel<T, PPT><<<b, T, smem, st>>>(n, m, bs, log2bs, S, dataset, temp, idxs);
    return check_launch("furthest_point_sampling");
}

template <int T>
static int dispatch_fps_ppt(int ppt, int b, int n, int m, int bs, int log2bs, int S,
                            const float *dataset, float *temp, int *idxs, cudaStream_t st) {
    switch (ppt) {
        case 1: return launch_fps<T, 1>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
        case 2: return launch_fps<T, 2>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
        case 4: return launch_fps<T, 4>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
        case 8: return launch_fps<T, 8>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
        default: return launch_fps<T, 16>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
    }
}

// cuda_utils.h:10-14, evaluated with the same double arithmetic
static int ref_opt_n_threads(int work_size) {
    const int pow_2 = (int)(std::log(static_cast<double>(work_size)) / std::log(2.0));
    int v = 1 << pow_2;
    if (v > 1024) v = 1024;
    return v < 1 ? 1 : v;
}

static int pow2_ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ---- gather -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_points_kernel(int c, int n, int m, const float *__restrict__ points,
                     const int *__restrict__ idx, float *__restrict__ out) {
    const int b = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int src = __ldg(idx + (size_t)b * m + j);
    for (int ci = blockIdx.y; ci < c; ci += gridDim.y)
        out[((size_t)b * c + ci) * m + j] = __ldg(points + ((size_t)b * c + ci) * n + src);
}

__global__ void __launch_bounds__(256)
gather_points_grad_kernel(int c, int n, int m, const float *__restrict__ grad_out,
                          const int *__restrict__ idx, float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int dst = __ldg(idx + (size_t)b * m + j);
    for (int ci = blockIdx.y; ci < c; ci += gridDim.y)
        atomicAdd(grad_points + ((size_t)b * c + ci) * n + dst,
                  __ldg(grad_out + ((size_t)b * c + ci) * m + j));
}

}  // namespace jmb

extern "C" int jmb_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp,
                                           int *idxs, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && n >= 0 && m >= 0, "fps: negative size");
    if (b == 0 || m == 0) return JMB_OK;
    JMB_REQUIRE(n > 0, "fps: empty point cloud");
    JMB_REQUIRE(dataset && idxs, "fps: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int bs = ref_opt_n_threads(n);
    int log2bs = 0;
    while ((1 << log2bs) < bs) ++log2bs;
    const int S = (n + bs - 1) / bs;
    const int slots = bs * S;
    if (slots > 16384) {
        JMB_REQUIRE(temp != nullptr, "fps: n=%d needs the temp buffer", n);
        fps_kernel_large<<<b, 1024, 0, st>>>(n, m, bs, log2bs, S, dataset, temp, idxs);
        return check_launch("furthest_point_sampling(large)");
    }
    int T = pow2_ceil((slots + 3) / 4);
    if (T < 32) T = 32;
    if (T > 1024) T = 1024;
    const int ppt = pow2_ceil((slots + T - 1) / T);
    switch (T) {
        case 32: return dispatch_fps_ppt<32>(ppt, b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
        case 64: return dispatch_fps_ppt<64>(ppt, b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
        case 128: return dispatch_fps_ppt<128>(ppt, b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
        case 256: return dispatch_fps_ppt<256>(ppt, b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
        case 512: return dispatch_fps_ppt<512>(ppt, b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
        default: return dispatch_fps_ppt<1024>(ppt, b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
    }
}

extern "C" int jmb_gather_points(int b, int c, int n, int npoints, const float *points,
                                 const int *idx, float *out, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0, "gather_points: negative size");
    if (b == 0 || c == 0 || npoints == 0) return JMB_OK;
    JMB_REQUIRE(points && idx && out, "gather_points: null pointer");
    JMB_REQUIRE(b <= 65535, "gather_points: batch %d exceeds grid.z limit", b);
    dim3 grid(div_up(npoints, 256), c < 64 ? c : 64, b);
    gather_points_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, npoints, points, idx, out);
    return check_launch("gather_points");
}

extern "C" int jmb_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out,
                                      const int *idx, float *grad_points, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0, "gather_points_grad: negative size");
    if (b == 0 || c == 0 || npoints == 0) return JMB_OK;
    JMB_REQUIRE(grad_out && idx && grad_points, "gather_points_grad: null pointer");
    JMB_REQUIRE(b <= 65535, "gather_points_grad: batch %d exceeds grid.z limit", b);
    dim3 grid(div_up(npoints, 256), c < 64 ? c : 64, b);
    gather_points_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, npoints, grad_out, idx,
                                                                      grad_points);
    return check_launch("gather_points_grad");
}
