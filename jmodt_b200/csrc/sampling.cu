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
#include <stdlib.h>

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

// ---- cluster FPS: one frame spread over CL thread blocks (distributed shared memory) --------------------
// For small batches one SM per frame leaves the chip idle while a 4095-step dependency chain runs.  Here the
// points of a frame live in the REGISTERS of CL*T threads (PPT each); every iteration each CTA reduces to its own
// best (distance, rank) and the owning thread pushes a 5-word record (distance, rank, x, y, z) into the shared
// memory of ALL CTAs of the cluster with st.shared::cluster, then arrives on their mbarriers; each CTA waits for CL
// arrivals and picks the global winner locally — one __syncthreads and one DSMEM exchange per iteration, no global
// memory traffic.  Same rank layout as fps_kernel, so the choice among ties is the reference's.
__device__ __forceinline__ uint32_t fps_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int CL, int T, int PPT>
__global__ void __launch_bounds__(T)
fps_cluster_kernel(int n, int m, int bs, int log2bs, int S, const float *__restrict__ dataset,
                   float *__restrict__ temp_out, int *__restrict__ idxs) {
    constexpr int STRIDE = CL * T;  // slots per j
    __shared__ int s_d[2][32];
    __shared__ unsigned s_r[2][32];
    __shared__ __align__(16) unsigned s_rec[2][CL][8];   // (dist bits, rank, x, y, z)
    __shared__ __align__(8) unsigned long long s_bar[2];

    unsigned cta;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta));
    const int frame = blockIdx.x / CL;
    const int t = threadIdx.x;
    const unsigned lane = lane_id();
    const int warp = t >> 5;
    const float *pts = dataset + (size_t)frame * n * 3;
    int *out = idxs + (size_t)frame * m;

    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fps_smem_u32(&s_bar[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fps_smem_u32(&s_bar[1])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float px[PPT], py[PPT], pz[PPT], tmp[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int r = j * STRIDE + (int)cta * T + t;
        const int k = fps_slot_to_point(r, bs, log2bs, S, n);
        const bool valid = k < n;
        px[j] = valid ? __ldg(pts + (size_t)k * 3) : 0.f;
        py[j] = valid ? __ldg(pts + (size_t)k * 3 + 1) : 0.f;
        pz[j] = valid ? __ldg(pts + (size_t)k * 3 + 2) : 0.f;
        tmp[j] = valid ? 1e10f : -1.0f;
    }
    // every CTA's barriers must be initialised before anyone arrives on them remotely
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");

    float x1 = __ldg(pts), y1 = __ldg(pts + 1), z1 = __ldg(pts + 2);  // point 0 (rank 0)
    if (cta == 0 && t == 0) out[0] = 0;

    for (int it = 1; it < m; ++it) {
        if (t == 0)  // arm this iteration's barrier: CL records of 20 bytes will land in s_rec[it & 1]
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fps_smem_u32(&s_bar[it & 1])), "r"(CL * 20) : "memory");
        float best = -1.0f;
        int bj = 0;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const float d = dist2_ref(px[j] - x1, py[j] - y1, pz[j] - z1);
            const float d2 = fminf(d, tmp[j]);
            tmp[j] = d2;
            if (d2 > best) { best = d2; bj = j; }
        }
        const int db = __float_as_int(best);
        const unsigned rr = (unsigned)(bj * STRIDE + (int)cta * T + t);
        int wm = __reduce_max_sync(0xffffffffu, db);
        unsigned wr = __reduce_min_sync(0xffffffffu, db == wm ? rr : 0xffffffffu);
        const int buf = it & 1;
        if (T > 32) {
            if (lane == 0) { s_d[buf][warp] = wm; s_r[buf][warp] = wr; }
            __syncthreads();
            const int vd = (int)lane < T / 32 ? s_d[buf][lane] : (int)0x80000000;
            const unsigned vr = (int)lane < T / 32 ? s_r[buf][lane] : 0xffffffffu;
            wm = __reduce_max_sync(0xffffffffu, vd);
            wr = __reduce_min_sync(0xffffffffu, vd == wm ? vr : 0xffffffffu);
        }
        // The warp that owns this CTA's best point publishes it: lanes 0..CL-1 each push the 20-byte record into one
        // CTA of the cluster with st.async, which completes transaction bytes on that CTA's mbarrier — no release
        // fence, no serialised remote arrives.
        {
            const bool iswin = (rr == wr) && (db == wm);
            const unsigned wmask = __ballot_sync(0xffffffffu, iswin);
            if (wmask) {
                const int src = __ffs(wmask) - 1;
                float wx = px[0], wy = py[0], wz = pz[0];
#pragma unroll
                for (int j = 1; j < PPT; ++j)
                    if (j == bj) { wx = px[j]; wy = py[j]; wz = pz[j]; }
                wx = __shfl_sync(0xffffffffu, wx, src);
                wy = __shfl_sync(0xffffffffu, wy, src);
                wz = __shfl_sync(0xffffffffu, wz, src);
                if (lane < (unsigned)CL) {
                    uint32_t rec_remote, bar_remote;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rec_remote) : "r"(fps_smem_u32(&s_rec[buf][cta][0])), "r"(lane));
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bar_remote) : "r"(fps_smem_u32(&s_bar[buf])), "r"(lane));
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(rec_remote),
                                 "r"((unsigned)wm), "r"(wr), "r"(__float_as_uint(wx)), "r"(__float_as_uint(wy)), "r"(bar_remote)
                                 : "memory");
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(rec_remote + 16),
                                 "r"(__float_as_uint(wz)), "r"(bar_remote)
                                 : "memory");
                }
            }
        }
        // wait for the CL records of this iteration (the k-th use of a buffer completes its phase k)
        {
            const uint32_t bar = fps_smem_u32(&s_bar[buf]);
            const uint32_t parity = (uint32_t)(((it - 1) >> 1) & 1);
            uint32_t done;
            do {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}\n"
                    : "=r"(done)
                    : "r"(bar), "r"(parity)
                    : "memory");
            } while (!done);
        }
        int gd = (int)0x80000000;
        unsigned gr = 0xffffffffu;
#pragma unroll
        for (int c = 0; c < CL; ++c) {
            const uint4 rec = *reinterpret_cast<const uint4 *>(&s_rec[buf][c][0]);
            const int d = (int)rec.x;
            if (d > gd || (d == gd && rec.y < gr)) {
                gd = d; gr = rec.y;
                x1 = __uint_as_float(rec.z); y1 = __uint_as_float(rec.w); z1 = __uint_as_float(s_rec[buf][c][4]);
            }
        }
        if (cta == 0 && t == 0) out[it] = fps_slot_to_point((int)gr, bs, log2bs, S, n);
    }

    if (temp_out != nullptr) {
        float *tp = temp_out + (size_t)frame * n;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const int k = fps_slot_to_point(j * STRIDE + (int)cta * T + t, bs, log2bs, S, n);
            if (k < n) tp[k] = tmp[j];
        }
    }
    // no CTA may exit while a peer can still write into its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CL, int T, int PPT>
static int launch_fps_cluster(int b, int n, int m, int bs, int log2bs, int S, const float *dataset, float *temp,
                              int *idxs, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)b * CL);
    cfg.blockDim = dim3(T);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (CL > 8) {
        int dev = 0, sms = 0;
        const int rc = device_info(&dev, &sms);
        if (rc != JMB_OK) return rc;
        JMB_FUNC_ATTR_ONCE((fps_cluster_kernel<CL, T, PPT>), cudaFuncAttributeNonPortableClusterSizeAllowed, 1, dev);
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, fps_cluster_kernel<CL, T, PPT>, n, m, bs, log2bs, S, dataset, temp, idxs);
    if (e != cudaSuccess) {
        set_error("fps(cluster): %s", cudaGetErrorString(e));
        return JMB_ERR_CUDA;
    }
    return check_launch("furthest_point_sampling(cluster)");
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

// ---- stream gate on FPS progress ---------------------------------------------------------------
// Furthest point sampling emits its indices one per iteration and is latency-bound on a fraction of the SMs; the
// consumers of sample k (ball query, grouping, the MLP stack) only need indices 0..k.  This kernel lets a second
// stream start on a prefix while the sampler is still running: it returns once every idx[f][k0..k1) is >= 0 (the
// caller pre-fills idx with -1), so work queued behind it sees a complete prefix.  The wait is bounded: if the
// producer never shows up the kernel traps and the stream reports an error instead of hanging the device.
__global__ void __launch_bounds__(256)
wait_indices_kernel(const int *idx, int b, int row_stride, int k0, int k1, unsigned long long timeout_ns) {
    const int w = k1 - k0;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int e = w * b - 1 - (int)threadIdx.x; e >= 0; e -= (int)blockDim.x) {   // newest entries first
        const volatile int *p = idx + (size_t)(e / w) * row_stride + k0 + e % w;
        while (*p < 0) {
            __nanosleep(256);
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) __trap();
        }
    }
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

extern "C" int jmb_wait_indices(const int *idx, int b, int row_stride, int k0, int k1, int timeout_ms, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && k0 >= 0 && k1 >= k0 && row_stride >= k1, "wait_indices: bad range");
    if (b == 0 || k1 == k0) return JMB_OK;
    JMB_REQUIRE(idx != nullptr, "wait_indices: null pointer");
    JMB_REQUIRE(timeout_ms > 0, "wait_indices: timeout must be positive");
    wait_indices_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(idx, b, row_stride, k0, k1,
                                                            (unsigned long long)timeout_ms * 1000000ULL);
    return check_launch("wait_indices");
}

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
    // small batches: spread each frame over a thread-block cluster (latency); large batches: one CTA per frame
    // (throughput).  sms*2 is the point where single-CTA kernels already fill the machine.
    int dev = 0, sms = 0;
    {
        const int rc = device_info(&dev, &sms);
        if (rc != JMB_OK) return rc;
    }
    if (m > 1) {
        static int cfg_override = -1;   // JMB_FPS_CFG: tuning aid for profiles/fps_tune.py
        if (cfg_override < 0) { const char *e = getenv("JMB_FPS_CFG"); cfg_override = e ? atoi(e) : 0; }
        if (cfg_override > 0 && slots == 16384) {
            switch (cfg_override) {
                case 1: return launch_fps_cluster<8, 512, 4>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 2: return launch_fps_cluster<8, 256, 8>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 3: return launch_fps_cluster<8, 128, 16>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 4: return launch_fps_cluster<16, 256, 4>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 5: return launch_fps_cluster<16, 128, 8>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 6: return launch_fps_cluster<4, 512, 8>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 7: return launch_fps_cluster<16, 64, 16>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 8: return launch_fps<1024, 16>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 9: return launch_fps_cluster<2, 1024, 8>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 10: return launch_fps_cluster<4, 1024, 4>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 11: return launch_fps_cluster<2, 512, 16>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 12: return launch_fps_cluster<2, 256, 32>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 13: return launch_fps_cluster<4, 256, 16>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 14: return launch_fps<512, 32>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 15: return launch_fps_cluster<4, 128, 32>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 16: return launch_fps_cluster<8, 128, 16>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 17: return launch_fps_cluster<8, 64, 32>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                default: break;
            }
        }
        static int cfg4k = -1;          // JMB_FPS_CFG4K: the same for 4 096-point clouds
        if (cfg4k < 0) { const char *e = getenv("JMB_FPS_CFG4K"); cfg4k = e ? atoi(e) : 0; }
        if (cfg4k > 0 && slots == 4096) {
            switch (cfg4k) {
                case 1: return launch_fps_cluster<4, 256, 4>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 2: return launch_fps_cluster<8, 128, 4>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 3: return launch_fps_cluster<4, 128, 8>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 4: return launch_fps_cluster<8, 64, 8>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 5: return launch_fps_cluster<2, 256, 8>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 6: return launch_fps_cluster<16, 64, 4>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 7: return launch_fps_cluster<2, 128, 16>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 8: return launch_fps<512, 8>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                case 9: return launch_fps<256, 16>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
                default: break;
            }
        }
        // measured on B200: alone, 8 CTAs x 256 threads x 8 points is the fastest shape for 16384 points (2.93 ms vs 6.0 ms
        // for one CTA, profiles/fps_tune.py) — but the sampler is pure dependency latency, and inside the pipelined step every
        // SM it holds is an SM the tensor-core kernels of the other stream cannot use: 4 CTAs x 256 threads x 16 points
        // (half the SMs, 3.15 ms alone) gives 7.62 ms per step against 8.25 (profiles/r02/fps_shapes_in_step.txt).
        // JMB_FPS_CFG=2 selects the latency shape.
        if (slots > 8192 && slots <= 16384 && (long long)b * 8 <= 2LL * sms)
            return launch_fps_cluster<4, 256, 16>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
        if (slots > 4096 && slots <= 8192 && (long long)b * 8 <= 2LL * sms)
            return launch_fps_cluster<8, 256, 4>(b, n, m, bs, log2bs, S, dataset, temp, idxs, st);
    }
    // one CTA per cloud.  Few clouds (a batch of frames): 4 points per thread, the lowest latency per iteration.  Many
    // clouds (1 024 proposals x 512 points): 16 points per thread — a 512-point cloud is then ONE warp, the argmax is two
    // shuffles reductions with no shared-memory round trip or barrier (72 vs 106 us).  JMB_FPS_PPT overrides.
    static int ppt_override = -1;
    if (ppt_override < 0) { const char *e = getenv("JMB_FPS_PPT"); ppt_override = e ? atoi(e) : 0; if (ppt_override < 0) ppt_override = 0; }
    const int small_ppt = ppt_override > 0 ? ppt_override : (b >= sms ? 16 : 4);
    int T = pow2_ceil((slots + small_ppt - 1) / small_ppt);
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
