// Rotated BEV overlap / IoU, 3-D IoU and NMS for sm_100a.
//
// Replaces boxes_overlap_kernel, boxes_iou_bev_kernel, nms_kernel, nms_normal_kernel
// (reference jmodt/ops/iou3d/src/iou3d_kernel.cu:223-348) and the host side of
// nms_gpu / nms_normal_gpu (iou3d.cpp:73-166: cudaMalloc, blocking D2H copy of the
// n x ceil(n/64) bitmask — 5 MB at n=6300 — serial host sweep, cudaFree).
// Differences in structure:
//   * the suppression bitmask is only computed for column blocks >= row block (the sweep
//     never reads the rest, iou3d.cpp:104-110);
//   * the greedy sweep runs on the device in one CTA (serial only over the 64 boxes of a
//     block, parallel over the remaining column words) and can stop after max_keep boxes;
//   * polygon vertices are sorted by an angle computed once per vertex instead of two
//     atan2f per comparison (same values, same comparison sequence, same order).
// Geometry arithmetic follows the reference binaries operation for operation (fused
// multiply-add placement read from their SASS; see oracle/jmodt_oracle.c).
#include "iou3d_device.cuh"

namespace jmb {

// MODE 0: BEV overlap, 1: BEV IoU, 2: 3-D IoU from (n,7) boxes (iou3d_utils.py:22-54)
template <int MODE>
__global__ void __launch_bounds__(128)
pairwise_kernel(int na, const float *__restrict__ boxes_a, int nb, const float *__restrict__ boxes_b,
                float *__restrict__ ans) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)na * nb) return;
    const int ia = (int)(e / nb), ib = (int)(e % nb);
    if (MODE == 2) {
        const float *A7 = boxes_a + (size_t)ia * 7, *B7 = boxes_b + (size_t)ib * 7;
        float a[7], b[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) { a[k] = __ldg(A7 + k); b[k] = __ldg(B7 + k); }
        // boxes3d_to_bev_torch (kitti_utils.py:136-149)
        const float ahl = __fmul_rn(a[5], 0.5f), ahw = __fmul_rn(a[4], 0.5f);
        const float bhl = __fmul_rn(b[5], 0.5f), bhw = __fmul_rn(b[4], 0.5f);
        const float ba[5] = {__fsub_rn(a[0], ahl), __fsub_rn(a[2], ahw), __fadd_rn(a[0], ahl),
                             __fadd_rn(a[2], ahw), a[6]};
        const float bb[5] = {__fsub_rn(b[0], bhl), __fsub_rn(b[2], bhw), __fadd_rn(b[0], bhl),
                             __fadd_rn(b[2], bhw), b[6]};
        const float ov = box_overlap(ba, bb);
        const float max_of_min = fmaxf(__fsub_rn(a[1], a[3]), __fsub_rn(b[1], b[3]));
        const float min_of_max = fminf(a[1], b[1]);
        const float oh = fmaxf(__fsub_rn(min_of_max, max_of_min), 0.f);
        const float o3 = __fmul_rn(ov, oh);
        const float va = __fmul_rn(__fmul_rn(a[3], a[4]), a[5]);
        const float vb = __fmul_rn(__fmul_rn(b[3], b[4]), b[5]);
        ans[e] = __fdiv_rn(o3, fmaxf(__fsub_rn(__fadd_rn(va, vb), o3), 1e-7f));
    } else {
        float a[5], b[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            a[k] = __ldg(boxes_a + (size_t)ia * 5 + k);
            b[k] = __ldg(boxes_b + (size_t)ib * 5 + k);
        }
        ans[e] = MODE == 0 ? box_overlap(a, b) : iou_bev(a, b);
    }
}

// Upper-triangular 64x64-tile suppression bitmask (iou3d_kernel.cu:250-292 / 306-348).
template <bool ROTATED>
__global__ void __launch_bounds__(64)
nms_mask_kernel(int n, float thresh, const float *__restrict__ boxes,
                unsigned long long *__restrict__ mask) {
    const int row_start = blockIdx.y, col_start = blockIdx.x;
    if (col_start < row_start) return;  // never read by the sweep
    const int row_size = min(n - row_start * 64, 64);
    const int col_size = min(n - col_start * 64, 64);
    __shared__ float block_boxes[64 * 5];
    if ((int)threadIdx.x < col_size) {
#pragma unroll
        for (int k = 0; k < 5; ++k)
            block_boxes[threadIdx.x * 5 + k] = __ldg(boxes + (size_t)(64 * col_start + threadIdx.x) * 5 + k);
    }
    __syncthreads();
    if ((int)threadIdx.x < row_size) {
        const int cur = 64 * row_start + threadIdx.x;
        float a[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) a[k] = __ldg(boxes + (size_t)cur * 5 + k);
        unsigned long long t = 0;
        const int start = (row_start == col_start) ? threadIdx.x + 1 : 0;
        for (int i = start; i < col_size; ++i) {
            const float v = ROTATED ? iou_bev(a, block_boxes + i * 5) : iou_normal(a, block_boxes + i * 5);
            if (v > thresh) t |= 1ULL << i;
        }
        const int col_blocks = (n + 63) / 64;
        mask[(size_t)cur * col_blocks + col_start] = t;
    }
}

// Greedy sweep of iou3d.cpp:98-113 on the device.
__global__ void __launch_bounds__(256)
nms_sweep_kernel(int n, int col_blocks, const unsigned long long *__restrict__ mask,
                 long long *__restrict__ keep, int *__restrict__ num_keep, int max_keep) {
    extern __shared__ unsigned long long s_remv[];  // [col_blocks]
    __shared__ unsigned long long s_diag[64];
    __shared__ unsigned long long s_kept;
    __shared__ int s_nkeep, s_stop;
    for (int j = threadIdx.x; j < col_blocks; j += blockDim.x) s_remv[j] = 0ULL;
    if (threadIdx.x == 0) { s_nkeep = 0; s_stop = 0; }
    __syncthreads();
    for (int blk = 0; blk < col_blocks; ++blk) {
        if (threadIdx.x < 64) {
            const int row = blk * 64 + threadIdx.x;
            s_diag[threadIdx.x] = row < n ? mask[(size_t)row * col_blocks + blk] : 0ULL;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long cur = s_remv[blk], kept = 0ULL;
            int nk = s_nkeep;
            const int rows = min(64, n - blk * 64);
            for (int i = 0; i < rows; ++i) {
                if (!((cur >> i) & 1ULL)) {
                    if (max_keep > 0 && nk >= max_keep) { s_stop = 1; break; }
                    keep[nk++] = (long long)blk * 64 + i;
                    kept |= 1ULL << i;
                    cur |= s_diag[i];
                }
            }
            s_kept = kept;
            s_nkeep = nk;
        }
        __syncthreads();
        if (s_stop) break;
        const unsigned long long kept = s_kept;
        for (int j = blk + 1 + threadIdx.x; j < col_blocks; j += blockDim.x) {
            unsigned long long acc = s_remv[j], kk = kept;
            while (kk) {
                const int i = __ffsll((long long)kk) - 1;
                kk &= kk - 1;
                acc |= mask[(size_t)(blk * 64 + i) * col_blocks + j];
            }
            s_remv[j] = acc;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *num_keep = s_nkeep;
}

template <int MODE>
static int launch_pairwise(int na, const float *a, int nb, const float *b, float *ans, void *stream,
                           const char *what) {
    JMB_REQUIRE(na >= 0 && nb >= 0, "%s: negative size", what);
    if (na == 0 || nb == 0) return JMB_OK;
    JMB_REQUIRE(a && b && ans, "%s: null pointer", what);
    const long long total = (long long)na * nb;
    pairwise_kernel<MODE><<<(unsigned)div_up_ll(total, 128), 128, 0, (cudaStream_t)stream>>>(na, a, nb, b, ans);
    return check_launch(what);
}

static int launch_nms(bool rotated, int n, const float *boxes, float thresh, int64_t *keep,
                      int *num_keep, int max_keep, void *workspace, size_t workspace_bytes,
                      void *stream) {
    JMB_REQUIRE(n >= 0, "nms: negative size");
    JMB_REQUIRE(num_keep, "nms: null num_keep");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        JMB_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int), st));
        return JMB_OK;
    }
    JMB_REQUIRE(boxes && keep, "nms: null pointer");
    if (!workspace || workspace_bytes < jmb_nms_workspace_bytes(n)) {
        set_error("nms: workspace of %zu bytes required", jmb_nms_workspace_bytes(n));
        return JMB_ERR_WORKSPACE;
    }
    const int col_blocks = div_up(n, 64);
    JMB_REQUIRE(col_blocks <= 65535, "nms: n=%d too large", n);
    unsigned long long *mask = (unsigned long long *)workspace;
    dim3 grid(col_blocks, col_blocks);
    if (rotated) nms_mask_kernel<true><<<grid, 64, 0, st>>>(n, thresh, boxes, mask);
    else nms_mask_kernel<false><<<grid, 64, 0, st>>>(n, thresh, boxes, mask);
    int rc = check_launch("nms(mask)");
    if (rc) return rc;
    const size_t smem = (size_t)col_blocks * sizeof(unsigned long long);
    JMB_REQUIRE(smem <= 48 * 1024, "nms: n=%d too large for the sweep", n);
    nms_sweep_kernel<<<1, 256, smem, st>>>(n, col_blocks, mask, (long long *)keep, num_keep, max_keep);
    return check_launch("nms(sweep)");
}

}  // namespace jmb

extern "C" int jmb_boxes_overlap_bev(int na, const float *boxes_a, int nb, const float *boxes_b,
                                     float *ans_overlap, void *stream) {
    return jmb::launch_pairwise<0>(na, boxes_a, nb, boxes_b, ans_overlap, stream, "boxes_overlap_bev");
}
extern "C" int jmb_boxes_iou_bev(int na, const float *boxes_a, int nb, const float *boxes_b,
                                 float *ans_iou, void *stream) {
    return jmb::launch_pairwise<1>(na, boxes_a, nb, boxes_b, ans_iou, stream, "boxes_iou_bev");
}
extern "C" int jmb_boxes_iou3d(int na, const float *boxes_a, int nb, const float *boxes_b,
                               float *ans_iou, void *stream) {
    return jmb::launch_pairwise<2>(na, boxes_a, nb, boxes_b, ans_iou, stream, "boxes_iou3d");
}
extern "C" size_t jmb_nms_workspace_bytes(int n) {
    if (n <= 0) return 0;
    const size_t col_blocks = (size_t)(n + 63) / 64;
    return (size_t)n * col_blocks * sizeof(unsigned long long);
}
extern "C" int jmb_nms(int n, const float *boxes, float thresh, int64_t *keep, int *num_keep,
                       int max_keep, void *workspace, size_t workspace_bytes, void *stream) {
    return jmb::launch_nms(true, n, boxes, thresh, keep, num_keep, max_keep, workspace, workspace_bytes, stream);
}
extern "C" int jmb_nms_normal(int n, const float *boxes, float thresh, int64_t *keep, int *num_keep,
                              int max_keep, void *workspace, size_t workspace_bytes, void *stream) {
    return jmb::launch_nms(false, n, boxes, thresh, keep, num_keep, max_keep, workspace, workspace_bytes, stream);
}
