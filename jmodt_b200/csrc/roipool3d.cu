// 3-D RoI point pooling for sm_100a — one fused kernel.
//
// Replaces roipool3dLauncher (reference jmodt/ops/roipool3d/src/roipool3d_kernel.cu:209-237):
//   K10 assign_pts_to_box3d (:97-120)  writes a (B,N,M) int flag tensor (8.4 MB / frame),
//   K11 get_pooled_idx      (:123-160) ONE THREAD per box scans 16 384 flags with stride-M reads,
//   K12 roipool3d_forward   (:163-194) one thread copies a whole 133-float row,
//   plus cudaMalloc x2 / cudaFree x2 (implicit device syncs) per call.
// Here one CTA owns one (frame, box): its 8 warps test disjoint, contiguous slices of the
// point cloud and compact the hits in index order with ballot/popcount (no flag tensor, no
// scratch allocation); the per-warp lists are concatenated, wrapped around exactly like
// :152-158, and the rows are then copied by whole warps with coalesced reads and writes.
// Outputs are fully written, so the caller's .zero_() pass (roipool3d_utils.py:22-24,
// 35 MB / frame) is not needed.
//
// The in-box predicate follows the SASS of the reference kernel (see oracle/jmodt_oracle.c):
// float libdevice cosf/sinf, x_rot = fma(dz,-sina, fl(dx*cosa)), z_rot = fma(dx,sina, fl(dz*cosa)),
// cy = float(double(bottom_y) - double(h)/2).
#include "common.cuh"

namespace jmb {

constexpr int RP_WARPS = 8;
constexpr int RP_THREADS = RP_WARPS * 32;

struct BoxTest {
    float cx, cy, cz, hh, hl, hw, cosa, sina;
    __device__ __forceinline__ bool inside(float x, float y, float z) const {
        const float dx = x - cx;
        if (fabsf(dx) > 10.0f) return false;
        if (fabsf(y - cy) > hh) return false;
        const float dz = z - cz;
        if (fabsf(dz) > 10.0f) return false;
        const float x_rot = __fmaf_rn(dz, -sina, __fmul_rn(dx, cosa));
        const float z_rot = __fmaf_rn(dx, sina, __fmul_rn(dz, cosa));
        return (x_rot >= -hl) & (x_rot <= hl) & (z_rot >= -hw) & (z_rot <= hw);
    }
};

// CANON: also apply the eval-branch canonical transform of
// proposal_target_layer.py:107-112 (centre on the roi, rotate by ry about y), with the box
// enlarged in-kernel as kitti_utils.enlarge_box3d (kitti_utils.py:152-162) does.
// lead > 0 selects the "head layout" consumed by jmb_rcnn_input_fused: a row is
//   [feature lead .. feat_len-1 | x, y, z | feature 0 .. lead-1 | zero padding]  with pitch round_up(3 + feat_len, 8)
// i.e. the 128 RPN channels first (one 16-byte aligned block) and the 5 inputs of xyz_up_layer (rcnn.py:172-180:
// xyz, seg mask, depth) behind them.  lead == 0 is the reference layout [x, y, z | features], pitch 3 + feat_len.
// RP_UNROLL = 4 (head layout): four rows / four scan steps in flight per warp, 4 CTAs per SM; RP_UNROLL = 1 (reference
// layout, rows only 4-byte aligned): one row per warp iteration at full occupancy — measured faster there.
template <bool CANON, int RP_UNROLL>
__global__ void __launch_bounds__(RP_THREADS, RP_UNROLL == 1 ? 6 : 4)
roipool3d_kernel(int pts_num, int boxes_num, int feat_len, int sampled, float extra, int lead,
                 const float *__restrict__ xyz, const float *__restrict__ boxes3d,
                 const float *__restrict__ pts_feature, float *__restrict__ pooled,
                 int *__restrict__ empty_flag) {
    extern __shared__ int rp_smem[];
    int *s_final = rp_smem;                 // [sampled]
    int *s_warp = rp_smem + sampled;        // [RP_WARPS][sampled]
    __shared__ int s_cnt[RP_WARPS];

    const int box = blockIdx.x, b = blockIdx.y;
    const int warp = threadIdx.x >> 5;
    const unsigned lane = lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;

    const float *bx = boxes3d + ((size_t)b * boxes_num + box) * 7;
    float bx0 = __ldg(bx), by = __ldg(bx + 1), bz0 = __ldg(bx + 2), bh = __ldg(bx + 3),
          bw = __ldg(bx + 4), bl = __ldg(bx + 5), ry = __ldg(bx + 6);
    const float roi_y = by;
    if (CANON) {
        const float e2 = extra * 2;  // python: extra_width * 2, then cast to the tensor dtype
        bh = __fadd_rn(bh, e2); bw = __fadd_rn(bw, e2); bl = __fadd_rn(bl, e2);
        by = __fadd_rn(by, extra);
    }
    BoxTest T;
    T.cx = bx0; T.cz = bz0;
    T.cy = (float)((double)by - (double)bh / 2.0);
    T.hh = bh * 0.5f; T.hl = bl * 0.5f; T.hw = bw * 0.5f;
    T.cosa = cosf(ry); T.sina = sinf(ry);

    // ---- phase 1: ordered compaction, one contiguous slice of the cloud per warp ----------
    const float *pts = xyz + (size_t)b * pts_num * 3;
    const int seg = ((pts_num + RP_WARPS * 32 - 1) / (RP_WARPS * 32)) * 32;
    const int beg = warp * seg, end = min(pts_num, beg + seg);
    int *mine = s_warp + (size_t)warp * sampled;
    int cnt = 0;
    for (int p0 = beg; p0 < end && cnt < sampled; p0 += 32 * RP_UNROLL) {   // RP_UNROLL 32-point steps per trip: their loads overlap
        float px[RP_UNROLL], py[RP_UNROLL], pz[RP_UNROLL];
#pragma unroll
        for (int u = 0; u < RP_UNROLL; ++u) {
            const int p = p0 + u * 32 + (int)lane;
            px[u] = py[u] = pz[u] = 0.f;
            if (p < end) {
                const float *q = pts + (size_t)p * 3;
                px[u] = __ldg(q); py[u] = __ldg(q + 1); pz[u] = __ldg(q + 2);
            }
        }
#pragma unroll
        for (int u = 0; u < RP_UNROLL; ++u) {
            const int p = p0 + u * 32 + (int)lane;
            const bool hit = p < end && T.inside(px[u], py[u], pz[u]);
            const unsigned mk = __ballot_sync(0xffffffffu, hit);
            const int pos = cnt + __popc(mk & lt_mask);
            if (hit && pos < sampled) mine[pos] = p;
            cnt += __popc(mk);
        }
    }
    cnt = min(cnt, sampled);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();

    int total = 0, my_off = 0;
#pragma unroll
    for (int w = 0; w < RP_WARPS; ++w) {
        if (w == warp) my_off = total;
        total += s_cnt[w];
    }
    const int have = min(total, sampled);
    for (int i = lane; i < cnt; i += 32)
        if (my_off + i < sampled) s_final[my_off + i] = mine[i];
    __syncthreads();

    const int row = lead > 0 ? ((3 + feat_len + 7) / 8) * 8 : 3 + feat_len;
    const int xoff = lead > 0 ? feat_len - lead : 0;      // column of x inside a row
    float *dst_box = pooled + ((size_t)b * boxes_num + box) * (size_t)sampled * row;
    if (threadIdx.x == 0) empty_flag[(size_t)b * boxes_num + box] = (have == 0) ? 1 : 0;

    float ox = 0.f, oy = 0.f, oz = 0.f, rc = 1.f, rs = 0.f;
    if (CANON) { ox = bx0; oy = roi_y; oz = bz0; rc = T.cosa; rs = T.sina; }

    if (have == 0) {
        // reference: the caller's zero-init survives (roipool3d_kernel.cu:177-179); with the
        // canonical transform the (0,0,0) xyz of every row is still shifted and rotated
        // (proposal_target_layer.py:109-112 applies to all rows).
        float zx = 0.f, zy = 0.f, zz = 0.f;
        if (CANON) {
            const float tx = 0.f - ox, tz = 0.f - oz;
            zy = 0.f - oy;
            zx = __fmaf_rn(tz, -rs, __fmul_rn(tx, rc));
            zz = __fmaf_rn(tz, rc, __fmul_rn(tx, rs));
        }
        const size_t nflt = (size_t)sampled * row;
        for (size_t e = threadIdx.x; e < nflt; e += RP_THREADS) {
            const int j = (int)(e % row);
            dst_box[e] = j == xoff ? zx : (j == xoff + 1 ? zy : (j == xoff + 2 ? zz : 0.f));
        }
        return;
    }

    // ---- phase 2: row copies, wrap-around duplicates (:152-158) --------------------------------
    // One warp per row, RP_UNROLL rows per iteration: the loads of all rows are issued before the first store, so a
    // warp keeps ~2 KB in flight instead of one 0.5 KB row (the copy is latency-bound: 64 warps per SM, ~1.5 us per
    // L2-read / HBM-write round trip).  The head layout with 128 channels moves them as 8-byte pairs.
    const float *feat = pts_feature + (size_t)b * pts_num * feat_len;
    const bool vec = lead > 0 && xoff == 128 && row == 136 && (feat_len & 1) == 0 && (lead & 1) == 0 &&
                     (reinterpret_cast<uintptr_t>(feat) & 7u) == 0 && (reinterpret_cast<uintptr_t>(dst_box) & 7u) == 0;
    for (int s0 = warp; s0 < sampled; s0 += RP_WARPS * RP_UNROLL) {
        int src[RP_UNROLL];
        float xv[RP_UNROLL];                 // lanes 0..2: transformed coordinate `lane` of the row
#pragma unroll
        for (int u = 0; u < RP_UNROLL; ++u) {
            const int sidx = s0 + u * RP_WARPS;
            src[u] = sidx < sampled ? s_final[sidx < have ? sidx : sidx % have] : -1;
            xv[u] = 0.f;
            if (src[u] >= 0 && lane < 3) {
                const float *q = pts + (size_t)src[u] * 3;
                float v = __ldg(q + lane);
                if (CANON) {
                    const float tx = __ldg(q) - ox, tz = __ldg(q + 2) - oz;
                    // rotate_pc_along_y_torch: x' = x*cos - z*sin ; z' = x*sin + z*cos
                    if (lane == 0) v = __fmaf_rn(tz, -rs, __fmul_rn(tx, rc));
                    else if (lane == 1) v = v - oy;
                    else v = __fmaf_rn(tz, rc, __fmul_rn(tx, rs));
                }
                xv[u] = v;
            }
        }
        if (vec) {
            float2 a[RP_UNROLL], c[RP_UNROLL];
            float tl[RP_UNROLL];
#pragma unroll
            for (int u = 0; u < RP_UNROLL; ++u) {
                if (src[u] < 0) continue;
                const float *f = feat + (size_t)src[u] * feat_len;
                const float2 *f2 = reinterpret_cast<const float2 *>(f + lead);
                a[u] = __ldg(f2 + lane);
                c[u] = __ldg(f2 + 32 + lane);
                tl[u] = (lane >= 3 && (int)lane < 3 + lead) ? __ldg(f + lane - 3) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < RP_UNROLL; ++u) {
                if (src[u] < 0) continue;
                float *dst = dst_box + (size_t)(s0 + u * RP_WARPS) * row;
                float2 *d2 = reinterpret_cast<float2 *>(dst);
                d2[lane] = a[u];
                d2[32 + lane] = c[u];
                if (lane < 8) dst[128 + lane] = lane < 3 ? xv[u] : tl[u];
            }
        } else if (lead == 0 && RP_UNROLL == 1) {
            if (src[0] >= 0) {
                float *dst = dst_box + (size_t)s0 * row;
                const float *f = feat + (size_t)src[0] * feat_len;
                if (lane < 3) dst[lane] = xv[0];
                for (int j = lane; j < feat_len; j += 32) dst[3 + j] = __ldg(f + j);
            }
        } else if (lead == 0) {
            // reference layout [x, y, z | features]: rows are only 4-byte aligned; the same 64 columns of all rows are
            // loaded before any is stored
            if (lane < 3) {
#pragma unroll
                for (int u = 0; u < RP_UNROLL; ++u)
                    if (src[u] >= 0) dst_box[(size_t)(s0 + u * RP_WARPS) * row + lane] = xv[u];
            }
            for (int j0 = 0; j0 < feat_len; j0 += 64) {
                const int ja = j0 + (int)lane, jb = ja + 32;
                float va[RP_UNROLL], vb[RP_UNROLL];
#pragma unroll
                for (int u = 0; u < RP_UNROLL; ++u) {
                    const float *f = feat + (size_t)(src[u] >= 0 ? src[u] : 0) * feat_len;
                    va[u] = (src[u] >= 0 && ja < feat_len) ? __ldg(f + ja) : 0.f;
                    vb[u] = (src[u] >= 0 && jb < feat_len) ? __ldg(f + jb) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < RP_UNROLL; ++u) {
                    float *dst = dst_box + (size_t)(s0 + u * RP_WARPS) * row + 3;
                    if (src[u] >= 0 && ja < feat_len) dst[ja] = va[u];
                    if (src[u] >= 0 && jb < feat_len) dst[jb] = vb[u];
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < RP_UNROLL; ++u) {
                if (src[u] < 0) continue;
                float *dst = dst_box + (size_t)(s0 + u * RP_WARPS) * row;
                const float *f = feat + (size_t)src[u] * feat_len;
                if (lane < 3) dst[xoff + lane] = xv[u];
                for (int j = lane; j < xoff; j += 32) dst[j] = __ldg(f + lead + j);
                for (int j = lane; j < row - xoff - 3; j += 32) dst[xoff + 3 + j] = j < lead ? __ldg(f + j) : 0.f;
            }
        }
    }
}

static int launch_roipool(bool canon, int lead, int batch, int pts_num, int boxes_num, int feat_len,
                          int sampled, float extra, const float *xyz, const float *boxes3d,
                          const float *pts_feature, float *pooled, int *empty_flag, void *stream) {
    JMB_REQUIRE(batch >= 0 && pts_num >= 0 && boxes_num >= 0 && feat_len >= 0 && sampled >= 0,
                "roipool3d: negative size");
    if (batch == 0 || boxes_num == 0) return JMB_OK;
    JMB_REQUIRE(xyz || pts_num == 0, "roipool3d: null xyz");
    JMB_REQUIRE(boxes3d && empty_flag, "roipool3d: null pointer");
    JMB_REQUIRE(sampled == 0 || pooled, "roipool3d: null output");
    JMB_REQUIRE(feat_len == 0 || pts_feature || pts_num == 0, "roipool3d: null features");
    JMB_REQUIRE(batch <= 65535, "roipool3d: batch %d exceeds grid.y limit", batch);
    const size_t smem = (size_t)(RP_WARPS + 1) * sampled * sizeof(int);
    JMB_REQUIRE(smem <= 200 * 1024, "roipool3d: sampled_pt_num %d too large", sampled);
    auto kern = lead > 0 ? (canon ? roipool3d_kernel<true, 4> : roipool3d_kernel<false, 4>)
                         : (canon ? roipool3d_kernel<true, 1> : roipool3d_kernel<false, 1>);
    if (smem > 48 * 1024)
        JMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(boxes_num, batch);
    JMB_REQUIRE(lead >= 0 && lead <= feat_len, "roipool3d: bad head-layout split %d", lead);
    kern<<<grid, RP_THREADS, smem, (cudaStream_t)stream>>>(pts_num, boxes_num, feat_len, sampled, extra, lead,
                                                          xyz, boxes3d, pts_feature, pooled, empty_flag);
    return check_launch("roipool3d");
}

}  // namespace jmb

extern "C" int jmb_roipool3d(int batch, int pts_num, int boxes_num, int feat_len, int sampled,
                             const float *xyz, const float *boxes3d, const float *pts_feature,
                             float *pooled_features, int *pooled_empty_flag, void *stream) {
    return jmb::launch_roipool(false, 0, batch, pts_num, boxes_num, feat_len, sampled, 0.f, xyz, boxes3d,
                               pts_feature, pooled_features, pooled_empty_flag, stream);
}

extern "C" int jmb_roipool3d_canonical(int batch, int pts_num, int boxes_num, int feat_len,
                                       int sampled, float pool_extra_width, const float *xyz,
                                       const float *boxes3d, const float *pts_feature,
                                       float *pooled_features, int *pooled_empty_flag, void *stream) {
    return jmb::launch_roipool(true, 0, batch, pts_num, boxes_num, feat_len, sampled, pool_extra_width,
                               xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag, stream);
}

// jmb_roipool3d_canonical writing the head layout (see roipool3d_kernel): pooled_features is
// (B, M, sampled, round_up(3 + feat_len, 8)) with the first `lead` feature columns moved behind xyz.
extern "C" int jmb_roipool3d_canonical_head(int batch, int pts_num, int boxes_num, int feat_len,
                                            int sampled, float pool_extra_width, int lead, const float *xyz,
                                            const float *boxes3d, const float *pts_feature,
                                            float *pooled_features, int *pooled_empty_flag, void *stream) {
    JMB_REQUIRE(lead > 0, "roipool3d_canonical_head: lead must be positive");
    return jmb::launch_roipool(true, lead, batch, pts_num, boxes_num, feat_len, sampled, pool_extra_width,
                               xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag, stream);
}
