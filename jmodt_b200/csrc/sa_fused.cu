// Fused set-abstraction layer on tcgen05 tensor cores, sm_100a:
//     ball-query indices -> grouped [xyz - centre, features] -> 3-layer shared MLP (+bias, ReLU) -> max over nsample
// in ONE kernel.  Replaces, for one PointnetSAModule (reference jmodt/ops/pointnet2/pointnet2_modules.py:20-63):
// QueryAndGroup's two group_points launches + subtract + cat (pointnet2_utils.py:241-264), three cuDNN 1x1 convs
// over (B, C, npoint, nsample) (pytorch_utils.py:6-33) and F.max_pool2d.  At the RCNN SA0 shape the reference
// round-trips a 549 MB/frame grouped tensor and two 537 MB/frame activations through HBM; here a tile of 128
// (centre, sample) columns never leaves the SM, and NO weight is re-read after the prologue:
//
//   * THE FIRST LAYER IS APPLIED BEFORE THE GATHER.  It is linear up to its ReLU, and grouping only selects columns:
//         relu(W1 . [xyz_j - c ; f_j] + b1) = relu(Z_j + W1x . (xyz_j - c) + b1),     Z = W1f . F  over the n_pts points.
//     Z is one small dense GEMM over the POINTS (tc_gemm_kernel; 8x fewer columns than the npoint x nsample grouped
//     neighbours at the RCNN SA0 shape) and this kernel's gather threads finish the layer on the CUDA cores — three
//     FMAs per channel — while they convert the row to the bf16 hi/lo operand image anyway.  That takes a third of
//     the MMAs (27 of 75 per part at C_in = 128) off the tensor pipe, frees the 80 KB W1 image and its tensor-memory
//     columns, and removes the C_in <= 152 limit of the earlier versions;
//   * layer-2 / layer-3 weights (bf16 hi and lo parts) stay in TENSOR MEMORY and are the A operand of TS-mode MMAs;
//   * a tile is processed as two 64-column parts with separate accumulators, operand-image slices, barriers and
//     epilogue warps, walked interleaved by ONE elected issuer lane (see SF_NP below);
//   * every fp32 product is three bf16 MMAs accumulated in fp32: W_hi.X_hi + W_lo.X_hi + W_hi.X_lo;
//   * the gather reads POINT-MAJOR rows Z (G, n_pts, C1): one neighbour = one contiguous row, fetched with 16-byte
//     loads; the layer-2 operand is therefore staged K-major (B operand, b_major = K).
//
// Shapes: layer widths C1, C2 <= 128 (C1 % 8 == 0) and C3 <= 256 (weights zero-padded to the 128-row tile); any C_in
// (0 = coordinates only: Z is NULL); nsample in {8, 16, 32, 64}; npoint * nsample a multiple of 128.  The ROWS instantiation
// of the same kernel is the input stage of the per-proposal network (see the comment above sa_fused_kernel).
#include "tc_common.cuh"

#include <stdlib.h>

namespace jmb {

// warp roles: 0-7   epilogue (TMEM lane quadrant w&3, part w>>2: one warp drains its quadrant's 64 columns of the part),
//             8-23  gather (four threads per column, one block of four 8-channel k-groups each: a single batch of
//                   eight 16-byte loads in flight per thread),
//             24    MMA issuer (warp-uniform control flow, one elected lane issues), 25 tile scheduler (one thread)
constexpr int SF_EPI_WARPS = 8;
constexpr int SF_GATHER_WARP0 = 8, SF_GATHER_WARPS = 16;
constexpr int SF_ISSUER_WARP = 24, SF_SCHED_WARP = 25;
constexpr int SF_THREADS = (SF_SCHED_WARP + 1) * 32;      // 832
// A tile of 128 columns is processed as two PARTS of 64 columns with their own accumulator columns, operand-image
// slices, barriers and epilogue warps.  The dependent chain of a part is  gather (+ layer 1) -> L2 -> epilogue -> L3 ->
// pooling epilogue;  ONE issuer walks the two chains interleaved (A.L2 B.L2 A.L3 B.L3 ...), so while the tensor pipe runs
// one part's MMAs the other part's epilogue warps drain its accumulator.  When the last layer is a single 128-row block,
// tensor memory has room for a SECOND accumulator per part: layer 2 accumulates into X, layer 3 into Y, and the pooling
// epilogue of tile t (reading Y) overlaps layer 2 of tile t+1 (writing X).
// History (in-kernel timelines under profiles/): one issuer thread per part -> the issuers fell into lock-step and the
// pipe idled during every epilogue (r01/sa_fused_parts.txt); all epilogue warps on one part at a time -> no gain, an
// epilogue is a ~1 100-cycle latency chain whatever its width (r02/sa_fused_timeline_v5a.txt); `if (lane == 0)` around
// tcgen05.mma -> an ELECT/BRA serialisation loop per instruction, ~60 issue cycles per MMA (v5b -> v5c: 3.3 -> 2.2 ms).
constexpr int SF_NP = 2;
constexpr int SF_PART = TC_BN / SF_NP;                              // 64 columns per part
constexpr uint32_t SF_PART_OFF = (SF_PART / 8) * TC_SBO;            // byte offset of part 1 inside an operand image
// kind::f16, BF16 x BF16 -> F32, M=128, N=SF_PART, A K-major, B MN-major / B K-major (layer 1)
constexpr uint32_t SF_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) |
                              ((uint32_t)(SF_PART >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
constexpr uint32_t SF_IDESC_L1 = SF_IDESC & ~(1u << 16);
constexpr int SF_MAXKC1 = 5;
constexpr int SF_CHUNK = 2 * TC_IMG;          // hi + lo image of one 32-row chunk: 16 KB
constexpr int SF_SMEM = (SF_MAXKC1 + SF_MAXKC1 + 4) * SF_CHUNK;   // W1 + X1 + activations = 224 KB
constexpr int SF_TSLOTS = 4;                  // tile ring (dynamic scheduler)
constexpr uint32_t SF_TMEM_W2 = 128, SF_TMEM_W3 = 256;             // column bases (hi at +0, lo at +64 of each block)
constexpr uint32_t SF_TMEM_ACC_Y = 384;                            // second accumulator pair (free when W3 is one block)

struct SaFusedParams {
    const __nv_bfloat16 *w1, *w2, *w3;
    const float *b1, *b2, *b3;
    int K1, Kc1, Mt3;      // ROWS mode: layer-1 depth (8) and its chunk count; Mt3 = 128-row blocks of the last layer
    int C1, C2;            // real widths of layers 1 and 2: rows / k-steps beyond them are zero padding and are skipped
    int C3;                // real width of the last layer (<= 128 * Mt3; rows beyond it are zero padding)
    int G, npoint, nsample, n_pts;
    const float *z;        // SA mode: (G, n_pts, C1) POINT-MAJOR rows of W1f . F (no bias), or null when C_in == 0
    float4 w1x[TC_BM];     // SA mode: rows [w_dx, w_dy, w_dz, b1] of the first layer, by value (constant bank)
    const float *feats;    // ROWS mode: (rows, row_pitch) input rows
    const int *idx;        // (G, npoint, nsample)
    const float *xyz;      // (G, n_pts, 3)
    const float *centres;  // (G, npoint, 3)
    float *out;            // (G, 128*Mt3, npoint), or (G, npoint, 128*Mt3) if out_point_major; ROWS: (rows, 128) or, with
                           // rows_per_group > 0, channel-first (rows / rows_per_group, 128, rows_per_group)
    int rows_per_group;
    int out_point_major;
    int w3_blocks;         // 128 x 128 blocks of W3 resident in tensor memory (SA: Mt3 row blocks; ROWS: 2 K blocks)
    long long rows;        // ROWS mode: number of input rows (points), a multiple of 128
    int row_pitch;         // ROWS mode: floats per input row (multiple of 4; 128 channels, then <= 8 extra inputs)
    int *counter, *done;   // dynamic tile scheduler (tc_sched_slot): tiles are claimed with atomicAdd, so a CTA that shares
                           // its SM with another stream's kernel simply takes fewer tiles; re-armed by the last CTA
    long long *dbg;        // optional timeline buffer (profiling aid): CTA 0 writes clock64() stamps
};

template <uint32_t IDESC>
__device__ __forceinline__ void umma_ss_part(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
template <uint32_t IDESC>
__device__ __forceinline__ void umma_ts_part(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// Row m of a packed 128 x 128 weight block (4 chunk images of 128 x 32) -> 64 TMEM columns (bf16 pairs) of lane m.
__device__ __forceinline__ void weight_rows_to_tmem(const __nv_bfloat16 *wpack_block, int part /*0 hi, 1 lo*/, int m,
                                                    uint32_t taddr, int nchunks = 4 /* chunks present; the rest is zero */) {
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {        // 32 columns (two K chunks) per tcgen05.st
        uint32_t r[32];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int kc = half * 2 + cc;
            const uint8_t *img = reinterpret_cast<const uint8_t *>(wpack_block) + (size_t)kc * SF_CHUNK + (size_t)part * TC_IMG;
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
                const uint4 v = kc < nchunks
                                    ? __ldg(reinterpret_cast<const uint4 *>(img + k8 * TC_LBO + (m >> 3) * TC_SBO + (m & 7) * 16))
                                    : make_uint4(0u, 0u, 0u, 0u);
                r[cc * 16 + k8 * 4 + 0] = v.x; r[cc * 16 + k8 * 4 + 1] = v.y;
                r[cc * 16 + k8 * 4 + 2] = v.z; r[cc * 16 + k8 * 4 + 3] = v.w;
            }
        }
        tmem_st32(taddr + half * 32, r);
    }
}

// ROWS = false: set abstraction (gather through idx, three layers, max-pool over nsample).
// ROWS = true : the input stage of the per-proposal network (reference rcnn.py:172-186: xyz_up_layer 5 -> 128 -> 128,
//   cat with the 128 RPN channels, merge_down_layer 256 -> 128) on consecutive rows [128 channels | up to 8 extra
//   inputs] of pitch `row_pitch`: layer 1 reads only the extra inputs (k-groups 16, 17 of the row image), layer 3 is
//   W3[:, 0:128] . act2 + W3[:, 128:256] . channels (both A blocks in tensor memory, the second with the K-major row
//   image as B), and every column is written point-major (rows, 128) — no pooling.  The (G, 256, 512) concat, both
//   transposes of the pooled tensor and two activation round trips of the unfused path disappear.
template <bool ROWS>
__global__ void __launch_bounds__(SF_THREADS, 1)
sa_fused_kernel(const SaFusedParams p) {
    constexpr int PART = SF_PART;
    extern __shared__ __align__(1024) uint8_t sf_smem[];
    uint8_t *s_w1 = sf_smem;
    uint8_t *s_x1 = sf_smem + SF_MAXKC1 * SF_CHUNK;
    uint8_t *s_act = s_x1 + SF_MAXKC1 * SF_CHUNK;
    // per part: x1_full / x1_free (gather <-> issuer), and per accumulator c (0 = X, 1 = Y): acc_full (issuer -> epilogue),
    // epi_done (epilogue -> issuer)
    __shared__ __align__(8) uint64_t s_x1_full[SF_NP], s_x1_free[SF_NP], s_acc_full[SF_NP][2], s_epi_done[SF_NP][2], s_w1_full,
        s_tfull[SF_TSLOTS], s_tempty[SF_TSLOTS];
    __shared__ int s_tile[SF_TSLOTS];
    __shared__ uint32_t s_tmem_base;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int h = 0; h < SF_NP; ++h) {
            mbar_init(&s_x1_full[h], SF_GATHER_WARPS / SF_NP);   // one arrival per gather warp of the part
            mbar_init(&s_x1_free[h], 1);
            for (int c = 0; c < 2; ++c) {
                mbar_init(&s_acc_full[h][c], 1);
                mbar_init(&s_epi_done[h][c], SF_EPI_WARPS / SF_NP);
            }
        }
        for (int s = 0; s < SF_TSLOTS; ++s) {
            mbar_init(&s_tfull[s], 1);
            mbar_init(&s_tempty[s], SF_EPI_WARPS + SF_GATHER_WARPS + 1);
        }
        mbar_init(&s_w1_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == SF_ISSUER_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    // Image rows between the layer width and the next multiple of 16 are read by the last k-step's MMAs but never
    // written: clear the images once (their weights are zero, but 0 * garbage-NaN would poison the sum).
    for (int i = threadIdx.x; i < (SF_MAXKC1 + 4) * SF_CHUNK / 16; i += SF_THREADS)
        reinterpret_cast<uint4 *>(s_x1)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();

    // ---- prologue: all weights become resident (ROWS: W1 in shared memory; W2 / W3 in tensor memory) ----
    if (ROWS) {
        if (threadIdx.x == SF_ISSUER_WARP * 32) {
            mbar_arrive_expect_tx(&s_w1_full, (uint32_t)p.Kc1 * SF_CHUNK);
            for (int c = 0; c < p.Kc1; ++c)
                bulk_g2s(s_w1 + (size_t)c * SF_CHUNK, p.w1 + (size_t)c * (SF_CHUNK / 2), SF_CHUNK, &s_w1_full);
        }
    }
    if (warp < 4) {
        const int m = warp * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        weight_rows_to_tmem(p.w2, 0, m, lane_base + SF_TMEM_W2);
        weight_rows_to_tmem(p.w2, 1, m, lane_base + SF_TMEM_W2 + 64);
        for (int mt = 0; mt < p.w3_blocks; ++mt) {
            const __nv_bfloat16 *blk = p.w3 + (size_t)mt * 4 * (SF_CHUNK / 2);
            weight_rows_to_tmem(blk, 0, m, lane_base + SF_TMEM_W3 + mt * 128);
            weight_rows_to_tmem(blk, 1, m, lane_base + SF_TMEM_W3 + mt * 128 + 64);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const int N = ROWS ? TC_BN : p.npoint * p.nsample;
    const int Nt = N / TC_BN;
    const int total_tiles = ROWS ? (int)(p.rows / TC_BN) : p.G * Nt;     // < 2^31 (checked by the launcher)
    // SA: layer 2 accumulates into X (columns 0..127), the Mt3 row blocks of layer 3 into Y — a second pair of
    // accumulators when tensor memory has room for it (one W3 block), else X again
    const bool split_acc = !ROWS && p.w3_blocks == 1;
    const uint32_t acc_y = split_acc ? SF_TMEM_ACC_Y : 0u;

    // tile ring, consumer side: slot i is read by every consumer warp and released with one arrival per warp
    auto ring_read = [&](uint32_t i) -> int {
        const int slot = i % SF_TSLOTS;
        mbar_wait(&s_tfull[slot], (i / SF_TSLOTS) & 1);
        return *reinterpret_cast<volatile int *>(&s_tile[slot]);
    };

    if (warp < SF_EPI_WARPS) {
        // ====================================== epilogue warps ======================================
        // warp w drains part h = w >> 2: accumulator rows [32*quad, +32), all 64 columns of the part (two 32-column loads)
        const int quad = warp & 3, h = warp >> 2;
        const int m = quad * 32 + lane;  // accumulator row (output channel) of this thread
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(h * PART);
        uint32_t ph[2] = {0, 0};

        // accumulator part -> next layer's operand image (row m of the accumulator is row k = m of the operand)
        auto epilogue_act = [&](const float *bias_ptr) {
            const float bias = __ldg(bias_ptr + m);
            uint8_t *ahi = s_act + (size_t)quad * SF_CHUNK, *alo = ahi + TC_IMG;
            const uint32_t rowoff = (uint32_t)(lane >> 3) * TC_LBO + (uint32_t)(lane & 7) * 16;
#pragma unroll 1
            for (int sblk = 0; sblk < 2; ++sblk) {
                float v[32];
                tmem_ld32(taddr + (uint32_t)(sblk * 32), v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 hh, ll;
                    float *w = v + q * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[j] = fmaxf(w[j] + bias, 0.f);
                    split2(w[0], w[1], hh.x, ll.x);
                    split2(w[2], w[3], hh.y, ll.y);
                    split2(w[4], w[5], hh.z, ll.z);
                    split2(w[6], w[7], hh.w, ll.w);
                    const uint32_t off = (uint32_t)(h * (PART / 8) + sblk * 4 + q) * TC_SBO + rowoff;
                    *reinterpret_cast<uint4 *>(ahi + off) = hh;
                    *reinterpret_cast<uint4 *>(alo + off) = ll;
                }
            }
        };

        // max over the nsample columns of each centre (nsample divides 64: a window never leaves the part)
        auto epilogue_pool = [&](int tile, int mt) {
            const int g = tile / Nt;
            const int nt = tile - g * Nt;
            const float bias = __ldg(p.b3 + mt * TC_BM + m);
            const int c3 = p.C3, ch = mt * TC_BM + m, ctr0 = (nt * TC_BN + h * PART) / p.nsample;
            const bool live = ch < c3;     // zero-padded rows of a layer narrower than the 128-row MMA tile are not stored
                                           // (no early return: tcgen05.ld below is warp-collective)
            // element (centre w) of this channel: channel-first out[g][ch][ctr0 + w]  or  point-major out[g][ctr0 + w][ch]
            float *orow = p.out_point_major ? p.out + ((size_t)g * p.npoint + ctr0) * c3 + ch
                                            : p.out + ((size_t)g * c3 + ch) * p.npoint + ctr0;
            const size_t ostride = p.out_point_major ? (size_t)c3 : 1;
            const int sub = p.nsample < 32 ? p.nsample : 32;
            float run = -INFINITY;
#pragma unroll 1
            for (int c0 = 0; c0 < PART; c0 += 32) {
                // bias and ReLU AFTER the maximum: relu(. + b) is monotone and b is per row, so
                // max_j relu(v_j + b) = relu(max_j v_j + b) exactly — one add / max per window instead of per element
                float v[32];
                tmem_ld32(taddr + acc_y + (uint32_t)c0, v);
                if (sub == 32) {
                    float mx = v[0];
#pragma unroll
                    for (int j = 1; j < 32; ++j) mx = fmaxf(mx, v[j]);
                    run = fmaxf(run, mx);
                    if ((c0 + 32) % p.nsample == 0) {
                        if (live) orow[(size_t)((c0 + 32) / p.nsample - 1) * ostride] = fmaxf(run + bias, 0.f);
                        run = -INFINITY;
                    }
                } else {
                    for (int w0 = 0; w0 < 32; w0 += sub) {
                        float mx = -INFINITY;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j >= w0 && j < w0 + sub) mx = fmaxf(mx, v[j]);
                        if (live) orow[(size_t)((c0 + w0) / p.nsample) * ostride] = fmaxf(mx + bias, 0.f);
                    }
                }
            }
        };

        // ROWS: every column is an output row
        auto epilogue_rows = [&](int tile) {
            const float bias = __ldg(p.b3 + m);
#pragma unroll 1
            for (int sblk = 0; sblk < 2; ++sblk) {
                float v[32];
                tmem_ld32(taddr + (uint32_t)(sblk * 32), v);
                const size_t row0 = (size_t)tile * TC_BN + h * PART + sblk * 32;
                if (p.rows_per_group > 0) {
                    // channel-first (group, 128, rows_per_group): this thread's 32 consecutive rows of channel m are 128
                    // contiguous bytes (a tile never straddles a group: rows_per_group is a multiple of 128)
                    const size_t grp = row0 / (size_t)p.rows_per_group, r = row0 - grp * (size_t)p.rows_per_group;
                    float4 *dst4 = reinterpret_cast<float4 *>(p.out + (grp * TC_BM + m) * (size_t)p.rows_per_group + r);
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        dst4[j >> 2] = make_float4(fmaxf(v[j] + bias, 0.f), fmaxf(v[j + 1] + bias, 0.f),
                                                   fmaxf(v[j + 2] + bias, 0.f), fmaxf(v[j + 3] + bias, 0.f));
                } else {
                    float *dst = p.out + row0 * TC_BM + m;      // a warp's 32 channels of one row are one 128-byte store
#pragma unroll
                    for (int j = 0; j < 32; ++j) dst[(size_t)j * TC_BM] = fmaxf(v[j] + bias, 0.f);
                }
            }
        };

        long long *dbg = (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) ? p.dbg + 512 : nullptr;
        int di = 0;
#define SF_ESTAMP(tag) do { if (dbg && di < 480) { dbg[di++] = (tag); dbg[di++] = clock64(); } } while (0)
        auto wait_full = [&](int c) {
            mbar_wait(&s_acc_full[h][c], ph[c]); ph[c] ^= 1;
            tc_fence_after();
        };
        auto hand_back = [&](int c) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_epi_done[h][c]);
        };
        uint32_t ti = 0;
        int tile;
        while ((tile = ring_read(ti)) >= 0) {
            if (ROWS) {
                for (int step = 0; step < 3; ++step) {
                    wait_full(0);
                    if (step < 2) {
                        epilogue_act(step == 0 ? p.b1 : p.b2);
                        fence_proxy_async();
                    } else {
                        epilogue_rows(tile);
                    }
                    hand_back(0);
                }
            } else {
                SF_ESTAMP(20);
                wait_full(0);
                SF_ESTAMP(30);
                // accumulator rows >= the layer's real width are zero padding nobody reads (the next layer's MMAs stop at
                // its last real k-step): their quadrants' warps only hand the barrier on
                if (quad * 32 < ((p.C2 + 15) & ~15)) {
                    epilogue_act(p.b2);
                    fence_proxy_async();
                }
                hand_back(0);
                SF_ESTAMP(40);
                for (int mt = 0; mt < p.Mt3; ++mt) {
                    SF_ESTAMP(21 + mt);
                    wait_full(1);
                    SF_ESTAMP(31 + mt);
                    if (mt * TC_BM + quad * 32 < p.C3) epilogue_pool(tile, mt);
                    hand_back(1);
                    SF_ESTAMP(41 + mt);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_tempty[ti % SF_TSLOTS]);
            ++ti;
        }
    } else if (warp < SF_ISSUER_WARP) {
        // ====================================== gather warps ======================================
        const int tg = threadIdx.x - SF_GATHER_WARP0 * 32;
        const int sub = tg >> 7;                 // which block of four k-groups this thread fetches (0..3)
        const int col = tg & 127, h = col / PART;
        const int n_groups = ROWS ? p.row_pitch / 8 : p.C1 / 8;   // k-groups of 8 channels this column's image row holds
        // Four threads = one column: the contiguous point-major row is split to bf16 hi/lo and stored as 16-byte slots of
        // the K-major operand image (offset = (k/8)*LBO + (n/8)*SBO + (n%8)*16; consecutive threads -> consecutive slots:
        // conflict-free).
        const uint32_t noff = (uint32_t)(col >> 3) * TC_SBO + (uint32_t)(col & 7) * 16;
        auto put = [&](int kg, const float (&v)[8]) {
            uint4 hh, ll;
            split2(v[0], v[1], hh.x, ll.x);
            split2(v[2], v[3], hh.y, ll.y);
            split2(v[4], v[5], hh.z, ll.z);
            split2(v[6], v[7], hh.w, ll.w);
            uint8_t *img = s_x1 + (size_t)(kg >> 2) * SF_CHUNK + (uint32_t)(kg & 3) * TC_LBO + noff;
            *reinterpret_cast<uint4 *>(img) = hh;
            *reinterpret_cast<uint4 *>(img + TC_IMG) = ll;
        };
        auto publish = [&](uint32_t ti) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&s_x1_full[h]);
                mbar_arrive(&s_tempty[ti % SF_TSLOTS]);
            }
        };
        uint32_t ti = 0;
        int tile;
        if (ROWS) {
            // ROWS: consecutive rows of the input (pulled into L2 ahead of time by the scheduler), 16-byte loads
            while ((tile = ring_read(ti)) >= 0) {
                const float *frow = p.feats + ((size_t)tile * TC_BN + col) * p.row_pitch;
                bool waited = ti == 0;
                for (int kg0 = sub * 4; kg0 < n_groups; kg0 += 16) {
                    float4 a4[4], b4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (kg0 + u < n_groups) {
                            a4[u] = __ldg(reinterpret_cast<const float4 *>(frow + (kg0 + u) * 8));
                            b4[u] = __ldg(reinterpret_cast<const float4 *>(frow + (kg0 + u) * 8) + 1);
                        }
                    if (!waited) {     // the MMAs that read this part of the previous tile are done (loads already in flight)
                        mbar_wait(&s_x1_free[h], (ti - 1) & 1);
                        waited = true;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (kg0 + u < n_groups) {
                            const float v[8] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w, b4[u].x, b4[u].y, b4[u].z, b4[u].w};
                            put(kg0 + u, v);
                        }
                }
                if (!waited) mbar_wait(&s_x1_free[h], (ti - 1) & 1);
                publish(ti);
                ++ti;
            }
        } else {
            // SA: the row is Z = W1f . f of the neighbour idx[n]; the thread adds W1x . (xyz - centre) + b1 and applies the
            // ReLU, i.e. it emits layer 1's OUTPUT as layer 2's operand.  W1x / b1 come from the kernel-parameter
            // constant bank (a warp-uniform index: no shared-memory traffic — broadcast LDS.128 of a 2 KB table cost
            // 2 000 shared-memory cycles per tile and made this stage the bottleneck, profiles/r02/sa_fused_timeline_v6a.txt).
            // The neighbour index of the NEXT tile is fetched while this one is produced.
            auto col_index = [&](int tl) -> int {
                const int g = tl / Nt;
                const int n = (tl - g * Nt) * TC_BN + col;
                return __ldg(p.idx + (size_t)g * N + n);
            };
            const int kg0 = sub * 4;       // C1 <= 128: at most 16 k-groups, one block of four per thread
            tile = ring_read(0);
            int pi = tile >= 0 ? col_index(tile) : 0;
            while (tile >= 0) {
                const int next_tile = ring_read(ti + 1);
                const int pi_next = next_tile >= 0 ? col_index(next_tile) : 0;     // in flight while this tile is produced
                const int g = tile / Nt;
                const int n = (tile - g * Nt) * TC_BN + col;
                float4 a4[4], b4[4];
                float dx = 0.f, dy = 0.f, dz = 0.f;
                if (kg0 < n_groups) {
                    const float *frow = p.z ? p.z + ((size_t)g * p.n_pts + pi) * p.C1 : nullptr;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        a4[u] = b4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (frow && kg0 + u < n_groups) {
                            a4[u] = __ldg(reinterpret_cast<const float4 *>(frow + (kg0 + u) * 8));
                            b4[u] = __ldg(reinterpret_cast<const float4 *>(frow + (kg0 + u) * 8) + 1);
                        }
                    }
                    // relative coordinates (pointnet2_utils.py:252: grouped_xyz -= new_xyz)
                    const float *cen = p.centres + ((size_t)g * p.npoint + n / p.nsample) * 3;
                    const float *pt = p.xyz + ((size_t)g * p.n_pts + pi) * 3;
                    dx = __fsub_rn(__ldg(pt), __ldg(cen));
                    dy = __fsub_rn(__ldg(pt + 1), __ldg(cen + 1));
                    dz = __fsub_rn(__ldg(pt + 2), __ldg(cen + 2));
                }
                if (ti > 0) mbar_wait(&s_x1_free[h], (ti - 1) & 1);     // the MMAs that read this part of the previous tile are done
                if (kg0 < n_groups) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (kg0 + u < n_groups) {
                            float v[8] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w, b4[u].x, b4[u].y, b4[u].z, b4[u].w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 w = p.w1x[(kg0 + u) * 8 + j];     // constant bank, warp-uniform index
                                // b1 is already in Z (the points' GEMM adds it); without input features it comes from w.w
                                const float base = p.z ? v[j] : w.w;
                                v[j] = fmaxf(fmaf(w.x, dx, fmaf(w.y, dy, fmaf(w.z, dz, base))), 0.f);
                            }
                            put(kg0 + u, v);
                        }
                }
                publish(ti);
                tile = next_tile;
                pi = pi_next;
                ++ti;
            }
        }
    } else if (warp == SF_ISSUER_WARP) {
        // ====================================== MMA issuer ======================================
        // The whole warp walks the loops and waits on the barriers (uniform control flow); the MMAs and commits of one
        // (layer, part) are issued by one elected lane (see elect_one()).  Descriptors are formed by adding constants
        // to a base descriptor (the start-address field is the low 14 bits; images never cross it).
        if (ROWS) mbar_wait(&s_w1_full, 0);
        uint32_t dph[SF_NP][2] = {{0, 0}, {0, 0}};      // phases of epi_done[h][c]
        const uint64_t w1_desc = make_smem_desc(smem_u32(s_w1));
        const uint64_t x1_desc0 = make_smem_desc(smem_u32(s_x1));
        const uint64_t act_desc0 = make_smem_desc(smem_u32(s_act));
        constexpr uint64_t D_IMG = TC_IMG >> 4, D_K16 = (2 * TC_LBO) >> 4, D_CHUNK = SF_CHUNK >> 4, D_PART = SF_PART_OFF >> 4;
        const int nk2 = (p.C1 + 15) / 16, nk3 = (p.C2 + 15) / 16;     // k-steps that hold real input rows
        long long *dbg = (p.dbg && blockIdx.x == 0 && lane == 0) ? p.dbg : nullptr;
        int di = 0;
#define SF_STAMP(tag) do { if (dbg && di < 480) { dbg[di++] = (tag); dbg[di++] = clock64(); } } while (0)
        auto wait_done = [&](int h, int c) {
            mbar_wait(&s_epi_done[h][c], dph[h][c]); dph[h][c] ^= 1;
        };
        uint32_t ti = 0;
        int tile;
        while ((tile = ring_read(ti)) >= 0) {
            if (ROWS) {
                for (int step = 0; step < 3; ++step) {
#pragma unroll
                    for (int h = 0; h < SF_NP; ++h) {
                        const uint32_t acc = tmem_base + (uint32_t)h * PART;
                        const uint64_t x1_desc = x1_desc0 + (uint64_t)h * D_PART;
                        const uint64_t act_desc = act_desc0 + (uint64_t)h * D_PART;
                        if (ti > 0 || step > 0) wait_done(h, 0);     // accumulator free / activation image ready
                        if (step == 0) mbar_wait(&s_x1_full[h], ti & 1);
                        tc_fence_after();
                        if (elect_one()) {
                            if (step == 0) {
                                // layer 1 reads only the extra inputs: k-groups 16, 17 = first k16 step of image chunk 4
                                const uint64_t xd = x1_desc + 4 * D_CHUNK;
                                umma_ss_part<SF_IDESC_L1>(acc, w1_desc, xd, 0);
                                umma_ss_part<SF_IDESC_L1>(acc, w1_desc + D_IMG, xd, 1);
                                umma_ss_part<SF_IDESC_L1>(acc, w1_desc, xd + D_IMG, 1);
                            } else {
                                // layers 2 and 3 (TS): A = weights resident in tensor memory, B = the activation image
                                const uint32_t wcol = tmem_base + (step == 1 ? SF_TMEM_W2 : SF_TMEM_W3);
#pragma unroll 2
                                for (int k16 = 0; k16 < 8; ++k16) {
                                    const uint64_t xd = act_desc + (uint64_t)(k16 >> 1) * D_CHUNK + (uint64_t)(k16 & 1) * D_K16;
                                    const uint32_t ahi = wcol + (uint32_t)k16 * 8;
                                    umma_ts_part<SF_IDESC>(acc, ahi, xd, k16 != 0);
                                    umma_ts_part<SF_IDESC>(acc, ahi + 64, xd, 1);
                                    umma_ts_part<SF_IDESC>(acc, ahi, xd + D_IMG, 1);
                                }
                                if (step == 2) {   // + W3[:, 128:256] . channels: A block 1 in tensor memory, B = the K-major row image
#pragma unroll 2
                                    for (int k16 = 0; k16 < 8; ++k16) {
                                        const uint64_t xd = x1_desc + (uint64_t)(k16 >> 1) * D_CHUNK + (uint64_t)(k16 & 1) * D_K16;
                                        const uint32_t ahi = wcol + 128 + (uint32_t)k16 * 8;
                                        umma_ts_part<SF_IDESC_L1>(acc, ahi, xd, 1);
                                        umma_ts_part<SF_IDESC_L1>(acc, ahi + 64, xd, 1);
                                        umma_ts_part<SF_IDESC_L1>(acc, ahi, xd + D_IMG, 1);
                                    }
                                    umma_commit(&s_x1_free[h]);
                                }
                            }
                            umma_commit(&s_acc_full[h][0]);
                        }
                        __syncwarp();
                    }
                }
            } else {
                // ---- layer 2 of both parts: A = W2 in tensor memory, B = layer 1's output as the gather wrote it (K-major).
                // X is free: layer 3 of the previous tile (issued earlier) waited for the epilogue that drained it.  Without
                // the second accumulator pair X also holds layer 3, so its last pooling epilogue must be over.
#pragma unroll
                for (int h = 0; h < SF_NP; ++h) {
                    const uint32_t acc = tmem_base + (uint32_t)h * PART;
                    const uint64_t x1_desc = x1_desc0 + (uint64_t)h * D_PART;
                    SF_STAMP(1 + h);
                    if (!split_acc && ti > 0) wait_done(h, 1);
                    mbar_wait(&s_x1_full[h], ti & 1);
                    tc_fence_after();
                    SF_STAMP(10 + h);
                    if (elect_one()) {
#pragma unroll 2
                        for (int k16 = 0; k16 < nk2; ++k16) {
                            const uint64_t xd = x1_desc + (uint64_t)(k16 >> 1) * D_CHUNK + (uint64_t)(k16 & 1) * D_K16;
                            const uint32_t ahi = tmem_base + SF_TMEM_W2 + (uint32_t)k16 * 8;
                            umma_ts_part<SF_IDESC_L1>(acc, ahi, xd, k16 != 0);
                            umma_ts_part<SF_IDESC_L1>(acc, ahi + 64, xd, 1);
                            umma_ts_part<SF_IDESC_L1>(acc, ahi, xd + D_IMG, 1);
                        }
                        umma_commit(&s_x1_free[h]);      // the gather warps may refill this part for the next tile
                        umma_commit(&s_acc_full[h][0]);
                    }
                    __syncwarp();
                }
                // ---- the Mt3 row blocks of layer 3: A = W3 block in tensor memory, B = layer 2's output (MN-major)
                for (int mt = 0; mt < p.Mt3; ++mt) {
#pragma unroll
                    for (int h = 0; h < SF_NP; ++h) {
                        const uint32_t acc = tmem_base + acc_y + (uint32_t)h * PART;
                        const uint64_t act_desc = act_desc0 + (uint64_t)h * D_PART;
                        SF_STAMP(3 + mt * 2 + h);
                        if (mt == 0) {
                            wait_done(h, 0);                                  // layer 2's epilogue wrote the activation image
                            if (split_acc && ti > 0) wait_done(h, 1);         // Y drained by the previous tile's pooling
                        } else {
                            wait_done(h, 1);                                  // previous row block pooled
                        }
                        tc_fence_after();
                        SF_STAMP(12 + mt * 2 + h);
                        if (elect_one()) {
                            const uint32_t wcol = tmem_base + SF_TMEM_W3 + (uint32_t)mt * 128;
#pragma unroll 2
                            for (int k16 = 0; k16 < nk3; ++k16) {
                                const uint64_t xd = act_desc + (uint64_t)(k16 >> 1) * D_CHUNK + (uint64_t)(k16 & 1) * D_K16;
                                const uint32_t ahi = wcol + (uint32_t)k16 * 8;
                                umma_ts_part<SF_IDESC>(acc, ahi, xd, k16 != 0);
                                umma_ts_part<SF_IDESC>(acc, ahi + 64, xd, 1);
                                umma_ts_part<SF_IDESC>(acc, ahi, xd + D_IMG, 1);
                            }
                            umma_commit(&s_acc_full[h][1]);
                        }
                        __syncwarp();
                    }
                }
            }
            if (lane == 0) mbar_arrive(&s_tempty[ti % SF_TSLOTS]);
            ++ti;
        }
        SF_STAMP(99);
    } else {
        if (lane == 0) {
            // ====================================== tile scheduler ======================================
            uint32_t ti = 0;
            for (;;) {
                int t = atomicAdd(p.counter, 1);
                if (t >= total_tiles) t = -1;
                if (ROWS && t >= 0) {
                    // the rows of a tile are contiguous: pull them into L2 ahead of the gather (the scheduler runs up to
                    // SF_TSLOTS tiles ahead), whose threads alone cannot keep enough HBM requests in flight
                    const uint32_t bytes = (uint32_t)(TC_BN * p.row_pitch * 4);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(
                                     p.feats + (size_t)t * TC_BN * p.row_pitch), "r"(bytes) : "memory");
                }
                const int slot = ti % SF_TSLOTS;
                mbar_wait(&s_tempty[slot], ((ti / SF_TSLOTS) & 1) ^ 1);
                *reinterpret_cast<volatile int *>(&s_tile[slot]) = t;
                mbar_arrive(&s_tfull[slot]);
                ++ti;
                if (t < 0) break;
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == SF_ISSUER_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
    if (threadIdx.x == 0) {
        // the last CTA to leave re-arms the scheduler for the next launch that uses this slot
        __threadfence();
        const int old = atomicAdd(p.done, 1);
        if (old == (int)gridDim.x - 1) {
            *p.counter = 0;
            *p.done = 0;
            __threadfence();
        }
    }
}

}  // namespace jmb

extern "C" int jmb_sa_fused(const float *z, const float *w1x, const void *w2, const float *b2, const void *w3,
                            const float *b3, int C1, int C2, int C3, int G, int npoint, int nsample, int n_pts,
                            const int *idx, const float *xyz, const float *centres, float *out, int out_point_major,
                            void *stream) {
    using namespace jmb;
    static long long *dbg_buf = nullptr;          // JMB_SA_DEBUG=1: CTA 0 records a clock64() timeline (profiling aid)
    static int dbg_on = -1;
    if (dbg_on < 0) {
        const char *e = getenv("JMB_SA_DEBUG");
        dbg_on = (e && e[0] == '1') ? 1 : 0;
        if (dbg_on) { cudaMalloc(&dbg_buf, 1024 * sizeof(long long)); cudaMemset(dbg_buf, 0, 1024 * sizeof(long long)); }
    }
    JMB_REQUIRE(G >= 0 && npoint > 0 && nsample > 0 && n_pts > 0, "sa_fused: bad sizes");
    if (G == 0) return JMB_OK;
    JMB_REQUIRE(w1x && w2 && w3 && b2 && b3 && idx && xyz && centres && out, "sa_fused: null pointer");
    JMB_REQUIRE((reinterpret_cast<uintptr_t>(z) & 15u) == 0, "sa_fused: z must be 16-byte aligned");
    JMB_REQUIRE(C3 >= 1 && C3 <= 256, "sa_fused: last layer width %d must be in 1..256", C3);
    JMB_REQUIRE(C1 >= 8 && C1 <= 128 && C1 % 8 == 0, "sa_fused: first layer width %d must be a multiple of 8 in 8..128", C1);
    JMB_REQUIRE(C2 >= 1 && C2 <= 128, "sa_fused: second layer width %d must be in 1..128", C2);
    JMB_REQUIRE(nsample % 8 == 0 && 64 % nsample == 0, "sa_fused: nsample must be 8, 16, 32 or 64");
    JMB_REQUIRE(((long long)npoint * nsample) % TC_BN == 0, "sa_fused: npoint*nsample must be a multiple of 128");
    SaFusedParams p;
    p = SaFusedParams{};
    p.w2 = (const __nv_bfloat16 *)w2; p.w3 = (const __nv_bfloat16 *)w3;
    p.b2 = b2; p.b3 = b3; p.z = z;
    for (int k = 0; k < C1; ++k) p.w1x[k] = make_float4(w1x[4 * k], w1x[4 * k + 1], w1x[4 * k + 2], w1x[4 * k + 3]);
    p.Mt3 = div_up(C3, TC_BM); p.C3 = C3; p.C1 = C1; p.C2 = C2;
    p.w3_blocks = p.Mt3; p.rows = 0; p.row_pitch = 0;
    p.G = G; p.npoint = npoint; p.nsample = nsample; p.n_pts = n_pts;
    p.idx = idx; p.xyz = xyz; p.centres = centres; p.out = out; p.out_point_major = out_point_major; p.dbg = dbg_buf;
    int dev = 0, sms = 0;
    {
        const int rc = device_info(&dev, &sms);
        if (rc != JMB_OK) return rc;
    }
    JMB_FUNC_ATTR_ONCE((sa_fused_kernel<false>), cudaFuncAttributeMaxDynamicSharedMemorySize, SF_SMEM, dev);
    {
        const int rc = tc_sched_slot(dev, &p.counter, &p.done);
        if (rc != JMB_OK) return rc;
    }
    const long long tiles = (long long)G * ((long long)npoint * nsample / TC_BN);
    JMB_REQUIRE(tiles < (1LL << 31), "sa_fused: too many tiles");
    const int grid = (int)(tiles < sms ? tiles : sms);
    sa_fused_kernel<false><<<grid, SF_THREADS, SF_SMEM, (cudaStream_t)stream>>>(p);
    if (dbg_on) {
        long long hbuf[1024];
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaMemcpy(hbuf, dbg_buf, sizeof(hbuf), cudaMemcpyDeviceToHost);
        for (int part = 0; part < 2; ++part) {
            const long long *b = hbuf + part * 512;
            fprintf(stderr, "[sa_fused timeline %s] ", part ? "epilogue warp 0" : "issuer h0");
            for (int i = 0; i + 1 < 160 && b[i]; i += 2) fprintf(stderr, "%lld:%lld ", b[i], b[i + 1] - hbuf[1]);
            fprintf(stderr, "\n");
        }
        cudaMemset(dbg_buf, 0, 1024 * sizeof(long long));
    }
    return check_launch("sa_fused");
}

// Input stage of the per-proposal network in one kernel (reference rcnn.py:172-186):
//   in  (rows, row_pitch) fp32: [128 RPN channels | x, y, z, mask, depth | zero padding]  (roipool3d "head layout")
//   out (rows, 128) fp32 point-major = merge_down( cat( xyz_up(extra inputs), channels ) )
// w1: 128 x 8 (extra inputs, zero padded), w2: 128 x 128, w3: 128 x 256, all packed as tc.PackedLayer images.
extern "C" int jmb_rcnn_input_fused(const void *w1, const float *b1, const void *w2, const float *b2, const void *w3,
                                    const float *b3, long long rows, int row_pitch, const float *in, float *out,
                                    int rows_per_group, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(rows >= 0, "rcnn_input_fused: negative size");
    if (rows == 0) return JMB_OK;
    JMB_REQUIRE(w1 && w2 && w3 && b1 && b2 && b3 && in && out, "rcnn_input_fused: null pointer");
    JMB_REQUIRE(rows % TC_BN == 0, "rcnn_input_fused: rows = %lld must be a multiple of 128", rows);
    JMB_REQUIRE(row_pitch == 136, "rcnn_input_fused: row pitch %d (expected 128 channels + 8 extra inputs)", row_pitch);
    JMB_REQUIRE(rows_per_group >= 0 && rows_per_group % TC_BN == 0 && (rows_per_group == 0 || rows % rows_per_group == 0),
                "rcnn_input_fused: rows_per_group %d must be 0 or a multiple of 128 dividing the row count", rows_per_group);
    JMB_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15u) == 0, "rcnn_input_fused: input must be 16-byte aligned");
    SaFusedParams p = {};
    p.w1 = (const __nv_bfloat16 *)w1; p.w2 = (const __nv_bfloat16 *)w2; p.w3 = (const __nv_bfloat16 *)w3;
    p.b1 = b1; p.b2 = b2; p.b3 = b3;
    p.K1 = 8; p.Kc1 = 1; p.Mt3 = 1; p.C3 = 128; p.C1 = 128; p.C2 = 128; p.w3_blocks = 2; p.rows_per_group = rows_per_group;
    p.G = 1; p.npoint = 1; p.nsample = TC_BN; p.n_pts = 0;
    p.feats = in; p.out = out; p.out_point_major = 1; p.rows = rows; p.row_pitch = row_pitch;
    int dev = 0, sms = 0;
    {
        const int rc = device_info(&dev, &sms);
        if (rc != JMB_OK) return rc;
    }
    JMB_FUNC_ATTR_ONCE((sa_fused_kernel<true>), cudaFuncAttributeMaxDynamicSharedMemorySize, SF_SMEM, dev);
    {
        const int rc = tc_sched_slot(dev, &p.counter, &p.done);
        if (rc != JMB_OK) return rc;
    }
    const long long tiles = rows / TC_BN;
    JMB_REQUIRE(tiles < (1LL << 31), "rcnn_input_fused: too many rows");
    const int grid = (int)(tiles < sms ? tiles : sms);
    sa_fused_kernel<true><<<grid, SF_THREADS, SF_SMEM, (cudaStream_t)stream>>>(p);
    return check_launch("rcnn_input_fused");
}
