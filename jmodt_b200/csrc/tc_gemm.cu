// Pointwise (1x1-conv) MLP layer on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//     Y[g] = act( W · X[g] + bias )        W: (M, K) fp32,  X[g]: (K, N) channel-first,  Y[g]: (M, N)
//
// This is the building block of SharedMLP / Conv1d stacks (reference
// jmodt/ops/pointnet2/pytorch_utils.py:6-33,127-198 run through cuDNN) in the layout the reference
// already uses — channel-first (B, C, npoint, nsample) — so no transposes are needed: W is the A
// operand (K-major), X is the B operand in MN-major form (n contiguous), and one accumulator row per
// TMEM lane is one OUTPUT CHANNEL.  That makes the per-channel bias a per-thread scalar and turns the
// set-abstraction max-pool over nsample (pointnet2_modules.py:50-52) into a max over a thread's own
// registers — the (B, C, npoint, nsample) activation of the last layer is never written.
//
// fp32-grade results from bf16 tensor cores: every fp32 operand is split x = hi + lo (two bf16), and
// the product is accumulated in fp32 as  W_hi·X_hi + W_lo·X_hi + W_hi·X_lo  (the dropped lo·lo term and
// the split residuals are <= 2^-16 relative per product).
//
// Prologues:  dense X from global memory, or the fused grouping of QueryAndGroup / GroupAll
// (pointnet2_utils.py:241-290): X[k][n] = xyz[idx[n]][k] - centre[n / nsample][k] for k < 3 and
// feats[k-3][idx[n]] otherwise — the grouped tensor is never materialised.
//
// Structure (v4): one persistent CTA per SM, 19 warps, every hand-off through mbarriers (one arrival per warp).
//   warp 17      scheduler: claims tiles from a global atomic counter (dynamic scheduling: a CTA that shares its SM
//                with another stream's kernel simply claims fewer tiles) and publishes them in a shared-memory tile
//                ring;  warp 18 streams the weights: one 16 KB cp.async.bulk chunk image per K chunk.
//   warps 0-7    converters: raw fp32 chunk (shared memory) -> bf16 hi/lo split -> MN-major core-matrix images.
//                Dense X is staged by the converters themselves with 16-byte cp.async (LDGSTS) into a 4-stage raw ring
//                whose mbarrier counts the copies (cp.async.mbarrier.arrive.noinc), as two independent groups of
//                four warps on alternate K chunks, each with two chunks in flight; no registers held.  (v3 staged each 512-byte row with its own cp.async.bulk and
//                wrote each output row with another: ~780 bulk requests per tile at ~30 ns each through the SM's one
//                TMA unit — 23 us per 128x128x512 tile, ncu tensor pipe 15 %.  Bulk copies are now only used where one
//                request moves 16 KB.)  In gather mode they load through the neighbour index (register double-buffer).
//   warp 16      MMA issuer: six tcgen05.mma (128x128x16) per 32-deep K chunk into a double-buffered accumulator.
//   warps 8-15   epilogue: tcgen05.ld -> bias -> ReLU -> (a) staged in shared memory and written back by the same warp
//                as 256-byte row segments (16 lanes x 16 B per row, two rows per store instruction);
//                (b) max-pool over nsample in registers; (c) point-major rows.
// Tiles narrower than 128 columns (N = 8..64 per group, e.g. GroupAll over 32 points) pack 128/N groups into one
// tile, so the tensor cores and the converters never work on padding columns.
#include "tc_common.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace jmb {

constexpr int TC_XSTAGES = 2;                      // converted X ring: 16 KB per stage (hi + lo image)
constexpr int TC_WSTAGES = 3;                      // W ring: 16 KB per stage
constexpr int TC_RSTAGES = 4;                      // raw fp32 X ring
constexpr int TC_TSLOTS = 8;                       // tile ring
constexpr int TC_CHUNK = 2 * TC_IMG;               // hi + lo image of one 32-row chunk
constexpr int TC_RAW_ROW = TC_BN * 4 + 16;         // padded row: 8 consecutive rows hit 8 distinct bank groups
constexpr int TC_RAW_STAGE = TC_BK * TC_RAW_ROW;   // 16 896 B
constexpr int TC_OUT_STAGE = TC_BM * TC_RAW_ROW;   // 67 584 B
constexpr int TC_OFF_X = 0;
constexpr int TC_OFF_W = TC_OFF_X + TC_XSTAGES * TC_CHUNK;
constexpr int TC_OFF_RAW = TC_OFF_W + TC_WSTAGES * TC_CHUNK;
constexpr int TC_OFF_OUT = TC_OFF_RAW + TC_RSTAGES * TC_RAW_STAGE;
constexpr int TC_SMEM = TC_OFF_OUT + TC_OUT_STAGE;  // 217 088 B -> one CTA per SM
constexpr int TC_CONV_WARPS = 8, TC_EPI_WARPS = 8;
constexpr int TC_THREADS = (TC_CONV_WARPS + TC_EPI_WARPS + 3) * 32;   // 608
constexpr int TC_ACC_BUFS = 2;                     // double-buffered accumulator: epilogue overlaps the next tile's MMAs

struct TcGemmParams {
    const __nv_bfloat16 *wpack;   // [Mt][Kc][2][TC_IMG/2] chunk images (see pack_weights in tc.py)
    const float *bias;            // [Mt*128] (zero padded), may be null
    int M, K, Mt, Kc;
    int G, N;                     // groups, columns per group
    int P, Nt, Nshift;            // groups per tile (1 unless N < 128 divides 128), column tiles per group, log2(N) if P > 1
    long long col_tiles;          // G * Nt (P == 1) or ceil(G / P)
    int mode;                     // 0 dense, 1 grouped gather, 2 3x3 convolution (channels-last rows, K-major X images)
    int use_raw;                  // dense X staged by bulk copies (needs 16-byte aligned rows)
    int bulk_out;                 // dense Y staged in shared memory and written as 16-byte pieces (needs 16-byte aligned rows)
    const float *x;               // dense: (G, K, N)   gather: feats (G, K-3, n_pts)
    long long x_group_stride;     // elements
    int x_row_stride;             // elements between consecutive k rows
    const int *idx;               // gather: (G, N) point index per column (null: column n -> point n % n_pts)
    const float *xyz;             // gather: (G, n_pts, 3)
    const float *centres;         // gather: (G, N / nsample, 3) or null (GroupAll: no centring)
    int nsample, n_pts;
    // mode 2, 3x3 convolution (padding 1) as an implicit GEMM over a channels-last input: x (G, cv_H, cv_W, C) with
    // C = 1 << cv_cshift channels per pixel, column n of group g = output pixel (n / cv_OW, n % cv_OW), K index
    // k = tap * C + channel, tap = 3 dy + dx reading input pixel (oy * cv_stride + dy - 1, ox * cv_stride + dx - 1)
    int cv_cshift, cv_H, cv_W, cv_OW, cv_stride;
    int cv_taps;                  // 9, or 1: "rows" mode — x (G, N, C) point-major, column n reads row n (a 1x1 convolution)
    int out_mode;                 // 0 dense (G, M, N), 1 max over `pool` consecutive columns -> (G, M, N / pool),
                                  // 2 point-major (G, N, M): a warp's 32 channels of one column are one 128-byte store,
                                  // 3 dot: the NEXT layer when it has one output channel, y (G, 4 Mt, N) = per 32-row
                                  //   block the partial sums  sum_m dot_w[m] * act(W x + b)[m]  (added up by the caller)
    const float *dot_w;           // out_mode 3: (Mt * 128) weights of the single-output layer that follows, zero padded
    int pool, relu;
    float *y;
    long long y_group_stride;     // elements between consecutive groups of y
    int *counter, *done;          // dynamic tile scheduler (reset by the last CTA to leave)
    long long *dbg;               // optional clock64 timeline of CTA 0 (JMB_TC_DEBUG=1): [0..511] MMA issuer, [512..] converter warp 0
};

#define TC_STAMP(buf, idx, tag) do { if ((buf) && (idx) < 500) { (buf)[(idx)++] = (tag); (buf)[(idx)++] = clock64(); } } while (0)

// column c (0..127) of column tile ct -> group g and column n inside the group
__device__ __forceinline__ void tc_col(const TcGemmParams &p, long long ct64, int c, int &g, int &n) {
    const int ct = (int)ct64;               // col_tiles * Mt < 2^30 (checked by the launcher): 32-bit division
    if (p.P == 1) {
        g = ct / p.Nt;
        n = (ct - g * p.Nt) * TC_BN + c;
    } else {
        g = ct * p.P + (c >> p.Nshift);
        n = c & (p.N - 1);
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const TcGemmParams p) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t s_xfull[TC_XSTAGES], s_xempty[TC_XSTAGES], s_wfull[TC_WSTAGES], s_wempty[TC_WSTAGES],
        s_rfull[TC_RSTAGES], s_rempty[TC_RSTAGES], s_acc_full[TC_ACC_BUFS], s_acc_empty[TC_ACC_BUFS],
        s_tfull[TC_TSLOTS], s_tempty[TC_TSLOTS];
    __shared__ int s_tile[TC_TSLOTS];
    __shared__ uint32_t s_tmem_base;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int W_MMA = TC_CONV_WARPS + TC_EPI_WARPS, W_LOAD = W_MMA + 1, W_WLOAD = W_MMA + 2;
    constexpr int N_CONSUMER_WARPS = TC_CONV_WARPS + TC_EPI_WARPS + 2;

    if (threadIdx.x == 0) {
        // one arrival per WARP (elected lane after __syncwarp): 256 per-thread arrivals on one shared-memory barrier per
        // K chunk serialise and were the bottleneck of the whole pipeline (~2 900 cycles per chunk, ncu: tensor pipe 15 %)
        const int conv_arrivals = p.use_raw ? TC_CONV_WARPS / 2 : TC_CONV_WARPS;    // dense: one group of four warps per stage
        for (int s = 0; s < TC_XSTAGES; ++s) { mbar_init(&s_xfull[s], conv_arrivals); mbar_init(&s_xempty[s], 1); }
        for (int s = 0; s < TC_WSTAGES; ++s) { mbar_init(&s_wfull[s], 1); mbar_init(&s_wempty[s], 1); }
        for (int s = 0; s < TC_RSTAGES; ++s) { mbar_init(&s_rfull[s], TC_CONV_WARPS * 16); mbar_init(&s_rempty[s], TC_CONV_WARPS / 2); }
        for (int b = 0; b < TC_ACC_BUFS; ++b) { mbar_init(&s_acc_full[b], 1); mbar_init(&s_acc_empty[b], TC_EPI_WARPS); }
        for (int s = 0; s < TC_TSLOTS; ++s) { mbar_init(&s_tfull[s], 1); mbar_init(&s_tempty[s], N_CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "r"((uint32_t)(TC_ACC_BUFS * TC_BN))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    const long long total_tiles = p.col_tiles * p.Mt;
    uint32_t tctr = 0;   // position in the tile ring (every role walks it independently)

    // tile ring consumer side: read slot i (blocking), release slot i (one arrive per warp)
    auto ring_read = [&](uint32_t i) -> int {
        const int slot = i % TC_TSLOTS;
        mbar_wait(&s_tfull[slot], (i / TC_TSLOTS) & 1);
        return *reinterpret_cast<volatile int *>(&s_tile[slot]);
    };
    auto ring_release = [&](uint32_t i) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_tempty[i % TC_TSLOTS]);
    };

    if (warp == W_LOAD) {
        // =============================== loader: tile scheduler + all bulk copies ===============================
        auto fetch = [&]() -> int {
            int t = -1;
            if (lane == 0) {
                t = atomicAdd(p.counter, 1);
                if ((long long)t >= total_tiles) t = -1;
            }
            return __shfl_sync(0xffffffffu, t, 0);
        };
        auto publish = [&](int t) {
            if (lane == 0) {
                const int slot = tctr % TC_TSLOTS;
                mbar_wait(&s_tempty[slot], ((tctr / TC_TSLOTS) & 1) ^ 1);
                *reinterpret_cast<volatile int *>(&s_tile[slot]) = t;
                mbar_arrive(&s_tfull[slot]);
            }
            ++tctr;
        };
        int cur = fetch();
        publish(cur);
        while (cur >= 0) {
            const int nxt = fetch();     // published one tile ahead: the converters prefetch across tile boundaries
            publish(nxt);
            cur = nxt;
        }
    } else if (warp == W_WLOAD) {
        // =============================== weight loader: one 16 KB cp.async.bulk per K chunk ===============================
        uint32_t wctr = 0;
        int cur = ring_read(tctr);
        while (cur >= 0) {
            if (lane == 0) {
                const int mt = cur % p.Mt;
                for (int kc = 0; kc < p.Kc; ++kc, ++wctr) {
                    const int s = wctr % TC_WSTAGES;
                    mbar_wait(&s_wempty[s], ((wctr / TC_WSTAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(&s_wfull[s], TC_CHUNK);
                    bulk_g2s(tc_smem + TC_OFF_W + (size_t)s * TC_CHUNK, p.wpack + ((size_t)mt * p.Kc + kc) * (size_t)TC_IMG,
                             TC_CHUNK, &s_wfull[s]);
                }
            }
            ring_release(tctr);
            ++tctr;
            cur = ring_read(tctr);
        }
        ring_release(tctr);
    } else if (warp < TC_CONV_WARPS) {
        // =============================== converters ===============================
        const int t = threadIdx.x;
        const int kk = t & 7;                // k row inside a group of 8
        const int ng = (t >> 3) & 15;        // group of 8 columns inside the tile
        const int kh = t >> 7;               // this thread converts k blocks kh and kh + 2 of the chunk
        uint32_t xctr = 0;
        long long *cdbg = (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) ? p.dbg + 512 : nullptr;
        int cdi = 0;

        auto store_images = [&](const float (&v)[2][8], uint64_t *release = nullptr) {
            const int s = xctr % TC_XSTAGES;
            mbar_wait(&s_xempty[s], ((xctr / TC_XSTAGES) & 1) ^ 1);
            TC_STAMP(cdbg, cdi, 12);
            uint8_t *xhi = tc_smem + TC_OFF_X + (size_t)s * TC_CHUNK, *xlo = xhi + TC_IMG;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint4 h, l;
                split2(v[q][0], v[q][1], h.x, l.x);
                split2(v[q][2], v[q][3], h.y, l.y);
                split2(v[q][4], v[q][5], h.z, l.z);
                split2(v[q][6], v[q][7], h.w, l.w);
                const uint32_t off = (uint32_t)ng * TC_SBO + (uint32_t)(kh + 2 * q) * TC_LBO + (uint32_t)kk * 16;
                *reinterpret_cast<uint4 *>(xhi + off) = h;
                *reinterpret_cast<uint4 *>(xlo + off) = l;
            }
            fence_proxy_async();        // this thread's image rows -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&s_xfull[s]);
                if (release) mbar_arrive(release);     // the raw stage the values came from is free again
            }
            ++xctr;
        };

        if (p.use_raw) {
            // Dense X: the eight converter warps work as TWO independent groups of four; group g stages, converts and
            // publishes the K chunks c = g, g+2, g+4, ... of the flattened (tile, chunk) sequence.  A chunk costs one
            // warp ~1 250 cycles of mostly latency (five barrier round trips, the copy issue, the image stores); with
            // the groups on alternate chunks those latencies overlap.  Raw stage c % 4 and image stage c % 2 belong to
            // group c % 2 alone, so the groups never wait for each other.
            const int grp = t >> 7, tl = t & 127;
            const int rk = tl & 7, rg = tl >> 3;          // k row inside a block of 8, group of 8 columns inside the tile
            struct Cur { uint32_t ring; int tile, kc, g0, n0; };
            auto norm = [&](Cur &c, bool release) {       // move to the tile that holds chunk index c.kc
                bool moved = false;
                while (c.tile >= 0 && c.kc >= p.Kc) {
                    c.kc -= p.Kc;
                    if (release) ring_release(c.ring);
                    c.tile = ring_read(++c.ring);
                    moved = true;
                }
                if (moved && c.tile >= 0) tc_col(p, c.tile / p.Mt, 0, c.g0, c.n0);
            };
            auto open_cur = [&](Cur &c, bool release) {
                c.ring = tctr; c.tile = ring_read(c.ring); c.kc = grp; c.g0 = c.n0 = 0;
                if (c.tile >= 0) tc_col(p, c.tile / p.Mt, 0, c.g0, c.n0);
                norm(c, release);
            };
            // mode 2: window origin (iy0, ix0) of this thread's eight pixels of the tile being STAGED (pixel rows
            // (tl >> 3) + 16 j of the raw stage; 0x7fffffff: column beyond N)
            int cv_tile = -2, cv_origin[8];
            // this thread's eight 16-byte pieces of raw chunk number ci -> stage ci % RSTAGES
            auto issue_raw = [&](const Cur &c, uint32_t ci) {
                const int s = ci % TC_RSTAGES;
                mbar_wait(&s_rempty[s], ((ci / TC_RSTAGES) & 1) ^ 1);       // the group's four warps have read the stage
                TC_STAMP(cdbg, cdi, 15);
                uint8_t *stage = tc_smem + TC_OFF_RAW + (size_t)s * TC_RAW_STAGE;
                if (p.mode == 2) {
                    // raw stage = [128 pixels][32 channels]: a pixel row is 128 contiguous bytes of the channels-last input
                    // (eight lanes per row, four rows per warp access); its 16-byte pieces are XOR-swizzled by the row so
                    // that the converter's per-pixel reads are conflict-free
                    if (cv_tile != c.tile) {
                        cv_tile = c.tile;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int n = c.n0 + (tl >> 3) + 16 * j;
                            if (p.cv_taps == 1) {
                                cv_origin[j] = n < p.N ? n : 0x7fffffff;
                            } else {
                                const int oy = n / p.cv_OW, ox = n - oy * p.cv_OW;
                                cv_origin[j] = n < p.N ? (((oy * p.cv_stride - 1) * 65536) | ((ox * p.cv_stride - 1) & 0xffff)) : 0x7fffffff;
                            }
                        }
                    }
                    const int c8 = tl & 7;
                    const int k = c.kc * TC_BK + c8 * 4;
                    const int tap = k >> p.cv_cshift, ch = k & ((1 << p.cv_cshift) - 1);
                    const int dy = tap / 3, dx = tap - dy * 3;
                    const float *xg = p.x + (size_t)c.g0 * p.x_group_stride + ch;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int r = (tl >> 3) + 16 * j;
                        const int iy = (cv_origin[j] >> 16) + dy, ix = (int)(short)(cv_origin[j] & 0xffff) + dx;
                        const bool rows_mode = p.cv_taps == 1;
                        const bool ok = tap < p.cv_taps && cv_origin[j] != 0x7fffffff &&
                                        (rows_mode || (iy >= 0 && iy < p.cv_H && ix >= 0 && ix < p.cv_W));
                        const size_t pix = rows_mode ? (size_t)cv_origin[j] : (size_t)iy * p.cv_W + ix;
                        const float *src = ok ? xg + (pix << p.cv_cshift) : p.x;
                        const uint32_t dst = smem_u32(stage + (size_t)r * 128 + (size_t)((c8 ^ (r & 7)) * 16));
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
                    }
                    TC_STAMP(cdbg, cdi, 16);
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&s_rfull[s])) : "memory");
                    return;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int piece = tl + j * 128;
                    const int r = piece >> 5, c4 = (piece & 31) * 4;           // row of the chunk, first of 4 columns
                    const int g = p.P == 1 ? c.g0 : c.g0 + (c4 >> p.Nshift);
                    const int n = p.P == 1 ? c.n0 + c4 : (c4 & (p.N - 1));
                    const int k = c.kc * TC_BK + r;
                    const bool ok = k < p.K && g < p.G && n < p.N;             // N % 4 == 0: a piece is all in or all out
                    const float *src = ok ? p.x + (size_t)g * p.x_group_stride + (size_t)k * p.x_row_stride + n : p.x;
                    const uint32_t dst = smem_u32(stage + (size_t)r * TC_RAW_ROW + (size_t)c4 * 4);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
                }
                TC_STAMP(cdbg, cdi, 16);
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&s_rfull[s])) : "memory");
            };
            Cur ic, cc;
            open_cur(ic, false);
            open_cur(cc, true);
            uint32_t ci = grp, c = grp;                    // chunk numbers of the next chunk to stage / to convert
            for (int i = 0; i < 2 && ic.tile >= 0; ++i) {  // two chunks of the group in flight
                issue_raw(ic, ci);
                ci += 2; ic.kc += 2;
                norm(ic, false);
            }
            while (cc.tile >= 0) {
                const int s = c % TC_RSTAGES;
                TC_STAMP(cdbg, cdi, 10);
                mbar_wait(&s_rfull[s], (c / TC_RSTAGES) & 1);
                TC_STAMP(cdbg, cdi, 11);
                float v[4][8];
                if (p.mode == 2) {
                    // pixel tl of the tile: its 32 channels -> four 8-channel runs, each one row of a K-major core matrix
                    const uint8_t *row = tc_smem + TC_OFF_RAW + (size_t)s * TC_RAW_STAGE + (size_t)tl * 128;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 a4 = *reinterpret_cast<const float4 *>(row + (((2 * q) ^ (tl & 7)) * 16));
                        const float4 b4 = *reinterpret_cast<const float4 *>(row + (((2 * q + 1) ^ (tl & 7)) * 16));
                        v[q][0] = a4.x; v[q][1] = a4.y; v[q][2] = a4.z; v[q][3] = a4.w;
                        v[q][4] = b4.x; v[q][5] = b4.y; v[q][6] = b4.z; v[q][7] = b4.w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 *src = reinterpret_cast<const float4 *>(
                            tc_smem + TC_OFF_RAW + (size_t)s * TC_RAW_STAGE + (size_t)(q * 8 + rk) * TC_RAW_ROW + rg * 32);
                        const float4 a4 = src[0], b4 = src[1];      // rows >= K and columns >= N were zero-filled by the copy
                        v[q][0] = a4.x; v[q][1] = a4.y; v[q][2] = a4.z; v[q][3] = a4.w;
                        v[q][4] = b4.x; v[q][5] = b4.y; v[q][6] = b4.z; v[q][7] = b4.w;
                    }
                }
                {   // image stage c % 2 (this group's own): wait until the MMAs of chunk c - 2 have read it
                    const int sx = c % TC_XSTAGES;
                    mbar_wait(&s_xempty[sx], ((c / TC_XSTAGES) & 1) ^ 1);
                    TC_STAMP(cdbg, cdi, 12);
                    uint8_t *xhi = tc_smem + TC_OFF_X + (size_t)sx * TC_CHUNK, *xlo = xhi + TC_IMG;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 h, l;
                        split2(v[q][0], v[q][1], h.x, l.x);
                        split2(v[q][2], v[q][3], h.y, l.y);
                        split2(v[q][4], v[q][5], h.z, l.z);
                        split2(v[q][6], v[q][7], h.w, l.w);
                        // MN-major image: core matrix = 8 k rows x 8 columns; K-major image (mode 2): 8 columns x 8 k
                        const uint32_t off = p.mode == 2
                            ? (uint32_t)(tl >> 3) * TC_SBO + (uint32_t)q * TC_LBO + (uint32_t)(tl & 7) * 16
                            : (uint32_t)rg * TC_SBO + (uint32_t)q * TC_LBO + (uint32_t)rk * 16;
                        *reinterpret_cast<uint4 *>(xhi + off) = h;
                        *reinterpret_cast<uint4 *>(xlo + off) = l;
                    }
                    fence_proxy_async();        // this thread's image rows -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&s_xfull[sx]);
                        mbar_arrive(&s_rempty[s]);             // the raw stage the values came from is free again
                    }
                }
                TC_STAMP(cdbg, cdi, 13);
                c += 2; cc.kc += 2;
                norm(cc, true);
                if (ic.tile >= 0) {
                    issue_raw(ic, ci);
                    ci += 2; ic.kc += 2;
                    norm(ic, false);
                }
                TC_STAMP(cdbg, cdi, 14);
            }
            ring_release(cc.ring);
        } else {
            // register path: gather mode, or dense rows that are not 16-byte aligned.  Software-pipelined: the global
            // loads of the next chunk (possibly of the next tile) are issued before the current one is converted.
            int tg = 0, tn0 = 0;          // group / first column (inside the group) of this thread's 8 columns, for the tile being LOADED
            int pidx[8];
            int loaded_tile = -1;
            auto load_chunk = [&](int tile, int kc, float (&v)[2][8]) {
                if (loaded_tile != tile) {
                    loaded_tile = tile;
                    tc_col(p, tile / p.Mt, ng * 8, tg, tn0);
                    if (p.mode == 1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int n = tn0 + j;
                            pidx[j] = 0;
                            if (tg < p.G && n < p.N) pidx[j] = p.idx ? __ldg(p.idx + (size_t)tg * p.N + n) : (n % p.n_pts);
                        }
                    }
                }
                const bool gok = tg < p.G;
                const float *xg = p.x + (size_t)(gok ? tg : 0) * p.x_group_stride;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int k = kc * TC_BK + (kh + 2 * q) * 8 + kk;
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[q][j] = 0.f;
                    if (k < p.K && gok) {
                        if (p.mode == 0) {
                            const float *row = xg + (size_t)k * p.x_row_stride + tn0;
                            if (tn0 + 8 <= p.N && ((reinterpret_cast<uintptr_t>(row) & 15u) == 0)) {
                                const float4 a4 = __ldg(reinterpret_cast<const float4 *>(row));
                                const float4 b4 = __ldg(reinterpret_cast<const float4 *>(row) + 1);
                                v[q][0] = a4.x; v[q][1] = a4.y; v[q][2] = a4.z; v[q][3] = a4.w;
                                v[q][4] = b4.x; v[q][5] = b4.y; v[q][6] = b4.z; v[q][7] = b4.w;
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    if (tn0 + j < p.N) v[q][j] = __ldg(row + j);
                            }
                        } else if (k < 3) {
                            const float *pts = p.xyz + (size_t)tg * p.n_pts * 3;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int n = tn0 + j;
                                if (n < p.N) {
                                    float c = 0.f;
                                    if (p.centres) c = __ldg(p.centres + ((size_t)tg * (p.N / p.nsample) + n / p.nsample) * 3 + k);
                                    v[q][j] = __fsub_rn(__ldg(pts + (size_t)pidx[j] * 3 + k), c);
                                }
                            }
                        } else {
                            const float *row = xg + (size_t)(k - 3) * p.x_row_stride;
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (tn0 + j < p.N) v[q][j] = __ldg(row + pidx[j]);
                        }
                    }
                }
            };

            float va[2][8], vb[2][8];
            int cur = ring_read(tctr);
            if (cur >= 0) load_chunk(cur, 0, va);
            while (cur >= 0) {
                const int nxt = ring_read(tctr + 1);
                for (int kc = 0; kc < p.Kc; kc += 2) {
                    // item (cur, kc) is in va; prefetch (cur, kc+1) or the next tile's first chunk into vb
                    if (kc + 1 < p.Kc) load_chunk(cur, kc + 1, vb);
                    else if (nxt >= 0) load_chunk(nxt, 0, vb);
                    store_images(va);
                    if (kc + 1 < p.Kc) {
                        if (kc + 2 < p.Kc) load_chunk(cur, kc + 2, va);
                        else if (nxt >= 0) load_chunk(nxt, 0, va);
                        store_images(vb);
                    } else {
                        // odd chunk count: the prefetched first chunk of the next tile sits in vb; move it to va
#pragma unroll
                        for (int q = 0; q < 2; ++q)
#pragma unroll
                            for (int j = 0; j < 8; ++j) va[q][j] = vb[q][j];
                    }
                }
                ring_release(tctr);
                ++tctr;
                cur = nxt;
            }
            ring_release(tctr);
        }
    } else if (warp < W_MMA) {
        // =============================== epilogue: one output channel per thread, 64 columns per warp ===============================
        const int e = warp - TC_CONV_WARPS;
        const int quad = e & 3, half = e >> 2;
        const int r = quad * 32 + lane;                 // accumulator row = TMEM lane
        uint8_t *stage_row = tc_smem + TC_OFF_OUT + (size_t)r * TC_RAW_ROW;
        uint32_t tile_ctr = 0;
        int cur = ring_read(tctr);
        while (cur >= 0) {
            const int mt = cur % p.Mt;
            const long long ct = cur / p.Mt;
            const int buf = tile_ctr % TC_ACC_BUFS;
            const int m = mt * TC_BM + r;
            const float bias = (p.bias && m < p.M) ? __ldg(p.bias + m) : 0.f;
            const float dot_w = (p.out_mode == 3 && m < p.M) ? __ldg(p.dot_w + m) : 0.f;
            // pool windows of 128 columns are reduced by the half-0 warps alone
            const bool whole = (p.out_mode == 1 && p.pool > 64);
            const int c_begin = whole ? 0 : half * 64;
            const int c_end = whole ? (half == 0 ? TC_BN : 0) : c_begin + 64;
            if (p.out_mode == 0 && p.bulk_out) __syncwarp();     // this warp's staging rows of the previous tile have been read
            mbar_wait(&s_acc_full[buf], (tile_ctr / TC_ACC_BUFS) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * TC_BN;
            float run = -INFINITY;
#pragma unroll 1
            for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                float v[32];
                tmem_ld32(taddr + c0, v);
                if (c0 + 32 >= c_end) {     // last read of this accumulator by this warp: hand it back to the MMA warp early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_acc_empty[buf]);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float o = v[j] + bias;
                    v[j] = p.relu ? fmaxf(o, 0.f) : o;
                }
                if (p.out_mode == 3) {
                    // the 32 rows of this warp times the following layer's weights, summed over the rows with a fixed
                    // butterfly (lane l ends up with column c0 + l): deterministic, independent of how columns are tiled
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= dot_w;
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        const bool up = (lane & off) != 0;
#pragma unroll
                        for (int i = 0; i < off; ++i) {
                            const float send = up ? v[i] : v[i + off];
                            const float keep = up ? v[i + off] : v[i];
                            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                        }
                    }
                    int g, n;
                    tc_col(p, ct, c0 + lane, g, n);
                    if (g < p.G && n < p.N)
                        p.y[(size_t)g * p.y_group_stride + (size_t)(mt * 4 + quad) * p.N + n] = v[0];
                } else if (p.out_mode == 0) {
                    if (p.bulk_out) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4 *>(stage_row + (size_t)(c0 + j) * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else if (m < p.M) {
                        if (p.P == 1 || p.N >= 32) {       // the 32 columns of the block belong to one group
                            int g, n;
                            tc_col(p, ct, c0, g, n);
                            float *dst = p.y + (size_t)g * p.y_group_stride + (size_t)m * p.N + n;
                            if (g < p.G) {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (n + j < p.N) dst[j] = v[j];
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                int g, n;
                                tc_col(p, ct, c0 + j, g, n);
                                if (g < p.G && n < p.N) p.y[(size_t)g * p.y_group_stride + (size_t)m * p.N + n] = v[j];
                            }
                        }
                    }
                } else if (p.out_mode == 2) {
                    if (m < p.M) {
                        if (p.P == 1 || p.N >= 32) {
                            int g, n;
                            tc_col(p, ct, c0, g, n);
                            float *dst = p.y + (size_t)g * p.y_group_stride + (size_t)n * p.M + m;
                            if (g < p.G) {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (n + j < p.N) dst[(size_t)j * p.M] = v[j];
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                int g, n;
                                tc_col(p, ct, c0 + j, g, n);
                                if (g < p.G && n < p.N) p.y[(size_t)g * p.y_group_stride + (size_t)n * p.M + m] = v[j];
                            }
                        }
                    }
                } else {
                    // max-pool over windows of `pool` columns (pool divides 128 and N: a window never straddles a tile or
                    // a group, and is either fully valid or fully outside)
                    const int sub = p.pool < 32 ? p.pool : 32;      // window length inside one 32-column block
                    const int wins = p.N / p.pool;
                    if (sub == 32) {
                        float mx = v[0];
#pragma unroll
                        for (int j = 1; j < 32; ++j) mx = fmaxf(mx, v[j]);
                        run = fmaxf(run, mx);
                        if ((c0 + 32) % p.pool == 0) {
                            int g, n;
                            tc_col(p, ct, c0 + 32 - p.pool, g, n);
                            if (m < p.M && g < p.G && n < p.N)
                                p.y[(size_t)g * p.y_group_stride + (size_t)m * wins + n / p.pool] = run;
                            run = -INFINITY;
                        }
                    } else {
                        for (int w0 = 0; w0 < 32; w0 += sub) {
                            float mx = -INFINITY;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j >= w0 && j < w0 + sub) mx = fmaxf(mx, v[j]);
                            int g, n;
                            tc_col(p, ct, c0 + w0, g, n);
                            if (m < p.M && g < p.G && n < p.N)
                                p.y[(size_t)g * p.y_group_stride + (size_t)m * wins + n / p.pool] = mx;
                        }
                    }
                }
            }
            if (whole && half != 0) {       // nothing read: still hand the accumulator back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_acc_empty[buf]);
            }
            if (p.out_mode == 0 && p.bulk_out) {
                // the warp's 32 staged rows x 64 columns go out as 256-byte row segments: lanes 0-15 one row, 16-31 the next
                __syncwarp();
                const int sub = lane >> 4, c4 = c_begin + (lane & 15) * 4;
                int g, n;
                tc_col(p, ct, c4, g, n);
                if (g < p.G && n < p.N) {
                    float *ybase = p.y + (size_t)g * p.y_group_stride + n;
#pragma unroll 4
                    for (int rr = 0; rr < 32; rr += 2) {
                        const int row = quad * 32 + rr + sub, mm = mt * TC_BM + row;
                        if (mm < p.M)
                            *reinterpret_cast<float4 *>(ybase + (size_t)mm * p.N) = *reinterpret_cast<const float4 *>(
                                tc_smem + TC_OFF_OUT + (size_t)row * TC_RAW_ROW + (size_t)c4 * 4);
                    }
                }
            }
            ++tile_ctr;
            ring_release(tctr);
            ++tctr;
            cur = ring_read(tctr);
        }
        ring_release(tctr);
    } else if (warp == W_MMA) {
        // =============================== MMA issuer ===============================
        constexpr uint64_t D_IMG = TC_IMG >> 4, D_K16 = (2 * TC_LBO) >> 4;
        const uint32_t idesc = p.mode == 2 ? TC_IDESC_KK : TC_IDESC;      // convolution: X images are K-major like W's
        uint32_t xctr = 0, wctr = 0, tile_ctr = 0;
        long long *mdbg = (p.dbg && blockIdx.x == 0 && lane == 0) ? p.dbg : nullptr;
        int mdi = 0;
        // The whole warp walks the loop and waits on the barriers (uniform control flow); the six MMAs and the commits of
        // a K chunk are issued by one elected lane (elect_one(): under a `lane == 0` test every tcgen05 instruction is
        // wrapped in an ELECT / BRA.U.ANY serialisation loop that costs ~60 issue cycles per MMA).
        int cur = ring_read(tctr);
        while (cur >= 0) {
            const int buf = tile_ctr % TC_ACC_BUFS;
            TC_STAMP(mdbg, mdi, 0);
            mbar_wait(&s_acc_empty[buf], ((tile_ctr / TC_ACC_BUFS) & 1) ^ 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)buf * TC_BN;
            for (int kc = 0; kc < p.Kc; ++kc, ++xctr, ++wctr) {
                const int sx = xctr % TC_XSTAGES, sw = wctr % TC_WSTAGES;
                TC_STAMP(mdbg, mdi, 1);
                mbar_wait(&s_wfull[sw], (wctr / TC_WSTAGES) & 1);
                TC_STAMP(mdbg, mdi, 2);
                mbar_wait(&s_xfull[sx], (xctr / TC_XSTAGES) & 1);
                TC_STAMP(mdbg, mdi, 3);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t xd = make_smem_desc(smem_u32(tc_smem + TC_OFF_X + (size_t)sx * TC_CHUNK));
                    const uint64_t wd = make_smem_desc(smem_u32(tc_smem + TC_OFF_W + (size_t)sw * TC_CHUNK));
                    umma_ss_i(acc, wd, xd, idesc, kc != 0);
                    umma_ss_i(acc, wd + D_IMG, xd, idesc, 1);
                    umma_ss_i(acc, wd, xd + D_IMG, idesc, 1);
                    umma_ss_i(acc, wd + D_K16, xd + D_K16, idesc, 1);
                    umma_ss_i(acc, wd + D_K16 + D_IMG, xd + D_K16, idesc, 1);
                    umma_ss_i(acc, wd + D_K16, xd + D_K16 + D_IMG, idesc, 1);
                    umma_commit(&s_xempty[sx]);
                    umma_commit(&s_wempty[sw]);
                    if (kc == p.Kc - 1) umma_commit(&s_acc_full[buf]);
                }
                __syncwarp();
                TC_STAMP(mdbg, mdi, 4);
            }
            ++tile_ctr;
            ring_release(tctr);
            ++tctr;
            cur = ring_read(tctr);
        }
        ring_release(tctr);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(TC_ACC_BUFS * TC_BN)) : "memory");
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

// Scheduler counters: TC_SCHED_SLOTS (counter, done) pairs in a __device__ array — every device the library runs on gets
// its own zero-initialised copy with the module, so there is no allocation and no synchronisation on first use.  A launch
// takes the next slot of a process-wide atomic sequence: two launches share a pair only if they are TC_SCHED_SLOTS
// launches apart, i.e. never while both can be in flight (a captured graph bakes its slots in; graphs captured one
// after the other hold disjoint slot ranges).  The last CTA of a launch re-arms its pair.
constexpr int TC_SCHED_SLOTS = 4096;
__device__ int g_tc_sched[2 * TC_SCHED_SLOTS];

int tc_sched_slot(int dev, int **counter, int **done);

static int *tc_sched_buffer(int dev) {
    static int *bufs[JMB_MAX_DEVICES] = {nullptr};     // symbol address per device (same value written by every thread)
    int *b = __atomic_load_n(&bufs[dev], __ATOMIC_ACQUIRE);
    if (!b) {
        void *sym = nullptr;
        if (cudaGetSymbolAddress(&sym, g_tc_sched) != cudaSuccess) return nullptr;
        b = static_cast<int *>(sym);
        __atomic_store_n(&bufs[dev], b, __ATOMIC_RELEASE);
    }
    return b;
}

// (counter, done) pair for one launch of a dynamically scheduled kernel on device `dev` (shared by tc_gemm_kernel and
// sa_fused_kernel)
int tc_sched_slot(int dev, int **counter, int **done) {
    int *sched = tc_sched_buffer(dev);
    if (!sched) {
        set_error("cannot resolve the tile-scheduler counters");
        return JMB_ERR_CUDA;
    }
    static unsigned launch_seq = 0;
    const unsigned slot = __atomic_fetch_add(&launch_seq, 1u, __ATOMIC_RELAXED) % TC_SCHED_SLOTS;
    *counter = sched + 2 * slot;
    *done = sched + 2 * slot + 1;
    return JMB_OK;
}

}  // namespace jmb

namespace jmb {
int tc_gemm_launch(TcGemmParams &p, long long tiles, void *stream);
}

extern "C" int jmb_tc_mlp_layer(const void *wpack, const float *bias, int M, int K, int G, int N, int mode,
                                        const float *x, long long x_group_stride, int x_row_stride, const int *idx,
                                        const float *xyz, const float *centres, int nsample, int n_pts, int out_mode,
                                        int pool, int relu, float *y, long long y_group_stride, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(M > 0 && K > 0 && G >= 0 && N >= 0, "tc_mlp_layer: bad sizes");
    if (G == 0 || N == 0) return JMB_OK;
    JMB_REQUIRE(wpack && x && y, "tc_mlp_layer: null pointer");
    JMB_REQUIRE(mode == 0 || mode == 1, "tc_mlp_layer: bad mode");
    JMB_REQUIRE(mode == 0 || (xyz && n_pts > 0 && K >= 3 && (centres == nullptr || (nsample > 0 && N % nsample == 0))),
                "tc_mlp_layer: gather mode needs xyz / nsample");
    JMB_REQUIRE(out_mode >= 0 && out_mode <= 2, "tc_mlp_layer: bad out_mode");
    JMB_REQUIRE(out_mode != 1 || (pool > 0 && pool <= TC_BN && TC_BN % pool == 0 && N % pool == 0),
                "tc_mlp_layer: pool must divide 128 and N");
    TcGemmParams p{};
    p.wpack = (const __nv_bfloat16 *)wpack; p.bias = bias;
    p.M = M; p.K = K; p.Mt = div_up(M, TC_BM); p.Kc = div_up(K, TC_BK);
    p.G = G; p.N = N; p.mode = mode; p.x = x; p.x_group_stride = x_group_stride; p.x_row_stride = x_row_stride;
    p.idx = idx; p.xyz = xyz; p.centres = centres; p.nsample = nsample; p.n_pts = n_pts;
    p.out_mode = out_mode; p.pool = pool; p.relu = relu; p.y = y;
    p.y_group_stride = y_group_stride > 0 ? y_group_stride : (long long)M * (out_mode == 1 ? N / pool : N);
    // narrow groups (N = 8, 16, 32, 64) are packed 128/N to a tile
    p.P = 1; p.Nshift = 0;
    if (N >= 8 && N < TC_BN && (N & (N - 1)) == 0 && G > 1) {
        p.P = TC_BN / N;
        while ((1 << p.Nshift) < N) ++p.Nshift;
    }
    p.Nt = p.P == 1 ? div_up(N, TC_BN) : 1;
    p.col_tiles = p.P == 1 ? (long long)G * p.Nt : (long long)div_up(G, p.P);
    const long long tiles = p.col_tiles * p.Mt;
    JMB_REQUIRE(tiles < (1LL << 30), "tc_mlp_layer: too many tiles");
    auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    p.use_raw = (mode == 0 && al16(x) && N % 4 == 0 && x_row_stride % 4 == 0 && x_group_stride % 4 == 0) ? 1 : 0;
    p.bulk_out = (out_mode == 0 && al16(y) && N % 4 == 0 && p.y_group_stride % 4 == 0) ? 1 : 0;

    return tc_gemm_launch(p, tiles, stream);
}

namespace jmb {
int tc_gemm_launch(TcGemmParams &p, long long tiles, void *stream) {
    const int M = p.M, K = p.K, N = p.N;
    int dev = 0, sms = 0;
    {
        const int rc = device_info(&dev, &sms);
        if (rc != JMB_OK) return rc;
    }
    {
        const int rc = tc_sched_slot(dev, &p.counter, &p.done);
        if (rc != JMB_OK) return rc;
    }

    static long long *dbg_buf = nullptr;          // JMB_TC_DEBUG=1: CTA 0 records a clock64() timeline (profiling aid)
    static int dbg_on = -1;
    if (dbg_on < 0) {
        const char *e = getenv("JMB_TC_DEBUG");
        dbg_on = (e && e[0] == '1') ? 1 : 0;
        if (dbg_on) { cudaMalloc(&dbg_buf, 1024 * sizeof(long long)); cudaMemset(dbg_buf, 0, 1024 * sizeof(long long)); }
    }
    p.dbg = dbg_buf;
    const size_t smem = (size_t)TC_SMEM;
    JMB_FUNC_ATTR_ONCE(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem, dev);
    const int grid = (int)(tiles < (long long)sms ? tiles : (long long)sms);
    tc_gemm_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(p);
    if (dbg_on) {
        long long hbuf[1024];
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaMemcpy(hbuf, dbg_buf, sizeof(hbuf), cudaMemcpyDeviceToHost);
        for (int part = 0; part < 2; ++part) {
            const long long *b = hbuf + part * 512;
            fprintf(stderr, "[tc_gemm timeline %s M=%d K=%d N=%d] ", part ? "converter warp 0" : "mma issuer", M, K, N);
            for (int i = 0; i + 1 < 500 && (b[i] || b[i + 1]); i += 2) fprintf(stderr, "%lld:%lld ", b[i], b[i + 1] - hbuf[1]);
            fprintf(stderr, "\n");
        }
        cudaMemset(dbg_buf, 0, 1024 * sizeof(long long));
    }
    return check_launch("tc_mlp_layer");
}
}  // namespace jmb

// Two layers in one launch when the second has ONE output channel (the cls / link / start-end heads end in a C -> 1 layer:
// reference rpn.py:40-47, rcnn.py:91-111): the epilogue of layer 1 multiplies its activated rows by the second layer's
// weights and reduces over the rows in fp32.  partial (G, 4 * ceil(M / 128), N): the caller adds the rows and the bias.
extern "C" int jmb_tc_mlp_layer_dot(const void *wpack, const float *bias, int M, int K, int G, int N, const float *x,
                                    long long x_group_stride, int x_row_stride, int relu, const float *dot_w,
                                    float *partial, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(M > 0 && K > 0 && G >= 0 && N >= 0, "tc_mlp_layer_dot: bad sizes");
    if (G == 0 || N == 0) return JMB_OK;
    JMB_REQUIRE(wpack && x && dot_w && partial, "tc_mlp_layer_dot: null pointer");
    TcGemmParams p{};
    p.wpack = (const __nv_bfloat16 *)wpack; p.bias = bias;
    p.M = M; p.K = K; p.Mt = div_up(M, TC_BM); p.Kc = div_up(K, TC_BK);
    p.G = G; p.N = N; p.mode = 0; p.x = x; p.x_group_stride = x_group_stride; p.x_row_stride = x_row_stride;
    p.out_mode = 3; p.pool = 0; p.relu = relu; p.y = partial; p.dot_w = dot_w;
    p.y_group_stride = (long long)4 * p.Mt * N;
    p.P = 1; p.Nshift = 0;
    if (N >= 8 && N < TC_BN && (N & (N - 1)) == 0 && G > 1) {
        p.P = TC_BN / N;
        while ((1 << p.Nshift) < N) ++p.Nshift;
    }
    p.Nt = p.P == 1 ? div_up(N, TC_BN) : 1;
    p.col_tiles = p.P == 1 ? (long long)G * p.Nt : (long long)div_up(G, p.P);
    const long long tiles = p.col_tiles * p.Mt;
    JMB_REQUIRE(tiles < (1LL << 30), "tc_mlp_layer_dot: too many tiles");
    auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    p.use_raw = (al16(x) && N % 4 == 0 && x_row_stride % 4 == 0 && x_group_stride % 4 == 0) ? 1 : 0;
    p.bulk_out = 0;
    return tc_gemm_launch(p, tiles, stream);
}

// Pointwise layer over POINT-MAJOR rows: x (G, N, C) -> y (G, N, M), the layout the fused set-abstraction kernel gathers
// from and the per-proposal input stage writes.  The convolution mode with one tap: a column's K values are one contiguous
// 4 C-byte run, staged with 16-byte copies and converted to K-major operand images.  C a power of two >= 32.
extern "C" int jmb_tc_mlp_rows(const void *wpack, const float *bias, int M, int C, int G, int N, const float *x, int relu,
                               float *y, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(M > 0 && C >= TC_BK && (C & (C - 1)) == 0 && G >= 0 && N >= 0, "tc_mlp_rows: bad sizes (C must be a power of two >= 32)");
    if (G == 0 || N == 0) return JMB_OK;
    JMB_REQUIRE(wpack && x && y, "tc_mlp_rows: null pointer");
    JMB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15u) == 0, "tc_mlp_rows: input must be 16-byte aligned");
    JMB_REQUIRE(N < 0x7fffffff, "tc_mlp_rows: too many rows per group");
    TcGemmParams p{};
    p.wpack = (const __nv_bfloat16 *)wpack; p.bias = bias;
    p.M = M; p.K = C; p.Mt = div_up(M, TC_BM); p.Kc = div_up(C, TC_BK);
    p.G = G; p.N = N; p.mode = 2; p.x = x; p.x_group_stride = (long long)N * C; p.x_row_stride = 0;
    p.cv_cshift = 0;
    while ((1 << p.cv_cshift) < C) ++p.cv_cshift;
    p.cv_H = 1; p.cv_W = N; p.cv_OW = N; p.cv_stride = 1; p.cv_taps = 1;
    p.out_mode = 2; p.pool = 0; p.relu = relu; p.y = y; p.y_group_stride = (long long)N * M;
    p.P = 1; p.Nshift = 0; p.Nt = div_up(N, TC_BN);
    p.col_tiles = (long long)G * p.Nt;
    const long long tiles = p.col_tiles * p.Mt;
    JMB_REQUIRE(tiles < (1LL << 30), "tc_mlp_rows: too many tiles");
    p.use_raw = 1; p.bulk_out = 0;
    return tc_gemm_launch(p, tiles, stream);
}

namespace jmb {
// out[g][n] = act(bias + partial[g][0][n] + partial[g][1][n] + ...): the rows are added one after the other, so the result
// does not depend on the shape (a library reduction picks its order by layout)
__global__ void __launch_bounds__(256)
tc_dot_finish_kernel(int rows, int N, long long total, float bias, int relu, const float *__restrict__ partial,
                     float *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long g = i / N;
    const int n = (int)(i - g * N);
    const float *src = partial + (size_t)g * rows * N + n;
    float acc = __ldg(src);
    for (int r = 1; r < rows; ++r) acc = __fadd_rn(acc, __ldg(src + (size_t)r * N));
    acc = __fadd_rn(acc, bias);
    out[i] = relu ? fmaxf(acc, 0.f) : acc;
}
}  // namespace jmb

extern "C" int jmb_tc_dot_finish(int G, int rows, int N, const float *partial, float bias, int relu, float *out,
                                 void *stream) {
    using namespace jmb;
    JMB_REQUIRE(G >= 0 && rows > 0 && N >= 0, "tc_dot_finish: bad sizes");
    if (G == 0 || N == 0) return JMB_OK;
    JMB_REQUIRE(partial && out, "tc_dot_finish: null pointer");
    const long long total = (long long)G * N;
    JMB_REQUIRE(total < (1LL << 40), "tc_dot_finish: too many columns");
    tc_dot_finish_kernel<<<(unsigned)div_up_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(rows, N, total, bias, relu, partial, out);
    return check_launch("tc_dot_finish");
}

// 3x3 convolution, padding 1, stride 1 or 2, bias + optional ReLU, channels-last in and out, on the same kernel: an implicit
// GEMM whose X operand rows are the 128-byte channel runs of the input pixels under the nine taps (zero-filled outside the
// image), converted to bf16 hi / lo K-major images by the same converter warps; replaces the `conv3x3` layers of
// BasicBlock (reference jmodt/detection/modeling/backbone.py:9-30; cuDNN there).  x (B, H, W, C) with C a power of two >= 4,
// wpack = tc.pack_weights of the (Cout, 9 * C) matrix [cout][3 dy + dx][c]; y (B, OH, OW, Cout), OH = (H - 1) / stride + 1.
extern "C" int jmb_tc_conv3x3(const void *wpack, const float *bias, int Cout, int C, int B, int H, int W, int stride,
                              const float *x, int relu, float *y, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(Cout > 0 && C >= 4 && (C & (C - 1)) == 0 && B >= 0 && H > 0 && W > 0, "tc_conv3x3: bad sizes");
    JMB_REQUIRE(C == 4 || C % TC_BK == 0, "tc_conv3x3: channels per pixel must be 4 or a multiple of 32");
    JMB_REQUIRE(stride == 1 || stride == 2, "tc_conv3x3: stride must be 1 or 2");
    JMB_REQUIRE(H < 32768 && W < 32768, "tc_conv3x3: image too large");
    if (B == 0) return JMB_OK;
    JMB_REQUIRE(wpack && x && y, "tc_conv3x3: null pointer");
    JMB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15u) == 0, "tc_conv3x3: input must be 16-byte aligned");
    const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1;
    JMB_REQUIRE((long long)OH * OW < (1LL << 30), "tc_conv3x3: too many pixels");
    TcGemmParams p{};
    p.wpack = (const __nv_bfloat16 *)wpack; p.bias = bias;
    p.M = Cout; p.K = 9 * C; p.Mt = div_up(p.M, TC_BM); p.Kc = div_up(p.K, TC_BK);
    p.G = B; p.N = OH * OW; p.mode = 2; p.x = x; p.x_group_stride = (long long)H * W * C; p.x_row_stride = 0;
    p.cv_cshift = 0;
    while ((1 << p.cv_cshift) < C) ++p.cv_cshift;
    p.cv_H = H; p.cv_W = W; p.cv_OW = OW; p.cv_stride = stride; p.cv_taps = 9;
    p.out_mode = 2; p.pool = 0; p.relu = relu; p.y = y; p.y_group_stride = (long long)p.N * Cout;
    p.P = 1; p.Nshift = 0; p.Nt = div_up(p.N, TC_BN);
    p.col_tiles = (long long)B * p.Nt;
    const long long tiles = p.col_tiles * p.Mt;
    JMB_REQUIRE(tiles < (1LL << 30), "tc_conv3x3: too many tiles");
    p.use_raw = 1; p.bulk_out = 0;
    return tc_gemm_launch(p, tiles, stream);
}
