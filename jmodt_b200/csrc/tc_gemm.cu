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
// the split residuals are <= 2^-16 relative per product).  The reference's own cuDNN path uses single
// TF32 (10-bit mantissa) by default; the 3-term split is ~64x more accurate than that.
//
// Prologues:  dense X from global memory, or the fused grouping of QueryAndGroup / GroupAll
// (pointnet2_utils.py:241-290): X[k][n] = xyz[idx[n]][k] - centre[n / nsample][k] for k < 3 and
// feats[k-3][idx[n]] otherwise — the 549 MB/frame grouped tensor of RCNN SA0 is never materialised.
//
// Structure: persistent CTAs of 5 warps.  Warps 0-3 stage operands (W chunk images by one 16 KB
// cp.async.bulk, X chunks by vector loads + in-register bf16 split) into an mbarrier ring and run the
// epilogue; warp 4 lane 0 issues tcgen05.mma and signals through tcgen05.commit.  64 KB smem + 128 TMEM
// columns per CTA -> 3 CTAs per SM, so one CTA's epilogue overlaps another's MMAs.
#include "tc_common.cuh"

namespace jmb {

// warp roles: 0-3 operand producers (X), 4-7 epilogue (TMEM lane quadrant = warp & 3), 8 MMA issuer, 9 weight loader
constexpr int TC_XSTAGES = 3;                      // X ring: 16 KB per stage (hi + lo image)
constexpr int TC_WSTAGES = 4;                      // W ring: 16 KB per stage, filled by cp.async.bulk, runs ahead of X
constexpr int TC_CHUNK = 2 * TC_IMG;               // hi + lo image of one 32-row chunk
constexpr int TC_SMEM = (TC_XSTAGES + TC_WSTAGES) * TC_CHUNK;   // 112 KB -> 2 CTAs per SM
constexpr int TC_THREADS = 320;
constexpr int TC_ACC_BUFS = 2;                     // double-buffered accumulator: epilogue overlaps the next tile's MMAs

struct TcGemmParams {
    const __nv_bfloat16 *wpack;   // [Mt][Kc][2][TC_IMG/2] chunk images (see pack_weights in tc.py)
    const float *bias;            // [Mt*128] (zero padded), may be null
    int M, K, Mt, Kc;
    int G, N;                     // groups, columns per group
    int mode;                     // 0 dense, 1 grouped gather
    const float *x;               // dense: (G, K, N)   gather: feats (G, K-3, n_pts)
    long long x_group_stride;     // elements
    int x_row_stride;             // elements between consecutive k rows
    const int *idx;               // gather: (G, N) point index per column (null: column n -> point n % n_pts)
    const float *xyz;             // gather: (G, n_pts, 3)
    const float *centres;         // gather: (G, N / nsample, 3) or null (GroupAll: no centring)
    int nsample, n_pts;
    int out_mode;                 // 0 dense (G, M, N), 1 max over `pool` consecutive columns -> (G, M, N / pool),
                                  // 2 point-major (G, N, M): a warp's 32 channels of one column are one 128-byte store
    int pool, relu;
    float *y;
    long long y_group_stride;     // elements between consecutive groups of y (>= M*N, lets a layer write into a slice of a wider tensor)
};

__global__ void __launch_bounds__(TC_THREADS, 2)
tc_gemm_kernel(const TcGemmParams p) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t s_xfull[TC_XSTAGES], s_xempty[TC_XSTAGES], s_wfull[TC_WSTAGES], s_wempty[TC_WSTAGES],
        s_acc_full[TC_ACC_BUFS], s_acc_empty[TC_ACC_BUFS];
    __shared__ uint32_t s_tmem_base;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_XSTAGES; ++s) { mbar_init(&s_xfull[s], 128); mbar_init(&s_xempty[s], 1); }
        for (int s = 0; s < TC_WSTAGES; ++s) { mbar_init(&s_wfull[s], 1); mbar_init(&s_wempty[s], 1); }
        for (int b = 0; b < TC_ACC_BUFS; ++b) { mbar_init(&s_acc_full[b], 1); mbar_init(&s_acc_empty[b], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "r"((uint32_t)(TC_ACC_BUFS * TC_BN))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    const int Nt = (p.N + TC_BN - 1) / TC_BN;
    const long long total_tiles = (long long)p.G * Nt * p.Mt;
    uint32_t chunk_ctr = 0;  // global k-chunk counter: stage = ctr % STAGES, phase = (ctr / STAGES) & 1
    uint32_t tile_ctr = 0;

    if (warp < 4) {
        // =============================== operand producers ===============================
        const int t = threadIdx.x;
        const int kk = t & 7;               // k row inside a group of 8
        const int ng = warp * 4 + ((t >> 3) & 3);   // n group of 8 columns inside the tile
        // Operand staging is software-pipelined: the global loads of work item i+1 (the next K chunk, possibly of the
        // next tile) are issued into registers before item i is converted and stored, so one load round trip is hidden
        // behind the split / store / MMA of the previous chunk and behind the epilogue.
        struct TileCoord { int mt, nt, g, n0; };
        auto coord = [&](long long tile) {
            TileCoord c;
            c.mt = (int)(tile % p.Mt);
            const long long gn = tile / p.Mt;
            c.nt = (int)(gn % Nt);
            c.g = (int)(gn / Nt);
            c.n0 = c.nt * TC_BN + ng * 8;
            return c;
        };
        int pidx[8];              // gather mode: point index of this thread's 8 columns, for the tile being LOADED
        long long pidx_tile = -1;
        auto load_chunk = [&](long long tile, int kc, float (&v)[TC_BK / 8][8]) {
            const TileCoord tc_ = coord(tile);
            const float *xg = p.x + (size_t)tc_.g * p.x_group_stride;
            if (p.mode == 1 && pidx_tile != tile) {
                pidx_tile = tile;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = tc_.n0 + j;
                    pidx[j] = 0;
                    if (n < p.N) pidx[j] = p.idx ? __ldg(p.idx + (size_t)tc_.g * p.N + n) : (n % p.n_pts);
                }
            }
#pragma unroll
            for (int kb = 0; kb < TC_BK / 8; ++kb) {
                const int k = kc * TC_BK + kb * 8 + kk;
#pragma unroll
                for (int j = 0; j < 8; ++j) v[kb][j] = 0.f;
                if (k < p.K) {
                    if (p.mode == 0) {
                        const float *row = xg + (size_t)k * p.x_row_stride + tc_.n0;
                        if (tc_.n0 + 8 <= p.N && ((reinterpret_cast<uintptr_t>(row) & 15u) == 0)) {
                            const float4 a4 = __ldg(reinterpret_cast<const float4 *>(row));
                            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(row) + 1);
                            v[kb][0] = a4.x; v[kb][1] = a4.y; v[kb][2] = a4.z; v[kb][3] = a4.w;
                            v[kb][4] = b4.x; v[kb][5] = b4.y; v[kb][6] = b4.z; v[kb][7] = b4.w;
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (tc_.n0 + j < p.N) v[kb][j] = __ldg(row + j);
                        }
                    } else if (k < 3) {
                        const float *pts = p.xyz + (size_t)tc_.g * p.n_pts * 3;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int n = tc_.n0 + j;
                            if (n < p.N) {
                                float c = 0.f;
                                if (p.centres) c = __ldg(p.centres + ((size_t)tc_.g * (p.N / p.nsample) + n / p.nsample) * 3 + k);
                                v[kb][j] = __fsub_rn(__ldg(pts + (size_t)pidx[j] * 3 + k), c);
                            }
                        }
                    } else {
                        const float *row = xg + (size_t)(k - 3) * p.x_row_stride;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (tc_.n0 + j < p.N) v[kb][j] = __ldg(row + pidx[j]);
                    }
                }
            }
        };
        auto store_chunk = [&](const float (&v)[TC_BK / 8][8]) {
            const int s = chunk_ctr % TC_XSTAGES;
            const uint32_t ph = (chunk_ctr / TC_XSTAGES) & 1;
            mbar_wait(&s_xempty[s], ph ^ 1);
            uint8_t *xhi = tc_smem + (size_t)s * TC_CHUNK, *xlo = xhi + TC_IMG;
#pragma unroll
            for (int kb = 0; kb < TC_BK / 8; ++kb) {
                uint4 h, l;
                split2(v[kb][0], v[kb][1], h.x, l.x);
                split2(v[kb][2], v[kb][3], h.y, l.y);
                split2(v[kb][4], v[kb][5], h.z, l.z);
                split2(v[kb][6], v[kb][7], h.w, l.w);
                const uint32_t off = (uint32_t)ng * TC_SBO + (uint32_t)kb * TC_LBO + (uint32_t)kk * 16;
                *reinterpret_cast<uint4 *>(xhi + off) = h;
                *reinterpret_cast<uint4 *>(xlo + off) = l;
            }
            fence_proxy_async();
            mbar_arrive(&s_xfull[s]);
            ++chunk_ctr;
        };

        float va[TC_BK / 8][8], vb[TC_BK / 8][8];
        if ((long long)blockIdx.x < total_tiles) load_chunk(blockIdx.x, 0, va);
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_ctr) {
            const TileCoord tcur = coord(tile);
            const int mt = tcur.mt, nt = tcur.nt, g = tcur.g;
            const long long next_tile = tile + gridDim.x;
            for (int kc = 0; kc < p.Kc; kc += 2) {
                // item (tile, kc) is in va; prefetch (tile, kc+1) or the next tile's first chunk into vb
                if (kc + 1 < p.Kc) load_chunk(tile, kc + 1, vb);
                else if (next_tile < total_tiles) load_chunk(next_tile, 0, vb);
                store_chunk(va);
                if (kc + 1 < p.Kc) {
                    if (kc + 2 < p.Kc) load_chunk(tile, kc + 2, va);
                    else if (next_tile < total_tiles) load_chunk(next_tile, 0, va);
                    store_chunk(vb);
                } else {
                    // odd chunk count: the prefetched first chunk of the next tile sits in vb; move it to va
#pragma unroll
                    for (int kb = 0; kb < TC_BK / 8; ++kb)
#pragma unroll
                        for (int j = 0; j < 8; ++j) va[kb][j] = vb[kb][j];
                }
            }

        }
    } else if (warp < 8) {
        // =============================== epilogue: one output channel per thread ===============================
        const int quad = warp & 3;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_ctr) {
            const int mt = (int)(tile % p.Mt);
            const long long gn = tile / p.Mt;
            const int nt = (int)(gn % Nt);
            const int g = (int)(gn / Nt);
            // ---- epilogue: one output channel per thread ----
            const int buf = tile_ctr % TC_ACC_BUFS;
            mbar_wait(&s_acc_full[buf], (tile_ctr / TC_ACC_BUFS) & 1);
            tc_fence_after();
            const int m = mt * TC_BM + quad * 32 + lane;
            const float bias = (p.bias && m < p.M) ? __ldg(p.bias + m) : 0.f;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * TC_BN;
            if (p.out_mode == 0) {
                float *yrow = p.y + (size_t)g * p.y_group_stride + (size_t)m * p.N + (size_t)nt * TC_BN;
#pragma unroll 1
                for (int c0 = 0; c0 < TC_BN; c0 += 32) {
                    float v[32];
                    tmem_ld32(taddr + c0, v);
                    if (m < p.M) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float o = v[j] + bias;
                            if (p.relu) o = fmaxf(o, 0.f);
                            v[j] = o;
                        }
                        const int ncol = nt * TC_BN + c0;
                        if (ncol + 32 <= p.N && ((reinterpret_cast<uintptr_t>(yrow + c0) & 15u) == 0)) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4 *>(yrow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (ncol + j < p.N) yrow[c0 + j] = v[j];
                        }
                    }
                }
            } else if (p.out_mode == 2) {
                float *ycol = p.y + (size_t)g * p.y_group_stride + (size_t)nt * TC_BN * p.M + m;
#pragma unroll 1
                for (int c0 = 0; c0 < TC_BN; c0 += 32) {
                    float v[32];
                    tmem_ld32(taddr + c0, v);
                    if (m < p.M) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float o = v[j] + bias;
                            if (p.relu) o = fmaxf(o, 0.f);
                            if (nt * TC_BN + c0 + j < p.N) ycol[(size_t)(c0 + j) * p.M] = o;
                        }
                    }
                }
            } else {
                // max-pool over windows of `pool` columns (pool divides 128; windows never straddle a tile).
                // N % pool == 0, so a window is either fully inside [0, N) or fully outside.
                const int groups_per_row = p.N / p.pool;
                float *yrow = p.y + (size_t)g * p.y_group_stride + (size_t)m * groups_per_row + (size_t)(nt * TC_BN) / p.pool;
                const int sub = p.pool < 32 ? p.pool : 32;      // window length inside one 32-column chunk
                float run = -INFINITY;
#pragma unroll 1
                for (int c0 = 0; c0 < TC_BN; c0 += 32) {
                    float v[32];
                    tmem_ld32(taddr + c0, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float o = v[j] + bias;
                        v[j] = p.relu ? fmaxf(o, 0.f) : o;
                    }
                    if (sub == 32) {
                        float mx = v[0];
#pragma unroll
                        for (int j = 1; j < 32; ++j) mx = fmaxf(mx, v[j]);
                        run = fmaxf(run, mx);
                        if ((c0 + 32) % p.pool == 0) {
                            const int w = (c0 + 32) / p.pool - 1;
                            if (m < p.M && nt * TC_BN + c0 < p.N) yrow[w] = run;
                            run = -INFINITY;
                        }
                    } else {
                        for (int w0 = 0; w0 < 32; w0 += sub) {
                            float mx = -INFINITY;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j >= w0 && j < w0 + sub) mx = fmaxf(mx, v[j]);
                            if (m < p.M && nt * TC_BN + c0 + w0 < p.N) yrow[(c0 + w0) / p.pool] = mx;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&s_acc_empty[buf]);
        }
    } else if (warp == 9) {
        if (lane == 0) {
            // =============================== weight loader: one 16 KB cp.async.bulk per K chunk ===============================
            uint32_t wctr = 0;
            for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int mt = (int)(tile % p.Mt);
                for (int kc = 0; kc < p.Kc; ++kc, ++wctr) {
                    const int s = wctr % TC_WSTAGES;
                    mbar_wait(&s_wempty[s], ((wctr / TC_WSTAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(&s_wfull[s], TC_CHUNK);
                    bulk_g2s(tc_smem + (size_t)(TC_XSTAGES + s) * TC_CHUNK, p.wpack + ((size_t)mt * p.Kc + kc) * (size_t)TC_IMG,
                             TC_CHUNK, &s_wfull[s]);
                }
            }
        }
        __syncwarp();
    } else {
      if (lane == 0) {
        // =============================== MMA issuer ===============================
        constexpr uint64_t D_IMG = TC_IMG >> 4, D_K16 = (2 * TC_LBO) >> 4;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_ctr) {
            const int buf = tile_ctr % TC_ACC_BUFS;
            mbar_wait(&s_acc_empty[buf], ((tile_ctr / TC_ACC_BUFS) & 1) ^ 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)buf * TC_BN;
            for (int kc = 0; kc < p.Kc; ++kc, ++chunk_ctr) {
                const int sx = chunk_ctr % TC_XSTAGES, sw = chunk_ctr % TC_WSTAGES;
                mbar_wait(&s_wfull[sw], (chunk_ctr / TC_WSTAGES) & 1);
                mbar_wait(&s_xfull[sx], (chunk_ctr / TC_XSTAGES) & 1);
                tc_fence_after();
                const uint64_t xd = make_smem_desc(smem_u32(tc_smem + (size_t)sx * TC_CHUNK));
                const uint64_t wd = make_smem_desc(smem_u32(tc_smem + (size_t)(TC_XSTAGES + sw) * TC_CHUNK));
                umma_ss(acc, wd, xd, kc != 0);
                umma_ss(acc, wd + D_IMG, xd, 1);
                umma_ss(acc, wd, xd + D_IMG, 1);
                umma_ss(acc, wd + D_K16, xd + D_K16, 1);
                umma_ss(acc, wd + D_K16 + D_IMG, xd + D_K16, 1);
                umma_ss(acc, wd + D_K16, xd + D_K16 + D_IMG, 1);
                umma_commit(&s_xempty[sx]);
                umma_commit(&s_wempty[sw]);
            }
            umma_commit(&s_acc_full[buf]);
        }
      }
      __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(TC_ACC_BUFS * TC_BN)) : "memory");
    }
}

}  // namespace jmb

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
    TcGemmParams p;
    p.wpack = (const __nv_bfloat16 *)wpack; p.bias = bias;
    p.M = M; p.K = K; p.Mt = div_up(M, TC_BM); p.Kc = div_up(K, TC_BK);
    p.G = G; p.N = N; p.mode = mode; p.x = x; p.x_group_stride = x_group_stride; p.x_row_stride = x_row_stride;
    p.idx = idx; p.xyz = xyz; p.centres = centres; p.nsample = nsample; p.n_pts = n_pts;
    p.out_mode = out_mode; p.pool = pool; p.relu = relu; p.y = y;
    p.y_group_stride = y_group_stride > 0 ? y_group_stride : (long long)M * (out_mode == 1 ? N / pool : N);
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        JMB_CUDA(cudaGetDevice(&dev));
        JMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const size_t smem = (size_t)TC_SMEM;
    static bool attr_set = false;
    if (!attr_set) {
        JMB_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const long long tiles = (long long)G * div_up(N, TC_BN) * p.Mt;
    const int grid = (int)(tiles < (long long)sms * 2 ? tiles : (long long)sms * 2);
    tc_gemm_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(p);
    return check_launch("tc_mlp_layer");
}
