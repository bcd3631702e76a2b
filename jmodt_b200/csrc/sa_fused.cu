// Fused set-abstraction layer on tcgen05 tensor cores, sm_100a:
//     ball-query indices -> grouped [xyz - centre, features] -> 3-layer shared MLP (+bias, ReLU) -> max over nsample
// in ONE kernel.  Replaces, for one PointnetSAModule (reference jmodt/ops/pointnet2/pointnet2_modules.py:20-63):
// QueryAndGroup's two group_points launches + subtract + cat (pointnet2_utils.py:241-264), three cuDNN 1x1 convs
// over (B, C, npoint, nsample) (pytorch_utils.py:6-33) and F.max_pool2d.  At the RCNN SA0 shape the reference
// round-trips a 549 MB/frame grouped tensor and two 537 MB/frame activations through HBM; here a tile of 128
// (centre, sample) columns never leaves the SM:
//
//   workers (8 warps)   gather the layer-1 operand X1 (K1 x 128) straight into its UMMA shared-memory image
//                       (bf16 hi/lo split in registers), later turn each layer's TMEM accumulator into the next
//                       layer's operand image (bias, ReLU, split, 16-byte stores), and max-pool the last layer;
//   issuer (1 thread)   streams the weight chunk images with cp.async.bulk through a 4-stage ring and issues
//                       tcgen05.mma (three bf16 MMAs per fp32 product: W_hi.X_hi + W_lo.X_hi + W_hi.X_lo);
//   the next tile's gather overlaps the current tile's layer-2 MMAs.
//
// Shapes: layer widths C1 = C2 = 128, C3 in {128, 256}; 3 + C_in = K1 <= 160; nsample in {8,...,64} dividing 64.
#include "tc_common.cuh"

namespace jmb {

constexpr int SF_WORKERS = 256;
constexpr int SF_THREADS = SF_WORKERS + 64;  // + warp 8: MMA issuer, warp 9: weight loader
constexpr int SF_WSTAGES = 4;
constexpr int SF_MAXKC1 = 5;
constexpr int SF_CHUNK = 2 * TC_IMG;  // hi + lo image of one 32-row chunk: 16 KB
constexpr int SF_SMEM = (SF_WSTAGES + SF_MAXKC1 + 4) * SF_CHUNK;  // W ring + X1 + activations = 208 KB

struct SaFusedParams {
    const __nv_bfloat16 *w1, *w2, *w3;
    const float *b1, *b2, *b3;
    int K1, Kc1, Mt3;
    int G, npoint, nsample, n_pts;
    const float *feats;    // (G, K1-3, n_pts)
    const int *idx;        // (G, npoint, nsample)
    const float *xyz;      // (G, n_pts, 3)
    const float *centres;  // (G, npoint, 3)
    float *out;            // (G, 128*Mt3, npoint)
};

__global__ void __launch_bounds__(SF_THREADS, 1)
sa_fused_kernel(const SaFusedParams p) {
    extern __shared__ __align__(1024) uint8_t sf_smem[];
    uint8_t *s_w = sf_smem;
    uint8_t *s_x1 = sf_smem + SF_WSTAGES * SF_CHUNK;
    uint8_t *s_act = s_x1 + SF_MAXKC1 * SF_CHUNK;
    __shared__ __align__(8) uint64_t s_x1_full[SF_MAXKC1], s_w_full[SF_WSTAGES], s_w_empty[SF_WSTAGES], s_acc_full,
        s_epi_done;
    __shared__ uint32_t s_tmem_base;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int c = 0; c < SF_MAXKC1; ++c) mbar_init(&s_x1_full[c], SF_WORKERS);
        for (int s = 0; s < SF_WSTAGES; ++s) { mbar_init(&s_w_full[s], 1); mbar_init(&s_w_empty[s], 1); }
        mbar_init(&s_acc_full, 1);
        mbar_init(&s_epi_done, SF_WORKERS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "r"((uint32_t)TC_BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    const int N = p.npoint * p.nsample;
    const int Nt = N / TC_BN;
    const long long total_tiles = (long long)p.G * Nt;
    const int kmax16 = ((p.K1 + 15) / 16) * 16;  // rows the layer-1 MMAs actually read

    if (warp < 8) {
        // ====================================== workers ======================================
        const int t = threadIdx.x;
        const int kk = t & 7, ng = (t >> 3) & 15, kbsel = t >> 7;
        const int quad = warp & 3, half = warp >> 2;
        const int m = quad * 32 + lane;  // accumulator row (output channel) of this thread in the epilogues
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)half * 64;
        uint32_t acc_phase = 0;

        auto produce_x1 = [&](long long tile, int c_begin, int c_end) {
            const int nt = (int)(tile % Nt);
            const int g = (int)(tile / Nt);
            const int n0 = nt * TC_BN + ng * 8;  // 8 columns of one centre (nsample % 8 == 0)
            int pidx[8];
            {
                const int4 a = __ldg(reinterpret_cast<const int4 *>(p.idx + (size_t)g * N + n0));
                const int4 b = __ldg(reinterpret_cast<const int4 *>(p.idx + (size_t)g * N + n0) + 1);
                pidx[0] = a.x; pidx[1] = a.y; pidx[2] = a.z; pidx[3] = a.w;
                pidx[4] = b.x; pidx[5] = b.y; pidx[6] = b.z; pidx[7] = b.w;
            }
            const float *cen = p.centres + ((size_t)g * p.npoint + n0 / p.nsample) * 3;
            const float *pts = p.xyz + (size_t)g * p.n_pts * 3;
            const float *fg = p.feats + (size_t)g * (p.K1 - 3) * p.n_pts;
            for (int c = c_begin; c < c_end; ++c) {
                uint8_t *xhi = s_x1 + (size_t)c * SF_CHUNK, *xlo = xhi + TC_IMG;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int kbl = 2 * i + kbsel;
                    const int kbase = c * TC_BK + kbl * 8;
                    if (kbase >= kmax16) continue;
                    const int k = kbase + kk;
                    float v[8];
                    if (k >= p.K1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = 0.f;
                    } else if (k < 3) {
                        const float cv = __ldg(cen + k);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = __fsub_rn(__ldg(pts + (size_t)pidx[j] * 3 + k), cv);
                    } else {
                        const float *row = fg + (size_t)(k - 3) * p.n_pts;
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = __ldg(row + pidx[j]);
                    }
                    uint4 h, l;
                    split2(v[0], v[1], h.x, l.x);
                    split2(v[2], v[3], h.y, l.y);
                    split2(v[4], v[5], h.z, l.z);
                    split2(v[6], v[7], h.w, l.w);
                    const uint32_t off = (uint32_t)ng * TC_SBO + (uint32_t)kbl * TC_LBO + (uint32_t)kk * 16;
                    *reinterpret_cast<uint4 *>(xhi + off) = h;
                    *reinterpret_cast<uint4 *>(xlo + off) = l;
                }
                fence_proxy_async();
                mbar_arrive(&s_x1_full[c]);
            }
        };

        // accumulator -> next layer's operand image (row m of the accumulator is row k = m of the operand)
        auto epilogue_act = [&](const float *bias_ptr) {
            const float bias = __ldg(bias_ptr + m);
            uint8_t *ahi = s_act + (size_t)quad * SF_CHUNK, *alo = ahi + TC_IMG;
            const uint32_t rowoff = (uint32_t)(lane >> 3) * TC_LBO + (uint32_t)(lane & 7) * 16;
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
                float v[32];
                tmem_ld32(taddr + c0, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 h, l;
                    float *w = v + q * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[j] = fmaxf(w[j] + bias, 0.f);
                    split2(w[0], w[1], h.x, l.x);
                    split2(w[2], w[3], h.y, l.y);
                    split2(w[4], w[5], h.z, l.z);
                    split2(w[6], w[7], h.w, l.w);
                    const uint32_t off = (uint32_t)(half * 8 + (c0 >> 3) + q) * TC_SBO + rowoff;
                    *reinterpret_cast<uint4 *>(ahi + off) = h;
                    *reinterpret_cast<uint4 *>(alo + off) = l;
                }
            }
        };

        auto epilogue_pool = [&](long long tile, int mt) {
            const int nt = (int)(tile % Nt);
            const int g = (int)(tile / Nt);
            const float bias = __ldg(p.b3 + mt * TC_BM + m);
            float *orow = p.out + ((size_t)g * (TC_BM * p.Mt3) + mt * TC_BM + m) * p.npoint +
                          (nt * TC_BN + half * 64) / p.nsample;
            const int sub = p.nsample < 32 ? p.nsample : 32;
            float run = -INFINITY;
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
                float v[32];
                tmem_ld32(taddr + c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + bias, 0.f);
                if (sub == 32) {
                    float mx = v[0];
#pragma unroll
                    for (int j = 1; j < 32; ++j) mx = fmaxf(mx, v[j]);
                    run = fmaxf(run, mx);
                    if ((c0 + 32) % p.nsample == 0) {
                        orow[(c0 + 32) / p.nsample - 1] = run;
                        run = -INFINITY;
                    }
                } else {
                    for (int w0 = 0; w0 < 32; w0 += sub) {
                        float mx = -INFINITY;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j >= w0 && j < w0 + sub) mx = fmaxf(mx, v[j]);
                        orow[(c0 + w0) / p.nsample] = mx;
                    }
                }
            }
        };

        const long long first = blockIdx.x;
        const int csplit = (p.Kc1 + 1) / 2;   // next tile's gather is split over the layer-2 and layer-3 MMA windows
        if (first < total_tiles) produce_x1(first, 0, p.Kc1);
        for (long long tile = first; tile < total_tiles; tile += gridDim.x) {
            mbar_wait(&s_acc_full, acc_phase); acc_phase ^= 1;
            tc_fence_after();
            epilogue_act(p.b1);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(&s_epi_done);
            const bool more = tile + gridDim.x < total_tiles;
            if (more) produce_x1(tile + gridDim.x, 0, csplit);          // overlaps the layer-2 MMAs

            mbar_wait(&s_acc_full, acc_phase); acc_phase ^= 1;
            tc_fence_after();
            epilogue_act(p.b2);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(&s_epi_done);
            if (more) produce_x1(tile + gridDim.x, csplit, p.Kc1);      // overlaps the layer-3 MMAs

            for (int mt = 0; mt < p.Mt3; ++mt) {
                mbar_wait(&s_acc_full, acc_phase); acc_phase ^= 1;
                tc_fence_after();
                epilogue_pool(tile, mt);
                tc_fence_before();
                mbar_arrive(&s_epi_done);
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            // ====================================== weight loader ======================================
            // The weight stream of a tile is fixed (Kc1 + 4 + 4*Mt3 chunk images); run ahead of the issuer through
            // the ring so a layer's first chunk is already in shared memory when its MMAs may start.
            const int CH = p.Kc1 + 4 + 4 * p.Mt3;
            long long my_tiles = 0;
            for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) ++my_tiles;
            const long long total_chunks = my_tiles * CH;
            for (long long wreq = 0; wreq < total_chunks; ++wreq) {
                const int s = (int)(wreq % SF_WSTAGES);
                int j = (int)(wreq % CH);
                const __nv_bfloat16 *src;
                if (j < p.Kc1) src = p.w1 + (size_t)j * (SF_CHUNK / 2);
                else if (j < p.Kc1 + 4) src = p.w2 + (size_t)(j - p.Kc1) * (SF_CHUNK / 2);
                else src = p.w3 + (size_t)(j - p.Kc1 - 4) * (SF_CHUNK / 2);
                mbar_wait(&s_w_empty[s], (uint32_t)(((wreq / SF_WSTAGES) & 1) ^ 1));
                mbar_arrive_expect_tx(&s_w_full[s], SF_CHUNK);
                bulk_g2s(s_w + (size_t)s * SF_CHUNK, src, SF_CHUNK, &s_w_full[s]);
            }
        }
        __syncwarp();
    } else {
        if (lane == 0) {
            // ====================================== MMA issuer ======================================
            long long wuse = 0;
            auto mma_chunk = [&](uint32_t b_base, int k16_steps, bool first_of_layer) {
                const int s = (int)(wuse % SF_WSTAGES);
                mbar_wait(&s_w_full[s], (uint32_t)((wuse / SF_WSTAGES) & 1));
                tc_fence_after();
                const uint32_t a_base = smem_u32(s_w + (size_t)s * SF_CHUNK);
                for (int k16 = 0; k16 < k16_steps; ++k16) {
                    const uint32_t koff = (uint32_t)k16 * 2 * TC_LBO;
                    const uint64_t whi = make_smem_desc(a_base + koff), wlo = make_smem_desc(a_base + TC_IMG + koff);
                    const uint64_t xhi = make_smem_desc(b_base + koff), xlo = make_smem_desc(b_base + TC_IMG + koff);
                    umma_ss(tmem_base, whi, xhi, !(first_of_layer && k16 == 0));
                    umma_ss(tmem_base, wlo, xhi, 1);
                    umma_ss(tmem_base, whi, xlo, 1);
                }
                umma_commit(&s_w_empty[s]);
                ++wuse;
            };
            uint32_t epi_phase = 0, tile_ctr = 0;
            bool first_layer_ever = true;
            auto wait_epilogue = [&]() {
                if (first_layer_ever) { first_layer_ever = false; return; }
                mbar_wait(&s_epi_done, epi_phase); epi_phase ^= 1;
                tc_fence_after();
            };
            for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_ctr) {
                // layer 1: operand = gathered X1
                wait_epilogue();
                for (int c = 0; c < p.Kc1; ++c) {
                    mbar_wait(&s_x1_full[c], tile_ctr & 1);
                    const int rows = kmax16 - c * TC_BK;
                    mma_chunk(smem_u32(s_x1 + (size_t)c * SF_CHUNK), rows >= 32 ? 2 : 1, c == 0);
                }
                umma_commit(&s_acc_full);
                // layer 2 and the Mt3 row blocks of layer 3: operand = activation image
                for (int l = 0; l < 1 + p.Mt3; ++l) {
                    wait_epilogue();
                    for (int c = 0; c < 4; ++c) mma_chunk(smem_u32(s_act + (size_t)c * SF_CHUNK), 2, c == 0);
                    umma_commit(&s_acc_full);
                }
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_BN) : "memory");
    }
}

}  // namespace jmb

extern "C" int jmb_sa_fused(const void *w1, const float *b1, const void *w2, const float *b2, const void *w3,
                            const float *b3, int K1, int C3, int G, int npoint, int nsample, int n_pts,
                            const float *feats, const int *idx, const float *xyz, const float *centres, float *out,
                            void *stream) {
    using namespace jmb;
    JMB_REQUIRE(G >= 0 && npoint > 0 && nsample > 0 && n_pts > 0, "sa_fused: bad sizes");
    if (G == 0) return JMB_OK;
    JMB_REQUIRE(w1 && w2 && w3 && b1 && b2 && b3 && feats && idx && xyz && centres && out, "sa_fused: null pointer");
    JMB_REQUIRE(K1 > 3 && K1 <= SF_MAXKC1 * TC_BK, "sa_fused: 3 + C_in = %d must be in (3, 160]", K1);
    JMB_REQUIRE(C3 == 128 || C3 == 256, "sa_fused: last layer width must be 128 or 256");
    JMB_REQUIRE(nsample % 8 == 0 && 64 % nsample == 0, "sa_fused: nsample must be 8, 16, 32 or 64");
    JMB_REQUIRE(((long long)npoint * nsample) % TC_BN == 0, "sa_fused: npoint*nsample must be a multiple of 128");
    JMB_REQUIRE((reinterpret_cast<uintptr_t>(idx) & 15u) == 0, "sa_fused: idx must be 16-byte aligned");
    SaFusedParams p;
    p.w1 = (const __nv_bfloat16 *)w1; p.w2 = (const __nv_bfloat16 *)w2; p.w3 = (const __nv_bfloat16 *)w3;
    p.b1 = b1; p.b2 = b2; p.b3 = b3;
    p.K1 = K1; p.Kc1 = div_up(K1, TC_BK); p.Mt3 = C3 / TC_BM;
    p.G = G; p.npoint = npoint; p.nsample = nsample; p.n_pts = n_pts;
    p.feats = feats; p.idx = idx; p.xyz = xyz; p.centres = centres; p.out = out;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        JMB_CUDA(cudaGetDevice(&dev));
        JMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        JMB_CUDA(cudaFuncSetAttribute(sa_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SF_SMEM));
    }
    const long long tiles = (long long)G * ((long long)npoint * nsample / TC_BN);
    const int grid = (int)(tiles < sms ? tiles : sms);
    sa_fused_kernel<<<grid, SF_THREADS, SF_SMEM, (cudaStream_t)stream>>>(p);
    return check_launch("sa_fused");
}
