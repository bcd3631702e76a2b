// The LI-Fusion image decoder evaluated only where LiDAR points sample it, sm_100a.
//
// Replaces, for inference, the tail of PointNet2MSG.forward (reference jmodt/detection/modeling/backbone.py:187-196):
//     de_i   = ConvTranspose2d(C_i -> 16, kernel = stride = 2^(i+1))(img_i)          i = 0..3   (B, 16, 384, 1280) each
//     fused  = relu(BatchNorm2d(Conv2d 1x1 (64 -> 32)(cat(de_0..de_3))))                         (B, 32, 384, 1280)
//     out    = grid_sample(fused, xy, bilinear, zeros, align_corners=True)                       (B, 32, N)
// The reference materialises both full-resolution maps (126 + 63 MB per frame) to read 4 pixels per point from them: the
// 16 384 points of a frame touch at most 65 536 of the 491 520 pixels.  Because kernel == stride, a decoded pixel (y, x)
// depends on ONE source pixel per level, (y >> (i+1), x >> (i+1)), through the weight slice of its phase
// (y mod 2^(i+1), x mod 2^(i+1)); everything between the source maps and the ReLU is linear.  So:
//   1. plan     every (point, bilinear tap) sample is binned by its phase (y mod 16, x mod 16) — which fixes the weight
//               slice of all four levels — with a counting sort (histogram, scan, fill);
//   2. decode   a CTA takes 512 samples of ONE phase, keeps that phase's 16 x 960 decoder weights moving through shared
//               memory in 32-channel chunks next to the samples' source rows (channels-last maps: a row is one
//               contiguous 128-byte read), register-tiles the 4 x (16 x C_i) products in fp32 FFMA, then applies the
//               folded 1x1 convolution + BatchNorm + ReLU and writes 32 floats per sample;
//   3. combine  a thread per point blends its four samples with the bilinear weights in the reference's nw, ne, sw, se
//               order and writes the channel-first (B, 32, N) result the fusion layer consumes.
// Work drops from 17.1 GFLOP per frame (dense) to <= 2.3 GFLOP, and nothing of full resolution is ever written.
#include "tc_common.cuh"

namespace jmb {

constexpr int DG_LEVELS = 4;
constexpr int DG_PH = 16;                 // largest stride: 16 x 16 phases
constexpr int DG_BINS = DG_PH * DG_PH;
constexpr int DG_PTS = 128;               // points per CTA
constexpr int DG_TILE = 4 * DG_PTS;       // samples (point, tap) per CTA = staged source rows per level at most
constexpr int DG_THREADS = 512;
constexpr int DG_KC = 32;                 // channels per staged chunk
constexpr int DG_R = 16;                  // decoder outputs per level (cfg.LI_FUSION.DeConv_Reduce)
constexpr int DG_CAT = DG_LEVELS * DG_R;  // concatenated decoder channels
constexpr int DG_OUT = 32;                // fused channels (cfg.LI_FUSION.IMG_FEATURES_CHANNEL / 4)
constexpr int DG_ROW = DG_KC + 8;         // padded shared-memory row, floats (conflict-free 64-bit fragment reads)
constexpr int DG_MAX_CHUNKS = 32;
// bins workspace (ints): histogram, first sample of a bin, first tile of a bin, fill cursor
constexpr int DG_OFF_HIST = 0, DG_OFF_BIN = 256, DG_OFF_TILE = 256 + 257, DG_OFF_CUR = 256 + 2 * 257;
constexpr int DG_BINS_INTS = 256 + 2 * 257 + 256;

struct DgTaps {
    int x[4], y[4];          // nw, ne, sw, se
    float w[4];
    bool valid[4];
};

// align_corners=True un-normalisation and the four bilinear taps, as feature_gather.cu
__device__ __forceinline__ DgTaps dg_taps(float gx, float gy, int h, int w) {
    DgTaps t;
    const float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)(w - 1));
    const float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)(h - 1));
    const float fx = floorf(ix), fy = floorf(iy);
    // far-away coordinates (or NaN) have no tap inside the image; clamp before the int conversion
    const float cfx = fminf(fmaxf(fx, -2.f), (float)w + 1.f), cfy = fminf(fmaxf(fy, -2.f), (float)h + 1.f);
    const int x0 = (fx == fx) ? (int)cfx : -2, y0 = (fy == fy) ? (int)cfy : -2;
    const float wx1 = ix - fx, wy1 = iy - fy, wx0 = (fx + 1.f) - ix, wy0 = (fy + 1.f) - iy;
    t.x[0] = x0; t.y[0] = y0; t.w[0] = wx0 * wy0;
    t.x[1] = x0 + 1; t.y[1] = y0; t.w[1] = wx1 * wy0;
    t.x[2] = x0; t.y[2] = y0 + 1; t.w[2] = wx0 * wy1;
    t.x[3] = x0 + 1; t.y[3] = y0 + 1; t.w[3] = wx1 * wy1;
#pragma unroll
    for (int i = 0; i < 4; ++i) t.valid[i] = t.x[i] >= 0 && t.x[i] < w && t.y[i] >= 0 && t.y[i] < h;
    return t;
}

__device__ __forceinline__ int dg_phase(int x, int y) { return (y & (DG_PH - 1)) * DG_PH + (x & (DG_PH - 1)); }

// ---- 1. plan: counting sort of the points (those with a tap inside the image) by the phase of their nw tap --------------
template <bool FILL>
__global__ void __launch_bounds__(256)
dg_bin_kernel(int total, int h, int w, const float *__restrict__ xy, int *__restrict__ bins, int *__restrict__ items) {
    __shared__ int s_cnt[DG_BINS], s_base[DG_BINS];
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    int ph = 0, rank = 0;
    bool ok = false;
    if (p < total) {
        const DgTaps t = dg_taps(__ldg(xy + (size_t)p * 2), __ldg(xy + (size_t)p * 2 + 1), h, w);
        ok = t.valid[0] || t.valid[1] || t.valid[2] || t.valid[3];
        if (ok) {
            ph = dg_phase(t.x[0], t.y[0]);
            rank = atomicAdd(&s_cnt[ph], 1);
        }
    }
    __syncthreads();
    const int c = s_cnt[threadIdx.x];
    if (!FILL) {
        if (c) atomicAdd(bins + DG_OFF_HIST + threadIdx.x, c);
        return;
    }
    s_base[threadIdx.x] = c ? atomicAdd(bins + DG_OFF_CUR + threadIdx.x, c) : 0;
    __syncthreads();
    if (ok) items[s_base[ph] + rank] = p;
}

__global__ void __launch_bounds__(DG_BINS) dg_scan_kernel(int *__restrict__ bins) {
    __shared__ int s_a[DG_BINS], s_t[DG_BINS];
    const int t = threadIdx.x;
    const int c = bins[DG_OFF_HIST + t];
    s_a[t] = c;
    s_t[t] = (c + DG_PTS - 1) / DG_PTS;
    __syncthreads();
    for (int d = 1; d < DG_BINS; d <<= 1) {       // inclusive Hillis-Steele scans
        const int a = t >= d ? s_a[t - d] : 0, b = t >= d ? s_t[t - d] : 0;
        __syncthreads();
        s_a[t] += a; s_t[t] += b;
        __syncthreads();
    }
    bins[DG_OFF_BIN + t + 1] = s_a[t];
    bins[DG_OFF_TILE + t + 1] = s_t[t];
    bins[DG_OFF_CUR + t] = s_a[t] - c;
    if (t == 0) { bins[DG_OFF_BIN] = 0; bins[DG_OFF_TILE] = 0; }
}

// ---- 2. decode ----------------------------------------------------------------------------------------------------------
struct DgParams {
    const float *map[DG_LEVELS];     // channels-last (B, H >> (l+1), W >> (l+1), C_l)
    int C[DG_LEVELS];
    int chunk_level[DG_MAX_CHUNKS], chunk_ch[DG_MAX_CHUNKS];
    int n_chunks;
    int B, N, H, W;
    const float *xy;
    const uint32_t *wexp;            // (256 phases, n_chunks, [hi, lo], 16, 16) packed bf16 pairs (k even in the low half)
    const float *w1;                 // (32, 64): 1x1 convolution with the BatchNorm scale folded in
    const float *b1;                 // (32): its bias with BatchNorm shift and the decoder biases folded in
    const int *bins, *items;
    float *taps;                     // (B * N * 4, 32)
};

constexpr int DG_WP = DG_KC / 2 + 4;      // padded weight row in 32-bit words (two bf16 each): conflict-free fragment reads
constexpr int DG_SRC_FLOATS = DG_TILE * DG_ROW, DG_W_WORDS = 4 * 2 * DG_R * DG_WP;      // four tap phases, hi and lo planes
constexpr int DG_W1P = DG_OUT + 8;        // padded row of the packed 1x1 weights
constexpr int DG_W1_WORDS = 2 * (DG_CAT / 2) * DG_W1P;
constexpr size_t DG_SMEM = (size_t)(2 * (DG_SRC_FLOATS + DG_W_WORDS) + DG_W1_WORDS) * 4 +
                           (size_t)3 * DG_PTS * 4 + (size_t)DG_TILE * 4;

__device__ __forceinline__ void dg_cp16(void *dst, const void *src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}

// D (16 x 8) += A (16 x 16, row) * B (16 x 8, col) on the warp-level tensor-core path, bf16 operands (two per register,
// the lower k in the low half), fp32 accumulation.  lane = 4 g + q:
//   a0 (g, 2q..2q+1)  a1 (g + 8, 2q..)  a2 (g, 2q + 8..)  a3 (g + 8, 2q + 8..);   b0 (k = 2q..2q+1, n = g)  b1 (k = 2q + 8.., n = g);
//   c0 (g, 2q)  c1 (g, 2q + 1)  c2 (g + 8, 2q)  c3 (g + 8, 2q + 1)
__device__ __forceinline__ void dg_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                 "{%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// fp32-grade products from bf16 tensor-core instructions, as in tc_gemm.cu: x = xh + xl (two bf16, residual 2^-17),
// x w = xh wh + xl wh + xh wl (+ xl wl <= 2^-16 |x w|, dropped).
// A CTA takes 128 POINTS whose nw taps share one phase (py, px) — so the phases of their four taps, and which of the taps
// fall into the neighbouring source pixel at each level, are the same for the whole tile.  Per level it stages the source
// rows of the 1, 2 or 4 distinct source pixels a point's taps touch (at stride 16 / 8 the four taps nearly always share
// one source pixel: 1 247 floats per point on average instead of 4 x 960) and the weight slices of the four tap phases.
// Warp w owns the 16-point tile w & 7 and the taps 2 (w >> 3), 2 (w >> 3) + 1, all 16 decoder outputs of the level being
// accumulated (two 8-column tiles): the source rows are the A operand straight out of the staging buffer (pitch 40 floats:
// every 64-bit fragment read is conflict-free), split in registers and shared by both taps when they read the same source
// pixel; the weights arrive pre-split and packed (hi / lo planes of bf16 pairs) as the B operand.  When a level is done
// its accumulator fragments ARE the A fragments of the folded 1x1 convolution (K = that level's 16 channels: accumulator
// columns 2q, 2q + 1 are exactly the k pairs of an A register), so the concatenated 64-channel vector never leaves
// registers.
// History (8 frames): FFMA 4 x 4 / 16 x 2 register tiles / packed fma.f32x2: 921 / 730 / 722 us, bound by shared-memory
// wavefronts plus FFMA issue; TF32 m16n8k8 with a hi / lo split, tiles of 512 (point, tap) samples of one phase: 503 us,
// with the source-row gathers alone at 342 us (2.0 GB of 128-byte L2 reads = 5.9 TB/s) and the MMAs alone at 342 us (the
// legacy HMMA pipe runs at ~1/9 of the tcgen05 rate); a 4-stage ring of 16-channel chunks was slower (624 us); bf16
// m16n8k16 (half the MMAs): 375 us, gather-bound.
__global__ void __launch_bounds__(DG_THREADS, 1) dg_decode_kernel(const __grid_constant__ DgParams p) {
    extern __shared__ __align__(16) float dg_smem[];
    float *s_src = dg_smem;                                                   // [2][4 source pixels][128 points][40]
    uint32_t *s_w = reinterpret_cast<uint32_t *>(s_src + 2 * DG_SRC_FLOATS);   // [2][4 taps][hi, lo][16][20]
    uint32_t *s_w1 = s_w + 2 * DG_W_WORDS;                                     // [hi, lo][32 channel pairs][40]
    int *s_item = reinterpret_cast<int *>(s_w1 + DG_W1_WORDS);                 // [128] point
    int *s_x0 = s_item + DG_PTS, *s_y0 = s_x0 + DG_PTS;                        // [128] nw tap
    uint32_t *s_rowoff = reinterpret_cast<uint32_t *>(s_y0 + DG_PTS);          // [4][128]: element offset of the source row
                                                                               // of the level being staged (~0: not needed)
    const int tile = blockIdx.x;
    const int *tile_start = p.bins + DG_OFF_TILE, *bin_start = p.bins + DG_OFF_BIN;
    if (tile >= __ldg(tile_start + DG_BINS)) return;
    int lo = 0, hi = DG_BINS - 1;              // last bin whose first tile is <= tile
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(tile_start + mid) <= tile) lo = mid; else hi = mid - 1;
    }
    const int bin = lo, py = bin >> 4, px = bin & 15;
    const int first = __ldg(bin_start + bin) + (tile - __ldg(tile_start + bin)) * DG_PTS;
    const int cnt = min(DG_PTS, __ldg(bin_start + bin + 1) - first);
    const int t = threadIdx.x;
    if (t < DG_PTS) {
        const int pg = __ldg(p.items + first + min(t, cnt - 1));
        s_item[t] = pg;
        const DgTaps tp = dg_taps(__ldg(p.xy + (size_t)pg * 2), __ldg(p.xy + (size_t)pg * 2 + 1), p.H, p.W);
        s_x0[t] = tp.x[0]; s_y0[t] = tp.y[0];
    }
    for (int e = t; e < (DG_CAT / 2) * DG_OUT; e += DG_THREADS) {       // w1 (32, 64) -> hi / lo planes of s_w1[k pair][o]
        const int kw = e / DG_OUT, o = e - kw * DG_OUT;
        uint32_t wh, wl;
        split2(__ldg(p.w1 + o * DG_CAT + 2 * kw), __ldg(p.w1 + o * DG_CAT + 2 * kw + 1), wh, wl);
        s_w1[kw * DG_W1P + o] = wh;
        s_w1[(DG_CAT / 2 + kw) * DG_W1P + o] = wl;
    }
    __syncthreads();

    // source rows of level l: slot (oy, ox) holds the rows of the source pixel (y0 >> (l+1)) + oy, (x0 >> (l+1)) + ox; a tap
    // (dy, dx) reads slot (((py & (s-1)) + dy) >> (l+1), ((px & (s-1)) + dx) >> (l+1)).  Pixels outside the map belong to
    // taps outside the image (their samples are ignored by the combine step): clamped.
    auto level_rows = [&](int l) {
        const int sh = l + 1, sm = (1 << sh) - 1;
        const int need_ox = ((px & sm) + 1) >> sh, need_oy = ((py & sm) + 1) >> sh;
        const int hl = p.H >> sh, wl = p.W >> sh;
        if (t < DG_TILE) {
            const int slot = t >> 7, i = t & (DG_PTS - 1);
            const int ox = slot & 1, oy = slot >> 1;
            uint32_t off = 0xffffffffu;
            if (ox <= need_ox && oy <= need_oy) {
                const int pg = s_item[i], b = pg / p.N;
                const int X = min(max((s_x0[i] >> sh) + ox, 0), wl - 1), Y = min(max((s_y0[i] >> sh) + oy, 0), hl - 1);
                off = (uint32_t)((((size_t)b * hl + Y) * wl + X) * p.C[l]);
            }
            s_rowoff[t] = off;
        }
    };
    // a chunk = 32 channels of one level: up to 512 source rows of 128 bytes (8 lanes per row, four rows per warp access)
    // and the two 16 x 16-word weight planes of each of the four tap phases
    auto stage = [&](int c, int buf) {
        const int l = p.chunk_level[c];
        const float *base = p.map[l] + p.chunk_ch[c];
        float *dst = s_src + buf * DG_SRC_FLOATS;
        const int part = t & 7;
#pragma unroll 4
        for (int r = t >> 3; r < DG_TILE; r += DG_THREADS / 8) {
            const uint32_t off = s_rowoff[r];
            if (off != 0xffffffffu) dg_cp16(dst + r * DG_ROW + part * 4, base + off + part * 4);
        }
        {
            const int tap = t >> 7, rem = t & 127;            // 4 taps x 128 pieces of 16 bytes
            const int phase = ((py + (tap >> 1)) & (DG_PH - 1)) * DG_PH + ((px + (tap & 1)) & (DG_PH - 1));
            dg_cp16(s_w + buf * DG_W_WORDS + tap * (2 * DG_R * DG_WP) + (rem >> 2) * DG_WP + (rem & 3) * 4,
                    p.wexp + ((size_t)phase * p.n_chunks + c) * (2 * DG_R * DG_KC / 2) + rem * 4);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int warp = t >> 5, lane = t & 31, g = lane >> 2, q = lane & 3;
    const int mt = warp & 7, tap0 = (warp >> 3) * 2;      // this warp: points mt * 16 .. + 15, taps tap0 and tap0 + 1
    float acc[2][2][4], h[2][4][4];         // [tap][8-output tile][fragment]
#pragma unroll
    for (int tt = 0; tt < 2; ++tt) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[tt][nt][i] = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) h[tt][nt][i] = 0.f;
    }
    level_rows(p.chunk_level[0]);
    __syncthreads();
    stage(0, 0);
    for (int c = 0; c < p.n_chunks; ++c) {
        const int buf = c & 1;
        const int l = p.chunk_level[c];
        if (c + 1 < p.n_chunks) {
            if (p.chunk_level[c + 1] != l) {          // the next chunk opens a level: its row table first
                __syncthreads();                      // everyone has issued the copies that read the current table
                level_rows(p.chunk_level[c + 1]);
                __syncthreads();
            }
            stage(c + 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const int sh = l + 1, sm = (1 << sh) - 1;
        // source-pixel slot of each of this warp's two taps (dy = tap0 >> 1 for both, dx = 0 / 1)
        const int oy = ((py & sm) + (tap0 >> 1)) >> sh;
        const int slot0 = oy * 2 + ((px & sm) >> sh), slot1 = oy * 2 + (((px & sm) + 1) >> sh);       // (px & sm) >> sh == 0
        const float *src0 = s_src + buf * DG_SRC_FLOATS + (slot0 * DG_PTS + mt * 16 + g) * DG_ROW + 2 * q;
        const float *src1 = s_src + buf * DG_SRC_FLOATS + (slot1 * DG_PTS + mt * 16 + g) * DG_ROW + 2 * q;
        const uint32_t *wbase = s_w + buf * DG_W_WORDS + tap0 * (2 * DG_R * DG_WP) + g * DG_WP + q;
#pragma unroll
        for (int ks = 0; ks < DG_KC / 16; ++ks) {
            uint32_t bh[2][2][2], bl[2][2][2];        // [tap][8-output tile][register]
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                const uint32_t *wh = wbase + tt * (2 * DG_R * DG_WP), *wl = wh + DG_R * DG_WP;
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    bh[tt][nt][0] = wh[nt * 8 * DG_WP + ks * 8]; bh[tt][nt][1] = wh[nt * 8 * DG_WP + ks * 8 + 4];
                    bl[tt][nt][0] = wl[nt * 8 * DG_WP + ks * 8]; bl[tt][nt][1] = wl[nt * 8 * DG_WP + ks * 8 + 4];
                }
            }
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                if (tt == 1 && slot1 == slot0) {      // both taps read the same source pixel (warp-uniform)
#pragma unroll
                    for (int i = 0; i < 4; ++i) { ah[1][i] = ah[0][i]; al[1][i] = al[0][i]; }
                } else {
                    const float *r0 = (tt ? src1 : src0) + ks * 16;
                    const float2 x0 = *reinterpret_cast<const float2 *>(r0), x1 = *reinterpret_cast<const float2 *>(r0 + 8 * DG_ROW);
                    const float2 x2 = *reinterpret_cast<const float2 *>(r0 + 8), x3 = *reinterpret_cast<const float2 *>(r0 + 8 * DG_ROW + 8);
                    split2(x0.x, x0.y, ah[tt][0], al[tt][0]);
                    split2(x1.x, x1.y, ah[tt][1], al[tt][1]);
                    split2(x2.x, x2.y, ah[tt][2], al[tt][2]);
                    split2(x3.x, x3.y, ah[tt][3], al[tt][3]);
                }
            }
            // three passes over the accumulator tiles: the MMAs that update the same accumulator are several issues apart
#pragma unroll
            for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) dg_mma(acc[tt][nt], al[tt], bh[tt][nt][0], bh[tt][nt][1]);
#pragma unroll
            for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) dg_mma(acc[tt][nt], ah[tt], bl[tt][nt][0], bl[tt][nt][1]);
#pragma unroll
            for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) dg_mma(acc[tt][nt], ah[tt], bh[tt][nt][0], bh[tt][nt][1]);
        }
        if (c + 1 == p.n_chunks || p.chunk_level[c + 1] != l) {
            // level finished: folded 1x1 convolution with this level's 16 concat channels as K (one k16 step)
            const uint32_t *w1h = s_w1 + (l * (DG_R / 2) + q) * DG_W1P + g;
            const uint32_t *w1l = w1h + (DG_CAT / 2) * DG_W1P;
            uint32_t bh[4][2], bl[4][2];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                bh[nt][0] = w1h[nt * 8]; bh[nt][1] = w1h[4 * DG_W1P + nt * 8];
                bl[nt][0] = w1l[nt * 8]; bl[nt][1] = w1l[4 * DG_W1P + nt * 8];
            }
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                uint32_t ah[4], al[4];
                split2(acc[tt][0][0], acc[tt][0][1], ah[0], al[0]);
                split2(acc[tt][0][2], acc[tt][0][3], ah[1], al[1]);
                split2(acc[tt][1][0], acc[tt][1][1], ah[2], al[2]);
                split2(acc[tt][1][2], acc[tt][1][3], ah[3], al[3]);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dg_mma(h[tt][nt], al, bh[nt][0], bh[nt][1]);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dg_mma(h[tt][nt], ah, bl[nt][0], bl[nt][1]);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dg_mma(h[tt][nt], ah, bh[nt][0], bh[nt][1]);
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[tt][nt][i] = 0.f;
            }
        }
        __syncthreads();       // buffer `buf` is refilled by the stage() of the next iteration
    }
    // bias (BatchNorm shift and decoder biases folded in) + ReLU; a thread holds fused channels 8 nt + 2q, + 1 of points
    // g and g + 8 of its 16-point tile, for its two taps
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const float2 bias = __ldg(reinterpret_cast<const float2 *>(p.b1 + nt * 8 + 2 * q));
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int i = mt * 16 + g + 8 * half;
                if (i < cnt)
                    *reinterpret_cast<float2 *>(p.taps + ((size_t)s_item[i] * 4 + tap0 + tt) * DG_OUT + nt * 8 + 2 * q) =
                        make_float2(fmaxf(h[tt][nt][2 * half] + bias.x, 0.f), fmaxf(h[tt][nt][2 * half + 1] + bias.y, 0.f));
            }
        }
    }
}

// ---- 3. combine -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
dg_combine_kernel(int n, int h, int w, const float *__restrict__ xy, const float *__restrict__ taps,
                  float *__restrict__ out) {
    const int b = blockIdx.y, pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= n) return;
    const size_t pg = (size_t)b * n + pt;
    const DgTaps t = dg_taps(__ldg(xy + pg * 2), __ldg(xy + pg * 2 + 1), h, w);
    float acc[DG_OUT];
#pragma unroll
    for (int c = 0; c < DG_OUT; ++c) acc[c] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {       // nw, ne, sw, se; taps outside the image contribute zero (padding_mode='zeros')
        if (!t.valid[i]) continue;
        const float4 *src = reinterpret_cast<const float4 *>(taps + (pg * 4 + i) * DG_OUT);
#pragma unroll
        for (int c4 = 0; c4 < DG_OUT / 4; ++c4) {
            const float4 v = __ldg(src + c4);
            acc[c4 * 4 + 0] += v.x * t.w[i]; acc[c4 * 4 + 1] += v.y * t.w[i];
            acc[c4 * 4 + 2] += v.z * t.w[i]; acc[c4 * 4 + 3] += v.w * t.w[i];
        }
    }
    float *dst = out + (size_t)b * DG_OUT * n + pt;
#pragma unroll
    for (int c = 0; c < DG_OUT; ++c) dst[(size_t)c * n] = acc[c];
}

// ---- channels-last feature_gather (reference backbone.py:79-89 on an NHWC map) ---------------------------------------
// 32 points x 64 channels per CTA: a warp reads the four taps of a point as contiguous channel runs, the tile is turned in
// shared memory and leaves as 128-byte runs of 32 consecutive points per channel row of the (B, C, N) output.
__global__ void __launch_bounds__(256)
feature_gather_nhwc_kernel(int c, int h, int w, int n, const float *__restrict__ fmap, const float *__restrict__ xy,
                           float *__restrict__ out) {
    __shared__ float s_t[64][33];
    const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int q = warp; q < 32; q += 8) {
        const int pt = p0 + q;
        float a0 = 0.f, a1 = 0.f;
        if (pt < n) {
            const size_t pg = (size_t)b * n + pt;
            const DgTaps t = dg_taps(__ldg(xy + pg * 2), __ldg(xy + pg * 2 + 1), h, w);
            const int ch = c0 + lane * 2;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!t.valid[i] || ch >= c) continue;
                const float *src = fmap + (((size_t)b * h + t.y[i]) * w + t.x[i]) * c + ch;
                if (ch + 1 < c) {
                    const float2 v = __ldg(reinterpret_cast<const float2 *>(src));
                    a0 += v.x * t.w[i]; a1 += v.y * t.w[i];
                } else {
                    a0 += __ldg(src) * t.w[i];
                }
            }
        }
        s_t[lane * 2][q] = a0;
        s_t[lane * 2 + 1][q] = a1;
    }
    __syncthreads();
    for (int r = warp; r < 64; r += 8) {
        const int ch = c0 + r, pt = p0 + lane;
        if (ch < c && pt < n) out[((size_t)b * c + ch) * n + pt] = s_t[r][lane];
    }
}

}  // namespace jmb

extern "C" long long jmb_decode_workspace_bytes(int b, int n) {
    // bins | items (B N ints, laid out for 4 B N) | taps (B N 4 x 32 floats)
    if (b < 0 || n < 0) return -1;
    const long long samples = (long long)b * n * 4;
    return (long long)jmb::DG_BINS_INTS * 4 + 16 + samples * 4 + 16 + samples * jmb::DG_OUT * 4;
}

extern "C" int jmb_feature_gather_nhwc(int b, int c, int h, int w, int n, const float *fmap, const float *xy,
                                       float *out, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && c >= 0 && h > 0 && w > 0 && n >= 0, "feature_gather_nhwc: bad sizes");
    if (b == 0 || c == 0 || n == 0) return JMB_OK;
    JMB_REQUIRE(fmap && xy && out, "feature_gather_nhwc: null pointer");
    JMB_REQUIRE(c % 2 == 0 && ((uintptr_t)fmap & 7) == 0, "feature_gather_nhwc: channel count must be even, map 8-byte aligned");
    JMB_REQUIRE(b <= 65535 && div_up(c, 64) <= 65535, "feature_gather_nhwc: batch / channel count too large");
    dim3 grid(div_up(n, 32), div_up(c, 64), b);
    feature_gather_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, h, w, n, fmap, xy, out);
    return check_launch("feature_gather_nhwc");
}

extern "C" int jmb_decode_gather(int b, int n, int h, int w, const float *xy, const float *m0, const float *m1,
                                 const float *m2, const float *m3, int c0, int c1, int c2, int c3, const void *wexp,
                                 const float *w1, const float *b1, void *workspace, float *out, void *stream) {
    using namespace jmb;
    JMB_REQUIRE(b >= 0 && n >= 0 && h > 0 && w > 0, "decode_gather: bad sizes");
    if (b == 0 || n == 0) return JMB_OK;
    JMB_REQUIRE(h % DG_PH == 0 && w % DG_PH == 0, "decode_gather: image size must be a multiple of 16");
    JMB_REQUIRE((long long)b * n * 4 < (1ll << 29), "decode_gather: too many samples");
    JMB_REQUIRE(xy && m0 && m1 && m2 && m3 && wexp && w1 && b1 && workspace && out, "decode_gather: null pointer");
    DgParams p{};
    const float *maps[DG_LEVELS] = {m0, m1, m2, m3};
    const int C[DG_LEVELS] = {c0, c1, c2, c3};
    p.n_chunks = 0;
    for (int l = 0; l < DG_LEVELS; ++l) {
        JMB_REQUIRE(C[l] > 0 && C[l] % 64 == 0, "decode_gather: level channels must be multiples of 64");
        JMB_REQUIRE(((uintptr_t)maps[l] & 15) == 0, "decode_gather: maps must be 16-byte aligned");
        JMB_REQUIRE((long long)b * (h >> (l + 1)) * (w >> (l + 1)) * C[l] < (1ll << 32), "decode_gather: map too large");
        p.map[l] = maps[l]; p.C[l] = C[l];
        for (int ch = 0; ch < C[l]; ch += DG_KC) {
            JMB_REQUIRE(p.n_chunks < DG_MAX_CHUNKS, "decode_gather: more than 1024 source channels");
            p.chunk_level[p.n_chunks] = l; p.chunk_ch[p.n_chunks] = ch; ++p.n_chunks;
        }
    }
    JMB_REQUIRE((((uintptr_t)wexp | (uintptr_t)w1 | (uintptr_t)workspace | (uintptr_t)out) & 15) == 0,
                "decode_gather: weights / workspace / output must be 16-byte aligned");
    const long long samples = (long long)b * n * 4;
    int *bins = static_cast<int *>(workspace);
    int *items = bins + ((DG_BINS_INTS + 3) / 4) * 4;
    float *taps = reinterpret_cast<float *>(items + ((samples + 3) / 4) * 4);
    p.B = b; p.N = n; p.H = h; p.W = w; p.xy = xy; p.wexp = static_cast<const uint32_t *>(wexp); p.w1 = w1; p.b1 = b1;
    p.bins = bins; p.items = items; p.taps = taps;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    {
        const int rc = device_info(&dev, &sms);
        if (rc != JMB_OK) return rc;
    }
    JMB_FUNC_ATTR_ONCE(dg_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DG_SMEM, dev);
    JMB_CUDA(cudaMemsetAsync(bins, 0, DG_BINS * sizeof(int), st));
    const int total = b * n;
    dg_bin_kernel<false><<<div_up(total, 256), 256, 0, st>>>(total, h, w, xy, bins, items);
    {
        const int rc = check_launch("decode_gather (histogram)");
        if (rc != JMB_OK) return rc;
    }
    dg_scan_kernel<<<1, DG_BINS, 0, st>>>(bins);
    {
        const int rc = check_launch("decode_gather (scan)");
        if (rc != JMB_OK) return rc;
    }
    dg_bin_kernel<true><<<div_up(total, 256), 256, 0, st>>>(total, h, w, xy, bins, items);
    {
        const int rc = check_launch("decode_gather (fill)");
        if (rc != JMB_OK) return rc;
    }
    const int max_tiles = total / DG_PTS + DG_BINS;
    dg_decode_kernel<<<max_tiles, DG_THREADS, DG_SMEM, st>>>(p);
    {
        const int rc = check_launch("decode_gather (decode)");
        if (rc != JMB_OK) return rc;
    }
    dg_combine_kernel<<<dim3(div_up(n, 128), b), 128, 0, st>>>(n, h, w, xy, taps, out);
    return check_launch("decode_gather (combine)");
}
