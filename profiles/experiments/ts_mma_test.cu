// Experiment: tcgen05.mma with the A operand in tensor memory (TS mode), to learn the TMEM layout of A.
// D (128 x 128 fp32) = A (128 x 16 bf16, TMEM) * B (16 x 128 bf16, smem MN-major no-swizzle image).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -I jmodt_b200/csrc -o /tmp/ts_test profiles/experiments/ts_mma_test.cu
#include "tc_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
namespace jmb { void set_error(const char *, ...) {} }
using namespace jmb;

__global__ void ts_test(const uint32_t *a_packed /*128 x 8 words*/, const __nv_bfloat16 *b_img /*8KB image*/, float *d_out) {
    __shared__ __align__(1024) uint8_t s_b[TC_IMG];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&s_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < TC_IMG / 16; i += blockDim.x) reinterpret_cast<uint4 *>(s_b)[i] = reinterpret_cast<const uint4 *>(b_img)[i];
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    // A row m -> lane m, columns [128, 136): 8 packed words
    {
        const int m = threadIdx.x;
        uint32_t r[8];
        for (int j = 0; j < 8; ++j) r[j] = a_packed[m * 8 + j];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 128;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                     "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const uint64_t bdesc = make_smem_desc(smem_u32(s_b));
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem), "r"(tmem + 128), "l"(bdesc), "r"(TC_IDESC), "r"(0u) : "memory");
        umma_commit(&s_bar);
    }
    mbar_wait(&s_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 128; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
        for (int j = 0; j < 32; ++j) d_out[(size_t)(warp * 32 + lane) * 128 + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFF + ((u >> 16) & 1); return (uint16_t)(u >> 16); }
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
    std::vector<float> A(128 * 16), B(16 * 128);
    for (auto &v : A) v = bf2f(f2bf((rand() % 200 - 100) / 64.0f));
    for (auto &v : B) v = bf2f(f2bf((rand() % 200 - 100) / 64.0f));
    std::vector<uint32_t> ap(128 * 8);
    for (int m = 0; m < 128; ++m)
        for (int j = 0; j < 8; ++j) ap[m * 8 + j] = (uint32_t)f2bf(A[m * 16 + 2 * j]) | ((uint32_t)f2bf(A[m * 16 + 2 * j + 1]) << 16);
    std::vector<uint16_t> bimg(TC_IMG / 2, 0);
    for (int k = 0; k < 16; ++k)
        for (int n = 0; n < 128; ++n) bimg[((n / 8) * 128 + (k / 8) * 2048 + (k % 8) * 16 + (n % 8) * 2) / 2] = f2bf(B[k * 128 + n]);
    uint32_t *dap; __nv_bfloat16 *db; float *dd;
    cudaMalloc(&dap, ap.size() * 4); cudaMalloc(&db, TC_IMG); cudaMalloc(&dd, 128 * 128 * 4);
    cudaMemcpy(dap, ap.data(), ap.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bimg.data(), TC_IMG, cudaMemcpyHostToDevice);
    ts_test<<<1, 128>>>(dap, db, dd);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    std::vector<float> D(128 * 128);
    cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
            double ref = 0;
            for (int k = 0; k < 16; ++k) ref += (double)A[m * 16 + k] * B[k * 128 + n];
            double err = fabs(ref - D[m * 128 + n]);
            if (err > 1e-3) ++bad;
            if (err > maxerr) maxerr = err;
        }
    printf("TS-mode MMA: max err %.3e, bad %d / 16384  (D[0][0]=%f D[5][7]=%f)\n", maxerr, bad, D[0], D[5 * 128 + 7]);
    return 0;
}
