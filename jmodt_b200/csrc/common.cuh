// Shared helpers for the jmodt_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/jmodt_b200.h"

namespace jmb {

// ---- error plumbing: entry points return codes, never exit() ---------------------------
void set_error(const char *fmt, ...);

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return JMB_ERR_CUDA;
    }
    return JMB_OK;
}

#define JMB_REQUIRE(cond, ...)              \
    do {                                    \
        if (!(cond)) {                      \
            ::jmb::set_error(__VA_ARGS__);  \
            return JMB_ERR_INVALID_ARG;     \
        }                                   \
    } while (0)

#define JMB_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) {                                             \
            ::jmb::set_error("%s: %s", #call, cudaGetErrorString(e__));      \
            return JMB_ERR_CUDA;                                              \
        }                                                                     \
    } while (0)

// ---- per-device state: the library may be driven on several GPUs from one process (the reference's only multi-GPU
// mode is single-process nn.DataParallel, tools/train.py:86-87), so nothing device-dependent is cached process-wide.
constexpr int JMB_MAX_DEVICES = 64;
// current device and its SM count (cached per device, thread-safe); returns JMB_OK or an error code
int device_info(int *dev, int *sms);
// cudaFuncSetAttribute once per (kernel, device): `mask` is the call site's own per-device "already set" bit set
int set_func_attr_once(const void *func, cudaFuncAttribute attr, int value, int dev, unsigned long long *mask);

#define JMB_FUNC_ATTR_ONCE(kernel, attr, value, dev)                                                         \
    do {                                                                                                     \
        static unsigned long long mask__ = 0;                                                                \
        int rc__ = ::jmb::set_func_attr_once(reinterpret_cast<const void *>(kernel), attr, value, dev, &mask__); \
        if (rc__ != JMB_OK) return rc__;                                                                     \
    } while (0)

inline int div_up(int a, int b) { return (a + b - 1) / b; }
inline long long div_up_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---- device helpers -------------------------------------------------------------------

// Squared distance in the reference's exact fp32 operation order (read from its SASS):
//   fma(dz, dz, fma(dx, dx, fl(dy*dy)))      ball_query_gpu.cu:33, sampling_gpu.cu:133,
//                                            interpolate_gpu.cu:36
// Intrinsics pin the rounding points so the compiler can neither fuse nor re-associate.
__device__ __forceinline__ float dist2_ref(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// a*b - c*d the way the reference binaries evaluate it: fma(a, b, -fl(c*d)).
__device__ __forceinline__ float fmsub2(float a, float b, float c, float d) {
    return __fmaf_rn(a, b, -__fmul_rn(c, d));
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ float ld_nc(const float *p) { return __ldg(p); }

}  // namespace jmb
