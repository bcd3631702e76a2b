// Library-level entry points of libjmodt_b200.so (see include/jmodt_b200.h).
#include "common.cuh"

#include <stdarg.h>
#include <string.h>

namespace jmb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int device_info(int *dev, int *sms) {
    static int cache[JMB_MAX_DEVICES] = {0};      // written once per device with the same value: benign race
    int d = 0;
    JMB_CUDA(cudaGetDevice(&d));
    JMB_REQUIRE(d >= 0 && d < JMB_MAX_DEVICES, "device index %d out of range", d);
    int n = __atomic_load_n(&cache[d], __ATOMIC_RELAXED);
    if (n == 0) {
        JMB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d));
        __atomic_store_n(&cache[d], n, __ATOMIC_RELAXED);
    }
    *dev = d;
    *sms = n;
    return JMB_OK;
}

int set_func_attr_once(const void *func, cudaFuncAttribute attr, int value, int dev, unsigned long long *mask) {
    const unsigned long long bit = 1ULL << dev;
    if (__atomic_load_n(mask, __ATOMIC_ACQUIRE) & bit) return JMB_OK;
    JMB_CUDA(cudaFuncSetAttribute(func, attr, value));      // idempotent: two threads racing here both succeed
    __atomic_fetch_or(mask, bit, __ATOMIC_RELEASE);
    return JMB_OK;
}

}  // namespace jmb

extern "C" int jmb_version(void) { return 1; }

extern "C" const char *jmb_last_error(void) { return jmb::g_err; }

extern "C" int jmb_set_device(int device) {
    JMB_REQUIRE(device >= 0, "set_device: negative device");
    JMB_CUDA(cudaSetDevice(device));
    return JMB_OK;
}

extern "C" int jmb_sm_count(void) {
    int dev = 0, sms = 0;
    const int rc = jmb::device_info(&dev, &sms);
    return rc == JMB_OK ? sms : rc;
}
