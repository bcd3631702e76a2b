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
    if (cudaGetDevice(&dev) != cudaSuccess) return JMB_ERR_CUDA;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return JMB_ERR_CUDA;
    return sms;
}
