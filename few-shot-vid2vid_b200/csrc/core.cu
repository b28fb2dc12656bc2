// Error reporting, version and device queries of the fsv_b200 C ABI.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void fsv_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fsv_sm_count() {
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (sms[dev] == 0) {
        if (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms[dev] <= 0) sms[dev] = 148;
    }
    return sms[dev];
}

bool fsv_first_on_device(unsigned long long* flags) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    const unsigned long long bit = 1ull << dev;
    if (*flags & bit) return false;
    *flags |= bit;
    return true;
}

extern "C" const char* fsv_last_error(void) { return g_err; }
extern "C" int fsv_version(void) { return 100; }
extern "C" int fsv_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    FSV_CUDA(cudaGetDevice(&dev));
    if (sm_count) FSV_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major) FSV_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) FSV_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    return FSV_OK;
}
