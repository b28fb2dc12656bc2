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
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
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
