// Dispatch between the SIMT and the tcgen05/TMA convolution paths.
#include "common.cuh"

extern "C" int fsv_conv2d_fwd_simt(const fsv_conv_desc* d, const float* x, const float* w, const float* bias,
                                   const float* residual, float* y, void* stream);
extern "C" int fsv_conv2d_fwd_tc(const fsv_conv_desc* d, const float* x, const float* w, const float* bias,
                                 const float* residual, float* y, void* stream);

extern "C" int fsv_conv2d_fwd(const fsv_conv_desc* d, const float* x, const float* w, const float* bias,
                              const float* residual, float* y, void* stream) {
    FSV_REQUIRE(d != nullptr, "conv2d_fwd: null descriptor");
    if (d->use_tc == 1) return fsv_conv2d_fwd_tc(d, x, w, bias, residual, y, stream);
    if (d->use_tc == -1 && fsv_conv2d_tc_eligible(d)) return fsv_conv2d_fwd_tc(d, x, w, bias, residual, y, stream);
    return fsv_conv2d_fwd_simt(d, x, w, bias, residual, y, stream);
}
