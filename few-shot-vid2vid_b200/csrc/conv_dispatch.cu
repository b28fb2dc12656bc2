// Dispatch between the SIMT and the tcgen05/TMA convolution paths.
#include "common.cuh"

extern "C" int fsv_conv2d_fwd_simt(const fsv_conv_desc* d, const float* x, const float* w, const float* bias,
                                   const float* residual, float* y, void* stream);
extern "C" int fsv_conv2d_fwd_tc(const fsv_conv_desc* d, const float* x, const float* w, const float* bias,
                                 const float* residual, float* y, void* stream);

extern "C" int fsv_conv2d_thin_kind(const fsv_conv_desc* d);
extern "C" int fsv_conv2d_fwd_thin(const fsv_conv_desc* d, const float* x, const float* w, const float* bias,
                                   const float* residual, float* y, void* stream);
int fsv_conv_validate(const fsv_conv_desc* d, const char* who);

extern "C" int fsv_conv2d_fwd(const fsv_conv_desc* d, const float* x, const float* w, const float* bias,
                              const float* residual, float* y, void* stream) {
    FSV_REQUIRE(d != nullptr, "conv2d_fwd: null descriptor");
    if (d->use_tc != 1 && fsv_conv2d_thin_kind(d)) {      // thin layers: dedicated exact-fp32 streaming kernels
        int rc = fsv_conv_validate(d, "conv2d_fwd");
        if (rc) return rc;
        return fsv_conv2d_fwd_thin(d, x, w, bias, residual, y, stream);
    }
    if (d->use_tc == 1) return fsv_conv2d_fwd_tc(d, x, w, bias, residual, y, stream);
    if (d->use_tc == -1 && fsv_conv2d_tc_eligible(d)) return fsv_conv2d_fwd_tc(d, x, w, bias, residual, y, stream);
    return fsv_conv2d_fwd_simt(d, x, w, bias, residual, y, stream);
}
