// tcgen05/TMA implicit-GEMM convolution (placeholder until the kernel lands: reports "not eligible").
#include "common.cuh"
extern "C" int fsv_conv2d_tc_eligible(const fsv_conv_desc* d) { (void)d; return 0; }
extern "C" int fsv_conv2d_fwd_tc(const fsv_conv_desc* d, const float* x, const float* w, const float* bias,
                                 const float* residual, float* y, void* stream) {
    (void)d; (void)x; (void)w; (void)bias; (void)residual; (void)y; (void)stream;
    fsv_set_error("conv2d_fwd_tc: tcgen05 path not built");
    return FSV_ENOTSUP;
}
