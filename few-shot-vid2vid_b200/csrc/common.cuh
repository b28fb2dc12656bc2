// Shared helpers for the fsv_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/fsv_b200.h"

#define FSV_LRELU_SLOPE 0.2f

void fsv_set_error(const char* fmt, ...);

#define FSV_REQUIRE(cond, ...)                  \
    do {                                        \
        if (!(cond)) {                          \
            fsv_set_error(__VA_ARGS__);         \
            return FSV_EINVAL;                  \
        }                                       \
    } while (0)

#define FSV_CHECK_LAUNCH(name)                                                   \
    do {                                                                         \
        cudaError_t e_ = cudaGetLastError();                                     \
        if (e_ != cudaSuccess) {                                                 \
            fsv_set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e_)); \
            return FSV_ECUDA;                                                    \
        }                                                                        \
    } while (0)

#define FSV_CUDA(call)                                                           \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess) {                                                 \
            fsv_set_error("%s failed: %s", #call, cudaGetErrorString(e_));       \
            return FSV_ECUDA;                                                    \
        }                                                                        \
    } while (0)

static inline int fsv_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
int fsv_sm_count();
// true the first time it is called for (flag set, current device): per-DEVICE one-time setup (cudaFuncSetAttribute is per device)
bool fsv_first_on_device(unsigned long long* flags);

// The transcendental activations live in ONE out-of-line copy per translation unit: inlined at every element of an unrolled epilogue
// (64 x tanhf in the SPADE kernel) they made 60-110 KB kernels whose warps stalled on instruction fetch (ncu: stall_no_inst was the
// top stall of k_spade_tc); the layers that use them (output head, flow mask) pay a call per element.
static __device__ __noinline__ float fsv_act_slow(float v, int act) {
    if (act == FSV_ACT_TANH) return tanhf(v);
    if (act == FSV_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    return v;
}
__device__ __forceinline__ float fsv_act(float v, int act) {
    if (act == FSV_ACT_LRELU) return v > 0.f ? v : v * FSV_LRELU_SLOPE;
    if (act == FSV_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == FSV_ACT_NONE) return v;
    return fsv_act_slow(v, act);
}
// derivative of act expressed through the post-activation value y (before out_scale)
__device__ __forceinline__ float fsv_act_grad(float y, int act) {
    if (act == FSV_ACT_LRELU) return y > 0.f ? 1.f : FSV_LRELU_SLOPE;
    if (act == FSV_ACT_TANH) return 1.f - y * y;
    if (act == FSV_ACT_SIGMOID) return y * (1.f - y);
    if (act == FSV_ACT_RELU) return y > 0.f ? 1.f : 0.f;
    return 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
