// Row-wise softmax over the channel dimension of the reference-label features
// (models/networks/generator.py:385, nn.Softmax(dim=1) on NCHW == over C for NHWC rows).
// The outer product that follows (generator.py:386) is a per-sample GEMM served by
// fsv_conv2d_wgrad (forward) and fsv_conv2d_fwd / fsv_conv2d_dgrad (backward), so the
// (b,c,c,hw) temporary of the reference is never materialised.
// One warp per row; rows are tiny (C <= 1024) and few (B*hw), the kernel is latency-bound.
#include "common.cuh"

__global__ void k_softmax_fwd(const float* __restrict__ x, float* __restrict__ y, long long rows, int C) {
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + row * C;
    float* yr = y + row * C;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, xr[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(xr[c] - m);
    s = warp_sum(s);
    float inv = 1.f / s;
    for (int c = lane; c < C; c += 32) yr[c] = expf(xr[c] - m) * inv;
}
__global__ void k_softmax_bwd(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, long long rows, int C) {
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* yr = y + row * C;
    const float* gr = dy + row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += yr[c] * gr[c];
    s = warp_sum(s);
    for (int c = lane; c < C; c += 32) dx[row * C + c] = yr[c] * (gr[c] - s);
}
extern "C" int fsv_softmax_rows_fwd(const float* x, float* y, long long rows, int C, void* stream) {
    FSV_REQUIRE(rows > 0 && C > 0, "softmax: bad dims");
    k_softmax_fwd<<<fsv_cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, y, rows, C);
    FSV_CHECK_LAUNCH("softmax_fwd");
    return FSV_OK;
}
extern "C" int fsv_softmax_rows_bwd(const float* y, const float* dy, float* dx, long long rows, int C, void* stream) {
    FSV_REQUIRE(rows > 0 && C > 0, "softmax: bad dims");
    k_softmax_bwd<<<fsv_cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(y, dy, dx, rows, C);
    FSV_CHECK_LAUNCH("softmax_bwd");
    return FSV_OK;
}
