// Fused flow warp + occlusion-mask composite.
//
// Reference: models/networks/base_network.py:13-37 (get_grid + resample: a linspace mesh built
// on the CPU and copied to the GPU every call, 5 small kernels, F.grid_sample bilinear /
// border / align_corners=True), the [warp, mask] concat of generator.py:439-443 and the linear
// blend raw*m + warp*(1-m) of generator.py:214-224.  One kernel here: flow -> coordinates in
// registers, 4-tap gather, optional blend, writes straight into the channel slice of the
// NHWC buffer the warped-image embedder reads (no grid tensor, no concat copy).
//
// Roofline: HBM-bound; algorithmic bytes = 4*(|img| + |flow| + |mask| + |raw| + |out|).
#include "common.cuh"

struct Tap {
    int x0, x1, y0, y1;
    float wx0, wx1, wy0, wy1;
    float gx, gy;   // d(ix)/d(flow_x), d(iy)/d(flow_y): 1 inside, 0 where the border clamp is active
};

// same arithmetic as the reference: normalised mesh + flow/((W-1)/2), un-normalised with align_corners=True
__device__ __forceinline__ float lin_coord(int j, int n) {
    // torch.linspace(-1, 1, n)[j]
    if (n == 1) return -1.f;
    float step = 2.f / (float)(n - 1);
    return (j < n / 2) ? (-1.f + step * (float)j) : (1.f - step * (float)(n - 1 - j));
}
__device__ __forceinline__ Tap make_tap(int h, int w, int H, int W, float fx, float fy) {
    Tap t;
    float gx = lin_coord(w, W) + fx / ((W - 1.0f) / 2.0f);
    float gy = lin_coord(h, H) + fy / ((H - 1.0f) / 2.0f);
    float ix = ((gx + 1.f) / 2.f) * (float)(W - 1);
    float iy = ((gy + 1.f) / 2.f) * (float)(H - 1);
    t.gx = (ix < 0.f || ix > (float)(W - 1)) ? 0.f : 1.f;
    t.gy = (iy < 0.f || iy > (float)(H - 1)) ? 0.f : 1.f;
    ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
    float fx0 = floorf(ix), fy0 = floorf(iy);
    t.x0 = (int)fx0; t.y0 = (int)fy0;
    t.wx1 = ix - fx0; t.wy1 = iy - fy0;
    t.wx0 = 1.f - t.wx1; t.wy0 = 1.f - t.wy1;
    t.x1 = min(t.x0 + 1, W - 1);   // weight is exactly 0 when x0+1 == W
    t.y1 = min(t.y0 + 1, H - 1);
    return t;
}

__global__ void k_warp_fwd(const float* __restrict__ img, const float* __restrict__ flow, const float* __restrict__ mask,
                           const float* __restrict__ raw, float* __restrict__ out, int N, int H, int W, int Ci, int out_ld,
                           int out_coff, int blend) {
    long long total = (long long)N * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int w = (int)(i % W);
        long long q = i / W;
        int h = (int)(q % H);
        long long n = q / H;
        Tap t = make_tap(h, w, H, W, flow[i * 2], flow[i * 2 + 1]);
        const float* b = img + n * H * W * Ci;
        const float* p00 = b + ((long long)t.y0 * W + t.x0) * Ci;
        const float* p01 = b + ((long long)t.y0 * W + t.x1) * Ci;
        const float* p10 = b + ((long long)t.y1 * W + t.x0) * Ci;
        const float* p11 = b + ((long long)t.y1 * W + t.x1) * Ci;
        float m = mask ? mask[i] : 0.f;
        float* o = out + i * out_ld + out_coff;
        for (int c = 0; c < Ci; ++c) {
            float v = p00[c] * (t.wy0 * t.wx0) + p01[c] * (t.wy0 * t.wx1) + p10[c] * (t.wy1 * t.wx0) + p11[c] * (t.wy1 * t.wx1);
            if (blend) v = raw[i * Ci + c] * m + v * (1.f - m);
            o[c] = v;
        }
        if (!blend && mask) o[Ci] = m;
    }
}

__global__ void k_warp_bwd(const float* __restrict__ img, const float* __restrict__ flow, const float* __restrict__ mask,
                           const float* __restrict__ raw, const float* __restrict__ dout, float* __restrict__ dflow,
                           float* __restrict__ dmask, float* __restrict__ draw, float* __restrict__ dimg, int N, int H, int W,
                           int Ci, int out_ld, int out_coff, int blend) {
    long long total = (long long)N * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int w = (int)(i % W);
        long long q = i / W;
        int h = (int)(q % H);
        long long n = q / H;
        Tap t = make_tap(h, w, H, W, flow[i * 2], flow[i * 2 + 1]);
        long long base = n * H * W * Ci;
        long long o00 = base + ((long long)t.y0 * W + t.x0) * Ci, o01 = base + ((long long)t.y0 * W + t.x1) * Ci;
        long long o10 = base + ((long long)t.y1 * W + t.x0) * Ci, o11 = base + ((long long)t.y1 * W + t.x1) * Ci;
        float m = mask ? mask[i] : 0.f;
        const float* go = dout + i * out_ld + out_coff;
        float dfx = 0.f, dfy = 0.f, dm = 0.f;
        for (int c = 0; c < Ci; ++c) {
            float v00 = img[o00 + c], v01 = img[o01 + c], v10 = img[o10 + c], v11 = img[o11 + c];
            float g = go[c];
            float gw = g;   // gradient w.r.t. the warped value
            if (blend) {
                float wv = v00 * (t.wy0 * t.wx0) + v01 * (t.wy0 * t.wx1) + v10 * (t.wy1 * t.wx0) + v11 * (t.wy1 * t.wx1);
                float rv = raw[i * Ci + c];
                dm += g * (rv - wv);
                if (draw) draw[i * Ci + c] = g * m;
                gw = g * (1.f - m);
            }
            dfx += gw * ((v01 - v00) * t.wy0 + (v11 - v10) * t.wy1);
            dfy += gw * ((v10 - v00) * t.wx0 + (v11 - v01) * t.wx1);
            if (dimg) {
                atomicAdd(dimg + o00 + c, gw * t.wy0 * t.wx0);
                atomicAdd(dimg + o01 + c, gw * t.wy0 * t.wx1);
                atomicAdd(dimg + o10 + c, gw * t.wy1 * t.wx0);
                atomicAdd(dimg + o11 + c, gw * t.wy1 * t.wx1);
            }
        }
        if (!blend && mask) dm = go[Ci];
        dflow[i * 2] = dfx * t.gx;
        dflow[i * 2 + 1] = dfy * t.gy;
        if (dmask) dmask[i] = dm;
    }
}

static inline int wgrid(long long items) {
    long long b = (items + 255) / 256, cap = (long long)fsv_sm_count() * 16;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

extern "C" int fsv_warp_fwd(const float* img, const float* flow, const float* mask, const float* raw, float* out,
                            int N, int H, int W, int Ci, int out_ld, int out_coff, int blend, void* stream) {
    FSV_REQUIRE(N > 0 && H > 1 && W > 1 && Ci > 0, "warp_fwd: bad dims");
    FSV_REQUIRE(!blend || (mask && raw), "warp_fwd: blend needs mask and raw");
    FSV_REQUIRE(out_ld >= out_coff + Ci + ((!blend && mask) ? 1 : 0), "warp_fwd: out_ld too small");
    k_warp_fwd<<<wgrid((long long)N * H * W), 256, 0, (cudaStream_t)stream>>>(img, flow, mask, raw, out, N, H, W, Ci, out_ld, out_coff, blend);
    FSV_CHECK_LAUNCH("warp_fwd");
    return FSV_OK;
}
extern "C" int fsv_warp_bwd(const float* img, const float* flow, const float* mask, const float* raw, const float* dout,
                            float* dflow, float* dmask, float* draw, float* dimg,
                            int N, int H, int W, int Ci, int out_ld, int out_coff, int blend, void* stream) {
    FSV_REQUIRE(N > 0 && H > 1 && W > 1 && Ci > 0 && dflow, "warp_bwd: bad args");
    FSV_REQUIRE(!blend || (mask && raw), "warp_bwd: blend needs mask and raw");
    k_warp_bwd<<<wgrid((long long)N * H * W), 256, 0, (cudaStream_t)stream>>>(img, flow, mask, raw, dout, dflow, dmask, draw, dimg,
                                                                              N, H, W, Ci, out_ld, out_coff, blend);
    FSV_CHECK_LAUNCH("warp_bwd");
    return FSV_OK;
}
