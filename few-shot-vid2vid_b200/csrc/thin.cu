// Thin-layer convolutions: layers with <= 8 input channels (the networks' first convs: label 1ch, image 3ch,
// [warp,mask] 4ch, flow input 5ch, discriminator input 8ch) or <= 4 output channels (conv_img 32->3, conv_flow
// 32->2, conv_mask 32->1, the discriminator's 512->1 head).  Reference call sites: generator.py:126,210,488-489,
// 523; discriminator.py:69,88; architecture.py:60.
//
// These layers carry <1% of the FLOPs but run at full image resolution; they are HBM/latency-bound and a 64x64
// GEMM tile wastes >90% of its lanes on them (measured: ~23 ms of a 124 ms step on the generic SIMT tile).
// Here every kernel streams the wide tensor exactly once with 16-byte accesses and keeps the thin side in
// registers / shared memory:
//   thin_cin_fwd   : one thread per output pixel, 32 output channels in registers, weights [tap][ci][co] in smem
//   thin_cout_fwd  : one thread per output pixel, <=4 accumulators, float4 loads of the 32+ input channels
//   thin_cin_wgrad : one warp per pixel chunk, lane = output channel, taps*Cin accumulators per lane
//   thin_cout_wgrad: one warp per pixel chunk, lane = input channel, taps*Cout accumulators per lane
// Roofline: HBM; algorithmic bytes = 4*(|x| + |y|) (+ weights, negligible).
#include "common.cuh"

struct ThinP {
    int N, H, W, Cin, x_ld, x_coff, Cout, kh, kw, stride, pad, Ho, Wo, y_ld, y_coff, act, in_act;
    float out_scale;
    int res_ld, res_coff;
};

static ThinP thin_p(const fsv_conv_desc* d) {
    ThinP p;
    p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.x_ld = d->x_ld; p.x_coff = d->x_coff; p.Cout = d->Cout;
    p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad = d->pad; p.Ho = d->Ho; p.Wo = d->Wo; p.y_ld = d->y_ld; p.y_coff = d->y_coff;
    p.act = d->act; p.in_act = d->in_act; p.out_scale = d->out_scale; p.res_ld = d->res_ld; p.res_coff = d->res_coff;
    return p;
}

// ------------------------------------------------------------------ Cin <= 8 forward
// grid.x = pixel blocks of 128, grid.y = 32-wide output-channel chunks
__global__ void __launch_bounds__(128) k_thin_cin_fwd(ThinP p, const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, const float* __restrict__ residual,
                                                      float* __restrict__ y) {
    extern __shared__ float ws[];                 // [tap][ci][32]
    const int taps = p.kh * p.kw;
    const int co0 = blockIdx.y * 32;
    for (int i = threadIdx.x; i < taps * p.Cin * 32; i += blockDim.x) {
        int c = i & 31, tc = i >> 5;
        int ci = tc % p.Cin, tap = tc / p.Cin;
        int co = co0 + c;
        ws[i] = co < p.Cout ? w[((long long)co * taps + tap) * p.Cin + ci] : 0.f;
    }
    __syncthreads();
    const long long total = (long long)p.N * p.Ho * p.Wo;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= total) return;
    const int wo = (int)(pix % p.Wo);
    const long long q = pix / p.Wo;
    const int ho = (int)(q % p.Ho);
    const long long n = q / p.Ho;
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    for (int r = 0; r < p.kh; ++r) {
        int ih = ho * p.stride + r - p.pad;
        if (ih < 0 || ih >= p.H) continue;
        for (int s = 0; s < p.kw; ++s) {
            int iw = wo * p.stride + s - p.pad;
            if (iw < 0 || iw >= p.W) continue;
            const float* xp = x + ((n * p.H + ih) * p.W + iw) * p.x_ld + p.x_coff;
            const float* wt = ws + (r * p.kw + s) * p.Cin * 32;
            for (int ci = 0; ci < p.Cin; ++ci) {
                float xv = xp[ci];
                if (p.in_act == FSV_ACT_LRELU) xv = fsv_act(xv, FSV_ACT_LRELU);
                const float4* w4 = reinterpret_cast<const float4*>(wt + ci * 32);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    float4 wv = w4[c4];
                    acc[c4 * 4 + 0] += xv * wv.x; acc[c4 * 4 + 1] += xv * wv.y;
                    acc[c4 * 4 + 2] += xv * wv.z; acc[c4 * 4 + 3] += xv * wv.w;
                }
            }
        }
    }
    float* yp = y + pix * p.y_ld + p.y_coff + co0;
    const float* rp = residual ? residual + pix * p.res_ld + p.res_coff + co0 : nullptr;
    const bool vec = ((p.y_ld & 3) == 0) && ((p.y_coff & 3) == 0) && (co0 + 32 <= p.Cout) && ((((uintptr_t)y) & 15) == 0);
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = c4 * 4 + j;
            float t = acc[c];
            if (co0 + c < p.Cout) {
                if (bias) t += bias[co0 + c];
                if (rp) t += rp[c];
            }
            v[j] = fsv_act(t, p.act) * p.out_scale;
        }
        if (vec) *reinterpret_cast<float4*>(yp + c4 * 4) = make_float4(v[0], v[1], v[2], v[3]);
        else
            for (int j = 0; j < 4; ++j)
                if (co0 + c4 * 4 + j < p.Cout) yp[c4 * 4 + j] = v[j];
    }
}

// ------------------------------------------------------------------ Cout <= 4 forward (Cin % 4 == 0)
__global__ void __launch_bounds__(128) k_thin_cout_fwd(ThinP p, const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, const float* __restrict__ residual,
                                                       float* __restrict__ y) {
    extern __shared__ float ws[];                 // [co][tap][ci]
    const int taps = p.kh * p.kw;
    const int wn = p.Cout * taps * p.Cin;
    for (int i = threadIdx.x; i < wn; i += blockDim.x) ws[i] = w[i];
    __syncthreads();
    const long long total = (long long)p.N * p.Ho * p.Wo;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= total) return;
    const int wo = (int)(pix % p.Wo);
    const long long q = pix / p.Wo;
    const int ho = (int)(q % p.Ho);
    const long long n = q / p.Ho;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < p.kh; ++r) {
        int ih = ho * p.stride + r - p.pad;
        if (ih < 0 || ih >= p.H) continue;
        for (int s = 0; s < p.kw; ++s) {
            int iw = wo * p.stride + s - p.pad;
            if (iw < 0 || iw >= p.W) continue;
            const float4* xp = reinterpret_cast<const float4*>(x + ((n * p.H + ih) * p.W + iw) * p.x_ld + p.x_coff);
            const int tap = r * p.kw + s;
            for (int c4 = 0; c4 < p.Cin / 4; ++c4) {
                float4 xv = xp[c4];
                if (p.in_act == FSV_ACT_LRELU) {
                    xv.x = fsv_act(xv.x, FSV_ACT_LRELU); xv.y = fsv_act(xv.y, FSV_ACT_LRELU);
                    xv.z = fsv_act(xv.z, FSV_ACT_LRELU); xv.w = fsv_act(xv.w, FSV_ACT_LRELU);
                }
#pragma unroll
                for (int co = 0; co < 4; ++co) {
                    if (co < p.Cout) {
                        float4 wv = *reinterpret_cast<const float4*>(ws + (co * taps + tap) * p.Cin + c4 * 4);
                        acc[co] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
                    }
                }
            }
        }
    }
    float* yp = y + pix * p.y_ld + p.y_coff;
    const float* rp = residual ? residual + pix * p.res_ld + p.res_coff : nullptr;
#pragma unroll
    for (int co = 0; co < 4; ++co) {
        if (co < p.Cout) {
            float t = acc[co];
            if (bias) t += bias[co];
            if (rp) t += rp[co];
            yp[co] = fsv_act(t, p.act) * p.out_scale;
        }
    }
}

// ------------------------------------------------------------------ Cin <= 8 weight gradient: lane = output channel
// dw[co][tap][ci] += sum_px dy[px][co] * x[px shifted][ci];  TC = compile-time bound on taps*Cin
template <int TC>
__global__ void __launch_bounds__(128) k_thin_cin_wgrad(ThinP p, const float* __restrict__ x, const float* __restrict__ dy,
                                                        float* __restrict__ dw, int pix_per_warp) {
    const int lane = threadIdx.x & 31;
    const int warp_g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int co = blockIdx.y * 32 + lane;
    const int taps = p.kh * p.kw;
    const int tc = taps * p.Cin;
    const long long total = (long long)p.N * p.Ho * p.Wo;
    long long px0 = (long long)warp_g * pix_per_warp;
    long long px1 = px0 + pix_per_warp;
    if (px1 > total) px1 = total;
    float acc[TC];
#pragma unroll
    for (int i = 0; i < TC; ++i) acc[i] = 0.f;
    for (long long px = px0; px < px1; ++px) {
        const int wo = (int)(px % p.Wo);
        const long long q = px / p.Wo;
        const int ho = (int)(q % p.Ho);
        const long long n = q / p.Ho;
        const float dv = co < p.Cout ? dy[px * p.y_ld + p.y_coff + co] : 0.f;
#pragma unroll
        for (int i = 0; i < TC; ++i) {
            if (i < tc) {
                int tap = i / p.Cin, ci = i - tap * p.Cin;
                int r = tap / p.kw, s = tap - r * p.kw;
                int ih = ho * p.stride + r - p.pad, iw = wo * p.stride + s - p.pad;
                float xv = 0.f;
                if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) xv = x[((n * p.H + ih) * p.W + iw) * p.x_ld + p.x_coff + ci];
                if (p.in_act == FSV_ACT_LRELU) xv = fsv_act(xv, FSV_ACT_LRELU);
                acc[i] += dv * xv;
            }
        }
    }
    if (co < p.Cout) {
#pragma unroll
        for (int i = 0; i < TC; ++i)
            if (i < tc) atomicAdd(dw + (long long)co * tc + i, acc[i]);
    }
}

// ------------------------------------------------------------------ Cout <= 4 weight gradient: lane = input channel
template <int TAPS>
__global__ void __launch_bounds__(128) k_thin_cout_wgrad(ThinP p, const float* __restrict__ x, const float* __restrict__ dy,
                                                         float* __restrict__ dw, int pix_per_warp) {
    const int lane = threadIdx.x & 31;
    const int warp_g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int ci = blockIdx.y * 32 + lane;
    const int taps = p.kh * p.kw;
    const long long total = (long long)p.N * p.Ho * p.Wo;
    long long px0 = (long long)warp_g * pix_per_warp;
    long long px1 = px0 + pix_per_warp;
    if (px1 > total) px1 = total;
    float acc[TAPS][4];
#pragma unroll
    for (int t = 0; t < TAPS; ++t)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[t][c] = 0.f;
    for (long long px = px0; px < px1; ++px) {
        const int wo = (int)(px % p.Wo);
        const long long q = px / p.Wo;
        const int ho = (int)(q % p.Ho);
        const long long n = q / p.Ho;
        float dv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) dv[c] = c < p.Cout ? dy[px * p.y_ld + p.y_coff + c] : 0.f;
#pragma unroll
        for (int t = 0; t < TAPS; ++t) {
            if (t < taps) {
                int r = t / p.kw, s = t - r * p.kw;
                int ih = ho * p.stride + r - p.pad, iw = wo * p.stride + s - p.pad;
                float xv = 0.f;
                if (ci < p.Cin && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
                    xv = x[((n * p.H + ih) * p.W + iw) * p.x_ld + p.x_coff + ci];
                if (p.in_act == FSV_ACT_LRELU) xv = fsv_act(xv, FSV_ACT_LRELU);
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[t][c] += dv[c] * xv;
            }
        }
    }
    if (ci < p.Cin) {
#pragma unroll
        for (int t = 0; t < TAPS; ++t)
            if (t < taps)
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < p.Cout) atomicAdd(dw + ((long long)c * taps + t) * p.Cin + ci, acc[t][c]);
    }
}

// ------------------------------------------------------------------ Cin <= 8 data gradient (thin OUTPUT dx, wide dy)
// dx[n,h,w,ci] = sum_{r,s,co} dy[n,(h+pad-r)/stride,(w+pad-s)/stride,co] * w[co][r][s][ci]; one thread per dx pixel.
// Needed for the discriminator's first conv (8 -> 32, k4 s2): its input carries the generated image.
__global__ void __launch_bounds__(128) k_thin_cin_dgrad(ThinP p, const float* __restrict__ dy, const float* __restrict__ w,
                                                        float* __restrict__ dx, int accumulate) {
    extern __shared__ float ws[];                 // [tap][co][ci8]
    const int taps = p.kh * p.kw;
    for (int i = threadIdx.x; i < taps * p.Cout * 8; i += blockDim.x) {
        int ci = i & 7, tco = i >> 3;
        int co = tco % p.Cout, tap = tco / p.Cout;
        ws[i] = ci < p.Cin ? w[((long long)co * taps + tap) * p.Cin + ci] : 0.f;
    }
    __syncthreads();
    const long long total = (long long)p.N * p.H * p.W;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= total) return;
    const int wq = (int)(pix % p.W);
    const long long q = pix / p.W;
    const int hq = (int)(q % p.H);
    const long long n = q / p.H;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    for (int r = 0; r < p.kh; ++r) {
        int th = hq + p.pad - r;
        if (th < 0 || (th % p.stride) != 0) continue;
        int oh = th / p.stride;
        if (oh >= p.Ho) continue;
        for (int s = 0; s < p.kw; ++s) {
            int tw = wq + p.pad - s;
            if (tw < 0 || (tw % p.stride) != 0) continue;
            int ow = tw / p.stride;
            if (ow >= p.Wo) continue;
            const float4* dp = reinterpret_cast<const float4*>(dy + ((n * p.Ho + oh) * p.Wo + ow) * p.y_ld + p.y_coff);
            const float* wt = ws + (r * p.kw + s) * p.Cout * 8;
            for (int c4 = 0; c4 < p.Cout / 4; ++c4) {
                float4 dv = dp[c4];
                float d4[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 w0 = *reinterpret_cast<const float4*>(wt + (c4 * 4 + j) * 8);
                    const float4 w1 = *reinterpret_cast<const float4*>(wt + (c4 * 4 + j) * 8 + 4);
                    acc[0] += d4[j] * w0.x; acc[1] += d4[j] * w0.y; acc[2] += d4[j] * w0.z; acc[3] += d4[j] * w0.w;
                    acc[4] += d4[j] * w1.x; acc[5] += d4[j] * w1.y; acc[6] += d4[j] * w1.z; acc[7] += d4[j] * w1.w;
                }
            }
        }
    }
    float* xp = dx + pix * p.x_ld + p.x_coff;
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if (c < p.Cin) { if (accumulate) xp[c] += acc[c]; else xp[c] = acc[c]; }
}

// ------------------------------------------------------------------ host side
static bool thin_common_ok(const fsv_conv_desc* d) {
    return d->up == 1 && d->w_nstride == 0 && d->b_nstride == 0 && d->kh * d->kw <= 16 && (long long)d->N * d->Ho * d->Wo >= 4096;
}
extern "C" int fsv_conv2d_thin_kind(const fsv_conv_desc* d) {
    if (!d || !thin_common_ok(d)) return 0;
    if (d->Cin <= 8 && d->Cout >= 16) return 1;                                         // thin input
    if (d->Cout <= 4 && d->Cin >= 16 && d->Cin % 4 == 0 && d->x_ld % 4 == 0 && d->x_coff % 4 == 0 &&
        (long long)d->Cout * d->kh * d->kw * d->Cin * 4 <= 40 * 1024) return 2;         // thin output
    return 0;
}

extern "C" int fsv_conv2d_fwd_thin(const fsv_conv_desc* d, const float* x, const float* w, const float* bias,
                                   const float* residual, float* y, void* stream) {
    int kind = fsv_conv2d_thin_kind(d);
    FSV_REQUIRE(kind != 0, "conv2d_fwd_thin: not a thin layer");
    ThinP p = thin_p(d);
    const long long total = (long long)d->N * d->Ho * d->Wo;
    const int taps = d->kh * d->kw;
    if (kind == 1) {
        dim3 grid(fsv_cdiv(total, 128), fsv_cdiv(d->Cout, 32));
        k_thin_cin_fwd<<<grid, 128, taps * d->Cin * 32 * sizeof(float), (cudaStream_t)stream>>>(p, x, w, bias, residual, y);
    } else {
        FSV_REQUIRE((((uintptr_t)x) & 15) == 0 && (((uintptr_t)w) & 15) == 0, "conv2d_fwd_thin: pointers must be 16-byte aligned");
        dim3 grid(fsv_cdiv(total, 128));
        k_thin_cout_fwd<<<grid, 128, (size_t)d->Cout * taps * d->Cin * sizeof(float), (cudaStream_t)stream>>>(p, x, w, bias, residual, y);
    }
    FSV_CHECK_LAUNCH("conv2d_fwd_thin");
    return FSV_OK;
}

// dw only (the bias gradient stays with the column-sum kernel of conv_simt.cu); dw must be zeroed by the caller
extern "C" int fsv_conv2d_wgrad_thin(const fsv_conv_desc* d, const float* x, const float* dy, float* dw, void* stream) {
    int kind = fsv_conv2d_thin_kind(d);
    FSV_REQUIRE(kind != 0, "conv2d_wgrad_thin: not a thin layer");
    ThinP p = thin_p(d);
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)d->N * d->Ho * d->Wo;
    const int taps = d->kh * d->kw;
    long long warps = (long long)fsv_sm_count() * 32;
    int pix_per_warp = (int)((total + warps - 1) / warps);
    if (pix_per_warp < 16) pix_per_warp = 16;
    long long nwarps = (total + pix_per_warp - 1) / pix_per_warp;
    dim3 block(128);
    if (kind == 1) {
        dim3 grid(fsv_cdiv(nwarps, 4), fsv_cdiv(d->Cout, 32));
        int tc = taps * d->Cin;
        if (tc <= 16) k_thin_cin_wgrad<16><<<grid, block, 0, st>>>(p, x, dy, dw, pix_per_warp);
        else if (tc <= 48) k_thin_cin_wgrad<48><<<grid, block, 0, st>>>(p, x, dy, dw, pix_per_warp);
        else k_thin_cin_wgrad<128><<<grid, block, 0, st>>>(p, x, dy, dw, pix_per_warp);
    } else {
        dim3 grid(fsv_cdiv(nwarps, 4), fsv_cdiv(d->Cin, 32));
        if (taps <= 9) k_thin_cout_wgrad<9><<<grid, block, 0, st>>>(p, x, dy, dw, pix_per_warp);
        else k_thin_cout_wgrad<16><<<grid, block, 0, st>>>(p, x, dy, dw, pix_per_warp);
    }
    FSV_CHECK_LAUNCH("conv2d_wgrad_thin");
    return FSV_OK;
}

extern "C" int fsv_conv2d_dgrad_thin_ok(const fsv_conv_desc* d) {
    return d && d->up == 1 && d->w_nstride == 0 && d->Cin <= 8 && d->Cout >= 16 && d->Cout % 4 == 0 && d->y_ld % 4 == 0 &&
           d->y_coff % 4 == 0 && d->kh * d->kw <= 16 && (long long)d->kh * d->kw * d->Cout * 8 * 4 <= 40 * 1024 &&
           (long long)d->N * d->H * d->W >= 4096;
}
extern "C" int fsv_conv2d_dgrad_thin(const fsv_conv_desc* d, const float* dy, const float* w, float* dx, int accumulate, void* stream) {
    FSV_REQUIRE(fsv_conv2d_dgrad_thin_ok(d), "conv2d_dgrad_thin: not eligible");
    FSV_REQUIRE((((uintptr_t)dy) & 15) == 0, "conv2d_dgrad_thin: dy must be 16-byte aligned");
    ThinP p = thin_p(d);
    const long long total = (long long)d->N * d->H * d->W;
    k_thin_cin_dgrad<<<fsv_cdiv(total, 128), 128, (size_t)d->kh * d->kw * d->Cout * 8 * sizeof(float), (cudaStream_t)stream>>>(p, dy, w, dx, accumulate);
    FSV_CHECK_LAUNCH("conv2d_dgrad_thin");
    return FSV_OK;
}
