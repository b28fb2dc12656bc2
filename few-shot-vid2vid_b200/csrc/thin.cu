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
//   thin_cin_dgrad : one thread per input pixel (<= 8 channels), float4 loads of the wide dy
//   tco_fwd/dgrad/wgrad (Cout <= 4): work item = (pixel, 4-channel group) -- LPP lanes per pixel, so a warp reads
//                    32/LPP neighbouring pixels x LPP float4 as ONE contiguous run (the first version had one thread
//                    per pixel: every 16-byte load touched 32 different 128-byte lines and the kernels ran at ~1/10 of
//                    HBM speed); blocks own 16x8-pixel tiles so the 3x3 / 4x4 halo re-reads hit L1
// Roofline: HBM; algorithmic bytes = 4*(|x| + |y|) (+ weights, negligible).
#include "common.cuh"
#include <stdlib.h>

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
// grid.x = pixel blocks of 128 * THIN_PPT, grid.y = 32-wide output-channel chunks
#define THIN_PPT 4
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
    // THIN_PPT pixels per thread, 128 apart: the weight staging above (taps*Cin*32 global loads per block) is amortised over 512 pixels
    for (int it = 0; it < THIN_PPT; ++it) {
    const long long pix = ((long long)blockIdx.x * THIN_PPT + it) * blockDim.x + threadIdx.x;
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
}

// ------------------------------------------------------------------ Cout <= 4 ("thin output"), Cin % 4 == 0
#define TT_W 16
#define TT_H 8
#define TT_THREADS 256

__device__ __forceinline__ float4 lrelu4(float4 v) {
    v.x = fsv_act(v.x, FSV_ACT_LRELU); v.y = fsv_act(v.y, FSV_ACT_LRELU);
    v.z = fsv_act(v.z, FSV_ACT_LRELU); v.w = fsv_act(v.w, FSV_ACT_LRELU);
    return v;
}

// forward: grid (x tiles, y tiles, N); weights [co][tap][ci] (= OHWI as stored) in shared memory
template <int COUT>
__global__ void __launch_bounds__(TT_THREADS) k_tco_fwd(ThinP p, int lpp, int tw, int th, const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, const float* __restrict__ residual,
                                                        float* __restrict__ y) {
    extern __shared__ __align__(16) float ws[];
    const int taps = p.kh * p.kw, cin4 = p.Cin >> 2;
    for (int i = threadIdx.x; i < COUT * taps * p.Cin; i += TT_THREADS) ws[i] = w[i];
    __syncthreads();
    const int sub = threadIdx.x & (lpp - 1), slot = threadIdx.x / lpp, slots = TT_THREADS / lpp;
    const long long n = blockIdx.z;
    const int h0 = blockIdx.y * th, w0 = blockIdx.x * tw;
    for (int pidx = slot; pidx < tw * th; pidx += slots) {
        const int ho = h0 + pidx / tw, wo = w0 + (pidx % tw);
        const bool valid = ho < p.Ho && wo < p.Wo;
        float acc[COUT];
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[co] = 0.f;
        if (valid) {
            for (int r = 0; r < p.kh; ++r) {
                const int ih = ho * p.stride + r - p.pad;
                if (ih < 0 || ih >= p.H) continue;
                for (int s = 0; s < p.kw; ++s) {
                    const int iw = wo * p.stride + s - p.pad;
                    if (iw < 0 || iw >= p.W) continue;
                    const float4* xp = reinterpret_cast<const float4*>(x + ((n * p.H + ih) * p.W + iw) * p.x_ld + p.x_coff);
                    const float* wt = ws + (r * p.kw + s) * p.Cin;
                    for (int c4 = sub; c4 < cin4; c4 += lpp) {
                        float4 xv = xp[c4];
                        if (p.in_act == FSV_ACT_LRELU) xv = lrelu4(xv);
#pragma unroll
                        for (int co = 0; co < COUT; ++co) {
                            const float4 wv = *reinterpret_cast<const float4*>(wt + co * taps * p.Cin + c4 * 4);
                            acc[co] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
                        }
                    }
                }
            }
        }
        for (int o = lpp >> 1; o > 0; o >>= 1)
#pragma unroll
            for (int co = 0; co < COUT; ++co) acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], o);
        if (valid && sub == 0) {
            const long long pix = (n * p.Ho + ho) * p.Wo + wo;
            float* yp = y + pix * p.y_ld + p.y_coff;
            const float* rp = residual ? residual + pix * p.res_ld + p.res_coff : nullptr;
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                float t = acc[co];
                if (bias) t += bias[co];
                if (rp) t += rp[co];
                yp[co] = fsv_act(t, p.act) * p.out_scale;
            }
        }
    }
}

// 4x4 kernels (the discriminator heads 512 -> 1, discriminator.py:88): same work split, taps unrolled at compile time.  The
// generic kernel above keeps ONE global load in flight per thread (run-time tap loops with `continue`s), which made the 34x34
// head 72 us and the face discriminator's 10x10 head 0.5 ms of pure load latency; here a lane issues all 16 tap loads of a
// channel quad back to back (predicated on the padding) before the FMAs.
template <int COUT, int KS>
__global__ void __launch_bounds__(TT_THREADS) k_tco_fwd_ks(ThinP p, int lpp, int tw, int th, const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, const float* __restrict__ residual,
                                                           float* __restrict__ y) {
    extern __shared__ __align__(16) float ws[];
    constexpr int taps = KS * KS;
    const int cin4 = p.Cin >> 2;
    for (int i = threadIdx.x; i < COUT * taps * p.Cin; i += TT_THREADS) ws[i] = w[i];
    __syncthreads();
    const int sub = threadIdx.x & (lpp - 1), slot = threadIdx.x / lpp, slots = TT_THREADS / lpp;
    const long long n = blockIdx.z;
    const int h0 = blockIdx.y * th, w0 = blockIdx.x * tw;
    for (int pidx = slot; pidx < tw * th; pidx += slots) {
        const int ho = h0 + pidx / tw, wo = w0 + (pidx % tw);
        const bool valid = ho < p.Ho && wo < p.Wo;
        float acc[COUT];
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[co] = 0.f;
        if (valid) {
            const int ih0 = ho * p.stride - p.pad, iw0 = wo * p.stride - p.pad;
            for (int c4 = sub; c4 < cin4; c4 += lpp) {
                float4 xv[taps];
#pragma unroll
                for (int r = 0; r < KS; ++r)
#pragma unroll
                    for (int s = 0; s < KS; ++s) {
                        const int ih = ih0 + r, iw = iw0 + s;
                        const bool ok = ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
                        xv[r * KS + s] = ok ? reinterpret_cast<const float4*>(x + ((n * p.H + ih) * p.W + iw) * p.x_ld + p.x_coff)[c4]
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                for (int t = 0; t < taps; ++t) {
                    float4 v = xv[t];
                    if (p.in_act == FSV_ACT_LRELU) v = lrelu4(v);
#pragma unroll
                    for (int co = 0; co < COUT; ++co) {
                        const float4 wv = *reinterpret_cast<const float4*>(ws + (co * taps + t) * p.Cin + c4 * 4);
                        acc[co] += v.x * wv.x + v.y * wv.y + v.z * wv.z + v.w * wv.w;
                    }
                }
            }
        }
        for (int o = lpp >> 1; o > 0; o >>= 1)
#pragma unroll
            for (int co = 0; co < COUT; ++co) acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], o);
        if (valid && sub == 0) {
            const long long pix = (n * p.Ho + ho) * p.Wo + wo;
            float* yp = y + pix * p.y_ld + p.y_coff;
            const float* rp = residual ? residual + pix * p.res_ld + p.res_coff : nullptr;
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                float t = acc[co];
                if (bias) t += bias[co];
                if (rp) t += rp[co];
                yp[co] = fsv_act(t, p.act) * p.out_scale;
            }
        }
    }
}

// data gradient: dx[n,h,w,ci] = sum_{r,s,co} dy[n,(h+pad-r)/stride,(w+pad-s)/stride,co] * w[co][r][s][ci]; tiles over (H, W)
template <int COUT>
__global__ void __launch_bounds__(TT_THREADS) k_tco_dgrad(ThinP p, int lpp, int tw, int th, const float* __restrict__ dy, const float* __restrict__ w,
                                                          float* __restrict__ dx, int accumulate) {
    extern __shared__ __align__(16) float ws[];
    const int taps = p.kh * p.kw, cin4 = p.Cin >> 2;
    for (int i = threadIdx.x; i < COUT * taps * p.Cin; i += TT_THREADS) ws[i] = w[i];
    __syncthreads();
    const int sub = threadIdx.x & (lpp - 1), slot = threadIdx.x / lpp, slots = TT_THREADS / lpp;
    const long long n = blockIdx.z;
    const int h0 = blockIdx.y * th, w0 = blockIdx.x * tw;
    for (int pidx = slot; pidx < tw * th; pidx += slots) {
        const int hq = h0 + pidx / tw, wq = w0 + (pidx % tw);
        if (hq >= p.H || wq >= p.W) continue;
        for (int c4 = sub; c4 < cin4; c4 += lpp) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < p.kh; ++r) {
                const int th = hq + p.pad - r;
                if (th < 0 || (th % p.stride) != 0) continue;
                const int oh = th / p.stride;
                if (oh >= p.Ho) continue;
                for (int s = 0; s < p.kw; ++s) {
                    const int tw = wq + p.pad - s;
                    if (tw < 0 || (tw % p.stride) != 0) continue;
                    const int ow = tw / p.stride;
                    if (ow >= p.Wo) continue;
                    const float* dp = dy + ((n * p.Ho + oh) * p.Wo + ow) * p.y_ld + p.y_coff;
                    const float* wt = ws + (r * p.kw + s) * p.Cin + c4 * 4;
#pragma unroll
                    for (int co = 0; co < COUT; ++co) {
                        const float dv = dp[co];
                        const float4 wv = *reinterpret_cast<const float4*>(wt + co * taps * p.Cin);
                        acc.x += dv * wv.x; acc.y += dv * wv.y; acc.z += dv * wv.z; acc.w += dv * wv.w;
                    }
                }
            }
            float4* xp = reinterpret_cast<float4*>(dx + ((n * p.H + hq) * p.W + wq) * p.x_ld + p.x_coff) + c4;
            if (accumulate) {
                float4 o = *xp;
                acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
            }
            *xp = acc;
        }
    }
}

// weight gradient: dw[co][tap][ci] += sum_px dy[px][co] * x[px + tap][ci].  Persistent blocks walk the tiles; every thread
// keeps taps x COUT float4 accumulators for its 4 channels; one warp-shuffle + shared-memory reduction per block, then
// taps*COUT*Cin global atomics per block.  grid.y = channel groups of lpp*4 (only > 1 when Cin > 128).
template <int TAPS, int COUT>
__global__ void __launch_bounds__(TT_THREADS) k_tco_wgrad(ThinP p, int lpp, int tw, int th, int tiles_x, int tiles_y, long long ntiles,
                                                          const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw) {
    __shared__ float red[TAPS * COUT * 128];
    const int taps = p.kh * p.kw, cin4 = p.Cin >> 2;
    const int sub = threadIdx.x & (lpp - 1), slot = threadIdx.x / lpp, slots = TT_THREADS / lpp;
    const int c4 = blockIdx.y * lpp + sub;
    const bool cvalid = c4 < cin4;
    float4 acc[TAPS][COUT];
#pragma unroll
    for (int t = 0; t < TAPS; ++t)
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[t][co] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tx = (int)(tile % tiles_x);
        const long long q = tile / tiles_x;
        const int ty = (int)(q % tiles_y);
        const long long n = q / tiles_y;
        for (int pidx = slot; pidx < tw * th; pidx += slots) {
            const int ho = ty * th + pidx / tw, wo = tx * tw + (pidx % tw);
            if (!cvalid || ho >= p.Ho || wo >= p.Wo) continue;
            const float* dp = dy + ((n * p.Ho + ho) * p.Wo + wo) * p.y_ld + p.y_coff;
            float dv[COUT];
#pragma unroll
            for (int co = 0; co < COUT; ++co) dv[co] = dp[co];
            int r = 0, s = 0;
#pragma unroll
            for (int t = 0; t < TAPS; ++t) {
                if (t < taps) {
                    const int ih = ho * p.stride + r - p.pad, iw = wo * p.stride + s - p.pad;
                    if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
                        float4 xv = reinterpret_cast<const float4*>(x + ((n * p.H + ih) * p.W + iw) * p.x_ld + p.x_coff)[c4];
                        if (p.in_act == FSV_ACT_LRELU) xv = lrelu4(xv);
#pragma unroll
                        for (int co = 0; co < COUT; ++co) {
                            acc[t][co].x += dv[co] * xv.x; acc[t][co].y += dv[co] * xv.y;
                            acc[t][co].z += dv[co] * xv.z; acc[t][co].w += dv[co] * xv.w;
                        }
                    }
                    if (++s == p.kw) { s = 0; ++r; }
                }
            }
        }
    }
    // lanes with the same sub (same channels) inside a warp
    for (int o = lpp; o < 32; o <<= 1) {
#pragma unroll
        for (int t = 0; t < TAPS; ++t)
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                acc[t][co].x += __shfl_xor_sync(0xffffffffu, acc[t][co].x, o);
                acc[t][co].y += __shfl_xor_sync(0xffffffffu, acc[t][co].y, o);
                acc[t][co].z += __shfl_xor_sync(0xffffffffu, acc[t][co].z, o);
                acc[t][co].w += __shfl_xor_sync(0xffffffffu, acc[t][co].w, o);
            }
    }
    const int row = lpp * 4;                      // channels this block covers
    for (int i = threadIdx.x; i < TAPS * COUT * row; i += TT_THREADS) red[i] = 0.f;
    __syncthreads();
    if ((threadIdx.x & 31) < lpp && cvalid) {
#pragma unroll
        for (int t = 0; t < TAPS; ++t)
            if (t < taps)
#pragma unroll
                for (int co = 0; co < COUT; ++co) {
                    float* rp = red + (t * COUT + co) * row + sub * 4;
                    atomicAdd(rp + 0, acc[t][co].x); atomicAdd(rp + 1, acc[t][co].y);
                    atomicAdd(rp + 2, acc[t][co].z); atomicAdd(rp + 3, acc[t][co].w);
                }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < taps * COUT * row; i += TT_THREADS) {
        const int cc = i % row, tc = i / row;
        const int co = tc % COUT, t = tc / COUT;
        const int ci = blockIdx.y * row + cc;
        if (ci < p.Cin) atomicAdd(dw + ((long long)co * taps + t) * p.Cin + ci, red[i]);
    }
}

// ------------------------------------------------------------------ Cout <= 4, 3x3 / stride 1 / pad 1 fast paths
// The generic kernels above are instruction-issue bound (ncu: 85 % issue-active, ~950 instructions per (pixel, lane):
// run-time tap loops, 64-bit addressing and divisions per tap).  All full-resolution thin-output layers are 3x3 s1 p1
// (conv_img / conv_flow / conv_mask), so these variants unroll the nine taps at compile time, form the pixel address
// once and reach the taps through 32-bit offsets; the weight gradient additionally spreads the three tap rows over
// blockIdx.y, which cuts its accumulators from 27 to 9 float4 per output (145 -> ~64 registers, 4x the occupancy).
template <int COUT, bool INACT>
__global__ void __launch_bounds__(TT_THREADS) k_tco_fwd_k3(ThinP p, int lpp, const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, const float* __restrict__ residual,
                                                           float* __restrict__ y) {
    extern __shared__ __align__(16) float ws[];   // [co][9][ci]
    const int cin4 = p.Cin >> 2;
    for (int i = threadIdx.x; i < COUT * 9 * p.Cin; i += TT_THREADS) ws[i] = w[i];
    __syncthreads();
    const int sub = threadIdx.x & (lpp - 1), slot = threadIdx.x / lpp, slots = TT_THREADS / lpp;
    const long long n = blockIdx.z;
    const int h0 = blockIdx.y * TT_H, w0 = blockIdx.x * TT_W;
    const int rowp = p.W * p.x_ld;
    for (int pidx = slot; pidx < TT_W * TT_H; pidx += slots) {
        const int ho = h0 + (pidx >> 4), wo = w0 + (pidx & 15);
        const bool valid = ho < p.Ho && wo < p.Wo;
        float acc[COUT];
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[co] = 0.f;
        if (valid) {
            const float* xb = x + ((n * p.H + ho) * p.W + wo) * p.x_ld + p.x_coff;
            const bool vr0 = ho > 0, vr2 = ho + 1 < p.H, vc0 = wo > 0, vc2 = wo + 1 < p.W;
            for (int c4 = sub; c4 < cin4; c4 += lpp) {
                const float* xc = xb + c4 * 4;
                const float* wc = ws + c4 * 4;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const bool vr = r == 0 ? vr0 : (r == 2 ? vr2 : true);
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        const bool vc = s == 0 ? vc0 : (s == 2 ? vc2 : true);
                        if (vr && vc) {
                            float4 xv = *reinterpret_cast<const float4*>(xc + (r - 1) * rowp + (s - 1) * p.x_ld);
                            if (INACT) xv = lrelu4(xv);
#pragma unroll
                            for (int co = 0; co < COUT; ++co) {
                                const float4 wv = *reinterpret_cast<const float4*>(wc + (co * 9 + r * 3 + s) * p.Cin);
                                acc[co] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
                            }
                        }
                    }
                }
            }
        }
        for (int o = lpp >> 1; o > 0; o >>= 1)
#pragma unroll
            for (int co = 0; co < COUT; ++co) acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], o);
        if (valid && sub == 0) {
            const long long pix = (n * p.Ho + ho) * p.Wo + wo;
            float* yp = y + pix * p.y_ld + p.y_coff;
            const float* rp = residual ? residual + pix * p.res_ld + p.res_coff : nullptr;
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                float t = acc[co];
                if (bias) t += bias[co];
                if (rp) t += rp[co];
                yp[co] = fsv_act(t, p.act) * p.out_scale;
            }
        }
    }
}

template <int COUT>
__global__ void __launch_bounds__(TT_THREADS) k_tco_dgrad_k3(ThinP p, int lpp, const float* __restrict__ dy, const float* __restrict__ w,
                                                             float* __restrict__ dx, int accumulate) {
    extern __shared__ __align__(16) float ws[];
    const int cin4 = p.Cin >> 2;
    for (int i = threadIdx.x; i < COUT * 9 * p.Cin; i += TT_THREADS) ws[i] = w[i];
    __syncthreads();
    const int sub = threadIdx.x & (lpp - 1), slot = threadIdx.x / lpp, slots = TT_THREADS / lpp;
    const long long n = blockIdx.z;
    const int h0 = blockIdx.y * TT_H, w0 = blockIdx.x * TT_W;
    const int rowp = p.Wo * p.y_ld;
    for (int pidx = slot; pidx < TT_W * TT_H; pidx += slots) {
        const int hq = h0 + (pidx >> 4), wq = w0 + (pidx & 15);
        if (hq >= p.H || wq >= p.W) continue;
        // tap (r, s) reads dy at (hq + 1 - r, wq + 1 - s)
        const float* db = dy + ((n * p.Ho + hq) * p.Wo + wq) * p.y_ld + p.y_coff;
        const bool vr0 = hq + 1 < p.Ho, vr2 = hq > 0, vc0 = wq + 1 < p.Wo, vc2 = wq > 0;
        float dv[9][COUT];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const bool vr = r == 0 ? vr0 : (r == 2 ? vr2 : true);
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const bool vc = s == 0 ? vc0 : (s == 2 ? vc2 : true);
#pragma unroll
                for (int co = 0; co < COUT; ++co) dv[r * 3 + s][co] = (vr && vc) ? db[(1 - r) * rowp + (1 - s) * p.y_ld + co] : 0.f;
            }
        }
        for (int c4 = sub; c4 < cin4; c4 += lpp) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* wc = ws + c4 * 4;
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
                for (int co = 0; co < COUT; ++co) {
                    const float4 wv = *reinterpret_cast<const float4*>(wc + (co * 9 + t) * p.Cin);
                    acc.x += dv[t][co] * wv.x; acc.y += dv[t][co] * wv.y; acc.z += dv[t][co] * wv.z; acc.w += dv[t][co] * wv.w;
                }
            float4* xp = reinterpret_cast<float4*>(dx + ((n * p.H + hq) * p.W + wq) * p.x_ld + p.x_coff) + c4;
            if (accumulate) {
                const float4 o = *xp;
                acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
            }
            *xp = acc;
        }
    }
}

// grid (persistent tile walkers, 3 tap rows x channel groups)
template <int COUT, bool INACT>
__global__ void __launch_bounds__(TT_THREADS) k_tco_wgrad_k3(ThinP p, int lpp, int tiles_x, int tiles_y, long long ntiles,
                                                             const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw) {
    __shared__ float red[3 * COUT * 128];
    const int cin4 = p.Cin >> 2;
    const int sub = threadIdx.x & (lpp - 1), slot = threadIdx.x / lpp, slots = TT_THREADS / lpp;
    const int tr = blockIdx.y % 3, cg = blockIdx.y / 3;
    const int c4 = cg * lpp + sub;
    const bool cvalid = c4 < cin4;
    float4 acc[3][COUT];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[s][co] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tx = (int)(tile % tiles_x);
        const long long q = tile / tiles_x;
        const int ty = (int)(q % tiles_y);
        const long long n = q / tiles_y;
        for (int pidx = slot; pidx < TT_W * TT_H; pidx += slots) {
            const int ho = ty * TT_H + (pidx >> 4), wo = tx * TT_W + (pidx & 15);
            const int ih = ho + tr - 1;
            if (!cvalid || ho >= p.Ho || wo >= p.Wo || ih < 0 || ih >= p.H) continue;
            const float* dp = dy + ((n * p.Ho + ho) * p.Wo + wo) * p.y_ld + p.y_coff;
            float dv[COUT];
#pragma unroll
            for (int co = 0; co < COUT; ++co) dv[co] = dp[co];
            const float* xr = x + ((n * p.H + ih) * p.W + wo) * p.x_ld + p.x_coff + c4 * 4;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int iw = wo + s - 1;
                if (iw >= 0 && iw < p.W) {
                    float4 xv = *reinterpret_cast<const float4*>(xr + (s - 1) * p.x_ld);
                    if (INACT) xv = lrelu4(xv);
#pragma unroll
                    for (int co = 0; co < COUT; ++co) {
                        acc[s][co].x += dv[co] * xv.x; acc[s][co].y += dv[co] * xv.y;
                        acc[s][co].z += dv[co] * xv.z; acc[s][co].w += dv[co] * xv.w;
                    }
                }
            }
        }
    }
    for (int o = lpp; o < 32; o <<= 1) {
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                acc[s][co].x += __shfl_xor_sync(0xffffffffu, acc[s][co].x, o);
                acc[s][co].y += __shfl_xor_sync(0xffffffffu, acc[s][co].y, o);
                acc[s][co].z += __shfl_xor_sync(0xffffffffu, acc[s][co].z, o);
                acc[s][co].w += __shfl_xor_sync(0xffffffffu, acc[s][co].w, o);
            }
    }
    const int row = lpp * 4;
    for (int i = threadIdx.x; i < 3 * COUT * row; i += TT_THREADS) red[i] = 0.f;
    __syncthreads();
    if ((threadIdx.x & 31) < lpp && cvalid) {
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                float* rp = red + (s * COUT + co) * row + sub * 4;
                atomicAdd(rp + 0, acc[s][co].x); atomicAdd(rp + 1, acc[s][co].y);
                atomicAdd(rp + 2, acc[s][co].z); atomicAdd(rp + 3, acc[s][co].w);
            }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * COUT * row; i += TT_THREADS) {
        const int cc = i % row, sc = i / row;
        const int co = sc % COUT, s = sc / COUT;
        const int ci = cg * row + cc;
        if (ci < p.Cin) atomicAdd(dw + ((long long)co * 9 + tr * 3 + s) * p.Cin + ci, red[i]);
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
    return d->up == 1 && d->w_nstride == 0 && d->b_nstride == 0 && d->kh * d->kw <= 16;
}
// Thin-output layers with few pixels but many input channels (the face discriminator's 512 -> 1 head on 10x10: 484 outputs of 8192
// MACs each) are far worse off on the generic 64x64 SIMT tile than the large ones: 8 CTAs walk the whole reduction serially (0.5 - 0.75 ms
// per call in the round-2 timeline; -0.9 ms per pose512 step with the thin-output kernels).  FSV_THIN_OUT_MIN_PX = their pixel threshold
// for layers with >= 64 input channels (default 64; 4096 restores the round-1 dispatch).
static long long thin_out_min_px() {
    static long long v = -1;
    if (v < 0) { const char* e = getenv("FSV_THIN_OUT_MIN_PX"); v = e ? atoll(e) : 64; if (v < 1) v = 1; }
    return v;
}
extern "C" int fsv_conv2d_thin_kind(const fsv_conv_desc* d) {
    if (!d || !thin_common_ok(d)) return 0;
    const long long px = (long long)d->N * d->Ho * d->Wo;
    if (d->Cin <= 8 && d->Cout >= 16 && px >= 4096) return 1;                                         // thin input
    if (d->Cout <= 4 && d->Cin >= 16 && d->Cin % 4 == 0 && d->x_ld % 4 == 0 && d->x_coff % 4 == 0 && d->N <= 65535 &&
        (long long)d->Cout * d->kh * d->kw * d->Cin * 4 <= 40 * 1024 &&
        (px >= 4096 || (px >= thin_out_min_px() && d->Cin >= 64))) return 2;         // thin output
    return 0;
}
// tile: 16x8 pixels, shrunk (16x4, 8x4) while the grid would not fill the GPU twice over; always a multiple of the
// block's pixel slots so every lane of a warp runs the same number of passes
static void tt_tile(int lpp, long long n, int hh, int ww, int* tw, int* th) {
    const int slots = TT_THREADS / lpp;
    *tw = TT_W; *th = TT_H;
    const long long want = 2LL * fsv_sm_count();
    while (n * fsv_cdiv(ww, *tw) * fsv_cdiv(hh, *th) < want) {
        if (*th > 4 && (*tw) * (*th / 2) >= slots) *th /= 2;
        else if (*tw > 8 && (*tw / 2) * (*th) >= slots) *tw /= 2;
        else break;
    }
}
// 3x3 / stride 1 / pad 1 on full 16x8 tiles with 32-bit tap offsets: the unrolled fast paths apply
static bool tco_k3(const fsv_conv_desc* d, int tw, int th) {
    return d->kh == 3 && d->kw == 3 && d->stride == 1 && d->pad == 1 && tw == TT_W && th == TT_H &&
           (long long)d->W * d->x_ld < (1 << 28) && (long long)d->Wo * d->y_ld < (1 << 28) &&
           (d->in_act == FSV_ACT_NONE || d->in_act == FSV_ACT_LRELU);
}
static int tt_lpp(int cin) {
    int l = 1;
    while (l * 2 <= cin / 4 && l < 32) l *= 2;
    return l;
}

extern "C" int fsv_conv2d_fwd_thin(const fsv_conv_desc* d, const float* x, const float* w, const float* bias,
                                   const float* residual, float* y, void* stream) {
    int kind = fsv_conv2d_thin_kind(d);
    FSV_REQUIRE(kind != 0, "conv2d_fwd_thin: not a thin layer");
    ThinP p = thin_p(d);
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)d->N * d->Ho * d->Wo;
    const int taps = d->kh * d->kw;
    if (kind == 1) {
        dim3 grid(fsv_cdiv(total, 128 * THIN_PPT), fsv_cdiv(d->Cout, 32));
        k_thin_cin_fwd<<<grid, 128, taps * d->Cin * 32 * sizeof(float), st>>>(p, x, w, bias, residual, y);
    } else {
        FSV_REQUIRE((((uintptr_t)x) & 15) == 0 && (((uintptr_t)w) & 15) == 0, "conv2d_fwd_thin: pointers must be 16-byte aligned");
        size_t sm = (size_t)d->Cout * taps * d->Cin * sizeof(float);
        int lpp = tt_lpp(d->Cin), tw, th;
        tt_tile(lpp, d->N, d->Ho, d->Wo, &tw, &th);
        dim3 grid(fsv_cdiv(d->Wo, tw), fsv_cdiv(d->Ho, th), d->N);
        if (tco_k3(d, tw, th)) {
#define TCO_FWD3(C) do { if (d->in_act == FSV_ACT_LRELU) k_tco_fwd_k3<C, true><<<grid, TT_THREADS, sm, st>>>(p, lpp, x, w, bias, residual, y); \
                         else k_tco_fwd_k3<C, false><<<grid, TT_THREADS, sm, st>>>(p, lpp, x, w, bias, residual, y); } while (0)
            switch (d->Cout) {
                case 1: TCO_FWD3(1); break;
                case 2: TCO_FWD3(2); break;
                case 3: TCO_FWD3(3); break;
                default: TCO_FWD3(4); break;
            }
#undef TCO_FWD3
            FSV_CHECK_LAUNCH("conv2d_fwd_thin_k3");
            return FSV_OK;
        }
        if (d->kh == 4 && d->kw == 4 && d->Cout == 1) {
            k_tco_fwd_ks<1, 4><<<grid, TT_THREADS, sm, st>>>(p, lpp, tw, th, x, w, bias, residual, y);
            FSV_CHECK_LAUNCH("conv2d_fwd_thin_k4");
            return FSV_OK;
        }
        switch (d->Cout) {
            case 1: k_tco_fwd<1><<<grid, TT_THREADS, sm, st>>>(p, lpp, tw, th, x, w, bias, residual, y); break;
            case 2: k_tco_fwd<2><<<grid, TT_THREADS, sm, st>>>(p, lpp, tw, th, x, w, bias, residual, y); break;
            case 3: k_tco_fwd<3><<<grid, TT_THREADS, sm, st>>>(p, lpp, tw, th, x, w, bias, residual, y); break;
            default: k_tco_fwd<4><<<grid, TT_THREADS, sm, st>>>(p, lpp, tw, th, x, w, bias, residual, y); break;
        }
    }
    FSV_CHECK_LAUNCH("conv2d_fwd_thin");
    return FSV_OK;
}

// thin-output weight gradient (dw only; the bias gradient stays with the column-sum kernel of conv_simt.cu); dw must be
// zeroed by the caller.  Returns FSV_OK with *handled = 0 when the (taps, Cout) pair has no instantiation.
extern "C" int fsv_conv2d_wgrad_thin(const fsv_conv_desc* d, const float* x, const float* dy, float* dw, int* handled, void* stream) {
    *handled = 0;
    if (fsv_conv2d_thin_kind(d) != 2 || (((uintptr_t)x) & 15) != 0) return FSV_OK;
    const int taps = d->kh * d->kw;
    if (taps > 9 && d->Cout > 2) return FSV_OK;
    ThinP p = thin_p(d);
    cudaStream_t st = (cudaStream_t)stream;
    const int lpp = tt_lpp(d->Cin);
    int tw, th;
    tt_tile(lpp, d->N, d->Ho, d->Wo, &tw, &th);
    const int tiles_x = fsv_cdiv(d->Wo, tw), tiles_y = fsv_cdiv(d->Ho, th);
    const long long ntiles = (long long)tiles_x * tiles_y * d->N;
    const int groups = fsv_cdiv(d->Cin / 4, lpp);
    long long gx = (2LL * fsv_sm_count() + groups - 1) / groups;
    if (gx > ntiles) gx = ntiles;
    if (tco_k3(d, tw, th)) {
        long long g3 = (4LL * fsv_sm_count() + 3 * groups - 1) / (3 * groups);
        if (g3 > ntiles) g3 = ntiles;
        dim3 grid3((unsigned)g3, 3 * groups);
#define TCO_WG3(C) do { if (d->in_act == FSV_ACT_LRELU) k_tco_wgrad_k3<C, true><<<grid3, TT_THREADS, 0, st>>>(p, lpp, tiles_x, tiles_y, ntiles, x, dy, dw); \
                        else k_tco_wgrad_k3<C, false><<<grid3, TT_THREADS, 0, st>>>(p, lpp, tiles_x, tiles_y, ntiles, x, dy, dw); } while (0)
        switch (d->Cout) {
            case 1: TCO_WG3(1); break;
            case 2: TCO_WG3(2); break;
            case 3: TCO_WG3(3); break;
            default: TCO_WG3(4); break;
        }
#undef TCO_WG3
        FSV_CHECK_LAUNCH("conv2d_wgrad_thin_k3");
        *handled = 1;
        return FSV_OK;
    }
    dim3 grid((unsigned)gx, groups);
#define TCO_WGRAD(T, C) k_tco_wgrad<T, C><<<grid, TT_THREADS, 0, st>>>(p, lpp, tw, th, tiles_x, tiles_y, ntiles, x, dy, dw)
    if (taps <= 9) {
        switch (d->Cout) {
            case 1: TCO_WGRAD(9, 1); break;
            case 2: TCO_WGRAD(9, 2); break;
            case 3: TCO_WGRAD(9, 3); break;
            default: TCO_WGRAD(9, 4); break;
        }
    } else {
        if (d->Cout == 1) TCO_WGRAD(16, 1); else TCO_WGRAD(16, 2);
    }
#undef TCO_WGRAD
    FSV_CHECK_LAUNCH("conv2d_wgrad_thin");
    *handled = 1;
    return FSV_OK;
}

// kind 1: thin dx (Cin <= 8); kind 2: thin dy (Cout <= 4)
extern "C" int fsv_conv2d_dgrad_thin_ok(const fsv_conv_desc* d) {
    if (!d) return 0;
    if (d->up == 1 && d->w_nstride == 0 && d->Cin <= 8 && d->Cout >= 16 && d->Cout % 4 == 0 && d->y_ld % 4 == 0 &&
        d->y_coff % 4 == 0 && d->kh * d->kw <= 16 && (long long)d->kh * d->kw * d->Cout * 8 * 4 <= 40 * 1024 &&
        (long long)d->N * d->H * d->W >= 4096) return 1;
    if (fsv_conv2d_thin_kind(d) == 2) return 2;
    return 0;
}
extern "C" int fsv_conv2d_dgrad_thin(const fsv_conv_desc* d, const float* dy, const float* w, float* dx, int accumulate, void* stream) {
    int kind = fsv_conv2d_dgrad_thin_ok(d);
    FSV_REQUIRE(kind != 0, "conv2d_dgrad_thin: not eligible");
    ThinP p = thin_p(d);
    cudaStream_t st = (cudaStream_t)stream;
    if (kind == 1) {
        FSV_REQUIRE((((uintptr_t)dy) & 15) == 0, "conv2d_dgrad_thin: dy must be 16-byte aligned");
        const long long total = (long long)d->N * d->H * d->W;
        k_thin_cin_dgrad<<<fsv_cdiv(total, 128), 128, (size_t)d->kh * d->kw * d->Cout * 8 * sizeof(float), st>>>(p, dy, w, dx, accumulate);
    } else {
        FSV_REQUIRE((((uintptr_t)dx) & 15) == 0 && (((uintptr_t)w) & 15) == 0, "conv2d_dgrad_thin: pointers must be 16-byte aligned");
        size_t sm = (size_t)d->Cout * d->kh * d->kw * d->Cin * sizeof(float);
        int lpp = tt_lpp(d->Cin), tw, th;
        tt_tile(lpp, d->N, d->H, d->W, &tw, &th);
        dim3 grid(fsv_cdiv(d->W, tw), fsv_cdiv(d->H, th), d->N);
        if (tco_k3(d, tw, th)) {
            switch (d->Cout) {
                case 1: k_tco_dgrad_k3<1><<<grid, TT_THREADS, sm, st>>>(p, lpp, dy, w, dx, accumulate); break;
                case 2: k_tco_dgrad_k3<2><<<grid, TT_THREADS, sm, st>>>(p, lpp, dy, w, dx, accumulate); break;
                case 3: k_tco_dgrad_k3<3><<<grid, TT_THREADS, sm, st>>>(p, lpp, dy, w, dx, accumulate); break;
                default: k_tco_dgrad_k3<4><<<grid, TT_THREADS, sm, st>>>(p, lpp, dy, w, dx, accumulate); break;
            }
            FSV_CHECK_LAUNCH("conv2d_dgrad_thin_k3");
            return FSV_OK;
        }
        switch (d->Cout) {
            case 1: k_tco_dgrad<1><<<grid, TT_THREADS, sm, st>>>(p, lpp, tw, th, dy, w, dx, accumulate); break;
            case 2: k_tco_dgrad<2><<<grid, TT_THREADS, sm, st>>>(p, lpp, tw, th, dy, w, dx, accumulate); break;
            case 3: k_tco_dgrad<3><<<grid, TT_THREADS, sm, st>>>(p, lpp, tw, th, dy, w, dx, accumulate); break;
            default: k_tco_dgrad<4><<<grid, TT_THREADS, sm, st>>>(p, lpp, tw, th, dy, w, dx, accumulate); break;
        }
    }
    FSV_CHECK_LAUNCH("conv2d_dgrad_thin");
    return FSV_OK;
}
