// Layout plumbing kernels: NCHW<->NHWC packing at the module boundary, channel-slice
// copies (concat), nearest x2 upsample, 3x3/s2 average pooling, activation backward.
// All are pure streaming kernels (HBM-bound, no reuse): coalesced along the contiguous
// dimension, grid-stride loops sized to a multiple of the SM count.
#include "common.cuh"
#include <stdlib.h>

static inline int stream_grid(long long work_items, int threads) {
    long long blocks = (work_items + threads - 1) / threads;
    long long cap = (long long)fsv_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// one thread per (n, hw): reads are coalesced along hw for every c, writes are C-contiguous per thread
__global__ void k_nchw_to_nhwc(const float* __restrict__ src, float* __restrict__ dst, int N, int C, long long HW,
                               int ld, int coff) {
    long long total = (long long)N * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long n = i / HW, p = i - n * HW;
        const float* s = src + n * C * HW + p;
        float* d = dst + i * ld + coff;
        for (int c = 0; c < C; ++c) d[c] = s[(long long)c * HW];
    }
}
__global__ void k_nhwc_to_nchw(const float* __restrict__ src, float* __restrict__ dst, int N, int C, long long HW,
                               int ld, int coff, int accumulate) {
    long long total = (long long)N * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long n = i / HW, p = i - n * HW;
        const float* s = src + i * ld + coff;
        float* d = dst + n * C * HW + p;
        for (int c = 0; c < C; ++c) {
            float v = s[c];
            if (accumulate) d[(long long)c * HW] += v; else d[(long long)c * HW] = v;
        }
    }
}
// Tiled form: a block moves PK_PX pixels x up to 32 channels through shared memory, so that the plane reads (128 consecutive pixels of one
// channel) AND the pixel-major writes (the channels of one pixel, consecutive pixels back to back when the destination is dense) are both
// contiguous.  The one-thread-per-pixel kernel above scatters every store instruction of a warp over 32 destination rows (round-2 timeline:
// 0.46 ms for the discriminator's 20-channel 4 x 512 x 512 input, 0.37 TB/s).  grid (pixel tiles, N).
#define PK_PX 128
__global__ void __launch_bounds__(256) k_nchw_to_nhwc_tiled(const float* __restrict__ src, float* __restrict__ dst, int C, int HW, int ld, int coff) {
    __shared__ float tile[32][PK_PX + 1];          // [channel][pixel]
    const int tid = threadIdx.x;
    const int p0 = blockIdx.x * PK_PX;
    const int np = min(PK_PX, HW - p0);
    const float* s = src + (long long)blockIdx.y * C * HW + p0;
    float* d = dst + ((long long)blockIdx.y * HW + p0) * ld + coff;
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int nc = min(32, C - c0);
        for (int i = tid; i < nc * PK_PX; i += 256) {
            const int c = i / PK_PX, px = i - c * PK_PX;
            if (px < np) tile[c][px] = s[(long long)(c0 + c) * HW + px];
        }
        __syncthreads();
        for (int i = tid; i < np * nc; i += 256) {
            const int px = i / nc, c = i - px * nc;
            d[(long long)px * ld + c0 + c] = tile[c][px];
        }
        __syncthreads();
    }
}
extern "C" int fsv_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, int dst_ld, int dst_coff, void* stream) {
    FSV_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && dst_ld >= dst_coff + C, "nchw_to_nhwc: bad dims");
    long long HW = (long long)H * W;
    static int tiled = -1;                          // FSV_PACK_TILED=0: the one-thread-per-pixel kernel
    if (tiled < 0) { const char* e = getenv("FSV_PACK_TILED"); tiled = (e && atoi(e) == 0) ? 0 : 1; }
    if (tiled && N <= 65535 && HW < (1LL << 31) - PK_PX && HW >= 1024) {
        dim3 grid((unsigned)((HW + PK_PX - 1) / PK_PX), N);
        k_nchw_to_nhwc_tiled<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, C, (int)HW, dst_ld, dst_coff);
    } else {
        k_nchw_to_nhwc<<<stream_grid(N * HW, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, N, C, HW, dst_ld, dst_coff);
    }
    FSV_CHECK_LAUNCH("nchw_to_nhwc");
    return FSV_OK;
}
extern "C" int fsv_nhwc_to_nchw(const float* src, float* dst, int N, int C, int H, int W, int src_ld, int src_coff,
                                int accumulate, void* stream) {
    FSV_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && src_ld >= src_coff + C, "nhwc_to_nchw: bad dims");
    long long HW = (long long)H * W;
    k_nhwc_to_nchw<<<stream_grid(N * HW, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, N, C, HW, src_ld, src_coff, accumulate);
    FSV_CHECK_LAUNCH("nhwc_to_nchw");
    return FSV_OK;
}

__global__ void k_copy_channels(const float* __restrict__ src, int sld, int scoff, float* __restrict__ dst, int dld, int dcoff,
                                long long rows, int C, int accumulate) {
    long long total = rows * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / C;
        int c = (int)(i - r * C);
        float v = src[r * sld + scoff + c];
        float* d = dst + r * dld + dcoff + c;
        if (accumulate) *d += v; else *d = v;
    }
}
extern "C" int fsv_copy_channels(const float* src, int src_ld, int src_coff, float* dst, int dst_ld, int dst_coff,
                                 long long rows, int C, int accumulate, void* stream) {
    FSV_REQUIRE(rows > 0 && C > 0 && src_ld >= src_coff + C && dst_ld >= dst_coff + C, "copy_channels: bad dims");
    k_copy_channels<<<stream_grid(rows * C, 256), 256, 0, (cudaStream_t)stream>>>(src, src_ld, src_coff, dst, dst_ld, dst_coff, rows, C, accumulate);
    FSV_CHECK_LAUNCH("copy_channels");
    return FSV_OK;
}

// y (N,2Hs,2Ws,C) <- x (N,Hs,Ws,C): one thread per output element (C contiguous => coalesced both sides)
__global__ void k_up2_fwd(const float* __restrict__ x, float* __restrict__ y, int N, int Hs, int Ws, int C) {
    long long total = (long long)N * Hs * 2 * Ws * 2 * C;
    int W = Ws * 2, H = Hs * 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long p = i / C;
        int w = (int)(p % W);
        long long q = p / W;
        int h = (int)(q % H);
        long long n = q / H;
        y[i] = x[((n * Hs + (h >> 1)) * Ws + (w >> 1)) * C + c];
    }
}
__global__ void k_up2_bwd(const float* __restrict__ dy, float* __restrict__ dx, int N, int Hs, int Ws, int C) {
    long long total = (long long)N * Hs * Ws * C;
    int W = Ws * 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long p = i / C;
        int w = (int)(p % Ws);
        long long q = p / Ws;
        int h = (int)(q % Hs);
        long long n = q / Hs;
        const float* b = dy + (((n * Hs * 2 + 2 * h) * W) + 2 * w) * C + c;
        dx[i] = b[0] + b[C] + b[(long long)W * C] + b[(long long)W * C + C];
    }
}
// float4 / 32-bit-index form (C % 4 == 0, fewer than 2^31 output float4s): one thread per output float4; the generic kernel above spends
// its time in three 64-bit divisions per element (round-2 timeline: 80 - 100 us for a 67 MB output, 5x the HBM time)
__global__ void k_up2_fwd4(const float4* __restrict__ x, float4* __restrict__ y, unsigned total4, unsigned Hs, unsigned Ws, unsigned C4) {
    const unsigned W = Ws * 2, H = Hs * 2;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
        const unsigned c = i % C4, p = i / C4;
        const unsigned w = p % W, q = p / W;
        const unsigned h = q % H, n = q / H;
        y[i] = x[((size_t)(n * Hs + (h >> 1)) * Ws + (w >> 1)) * C4 + c];
    }
}
extern "C" int fsv_upsample2x_fwd(const float* x, float* y, int N, int Hs, int Ws, int C, void* stream) {
    FSV_REQUIRE(N > 0 && Hs > 0 && Ws > 0 && C > 0, "upsample2x: bad dims");
    const long long total4 = (long long)N * Hs * Ws * C;           // = output elements / 4
    if (C % 4 == 0 && total4 < (1LL << 31) - (1 << 22) && ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0) {
        k_up2_fwd4<<<stream_grid(total4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (float4*)y, (unsigned)total4, Hs, Ws, C / 4);
        FSV_CHECK_LAUNCH("upsample2x_fwd");
        return FSV_OK;
    }
    k_up2_fwd<<<stream_grid((long long)N * Hs * Ws * 4 * C, 256), 256, 0, (cudaStream_t)stream>>>(x, y, N, Hs, Ws, C);
    FSV_CHECK_LAUNCH("upsample2x_fwd");
    return FSV_OK;
}
extern "C" int fsv_upsample2x_bwd(const float* dy, float* dx, int N, int Hs, int Ws, int C, void* stream) {
    FSV_REQUIRE(N > 0 && Hs > 0 && Ws > 0 && C > 0, "upsample2x: bad dims");
    k_up2_bwd<<<stream_grid((long long)N * Hs * Ws * C, 256), 256, 0, (cudaStream_t)stream>>>(dy, dx, N, Hs, Ws, C);
    FSV_CHECK_LAUNCH("upsample2x_bwd");
    return FSV_OK;
}

// AvgPool2d(3, s2, p1, count_include_pad=False): Ho = (H+2-3)/2+1
__global__ void k_avgpool_fwd(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C, int Ho, int Wo) {
    long long total = (long long)N * Ho * Wo * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long p = i / C;
        int wo = (int)(p % Wo);
        long long q = p / Wo;
        int ho = (int)(q % Ho);
        long long n = q / Ho;
        float s = 0.f;
        int cnt = 0;
        for (int r = -1; r <= 1; ++r) {
            int h = 2 * ho + r;
            if (h < 0 || h >= H) continue;
            for (int t = -1; t <= 1; ++t) {
                int w = 2 * wo + t;
                if (w < 0 || w >= W) continue;
                s += x[((n * H + h) * W + w) * C + c];
                ++cnt;
            }
        }
        y[i] = s / (float)cnt;
    }
}
// gather form of the adjoint: dx[h,w] = sum over outputs whose window covers (h,w) of dy/cnt(out)
__global__ void k_avgpool_bwd(const float* __restrict__ dy, float* __restrict__ dx, int N, int H, int W, int C, int Ho, int Wo) {
    long long total = (long long)N * H * W * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long p = i / C;
        int w = (int)(p % W);
        long long q = p / W;
        int h = (int)(q % H);
        long long n = q / H;
        float s = 0.f;
        for (int ho = (h) / 2; ho <= (h + 1) / 2; ++ho) {     // 2*ho-1 <= h <= 2*ho+1
            if (ho < 0 || ho >= Ho) continue;
            int ch = min(2 * ho + 1, H - 1) - max(2 * ho - 1, 0) + 1;
            for (int wo = (w) / 2; wo <= (w + 1) / 2; ++wo) {
                if (wo < 0 || wo >= Wo) continue;
                int cw = min(2 * wo + 1, W - 1) - max(2 * wo - 1, 0) + 1;
                s += dy[((n * Ho + ho) * Wo + wo) * C + c] / (float)(ch * cw);
            }
        }
        dx[i] = s;
    }
}
extern "C" int fsv_avgpool3s2_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream) {
    FSV_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, "avgpool: bad dims");
    int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    k_avgpool_fwd<<<stream_grid((long long)N * Ho * Wo * C, 256), 256, 0, (cudaStream_t)stream>>>(x, y, N, H, W, C, Ho, Wo);
    FSV_CHECK_LAUNCH("avgpool_fwd");
    return FSV_OK;
}
extern "C" int fsv_avgpool3s2_bwd(const float* dy, float* dx, int N, int H, int W, int C, void* stream) {
    FSV_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, "avgpool: bad dims");
    int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    k_avgpool_bwd<<<stream_grid((long long)N * H * W * C, 256), 256, 0, (cudaStream_t)stream>>>(dy, dx, N, H, W, C, Ho, Wo);
    FSV_CHECK_LAUNCH("avgpool_bwd");
    return FSV_OK;
}

__global__ void k_act_bwd(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ g, long long n,
                          int act, float out_scale) {
    float inv = 1.f / out_scale;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float yy = y[i] * inv;   // value before out_scale
        g[i] = dy[i] * out_scale * fsv_act_grad(yy, act);
    }
}
extern "C" int fsv_act_bwd(const float* y, const float* dy, float* g, long long n, int act, float out_scale, void* stream) {
    FSV_REQUIRE(n > 0 && out_scale != 0.f, "act_bwd: bad args");
    k_act_bwd<<<stream_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(y, dy, g, n, act, out_scale);
    FSV_CHECK_LAUNCH("act_bwd");
    return FSV_OK;
}


// Pre-summed weights of the upsample-collapsed convolution (conv_tc.cu: fsv_conv2d_fwd_tc_up2): conv3x3 over a nearest-x2-upsampled
// image equals, per output parity (p, q), a 2x2-tap convolution of the source image whose tap (a, b) sums the 3x3 taps that land on
// the same source pixel: even rows: {r0} | {r1, r2}; odd rows: {r0, r1} | {r2}.  w (Cout, 3, 3, Cin) -> w4 (Cout, 2, 2, 2, 2, Cin).
// (Replaces a torch.einsum = six tiny cuBLAS GEMMs per forward.)
__global__ void k_up2_weights(const float* __restrict__ w, float* __restrict__ w4, long long cout, int cin) {
    const long long total = cout * 16 * cin;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cin);
        long long t = i / cin;
        const int b = (int)(t & 1), a = (int)((t >> 1) & 1), q = (int)((t >> 2) & 1), p = (int)((t >> 3) & 1);
        const long long o = t >> 4;
        // rows r contributing to tap a of parity p: p=0: a=0 -> {0}, a=1 -> {1,2};  p=1: a=0 -> {0,1}, a=1 -> {2}
        const int r0 = p == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2), r1 = p == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
        const int s0 = q == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2), s1 = q == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
        float acc = 0.f;
        for (int r = r0; r <= r1; ++r)
            for (int s2 = s0; s2 <= s1; ++s2) acc += w[((o * 3 + r) * 3 + s2) * cin + c];
        w4[i] = acc;
    }
}
extern "C" int fsv_up2_weights(const float* w, float* w4, int Cout, int Cin, void* stream) {
    FSV_REQUIRE(w && w4 && Cout > 0 && Cin > 0, "up2_weights: bad args");
    k_up2_weights<<<stream_grid((long long)Cout * 16 * Cin, 256), 256, 0, (cudaStream_t)stream>>>(w, w4, Cout, Cin);
    FSV_CHECK_LAUNCH("up2_weights");
    return FSV_OK;
}


// Backward of y = conv3x3(nearest_up2(x)) at SOURCE resolution (the autograd of generator.py:484,537 runs the 3x3 data gradient at
// the upsampled resolution and then sums 2x2 blocks).  Source pixel i feeds the upsampled rows 2i and 2i+1, and kernel row r of
// upsampled row u reads dy row u + 1 - r, so dx[i] gathers dy rows 2i + k - 1, k = 0..3: a 4x4 / stride-2 / pad-1 convolution of dy
// whose tap row k sums the 3x3 rows r with k in {2 - r, 3 - r}:   k=0 <- {2}, k=1 <- {1,2}, k=2 <- {0,1}, k=3 <- {0}
// (same for columns).  16 taps at a quarter of the pixels = 4/9 of the MACs, no full-resolution dx and no 2x2 reduction pass.
//   data gradient  : dx = conv4x4s2(dy, wf),  wf[ci][k][l][co] = sum_{r in R(k), s in R(l)} wt[ci][r][s][co]     (fsv_up2_dgrad_weights)
//   weight gradient: dW16[ci][k][l][co] = sum_px x[px][ci] * dy[2 px + (k,l) - 1][co]  (the 4x4 / stride-2 weight gradient with the roles
//                    of the two tensors exchanged), dW[co][r][s][ci] = sum_{k in {2-r,3-r}, l in {2-s,3-s}} dW16[ci][k][l][co]  (fsv_up2_wgrad_fold)
__global__ void k_up2_dgrad_weights(const float* __restrict__ wt, float* __restrict__ wf, long long cin, int cout) {
    const long long total = cin * 16 * cout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % cout);
        const long long t = i / cout;
        const int l = (int)(t & 3), k = (int)((t >> 2) & 3);
        const long long ci = t >> 4;
        const int r0 = max(0, 2 - k), r1 = min(2, 3 - k);
        const int s0 = max(0, 2 - l), s1 = min(2, 3 - l);
        float acc = 0.f;
        for (int r = r0; r <= r1; ++r)
            for (int s = s0; s <= s1; ++s) acc += wt[((ci * 3 + r) * 3 + s) * cout + co];
        wf[i] = acc;
    }
}
extern "C" int fsv_up2_dgrad_weights(const float* wt, float* wf, int Cin, int Cout, void* stream) {
    FSV_REQUIRE(wt && wf && Cout > 0 && Cin > 0, "up2_dgrad_weights: bad args");
    k_up2_dgrad_weights<<<stream_grid((long long)Cin * 16 * Cout, 256), 256, 0, (cudaStream_t)stream>>>(wt, wf, Cin, Cout);
    FSV_CHECK_LAUNCH("up2_dgrad_weights");
    return FSV_OK;
}

// grid (co tiles, ci tiles, 9 taps); 32x32 tiles through shared memory so that both the (.., co) reads and the (.., ci) writes are coalesced
__global__ void __launch_bounds__(256) k_up2_wgrad_fold(const float* __restrict__ dw16, float* __restrict__ dw, int cout, int cin, int accumulate) {
    __shared__ float tile[32][33];
    const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
    const int r = blockIdx.z / 3, s = blockIdx.z - 3 * r;
    for (int i = threadIdx.y; i < 32; i += 8) {         // i = ci within the tile, threadIdx.x = co
        const int ci = ci0 + i, co = co0 + threadIdx.x;
        float acc = 0.f;
        if (ci < cin && co < cout) {
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b)
                    acc += dw16[(((long long)ci * 4 + (2 - r + a)) * 4 + (2 - s + b)) * cout + co];
        }
        tile[i][threadIdx.x] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {         // i = co within the tile, threadIdx.x = ci
        const int co = co0 + i, ci = ci0 + threadIdx.x;
        if (ci < cin && co < cout) {
            float* o = dw + (((long long)co * 3 + r) * 3 + s) * cin + ci;
            *o = accumulate ? *o + tile[threadIdx.x][i] : tile[threadIdx.x][i];
        }
    }
}
extern "C" int fsv_up2_wgrad_fold(const float* dw16, float* dw, int Cout, int Cin, int accumulate, void* stream) {
    FSV_REQUIRE(dw16 && dw && Cout > 0 && Cin > 0, "up2_wgrad_fold: bad args");
    dim3 grid(fsv_cdiv(Cout, 32), fsv_cdiv(Cin, 32), 9);
    FSV_REQUIRE(grid.y <= 65535, "up2_wgrad_fold: grid too large");
    k_up2_wgrad_fold<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(dw16, dw, Cout, Cin, accumulate);
    FSV_CHECK_LAUNCH("up2_wgrad_fold");
    return FSV_OK;
}


// MaxPool2d(kernel 2, stride 2) on NHWC (VGG19 feature stack of the perceptual loss, vgg.py:45-59).  One thread per output element.
__global__ void k_maxpool2_fwd(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C) {
    const int Ho = H / 2, Wo = W / 2;
    const long long total = (long long)N * Ho * Wo * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long p = i / C;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const long long n = p / Ho;
        const float* b = x + ((n * H + 2 * ho) * W + 2 * wo) * C + c;
        y[i] = fmaxf(fmaxf(b[0], b[C]), fmaxf(b[(long long)W * C], b[(long long)W * C + C]));
    }
}
// dx fully written: every input pixel belongs to at most one window; the gradient goes to the first maximum in window order
__global__ void k_maxpool2_bwd(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int N, int H, int W, int C) {
    const int Ho = H / 2, Wo = W / 2;
    const long long total = (long long)N * H * W * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long p = i / C;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const long long n = p / H;
        const int ho = h / 2, wo = w / 2;
        float g = 0.f;
        if (ho < Ho && wo < Wo) {
            const float* b = x + ((n * H + 2 * ho) * W + 2 * wo) * C + c;
            const float v[4] = {b[0], b[C], b[(long long)W * C], b[(long long)W * C + C]};
            int arg = 0;
            for (int k = 1; k < 4; ++k)
                if (v[k] > v[arg]) arg = k;
            if (arg == (h & 1) * 2 + (w & 1)) g = dy[((n * Ho + ho) * Wo + wo) * C + c];
        }
        dx[i] = g;
    }
}
extern "C" int fsv_maxpool2_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream) {
    FSV_REQUIRE(x && y && N > 0 && H > 1 && W > 1 && C > 0, "maxpool2_fwd: bad dims");
    k_maxpool2_fwd<<<stream_grid((long long)N * (H / 2) * (W / 2) * C, 256), 256, 0, (cudaStream_t)stream>>>(x, y, N, H, W, C);
    FSV_CHECK_LAUNCH("maxpool2_fwd");
    return FSV_OK;
}
extern "C" int fsv_maxpool2_bwd(const float* x, const float* dy, float* dx, int N, int H, int W, int C, void* stream) {
    FSV_REQUIRE(x && dy && dx && N > 0 && H > 1 && W > 1 && C > 0, "maxpool2_bwd: bad dims");
    k_maxpool2_bwd<<<stream_grid((long long)N * H * W * C, 256), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, N, H, W, C);
    FSV_CHECK_LAUNCH("maxpool2_bwd");
    return FSV_OK;
}
