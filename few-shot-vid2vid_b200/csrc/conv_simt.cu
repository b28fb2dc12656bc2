// SIMT (CUDA-core, fp32 FFMA) implicit-GEMM convolution: forward, data gradient, weight gradient.
//
// Role: (1) the exact-fp32 path for every shape -- thin layers (Cin in {1,3,4,5,6,8,15,20},
// Cout in {1,2,3}), odd hyper-network widths (33, 65, 66, 130, ...), per-sample 1x1 "batch_conv"
// weights (base_network.py:56-71) -- which are HBM-bound and do not belong on tensor cores;
// (2) the parity baseline that the tcgen05/TMA path (conv_tc.cu) is validated against.
// Replaces F.conv2d / F.linear / F.conv_transpose2d-as-autograd of the reference
// (architecture.py:60,81-84; generator.py:260-270,473-489,523-537; discriminator.py:69-88).
//
// GEMM view (per sample n, so per-sample weights never straddle a tile):
//   fwd   : M = Ho*Wo pixels, Ncol = Cout, K = taps x Cin ; B[k,ncol] = w[ncol][tap][k]
//   dgrad : M = H*W   pixels, Ncol = Cin , K = taps x Cout; B[k,ncol] = w[k][tap][ncol]
//   wgrad : M = Cout, Ncol = Cin (per tap), K = pixels (split over blocks, fp32 atomics)
// Tile 64x64x16, 256 threads, 4x4 register block per thread, operands staged through shared
// memory with 16-byte global loads whenever the channel slice is 16-byte aligned.
#include "common.cuh"

#define BM 64
#define BN 64
#define BK 16
#define PADM 4

struct ConvP {
    int N, H, W, Cin, x_ld, x_coff, up;
    int Cout, kh, kw, stride, pad, Ho, Wo, y_ld, y_coff;
    int act;
    float out_scale;
    long long w_nstride, b_nstride;
    int res_ld, res_coff;
    int in_act;
    int accumulate;
};

static ConvP make_p(const fsv_conv_desc* d, int accumulate) {
    ConvP p;
    p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.x_ld = d->x_ld; p.x_coff = d->x_coff; p.up = d->up;
    p.Cout = d->Cout; p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad = d->pad; p.Ho = d->Ho; p.Wo = d->Wo;
    p.y_ld = d->y_ld; p.y_coff = d->y_coff; p.act = d->act; p.out_scale = d->out_scale;
    p.w_nstride = d->w_nstride; p.b_nstride = d->b_nstride; p.res_ld = d->res_ld; p.res_coff = d->res_coff; p.in_act = d->in_act;
    p.accumulate = accumulate;
    return p;
}

int fsv_conv_validate(const fsv_conv_desc* d, const char* who) {
    FSV_REQUIRE(d != nullptr, "%s: null descriptor", who);
    FSV_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "%s: bad dims", who);
    FSV_REQUIRE(d->kh > 0 && d->kw > 0 && d->stride > 0 && d->pad >= 0, "%s: bad kernel geometry", who);
    FSV_REQUIRE(d->up == 1 || d->up == 2, "%s: up must be 1 or 2", who);
    FSV_REQUIRE(d->up == 1 || (d->H % 2 == 0 && d->W % 2 == 0), "%s: upsample-on-load needs even H,W", who);
    FSV_REQUIRE(d->Ho == (d->H + 2 * d->pad - d->kh) / d->stride + 1 && d->Wo == (d->W + 2 * d->pad - d->kw) / d->stride + 1,
                "%s: Ho/Wo inconsistent with H,W,k,stride,pad", who);
    FSV_REQUIRE(d->x_ld >= d->x_coff + d->Cin && d->y_ld >= d->y_coff + d->Cout, "%s: ld/coff inconsistent", who);
    FSV_REQUIRE(d->out_scale != 0.f, "%s: out_scale must be non-zero", who);
    FSV_REQUIRE(d->in_act == FSV_ACT_NONE || d->in_act == FSV_ACT_LRELU, "%s: in_act must be none or lrelu", who);
    return FSV_OK;
}

// MODE 0 = forward, 1 = dgrad
template <int MODE>
__global__ void __launch_bounds__(256) k_conv_simt(ConvP p, const float* __restrict__ src, const float* __restrict__ w,
                                                   const float* __restrict__ bias, const float* __restrict__ residual,
                                                   float* __restrict__ dst) {
    __shared__ __align__(16) float As[BK][BM + PADM];
    __shared__ __align__(16) float Bs[BK][BN + PADM];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int n = blockIdx.z;
    const int m0 = blockIdx.x * BM;
    const int c0 = blockIdx.y * BN;

    // GEMM-view dimensions
    const int MW = MODE == 0 ? p.Wo : p.W;                 // width of the M pixel grid
    const int MP = MODE == 0 ? p.Ho * p.Wo : p.H * p.W;    // pixels per sample
    const int NC = MODE == 0 ? p.Cout : p.Cin;             // GEMM columns
    const int KC = MODE == 0 ? p.Cin : p.Cout;             // reduction channels per tap
    const int s_ld = MODE == 0 ? p.x_ld : p.y_ld;
    const int s_coff = MODE == 0 ? p.x_coff : p.y_coff;
    const int SH = MODE == 0 ? p.H / p.up : p.Ho;          // source buffer spatial dims
    const int SW = MODE == 0 ? p.W / p.up : p.Wo;
    const float* wn = w + (long long)n * p.w_nstride;
    const float* srcn = src + (long long)n * SH * SW * s_ld + s_coff;
    const bool a_vec = ((s_ld & 3) == 0) && ((s_coff & 3) == 0) && ((KC & 3) == 0) && ((((uintptr_t)src) & 15) == 0);
    const bool b_vec = ((p.Cin & 3) == 0) && ((p.w_nstride & 3) == 0) && ((((uintptr_t)w) & 15) == 0);

    // A-load assignment: pixel ap, k-quad akq
    const int ap = tid >> 2, akq = (tid & 3) * 4;
    const int am = m0 + ap;
    const int ai = am / MW, aj = am - ai * MW;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int taps = p.kh * p.kw;
    for (int tap = 0; tap < taps; ++tap) {
        const int r = tap / p.kw, s = tap - r * p.kw;
        // source pixel of this thread's A row for this tap
        long long a_off = -1;
        if (am < MP) {
            if (MODE == 0) {
                int ih = ai * p.stride + r - p.pad, iw = aj * p.stride + s - p.pad;
                if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
                    if (p.up == 2) { ih >>= 1; iw >>= 1; }
                    a_off = ((long long)ih * SW + iw) * s_ld;
                }
            } else {
                int th = ai + p.pad - r, tw = aj + p.pad - s;
                if (th >= 0 && tw >= 0 && (th % p.stride) == 0 && (tw % p.stride) == 0) {
                    int oh = th / p.stride, ow = tw / p.stride;
                    if (oh < p.Ho && ow < p.Wo) a_off = ((long long)oh * SW + ow) * s_ld;
                }
            }
        }
        for (int k0 = 0; k0 < KC; k0 += BK) {
            // ---- A tile: As[k][pixel]
            {
                float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
                int k = k0 + akq;
                if (a_off >= 0 && k < KC) {
                    const float* ptr = srcn + a_off + k;
                    if (a_vec) {
                        float4 t = *reinterpret_cast<const float4*>(ptr);
                        v0 = t.x; v1 = t.y; v2 = t.z; v3 = t.w;
                    } else {
                        v0 = ptr[0];
                        if (k + 1 < KC) v1 = ptr[1];
                        if (k + 2 < KC) v2 = ptr[2];
                        if (k + 3 < KC) v3 = ptr[3];
                    }
                }
                if (MODE == 0 && p.in_act == FSV_ACT_LRELU) {
                    v0 = fsv_act(v0, FSV_ACT_LRELU); v1 = fsv_act(v1, FSV_ACT_LRELU);
                    v2 = fsv_act(v2, FSV_ACT_LRELU); v3 = fsv_act(v3, FSV_ACT_LRELU);
                }
                As[akq + 0][ap] = v0; As[akq + 1][ap] = v1; As[akq + 2][ap] = v2; As[akq + 3][ap] = v3;
            }
            // ---- B tile: Bs[k][col]
            if (MODE == 0) {
                int col = tid >> 2, kq = (tid & 3) * 4;
                int co = c0 + col, k = k0 + kq;
                float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
                if (co < NC && k < KC) {
                    const float* ptr = wn + ((long long)co * taps + tap) * p.Cin + k;
                    if (b_vec) {
                        float4 t = *reinterpret_cast<const float4*>(ptr);
                        v0 = t.x; v1 = t.y; v2 = t.z; v3 = t.w;
                    } else {
                        v0 = ptr[0];
                        if (k + 1 < KC) v1 = ptr[1];
                        if (k + 2 < KC) v2 = ptr[2];
                        if (k + 3 < KC) v3 = ptr[3];
                    }
                }
                Bs[kq + 0][col] = v0; Bs[kq + 1][col] = v1; Bs[kq + 2][col] = v2; Bs[kq + 3][col] = v3;
            } else {
                int kk = tid >> 4, cq = (tid & 15) * 4;
                int k = k0 + kk, ci = c0 + cq;
                float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
                if (k < KC && ci < NC) {
                    const float* ptr = wn + ((long long)k * taps + tap) * p.Cin + ci;
                    if (b_vec) {
                        float4 t = *reinterpret_cast<const float4*>(ptr);
                        v0 = t.x; v1 = t.y; v2 = t.z; v3 = t.w;
                    } else {
                        v0 = ptr[0];
                        if (ci + 1 < NC) v1 = ptr[1];
                        if (ci + 2 < NC) v2 = ptr[2];
                        if (ci + 3 < NC) v3 = ptr[3];
                    }
                }
                *reinterpret_cast<float4*>(&Bs[kk][cq]) = make_float4(v0, v1, v2, v3);
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                acc[0][0] += a.x * b.x; acc[0][1] += a.x * b.y; acc[0][2] += a.x * b.z; acc[0][3] += a.x * b.w;
                acc[1][0] += a.y * b.x; acc[1][1] += a.y * b.y; acc[1][2] += a.y * b.z; acc[1][3] += a.y * b.w;
                acc[2][0] += a.z * b.x; acc[2][1] += a.z * b.y; acc[2][2] += a.z * b.z; acc[2][3] += a.z * b.w;
                acc[3][0] += a.w * b.x; acc[3][1] += a.w * b.y; acc[3][2] += a.w * b.z; acc[3][3] += a.w * b.w;
            }
            __syncthreads();
        }
    }

    // ---- epilogue
    const int d_ld = MODE == 0 ? p.y_ld : p.x_ld;
    const int d_coff = MODE == 0 ? p.y_coff : p.x_coff;
    const int DH = MODE == 0 ? p.Ho : p.H / p.up;   // dgrad with up!=1 is rejected on the host
    const int DW = MODE == 0 ? p.Wo : p.W / p.up;
    float* dstn = dst + (long long)n * DH * DW * d_ld + d_coff;
    const float* bn = (MODE == 0 && bias) ? bias + (long long)n * p.b_nstride : nullptr;
    const float* resn = (MODE == 0 && residual) ? residual + (long long)n * DH * DW * p.res_ld + p.res_coff : nullptr;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= MP) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = c0 + tx * 4 + j;
            if (c >= NC) continue;
            float v = acc[i][j];
            if (MODE == 0) {
                if (bn) v += bn[c];
                if (resn) v += resn[(long long)m * p.res_ld + c];
                v = fsv_act(v, p.act) * p.out_scale;
                dstn[(long long)m * d_ld + c] = v;
            } else {
                float* q = dstn + (long long)m * d_ld + c;
                if (p.accumulate) *q += v; else *q = v;
            }
        }
    }
}

extern "C" int fsv_conv2d_fwd_simt(const fsv_conv_desc* d, const float* x, const float* w, const float* bias,
                                   const float* residual, float* y, void* stream) {
    int rc = fsv_conv_validate(d, "conv2d_fwd");
    if (rc) return rc;
    FSV_REQUIRE(!residual || d->res_ld >= d->res_coff + d->Cout, "conv2d_fwd: residual ld/coff inconsistent");
    ConvP p = make_p(d, 0);
    dim3 grid(fsv_cdiv((long long)d->Ho * d->Wo, BM), fsv_cdiv(d->Cout, BN), d->N);
    k_conv_simt<0><<<grid, 256, 0, (cudaStream_t)stream>>>(p, x, w, bias, residual, y);
    FSV_CHECK_LAUNCH("conv2d_fwd_simt");
    return FSV_OK;
}

extern "C" int fsv_conv2d_dgrad_thin_ok(const fsv_conv_desc* d);
extern "C" int fsv_conv2d_dgrad_thin(const fsv_conv_desc* d, const float* dy, const float* w, float* dx, int accumulate, void* stream);

extern "C" int fsv_conv2d_dgrad(const fsv_conv_desc* d, const float* dy, const float* w, float* dx, int accumulate, void* stream) {
    int rc = fsv_conv_validate(d, "conv2d_dgrad");
    if (rc) return rc;
    FSV_REQUIRE(d->up == 1, "conv2d_dgrad: up must be 1 (take the gradient at conv-input resolution, then fsv_upsample2x_bwd)");
    {
        int tk = fsv_conv2d_dgrad_thin_ok(d);
        if ((tk == 1 && (((uintptr_t)dy) & 15) == 0) || (tk == 2 && (((uintptr_t)dx) & 15) == 0 && (((uintptr_t)w) & 15) == 0))
            return fsv_conv2d_dgrad_thin(d, dy, w, dx, accumulate, stream);
    }
    ConvP p = make_p(d, accumulate);
    dim3 grid(fsv_cdiv((long long)d->H * d->W, BM), fsv_cdiv(d->Cin, BN), d->N);
    k_conv_simt<1><<<grid, 256, 0, (cudaStream_t)stream>>>(p, dy, w, nullptr, nullptr, dx);
    FSV_CHECK_LAUNCH("conv2d_dgrad");
    return FSV_OK;
}

// ------------------------------------------------------------------ weight gradient
// grid.x = co tiles * ci tiles, grid.y = taps, grid.z = pixel chunks (never straddling a sample)
__global__ void __launch_bounds__(256) k_conv_wgrad(ConvP p, const float* __restrict__ x, const float* __restrict__ dy,
                                                    float* __restrict__ dw, int ci_tiles, int chunks_per_sample, int pix_per_chunk) {
    __shared__ __align__(16) float As[BK][BM + PADM];   // dy tile  [pixel][co]
    __shared__ __align__(16) float Bs[BK][BN + PADM];   // x tile   [pixel][ci]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int co0 = (blockIdx.x / ci_tiles) * BM;
    const int ci0 = (blockIdx.x % ci_tiles) * BN;
    const int tap = blockIdx.y;
    const int r = tap / p.kw, s = tap - r * p.kw;
    const int n = blockIdx.z / chunks_per_sample;
    const int chunk = blockIdx.z - n * chunks_per_sample;
    const int MP = p.Ho * p.Wo;
    const int px0 = chunk * pix_per_chunk;
    const int px1 = min(px0 + pix_per_chunk, MP);
    const int Hs = p.H / p.up, Ws = p.W / p.up;
    const float* dyn = dy + (long long)n * MP * p.y_ld + p.y_coff;
    const float* xn = x + (long long)n * Hs * Ws * p.x_ld + p.x_coff;
    const bool a_vec = ((p.y_ld & 3) == 0) && ((p.y_coff & 3) == 0) && ((p.Cout & 3) == 0) && ((((uintptr_t)dy) & 15) == 0);
    const bool b_vec = ((p.x_ld & 3) == 0) && ((p.x_coff & 3) == 0) && ((p.Cin & 3) == 0) && ((((uintptr_t)x) & 15) == 0);

    const int kk = tid >> 4, cq = (tid & 15) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int p0 = px0; p0 < px1; p0 += BK) {
        int px = p0 + kk;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
        if (px < px1) {
            int co = co0 + cq;
            if (co < p.Cout) {
                const float* ptr = dyn + (long long)px * p.y_ld + co;
                if (a_vec) { float4 t = *reinterpret_cast<const float4*>(ptr); a0 = t.x; a1 = t.y; a2 = t.z; a3 = t.w; }
                else { a0 = ptr[0]; if (co + 1 < p.Cout) a1 = ptr[1]; if (co + 2 < p.Cout) a2 = ptr[2]; if (co + 3 < p.Cout) a3 = ptr[3]; }
            }
            int ho = px / p.Wo, wo = px - ho * p.Wo;
            int ih = ho * p.stride + r - p.pad, iw = wo * p.stride + s - p.pad;
            int ci = ci0 + cq;
            if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W && ci < p.Cin) {
                if (p.up == 2) { ih >>= 1; iw >>= 1; }
                const float* ptr = xn + ((long long)ih * Ws + iw) * p.x_ld + ci;
                if (b_vec) { float4 t = *reinterpret_cast<const float4*>(ptr); b0 = t.x; b1 = t.y; b2 = t.z; b3 = t.w; }
                else { b0 = ptr[0]; if (ci + 1 < p.Cin) b1 = ptr[1]; if (ci + 2 < p.Cin) b2 = ptr[2]; if (ci + 3 < p.Cin) b3 = ptr[3]; }
            }
        }
        if (p.in_act == FSV_ACT_LRELU) {
            b0 = fsv_act(b0, FSV_ACT_LRELU); b1 = fsv_act(b1, FSV_ACT_LRELU);
            b2 = fsv_act(b2, FSV_ACT_LRELU); b3 = fsv_act(b3, FSV_ACT_LRELU);
        }
        *reinterpret_cast<float4*>(&As[kk][cq]) = make_float4(a0, a1, a2, a3);
        *reinterpret_cast<float4*>(&Bs[kk][cq]) = make_float4(b0, b1, b2, b3);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            acc[0][0] += a.x * b.x; acc[0][1] += a.x * b.y; acc[0][2] += a.x * b.z; acc[0][3] += a.x * b.w;
            acc[1][0] += a.y * b.x; acc[1][1] += a.y * b.y; acc[1][2] += a.y * b.z; acc[1][3] += a.y * b.w;
            acc[2][0] += a.z * b.x; acc[2][1] += a.z * b.y; acc[2][2] += a.z * b.z; acc[2][3] += a.z * b.w;
            acc[3][0] += a.w * b.x; acc[3][1] += a.w * b.y; acc[3][2] += a.w * b.z; acc[3][3] += a.w * b.w;
        }
        __syncthreads();
    }
    float* dwn = dw + (long long)n * p.w_nstride;
    const int taps = p.kh * p.kw;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int co = co0 + ty * 4 + i;
        if (co >= p.Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int ci = ci0 + tx * 4 + j;
            if (ci >= p.Cin) continue;
            atomicAdd(dwn + ((long long)co * taps + tap) * p.Cin + ci, acc[i][j]);
        }
    }
}

// Thin-input variant (Cin <= 8): the (tap, ci) pairs are flattened into the GEMM column axis j = tap*Cin + ci, so one CTA
// covers 64 output channels x 64 (tap,ci) columns instead of wasting a 64-wide tile on <= 8 input channels per tap.
// grid.x = co tiles * j tiles, grid.z = pixel chunks.
__global__ void __launch_bounds__(256) k_conv_wgrad_flat(ConvP p, const float* __restrict__ x, const float* __restrict__ dy,
                                                         float* __restrict__ dw, int j_tiles, int chunks_per_sample, int pix_per_chunk) {
    __shared__ __align__(16) float As[BK][BM + PADM];   // dy tile  [pixel][co]
    __shared__ __align__(16) float Bs[BK][BN + PADM];   // x gather [pixel][j]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int co0 = (blockIdx.x / j_tiles) * BM;
    const int j0 = (blockIdx.x % j_tiles) * BN;
    const int n = blockIdx.z / chunks_per_sample;
    const int chunk = blockIdx.z - n * chunks_per_sample;
    const int MP = p.Ho * p.Wo;
    const int px0 = chunk * pix_per_chunk;
    const int px1 = min(px0 + pix_per_chunk, MP);
    const int taps = p.kh * p.kw, tc = taps * p.Cin;
    const float* dyn = dy + (long long)n * MP * p.y_ld + p.y_coff;
    const float* xn = x + (long long)n * p.H * p.W * p.x_ld + p.x_coff;
    const bool a_vec = ((p.y_ld & 3) == 0) && ((p.y_coff & 3) == 0) && ((p.Cout & 3) == 0) && ((((uintptr_t)dy) & 15) == 0);
    const int kk = tid >> 4, cq = (tid & 15) * 4;
    // this thread's 4 gather columns
    int jr[4], js[4], jc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int j = j0 + cq + e;
        int tap = j < tc ? j / p.Cin : 0;
        jc[e] = j < tc ? j - tap * p.Cin : -1;
        jr[e] = tap / p.kw;
        js[e] = tap - jr[e] * p.kw;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int p0 = px0; p0 < px1; p0 += BK) {
        int px = p0 + kk;
        float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
        if (px < px1) {
            int co = co0 + cq;
            if (co < p.Cout) {
                const float* ptr = dyn + (long long)px * p.y_ld + co;
                if (a_vec) { float4 t = *reinterpret_cast<const float4*>(ptr); a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w; }
                else { a[0] = ptr[0]; if (co + 1 < p.Cout) a[1] = ptr[1]; if (co + 2 < p.Cout) a[2] = ptr[2]; if (co + 3 < p.Cout) a[3] = ptr[3]; }
            }
            int ho = px / p.Wo, wo = px - ho * p.Wo;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (jc[e] >= 0) {
                    int ih = ho * p.stride + jr[e] - p.pad, iw = wo * p.stride + js[e] - p.pad;
                    if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
                        float v = xn[((long long)ih * p.W + iw) * p.x_ld + jc[e]];
                        b[e] = p.in_act == FSV_ACT_LRELU ? fsv_act(v, FSV_ACT_LRELU) : v;
                    }
                }
            }
        }
        *reinterpret_cast<float4*>(&As[kk][cq]) = make_float4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<float4*>(&Bs[kk][cq]) = make_float4(b[0], b[1], b[2], b[3]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            acc[0][0] += av.x * bv.x; acc[0][1] += av.x * bv.y; acc[0][2] += av.x * bv.z; acc[0][3] += av.x * bv.w;
            acc[1][0] += av.y * bv.x; acc[1][1] += av.y * bv.y; acc[1][2] += av.y * bv.z; acc[1][3] += av.y * bv.w;
            acc[2][0] += av.z * bv.x; acc[2][1] += av.z * bv.y; acc[2][2] += av.z * bv.z; acc[2][3] += av.z * bv.w;
            acc[3][0] += av.w * bv.x; acc[3][1] += av.w * bv.y; acc[3][2] += av.w * bv.z; acc[3][3] += av.w * bv.w;
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int co = co0 + ty * 4 + i;
        if (co >= p.Cout) continue;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int j = j0 + tx * 4 + e;
            if (j < tc) atomicAdd(dw + (long long)co * tc + j, acc[i][e]);
        }
    }
}

// Second version of the thin-input weight gradient: a 32 (co) x BNJ (tap,ci) tile instead of 64 x 64 -- the layers it
// serves have 32 output channels and 9..128 (tap,ci) columns, so the square tile ran at 7-50 % utilisation.  The 256
// threads are (BNJ/4 column quads) x (8 co quads) x KS pixel slices; every slice owns BK/KS of the chunk's pixels and
// the slices meet in the final atomics.  Used with BNJ = 64 when taps*Cin > 64 (see the dispatch in fsv_conv2d_wgrad).
template <int BNJ>
__global__ void __launch_bounds__(256) k_conv_wgrad_flat2(ConvP p, const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ dw, int j_tiles, int chunks_per_sample, int pix_per_chunk) {
    constexpr int QJ = BNJ / 4;              // column quads
    constexpr int KS = 256 / (QJ * 8);       // pixel slices: 2 (BNJ 64) or 8 (BNJ 16)
    constexpr int KPS = BK / KS;             // pixels per slice per chunk
    constexpr int BPT = BK * BNJ / 256;      // gathered x values per thread per chunk: 4 or 1
    __shared__ __align__(16) float As[BK][32 + PADM];    // dy tile  [pixel][co]
    __shared__ __align__(16) float Bs[BK][BNJ + PADM];   // x gather [pixel][j]
    const int tid = threadIdx.x;
    const int tx = tid % QJ, ty = (tid / QJ) & 7, kz = tid / (QJ * 8);
    const int co0 = (blockIdx.x / j_tiles) * 32;
    const int j0 = (blockIdx.x % j_tiles) * BNJ;
    const int n = blockIdx.z / chunks_per_sample;
    const int chunk = blockIdx.z - n * chunks_per_sample;
    const int MP = p.Ho * p.Wo;
    const int px0 = chunk * pix_per_chunk;
    const int px1 = min(px0 + pix_per_chunk, MP);
    const int taps = p.kh * p.kw, tc = taps * p.Cin;
    const float* dyn = dy + (long long)n * MP * p.y_ld + p.y_coff;
    const float* xn = x + (long long)n * p.H * p.W * p.x_ld + p.x_coff;
    const bool a_vec = ((p.y_ld & 3) == 0) && ((p.y_coff & 3) == 0) && ((p.Cout & 3) == 0) && ((((uintptr_t)dy) & 15) == 0);
    // loader roles: B -- thread owns BPT consecutive columns of pixel row kb; A -- threads < 128 own one float4 of dy
    const int kb = tid / (BNJ / BPT), cb = (tid % (BNJ / BPT)) * BPT;
    const int ka = tid >> 3, ca = (tid & 7) * 4;
    int jr[BPT], js[BPT], jc[BPT];
#pragma unroll
    for (int e = 0; e < BPT; ++e) {
        int j = j0 + cb + e;
        int tap = j < tc ? j / p.Cin : 0;
        jc[e] = j < tc ? j - tap * p.Cin : -1;
        jr[e] = tap / p.kw;
        js[e] = tap - jr[e] * p.kw;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int p0 = px0; p0 < px1; p0 += BK) {
        if (tid < 128) {
            float a[4] = {0.f, 0.f, 0.f, 0.f};
            const int px = p0 + ka, co = co0 + ca;
            if (px < px1 && co < p.Cout) {
                const float* ptr = dyn + (long long)px * p.y_ld + co;
                if (a_vec) { float4 t = *reinterpret_cast<const float4*>(ptr); a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w; }
                else { a[0] = ptr[0]; if (co + 1 < p.Cout) a[1] = ptr[1]; if (co + 2 < p.Cout) a[2] = ptr[2]; if (co + 3 < p.Cout) a[3] = ptr[3]; }
            }
            *reinterpret_cast<float4*>(&As[ka][ca]) = make_float4(a[0], a[1], a[2], a[3]);
        }
        {
            float b[BPT];
#pragma unroll
            for (int e = 0; e < BPT; ++e) b[e] = 0.f;
            const int px = p0 + kb;
            if (px < px1) {
                const int ho = px / p.Wo, wo = px - ho * p.Wo;
#pragma unroll
                for (int e = 0; e < BPT; ++e) {
                    if (jc[e] >= 0) {
                        const int ih = ho * p.stride + jr[e] - p.pad, iw = wo * p.stride + js[e] - p.pad;
                        if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
                            float v = xn[((long long)ih * p.W + iw) * p.x_ld + jc[e]];
                            b[e] = p.in_act == FSV_ACT_LRELU ? fsv_act(v, FSV_ACT_LRELU) : v;
                        }
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < BPT; ++e) Bs[kb][cb + e] = b[e];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < KPS; ++kk) {
            const int k = kz * KPS + kk;
            float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            acc[0][0] += av.x * bv.x; acc[0][1] += av.x * bv.y; acc[0][2] += av.x * bv.z; acc[0][3] += av.x * bv.w;
            acc[1][0] += av.y * bv.x; acc[1][1] += av.y * bv.y; acc[1][2] += av.y * bv.z; acc[1][3] += av.y * bv.w;
            acc[2][0] += av.z * bv.x; acc[2][1] += av.z * bv.y; acc[2][2] += av.z * bv.z; acc[2][3] += av.z * bv.w;
            acc[3][0] += av.w * bv.x; acc[3][1] += av.w * bv.y; acc[3][2] += av.w * bv.z; acc[3][3] += av.w * bv.w;
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int co = co0 + ty * 4 + i;
        if (co >= p.Cout) continue;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int j = j0 + tx * 4 + e;
            if (j < tc) atomicAdd(dw + (long long)co * tc + j, acc[i][e]);
        }
    }
}

// per-(sample-group, channel) column sums of dy -> dbias (fp32 atomics into a zeroed buffer)
__global__ void k_colsum(const float* __restrict__ dy, int ld, int coff, long long rows_per_group, int C, long long out_gstride,
                         float* __restrict__ out) {
    int c = blockIdx.y * 32 + threadIdx.x;
    int g = blockIdx.z;
    long long r0 = (long long)blockIdx.x * 256, r1 = min(r0 + 256, rows_per_group);
    float s = 0.f;
    if (c < C)
        for (long long r = r0 + threadIdx.y; r < r1; r += 8) s += dy[(g * rows_per_group + r) * ld + coff + c];
    __shared__ float sm[8][32];
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
        atomicAdd(out + g * out_gstride + c, t);
    }
}

extern "C" int fsv_conv2d_thin_kind(const fsv_conv_desc* d);
extern "C" int fsv_conv2d_wgrad_thin(const fsv_conv_desc* d, const float* x, const float* dy, float* dw, int* handled, void* stream);

extern "C" int fsv_conv2d_wgrad(const fsv_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias,
                                int accumulate, void* stream) {
    int rc = fsv_conv_validate(d, "conv2d_wgrad");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    ConvP p = make_p(d, accumulate);
    const int taps = d->kh * d->kw;
    const long long wsize = (long long)d->Cout * taps * d->Cin;
    const int MP = d->Ho * d->Wo;
    if (dw) {
        if (!accumulate) {
            if (d->w_nstride == 0) {
                FSV_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * wsize, st));
            } else if (d->w_nstride == wsize) {
                FSV_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * wsize * d->N, st));
            } else {
                FSV_CUDA(cudaMemset2DAsync(dw, sizeof(float) * d->w_nstride, 0, sizeof(float) * wsize, d->N, st));
            }
        }
        if (d->Cin <= 8 && d->up == 1 && d->w_nstride == 0) {
            // thin input: flatten (tap, ci) into the GEMM column axis
            const int tc = taps * d->Cin;
            // measured on the B200 (profiles/breakdown_r1_cuda_events.txt history): the 32 x 64 tile wins only when there is
            // more than one 64-column tile (8 ch x 4x4 taps: 0.39 -> 0.32 ms); for taps*Cin <= 64 the 64 x 64 kernel is as
            // fast, and a 32 x 16 tile (tried for the 1-channel label conv) was 2.8x slower (sync-bound), so it is gone.
            const int flat_v1 = tc <= 64 ? 1 : 0;
            const int bm = flat_v1 ? BM : 32;
            const int bnj = 64;
            int co_tiles = fsv_cdiv(d->Cout, bm), j_tiles = fsv_cdiv(tc, bnj);
            long long base = (long long)co_tiles * j_tiles;
            long long want = ((long long)fsv_sm_count() * 8 + base - 1) / base;
            long long per_sample = (want + d->N - 1) / d->N;
            if (per_sample < 1) per_sample = 1;
            long long max_per_sample = (MP + 63) / 64;
            if (per_sample > max_per_sample) per_sample = max_per_sample;
            int pix_per_chunk = (int)((MP + per_sample - 1) / per_sample);
            pix_per_chunk = ((pix_per_chunk + BK - 1) / BK) * BK;
            int chunks_per_sample = fsv_cdiv(MP, pix_per_chunk);
            dim3 grid(co_tiles * j_tiles, 1, d->N * chunks_per_sample);
            FSV_REQUIRE(grid.z <= 65535, "conv2d_wgrad: grid too large");
            if (flat_v1) k_conv_wgrad_flat<<<grid, 256, 0, st>>>(p, x, dy, dw, j_tiles, chunks_per_sample, pix_per_chunk);
            else k_conv_wgrad_flat2<64><<<grid, 256, 0, st>>>(p, x, dy, dw, j_tiles, chunks_per_sample, pix_per_chunk);
            FSV_CHECK_LAUNCH("conv2d_wgrad_flat");
            goto bias_part;
        }
        if (fsv_conv2d_thin_kind(d) == 2) {
            int handled = 0;
            int trc = fsv_conv2d_wgrad_thin(d, x, dy, dw, &handled, stream);
            if (trc) return trc;
            if (handled) goto bias_part;
        }
        {
        int co_tiles = fsv_cdiv(d->Cout, BM), ci_tiles = fsv_cdiv(d->Cin, BN);
        long long base = (long long)co_tiles * ci_tiles * taps;
        long long want = ((long long)fsv_sm_count() * 16 + base - 1) / base;      // ~16 CTAs per SM (short K loops: latency-bound otherwise)
        long long per_sample = (want + d->N - 1) / d->N;
        if (per_sample < 1) per_sample = 1;
        long long max_per_sample = (MP + 63) / 64;
        if (per_sample > max_per_sample) per_sample = max_per_sample;
        int pix_per_chunk = (int)((MP + per_sample - 1) / per_sample);
        pix_per_chunk = ((pix_per_chunk + BK - 1) / BK) * BK;
        int chunks_per_sample = fsv_cdiv(MP, pix_per_chunk);
        dim3 grid(co_tiles * ci_tiles, taps, d->N * chunks_per_sample);
        FSV_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv2d_wgrad: grid too large");
        k_conv_wgrad<<<grid, 256, 0, st>>>(p, x, dy, dw, ci_tiles, chunks_per_sample, pix_per_chunk);
        FSV_CHECK_LAUNCH("conv2d_wgrad");
        }
    }
bias_part:
    if (dbias) {
        int groups = d->b_nstride ? d->N : 1;
        long long rpg = d->b_nstride ? MP : (long long)d->N * MP;
        if (!accumulate) {
            if (groups == 1 || d->b_nstride == d->Cout) {
                FSV_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)d->Cout * groups, st));
            } else {
                FSV_CUDA(cudaMemset2DAsync(dbias, sizeof(float) * d->b_nstride, 0, sizeof(float) * d->Cout, groups, st));
            }
        }
        dim3 grid(fsv_cdiv(rpg, 256), fsv_cdiv(d->Cout, 32), groups);
        k_colsum<<<grid, dim3(32, 8), 0, st>>>(dy, d->y_ld, d->y_coff, rpg, d->Cout, d->b_nstride, dbias);
        FSV_CHECK_LAUNCH("conv2d_bias_grad");
    }
    return FSV_OK;
}
