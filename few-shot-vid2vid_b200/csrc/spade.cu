// Fused SPADE modulation (SIMT version).
//
// Reference: models/networks/normalization.py:37-52 (SPADE.forward) + the LeakyReLU its caller
// applies (architecture.py:96-97; not on the shortcut, :103), fed through the nearest x2
// upsample of generator.py:207.  The reference runs, per SPADE, a cuDNN BatchNorm, two conv2d
// per label map (gamma, beta), four elementwise kernels per map and -- for the adaptive map -- a
// python loop over samples (base_network.py:56-71).  Here one kernel does, per 64-pixel x
// 64-channel tile:
//     v  = (x[n, h/up, w/up, c] - mean) * rstd                    (normalise, upsample-on-load)
//     for each label map i:   gamma_i, beta_i = 1x1 GEMM(map_i tile, Wg_i / Wb_i) + bias
//                             v = v * (1 + gamma_i) + beta_i       (gamma/beta never reach HBM)
//     out = LeakyReLU(v)
// Per-sample hyper-weights (w_nstride != 0) are addressed straight inside the hyper-network's
// flat output (no reshape/copy): wg + n*w_nstride, etc.
//
// Roofline: HBM-bound.  Algorithmic bytes (fp32) = 4*(|x|/up^2 + sum_i |map_i| + |out|) in this
// kernel (+ one read of x for the statistics pass).  The gamma/beta GEMM runs on FFMA here; at
// C*K >= 64*32 per pixel it is FFMA-bound, which is what the tcgen05 version removes.
#include "common.cuh"

#define TM 64
#define TN 64
#define TK 16
#define TPAD 4

struct SpadeP {
    int N, H, W, C, up, instance, act, nmaps;
    int K[FSV_SPADE_MAX_MAPS], m_ld[FSV_SPADE_MAX_MAPS], m_coff[FSV_SPADE_MAX_MAPS];
    long long w_nstride[FSV_SPADE_MAX_MAPS];
    int dgb_ld[FSV_SPADE_MAX_MAPS];
    const float* maps[FSV_SPADE_MAX_MAPS];
    const float* wg[FSV_SPADE_MAX_MAPS];
    const float* bg[FSV_SPADE_MAX_MAPS];
    const float* wb[FSV_SPADE_MAX_MAPS];
    const float* bb[FSV_SPADE_MAX_MAPS];
    float* dgamma[FSV_SPADE_MAX_MAPS];
    float* dbeta[FSV_SPADE_MAX_MAPS];
};

// gamma/beta 4x4 register tiles for map i of the block's (pixel tile, channel tile)
__device__ __forceinline__ void spade_gemm(const SpadeP& p, int i, int n, int m0, int c0, int HW,
                                           float (*Ms)[TM + TPAD], float (*Gs)[TN + TPAD], float (*Bs)[TN + TPAD],
                                           float gam[4][4], float bet[4][4]) {
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int K = p.K[i];
    const int ld = p.m_ld[i];
    const float* mp = p.maps[i] + (long long)n * HW * ld + p.m_coff[i];
    const float* wg = p.wg[i] + (long long)n * p.w_nstride[i];
    const float* wb = p.wb[i] + (long long)n * p.w_nstride[i];
    const bool m_vec = ((ld & 3) == 0) && ((p.m_coff[i] & 3) == 0) && ((K & 3) == 0) && ((((uintptr_t)p.maps[i]) & 15) == 0);
    const bool w_vec = ((K & 3) == 0) && ((p.w_nstride[i] & 3) == 0) && ((((uintptr_t)p.wg[i]) & 15) == 0) &&
                       ((((uintptr_t)p.wb[i]) & 15) == 0);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { gam[a][b] = 0.f; bet[a][b] = 0.f; }
    const int lp = tid >> 2, lk = (tid & 3) * 4;   // loader: row (pixel or channel) and k-quad
    for (int k0 = 0; k0 < K; k0 += TK) {
        int k = k0 + lk;
        float4 mv = make_float4(0.f, 0.f, 0.f, 0.f), gv = mv, bv = mv;
        if (k < K) {
            int px = m0 + lp;
            if (px < HW) {
                const float* q = mp + (long long)px * ld + k;
                if (m_vec) mv = *reinterpret_cast<const float4*>(q);
                else { mv.x = q[0]; if (k + 1 < K) mv.y = q[1]; if (k + 2 < K) mv.z = q[2]; if (k + 3 < K) mv.w = q[3]; }
            }
            int c = c0 + lp;
            if (c < p.C) {
                const float* qg = wg + (long long)c * K + k;
                const float* qb = wb + (long long)c * K + k;
                if (w_vec) { gv = *reinterpret_cast<const float4*>(qg); bv = *reinterpret_cast<const float4*>(qb); }
                else {
                    gv.x = qg[0]; bv.x = qb[0];
                    if (k + 1 < K) { gv.y = qg[1]; bv.y = qb[1]; }
                    if (k + 2 < K) { gv.z = qg[2]; bv.z = qb[2]; }
                    if (k + 3 < K) { gv.w = qg[3]; bv.w = qb[3]; }
                }
            }
        }
        Ms[lk + 0][lp] = mv.x; Ms[lk + 1][lp] = mv.y; Ms[lk + 2][lp] = mv.z; Ms[lk + 3][lp] = mv.w;
        Gs[lk + 0][lp] = gv.x; Gs[lk + 1][lp] = gv.y; Gs[lk + 2][lp] = gv.z; Gs[lk + 3][lp] = gv.w;
        Bs[lk + 0][lp] = bv.x; Bs[lk + 1][lp] = bv.y; Bs[lk + 2][lp] = bv.z; Bs[lk + 3][lp] = bv.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&Ms[kk][ty * 4]);
            float4 g = *reinterpret_cast<const float4*>(&Gs[kk][tx * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w};
            float gq[4] = {g.x, g.y, g.z, g.w};
            float bq[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int s = 0; s < 4; ++s) { gam[r][s] += av[r] * gq[s]; bet[r][s] += av[r] * bq[s]; }
        }
        __syncthreads();
    }
    // biases may be absent: the reference's adaptive path calls batch_conv(m, weights[0][j]) with the weight
    // tensor only (normalization.py:48-50), i.e. the hyper-network's bias slots are never applied.
    const float* bgp = p.bg[i] ? p.bg[i] + (long long)n * p.w_nstride[i] : nullptr;
    const float* bbp = p.bb[i] ? p.bb[i] + (long long)n * p.w_nstride[i] : nullptr;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        int c = c0 + tx * 4 + s;
        float g0 = (bgp && c < p.C) ? bgp[c] : 0.f, b0 = (bbp && c < p.C) ? bbp[c] : 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) { gam[r][s] += g0; bet[r][s] += b0; }
    }
}

__device__ __forceinline__ void spade_load_xhat(const SpadeP& p, const float* __restrict__ x, const float* __restrict__ mean,
                                                const float* __restrict__ rstd, int n, int m0, int c0, float v[4][4]) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int HW = p.H * p.W, Hs = p.H / p.up, Ws = p.W / p.up;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int px = m0 + ty * 4 + r;
        int h = px / p.W, w = px - h * p.W;
        const float* xr = x + (((long long)n * Hs + h / p.up) * Ws + w / p.up) * p.C;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            int c = c0 + tx * 4 + s;
            float val = 0.f;
            if (px < HW && c < p.C) {
                int si = p.instance ? n * p.C + c : c;
                val = (xr[c] - mean[si]) * rstd[si];
            }
            v[r][s] = val;
        }
    }
}

__global__ void __launch_bounds__(256) k_spade_fwd(SpadeP p, const float* __restrict__ x, const float* __restrict__ mean,
                                                   const float* __restrict__ rstd, float* __restrict__ out) {
    __shared__ __align__(16) float Ms[TK][TM + TPAD];
    __shared__ __align__(16) float Gs[TK][TN + TPAD];
    __shared__ __align__(16) float Bs[TK][TN + TPAD];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int n = blockIdx.z, m0 = blockIdx.x * TM, c0 = blockIdx.y * TN;
    const int HW = p.H * p.W;
    float v[4][4], gam[4][4], bet[4][4];
    spade_load_xhat(p, x, mean, rstd, n, m0, c0, v);
    for (int i = 0; i < p.nmaps; ++i) {
        spade_gemm(p, i, n, m0, c0, HW, Ms, Gs, Bs, gam, bet);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int s = 0; s < 4; ++s) v[r][s] = v[r][s] * (1.f + gam[r][s]) + bet[r][s];
    }
    const bool o_vec = (p.C & 3) == 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int px = m0 + ty * 4 + r;
        if (px >= HW) continue;
        float* o = out + ((long long)n * HW + px) * p.C + c0 + tx * 4;
        if (o_vec && c0 + tx * 4 < p.C) {
            *reinterpret_cast<float4*>(o) = make_float4(fsv_act(v[r][0], p.act), fsv_act(v[r][1], p.act),
                                                        fsv_act(v[r][2], p.act), fsv_act(v[r][3], p.act));
        } else {
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (c0 + tx * 4 + s < p.C) o[s] = fsv_act(v[r][s], p.act);
        }
    }
}

__global__ void __launch_bounds__(256) k_spade_bwd(SpadeP p, const float* __restrict__ x, const float* __restrict__ mean,
                                                   const float* __restrict__ rstd, const float* __restrict__ dout,
                                                   float* __restrict__ dxhat) {
    __shared__ __align__(16) float Ms[TK][TM + TPAD];
    __shared__ __align__(16) float Gs[TK][TN + TPAD];
    __shared__ __align__(16) float Bs[TK][TN + TPAD];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int n = blockIdx.z, m0 = blockIdx.x * TM, c0 = blockIdx.y * TN;
    const int HW = p.H * p.W;
    float v[4][4], bet[4][4];
    float vprev[FSV_SPADE_MAX_MAPS][4][4], gam[FSV_SPADE_MAX_MAPS][4][4];
    spade_load_xhat(p, x, mean, rstd, n, m0, c0, v);
#pragma unroll
    for (int i = 0; i < FSV_SPADE_MAX_MAPS; ++i) {
        if (i < p.nmaps) {
            spade_gemm(p, i, n, m0, c0, HW, Ms, Gs, Bs, gam[i], bet);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    vprev[i][r][s] = v[r][s];
                    v[r][s] = v[r][s] * (1.f + gam[i][r][s]) + bet[r][s];
                }
        }
    }
    // g = dout * act'(v)
    float g[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int px = m0 + ty * 4 + r;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            int c = c0 + tx * 4 + s;
            float d = (px < HW && c < p.C) ? dout[((long long)n * HW + px) * p.C + c] : 0.f;
            float slope = (p.act == FSV_ACT_LRELU) ? (v[r][s] > 0.f ? 1.f : FSV_LRELU_SLOPE) : 1.f;
            g[r][s] = d * slope;
        }
    }
#pragma unroll
    for (int i = FSV_SPADE_MAX_MAPS - 1; i >= 0; --i) {
        if (i < p.nmaps) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int px = m0 + ty * 4 + r;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    int c = c0 + tx * 4 + s;
                    if (px < HW && c < p.C) {
                        long long o = ((long long)n * HW + px) * p.dgb_ld[i] + c;
                        p.dbeta[i][o] = g[r][s];
                        p.dgamma[i][o] = g[r][s] * vprev[i][r][s];
                    }
                    g[r][s] *= (1.f + gam[i][r][s]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int px = m0 + ty * 4 + r;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            int c = c0 + tx * 4 + s;
            if (px < HW && c < p.C) dxhat[((long long)n * HW + px) * p.C + c] = g[r][s];
        }
    }
}

static int fill_p(SpadeP& p, const fsv_spade_desc* d, const float* const* maps, const float* const* wg, const float* const* bg,
                  const float* const* wb, const float* const* bb, const char* who) {
    FSV_REQUIRE(d && d->N > 0 && d->H > 0 && d->W > 0 && d->C > 0, "%s: bad dims", who);
    FSV_REQUIRE(d->up == 1 || (d->up == 2 && d->H % 2 == 0 && d->W % 2 == 0), "%s: bad up", who);
    FSV_REQUIRE(d->nmaps >= 1 && d->nmaps <= FSV_SPADE_MAX_MAPS, "%s: nmaps must be 1..3", who);
    FSV_REQUIRE(d->act == FSV_ACT_NONE || d->act == FSV_ACT_LRELU, "%s: act must be none or lrelu", who);
    p.N = d->N; p.H = d->H; p.W = d->W; p.C = d->C; p.up = d->up; p.instance = d->mode == FSV_NORM_INSTANCE; p.act = d->act;
    p.nmaps = d->nmaps;
    for (int i = 0; i < FSV_SPADE_MAX_MAPS; ++i) {
        bool on = i < d->nmaps;
        p.K[i] = on ? d->K[i] : 0; p.m_ld[i] = on ? d->m_ld[i] : 0; p.m_coff[i] = on ? d->m_coff[i] : 0;
        p.w_nstride[i] = on ? d->w_nstride[i] : 0;
        p.dgb_ld[i] = (on && d->dgb_ld[i] > 0) ? d->dgb_ld[i] : d->C;
        p.maps[i] = on ? maps[i] : nullptr; p.wg[i] = on ? wg[i] : nullptr; p.bg[i] = on ? bg[i] : nullptr;
        p.wb[i] = on ? wb[i] : nullptr; p.bb[i] = on ? bb[i] : nullptr;
        p.dgamma[i] = nullptr; p.dbeta[i] = nullptr;
        if (on) {
            FSV_REQUIRE(p.K[i] > 0 && p.m_ld[i] >= p.m_coff[i] + p.K[i], "%s: map %d ld/coff/K inconsistent", who, i);
            FSV_REQUIRE(maps[i] && wg[i] && wb[i], "%s: map %d has null map/weight pointers", who, i);
        }
    }
    return FSV_OK;
}

extern "C" int fsv_spade_fwd(const fsv_spade_desc* d, const float* x, const float* mean, const float* rstd,
                             const float* const* maps, const float* const* wg, const float* const* bg,
                             const float* const* wb, const float* const* bb, float* out, void* stream) {
    SpadeP p;
    int rc = fill_p(p, d, maps, wg, bg, wb, bb, "spade_fwd");
    if (rc) return rc;
    dim3 grid(fsv_cdiv((long long)d->H * d->W, TM), fsv_cdiv(d->C, TN), d->N);
    k_spade_fwd<<<grid, 256, 0, (cudaStream_t)stream>>>(p, x, mean, rstd, out);
    FSV_CHECK_LAUNCH("spade_fwd");
    return FSV_OK;
}

extern "C" int fsv_spade_bwd(const fsv_spade_desc* d, const float* x, const float* mean, const float* rstd,
                             const float* const* maps, const float* const* wg, const float* const* bg,
                             const float* const* wb, const float* const* bb, const float* dout,
                             float* dxhat, float* const* dgamma, float* const* dbeta, void* stream) {
    SpadeP p;
    int rc = fill_p(p, d, maps, wg, bg, wb, bb, "spade_bwd");
    if (rc) return rc;
    for (int i = 0; i < d->nmaps; ++i) {
        FSV_REQUIRE(dgamma[i] && dbeta[i], "spade_bwd: null dgamma/dbeta for map %d", i);
        p.dgamma[i] = dgamma[i];
        p.dbeta[i] = dbeta[i];
    }
    dim3 grid(fsv_cdiv((long long)d->H * d->W, TM), fsv_cdiv(d->C, TN), d->N);
    k_spade_bwd<<<grid, 256, 0, (cudaStream_t)stream>>>(p, x, mean, rstd, dout, dxhat);
    FSV_CHECK_LAUNCH("spade_bwd");
    return FSV_OK;
}
