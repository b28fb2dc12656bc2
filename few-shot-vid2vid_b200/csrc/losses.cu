// Fused flow / mask losses of the generator step (one pass over the frame forward, one backward) and the fused L1 "feature
// matching" between the two halves of a discriminator feature map.
//
// Reference: models/loss_collector.py:131-204 (compute_flow_losses / compute_mask_losses / compute_mask_loss) + loss.py:130-138
// (MaskedL1Loss) -- about twenty full-frame elementwise kernels per term group in eager PyTorch (sub, abs, sum, clamp, mul, mul,
// sub, abs, mean ... and as many again in backward); loss_collector.py:206-215 (GAN_matching_loss: L1 between the fake and the
// real half of every intermediate discriminator feature, the real half detached).
//
// F_Warp / lambda_flow = sum_k mean|warp_k - tgt|  [+ pose: mean|rbw - body| + mean|rfw - fg|]
// F_Mask / lambda_mask = sum_k ( mean|m_k conf_k| + mean|m_k (1-conf_k) - (1-conf_k)| ),  conf_k = clamp(1 - sum_c |warp_k - tgt|, 0, 1)
//                        [+ pose: mean|m_0 fa| + mean|fake fa - warp_0.detach() fa| + mean|m_0 fgd - fgd| + mean|m_0 bd - bd|,
//                           bd = sum_c |rbw - body| (carries gradient), fa = AvgPool15(face mask), fgd = (ref_fg - fg > 0)]
// Gradients flow to warp_k (through the L1 term and through conf_k), m_k, fake, rbw (L1 term and bd) and rfw.
// HBM-bound streaming kernels; deterministic (per-block fp64 partials summed in a fixed order by a second tiny kernel).
#include "common.cuh"

#define LS_THREADS 256
#define LS_MAXBLOCKS 1024

struct FlowMaskP {
    // per branch k (0 = reference image, 1 = previous frame); NULL warp = branch absent
    const float* warp[2];   // (B,H,W,3) NHWC
    const float* m[2];      // (B,H,W,1)
    const float* tgt;       // (B,3,H,W) NCHW
    const float* fake;      // (B,H,W,3) NHWC or NULL (pose + spade_combine term)
    const float* rbw;       // (B,H,W,9) warped reference body-part masks or NULL
    const float* body;      // (B,H,W,9)
    const float* rfw;       // (B,H,W,1) warped reference foreground mask or NULL
    const float* fg;        // (B,H,W,1)
    const float* fa;        // (B,H,W,1) face average mask or NULL
    const float* fgd;       // (B,H,W,1) foreground disocclusion mask or NULL
    int B, H, W;
};

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) - (v < 0.f); }

__global__ void __launch_bounds__(LS_THREADS) k_flow_mask_fwd(FlowMaskP p, double* __restrict__ part) {
    const long long HW = (long long)p.H * p.W, total = (long long)p.B * HW;
    double a_warp3 = 0, a_body9 = 0, a_one = 0, a_mask1 = 0, a_fake3 = 0;   // sums by normaliser (3N, 9N, N (warp side), N (mask side), 3N (mask side))
    for (long long i = blockIdx.x * (long long)LS_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * LS_THREADS) {
        const long long n = i / HW, px = i - n * HW;
        float t[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c] = p.tgt[(n * 3 + c) * HW + px];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (!p.warp[k]) continue;
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) s += fabsf(p.warp[k][i * 3 + c] - t[c]);
            a_warp3 += s;
            const float conf = fminf(fmaxf(1.f - s, 0.f), 1.f), q = 1.f - conf, m = p.m[k][i];
            a_mask1 += fabsf(m * conf) + fabsf(m * q - q);
        }
        float bd = 0.f;
        if (p.rbw) {
#pragma unroll
            for (int c = 0; c < 9; ++c) bd += fabsf(p.rbw[i * 9 + c] - p.body[i * 9 + c]);
            a_body9 += bd;
        }
        if (p.rfw) a_one += fabsf(p.rfw[i] - p.fg[i]);
        if (p.fa) {
            const float fa = p.fa[i], m0 = p.m[0][i];
            a_mask1 += fabsf(m0 * fa);
            if (p.fake) {
#pragma unroll
                for (int c = 0; c < 3; ++c) a_fake3 += fabsf(p.fake[i * 3 + c] * fa - p.warp[0][i * 3 + c] * fa);
            }
            const float fgd = p.fgd[i];
            a_mask1 += fabsf(m0 * fgd - fgd);
            if (p.rbw) a_mask1 += fabsf(m0 * bd - bd);
        }
    }
    __shared__ double sh[5][LS_THREADS / 32];
    double v[5] = {a_warp3, a_body9, a_one, a_mask1, a_fake3};
#pragma unroll
    for (int j = 0; j < 5; ++j) v[j] = warp_sum_d(v[j]);
    if ((threadIdx.x & 31) == 0)
        for (int j = 0; j < 5; ++j) sh[j][threadIdx.x >> 5] = v[j];
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0;
        for (int w = 0; w < LS_THREADS / 32; ++w) s += sh[threadIdx.x][w];
        part[blockIdx.x * 5 + threadIdx.x] = s;
    }
}

__global__ void k_flow_mask_final(const double* __restrict__ part, int nblocks, double n1, float* __restrict__ out) {
    // one warp: lane j < 5 sums column j in block order
    __shared__ double tot[5];
    if (threadIdx.x < 5) {
        double s = 0;
        for (int b = 0; b < nblocks; ++b) s += part[b * 5 + threadIdx.x];
        tot[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        out[0] = (float)(tot[0] / (3.0 * n1) + tot[1] / (9.0 * n1) + tot[2] / n1);     // F_Warp / lambda_flow
        out[1] = (float)(tot[3] / n1 + tot[4] / (3.0 * n1));                           // F_Mask / lambda_mask
    }
}

// gw = d(total)/d(F_Warp term), gm = d(total)/d(F_Mask term) (device scalars, lambdas already applied by the caller's graph)
__global__ void __launch_bounds__(LS_THREADS) k_flow_mask_bwd(FlowMaskP p, const float* __restrict__ gwp, const float* __restrict__ gmp,
                                                              float* __restrict__ dwarp0, float* __restrict__ dwarp1, float* __restrict__ dm0,
                                                              float* __restrict__ dm1, float* __restrict__ dfake, float* __restrict__ drbw,
                                                              float* __restrict__ drfw) {
    const long long HW = (long long)p.H * p.W, total = (long long)p.B * HW;
    const float n1 = (float)total;
    const float gw = *gwp, gm = *gmp;
    for (long long i = blockIdx.x * (long long)LS_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * LS_THREADS) {
        const long long n = i / HW, px = i - n * HW;
        float t[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c] = p.tgt[(n * 3 + c) * HW + px];
        float dm0v = 0.f;
        float bd = 0.f;
        if (p.rbw) {
#pragma unroll
            for (int c = 0; c < 9; ++c) bd += fabsf(p.rbw[i * 9 + c] - p.body[i * 9 + c]);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (!p.warp[k]) continue;
            float d[3], s = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) { d[c] = p.warp[k][i * 3 + c] - t[c]; s += fabsf(d[c]); }
            const float u = 1.f - s;
            const float conf = fminf(fmaxf(u, 0.f), 1.f), q = 1.f - conf, m = p.m[k][i];
            const float sa = sgn(m * conf), sb = sgn(m * q - q);
            float dmk = gm * (sa * conf + sb * q) / n1;
            // d/dconf of |m conf| + |m q - q| with q = 1 - conf:  sa m - sb (m - 1);  clamp passes gradient for 0 <= u <= 1
            const float dconf = (u >= 0.f && u <= 1.f) ? gm * (sa * m - sb * (m - 1.f)) / n1 : 0.f;
            float* dw = k == 0 ? dwarp0 : dwarp1;
#pragma unroll
            for (int c = 0; c < 3; ++c) dw[i * 3 + c] = sgn(d[c]) * (gw / (3.f * n1) - dconf);
            if (k == 0) dm0v = dmk; else dm1[i] = dmk;
        }
        float dbd = 0.f;
        if (p.fa) {
            const float fa = p.fa[i], m0 = p.m[0][i], fgd = p.fgd[i];
            dm0v += gm * (sgn(m0 * fa) * fa + sgn(m0 * fgd - fgd) * fgd) / n1;
            if (p.rbw) {
                const float se = sgn(m0 * bd - bd);
                dm0v += gm * se * bd / n1;
                dbd = gm * se * (m0 - 1.f) / n1;
            }
            if (p.fake) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    dfake[i * 3 + c] = gm * sgn(p.fake[i * 3 + c] * fa - p.warp[0][i * 3 + c] * fa) * fa / (3.f * n1);
            }
        }
        if (p.warp[0]) dm0[i] = dm0v;
        if (p.rbw) {
#pragma unroll
            for (int c = 0; c < 9; ++c) drbw[i * 9 + c] = sgn(p.rbw[i * 9 + c] - p.body[i * 9 + c]) * (gw / (9.f * n1) + dbd);
        }
        if (p.rfw) drfw[i] = gw * sgn(p.rfw[i] - p.fg[i]) / n1;
    }
}

static int ls_blocks(long long total) {
    long long b = (total + LS_THREADS - 1) / LS_THREADS, cap = 4LL * fsv_sm_count();
    if (b > cap) b = cap;
    if (b > LS_MAXBLOCKS) b = LS_MAXBLOCKS;
    return (int)(b < 1 ? 1 : b);
}

static FlowMaskP make_fm(const fsv_flow_mask_desc* d) {
    FlowMaskP p;
    p.warp[0] = d->warp0; p.warp[1] = d->warp1; p.m[0] = d->mask0; p.m[1] = d->mask1; p.tgt = d->tgt; p.fake = d->fake;
    p.rbw = d->ref_body_warp; p.body = d->body; p.rfw = d->ref_fg_warp; p.fg = d->fg; p.fa = d->face_avg; p.fgd = d->fg_diff;
    p.B = d->B; p.H = d->H; p.W = d->W;
    return p;
}

extern "C" long long fsv_flow_mask_loss_work_doubles(void) { return (long long)LS_MAXBLOCKS * 5; }

extern "C" int fsv_flow_mask_loss_fwd(const fsv_flow_mask_desc* d, float* out2, double* work, void* stream) {
    FSV_REQUIRE(d && out2 && work && d->B > 0 && d->H > 0 && d->W > 0 && d->tgt, "flow_mask_loss_fwd: bad args");
    FSV_REQUIRE((!d->warp0 || d->mask0) && (!d->warp1 || d->mask1), "flow_mask_loss_fwd: a warped frame needs its mask");
    FSV_REQUIRE(!d->face_avg || (d->warp0 && d->fg_diff), "flow_mask_loss_fwd: the pose terms need the reference branch and fg_diff");
    FSV_REQUIRE(!d->ref_body_warp || d->body, "flow_mask_loss_fwd: body masks missing");
    FSV_REQUIRE(!d->ref_fg_warp || d->fg, "flow_mask_loss_fwd: fg mask missing");
    const long long total = (long long)d->B * d->H * d->W;
    const int nb = ls_blocks(total);
    k_flow_mask_fwd<<<nb, LS_THREADS, 0, (cudaStream_t)stream>>>(make_fm(d), work);
    FSV_CHECK_LAUNCH("flow_mask_loss_fwd");
    k_flow_mask_final<<<1, 32, 0, (cudaStream_t)stream>>>(work, nb, (double)total, out2);
    FSV_CHECK_LAUNCH("flow_mask_loss_final");
    return FSV_OK;
}

extern "C" int fsv_flow_mask_loss_bwd(const fsv_flow_mask_desc* d, const float* g_warp, const float* g_mask, float* dwarp0, float* dwarp1,
                                      float* dmask0, float* dmask1, float* dfake, float* dref_body_warp, float* dref_fg_warp, void* stream) {
    FSV_REQUIRE(d && g_warp && g_mask, "flow_mask_loss_bwd: bad args");
    FSV_REQUIRE((!d->warp0 || (dwarp0 && dmask0)) && (!d->warp1 || (dwarp1 && dmask1)), "flow_mask_loss_bwd: missing gradient buffers");
    FSV_REQUIRE((!(d->fake && d->face_avg) || dfake) && (!d->ref_body_warp || dref_body_warp) && (!d->ref_fg_warp || dref_fg_warp),
                "flow_mask_loss_bwd: missing gradient buffers (pose terms)");
    const long long total = (long long)d->B * d->H * d->W;
    k_flow_mask_bwd<<<ls_blocks(total), LS_THREADS, 0, (cudaStream_t)stream>>>(make_fm(d), g_warp, g_mask, dwarp0, dwarp1, dmask0, dmask1, dfake,
                                                                              dref_body_warp, dref_fg_warp);
    FSV_CHECK_LAUNCH("flow_mask_loss_bwd");
    return FSV_OK;
}

// ------------------------------------------------------------------------------------------------ feature matching
// x: one discriminator feature map for the batch [fake ; real] (2B samples, any layout, `half` = elements per half).
// out[0] = mean |x[:half] - x[half:]|;  backward: dx[:half] = g * sign(...) / half, dx[half:] = 0 (the real half is detached).
__global__ void __launch_bounds__(LS_THREADS) k_halves_l1_fwd(const float* __restrict__ x, long long half, double* __restrict__ part) {
    double a = 0;
    for (long long i = blockIdx.x * (long long)LS_THREADS + threadIdx.x; i < half; i += (long long)gridDim.x * LS_THREADS)
        a += fabsf(x[i] - x[half + i]);
    __shared__ double sh[LS_THREADS / 32];
    a = warp_sum_d(a);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int w = 0; w < LS_THREADS / 32; ++w) s += sh[w];
        part[blockIdx.x] = s;
    }
}
__global__ void k_halves_l1_final(const double* __restrict__ part, int nblocks, double half, float* __restrict__ out) {
    if (threadIdx.x == 0) {
        double s = 0;
        for (int b = 0; b < nblocks; ++b) s += part[b];
        out[0] = (float)(s / half);
    }
}
__global__ void __launch_bounds__(LS_THREADS) k_halves_l1_bwd(const float* __restrict__ x, long long half, const float* __restrict__ g,
                                                              float* __restrict__ dx) {
    const float gv = *g / (float)half;
    for (long long i = blockIdx.x * (long long)LS_THREADS + threadIdx.x; i < half; i += (long long)gridDim.x * LS_THREADS) {
        dx[i] = gv * sgn(x[i] - x[half + i]);
        dx[half + i] = 0.f;
    }
}

extern "C" int fsv_halves_l1_fwd(const float* x, long long half, float* out, double* work, void* stream) {
    FSV_REQUIRE(x && out && work && half > 0, "halves_l1_fwd: bad args");
    const int nb = ls_blocks(half);
    k_halves_l1_fwd<<<nb, LS_THREADS, 0, (cudaStream_t)stream>>>(x, half, work);
    FSV_CHECK_LAUNCH("halves_l1_fwd");
    k_halves_l1_final<<<1, 32, 0, (cudaStream_t)stream>>>(work, nb, (double)half, out);
    FSV_CHECK_LAUNCH("halves_l1_final");
    return FSV_OK;
}
extern "C" int fsv_halves_l1_bwd(const float* x, long long half, const float* g, float* dx, void* stream) {
    FSV_REQUIRE(x && g && dx && half > 0, "halves_l1_bwd: bad args");
    k_halves_l1_bwd<<<ls_blocks(half), LS_THREADS, 0, (cudaStream_t)stream>>>(x, half, g, dx);
    FSV_CHECK_LAUNCH("halves_l1_bwd");
    return FSV_OK;
}
