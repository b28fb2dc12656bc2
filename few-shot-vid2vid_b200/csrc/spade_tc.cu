// Fused SPADE forward on tcgen05 tensor cores (TF32 gamma/beta GEMM, fp32 modulation).  sm_100a only.
//
// Same contract as fsv_spade_fwd (spade.cu; reference normalization.py:37-52 + architecture.py:96-97), for the shapes
// that carry the cost: C % 64 == 0 and every label map with K % 32 == 0.  The SIMT kernel is FFMA-bound on the
// per-pixel 1x1 GEMM (measured 341 us = 0.78 TB/s at up_0, 8.6 GFMA); here that GEMM runs on the tensor cores
// and the kernel goes back to being the HBM stream it should be:
//
//   per CTA: 128 pixels (TN x TH x TW patch) x 64 channels
//     for each label map i, each 32-channel K block:  TMA  map tile (128 px x 32)            -> smem (K-major, SW128)
//                                                     TMA  Wgamma_i[64 x 32], Wbeta_i[64 x 32] -> smem (one 128-row B tile)
//                                                     tcgen05.mma  D_i[128 px x (64 gamma | 64 beta)] += A * B^T   (TMEM)
//     epilogue (4 warps, thread = pixel): v = (x[n, h/up, w/up, c] - mean) * rstd   (nearest x2 upsample folded into the load)
//                                          for each map: v = v * (1 + gamma_i + bg_i) + beta_i + bb_i   (gamma/beta from TMEM)
//                                          out = LeakyReLU(v)                                        (16-byte stores)
// Per-sample hyper-weights (map 0 of the adaptive layers) are addressed through a 3-D tensor map over the
// hyper-network's flat output (K, C, sample) -- still zero-copy.  gamma/beta never touch HBM.
// Roofline: HBM, algorithmic bytes 4*(|x|/up^2 + sum|map_i| + |out|) per launch (weights: L2-resident).
#include "tc_common.cuh"
#include <stdlib.h>

#define SP_STAGES 3
// channels per CTA: template parameter SP_CB (64, or 32 for the C=32 layer at full resolution)
#define SP_STAGE_BYTES_OF(CB) (TC_A_BYTES + 2 * (CB) * TC_BK * 4)   // 16 KB map tile + gamma rows + beta rows

struct __align__(64) SpTcParams {
    CUtensorMap mmap[FSV_SPADE_MAX_MAPS];
    CUtensorMap gmap[FSV_SPADE_MAX_MAPS];
    CUtensorMap bmap[FSV_SPADE_MAX_MAPS];
    const float* bg[FSV_SPADE_MAX_MAPS];
    const float* bb[FSV_SPADE_MAX_MAPS];
    long long b_nstride[FSV_SPADE_MAX_MAPS];
    int K[FSV_SPADE_MAX_MAPS], per_sample[FSV_SPADE_MAX_MAPS];
    int nmaps, N, H, W, C, up, instance, act;
    int TW, TH, TN, tiles_w, tiles_h;
    // backward only
    float* dgamma[FSV_SPADE_MAX_MAPS];
    float* dbeta[FSV_SPADE_MAX_MAPS];
    int dgb_ld[FSV_SPADE_MAX_MAPS];
};

template <int SP_CB, bool BWD, int MINB = (BWD ? 1 : 2)>
__global__ void __launch_bounds__(192, MINB) k_spade_tc(const __grid_constant__ SpTcParams p, const float* __restrict__ x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     float* __restrict__ out, const float* __restrict__ dout,
                                                     float* __restrict__ dxhat) {
    constexpr int SP_STAGE_BYTES = SP_STAGE_BYTES_OF(SP_CB);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + SP_STAGES * SP_STAGE_BYTES);   // full[S], empty[S], tmem_full
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * SP_STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int mt = blockIdx.x;
    const int tw_i = mt % p.tiles_w; mt /= p.tiles_w;
    const int th_i = mt % p.tiles_h; mt /= p.tiles_h;
    const int n0 = mt * p.TN, h0 = th_i * p.TH, w0 = tw_i * p.TW;
    const int c0 = blockIdx.y * SP_CB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SP_STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[SP_STAGES + s]), 1);
        }
        mbar_init(smem_u32(&bars[2 * SP_STAGES]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(p.nmaps * 2 * SP_CB)) tmem_cols <<= 1;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    int num_k = 0;
    for (int i = 0; i < p.nmaps; ++i) num_k += p.K[i] / TC_BK;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int i = 0; i < p.nmaps; ++i) {
                const int kbs = p.K[i] / TC_BK;
                const int wn = p.per_sample[i] ? n0 : 0;
                for (int kb = 0; kb < kbs; ++kb, ++it) {
                    const int s = it % SP_STAGES;
                    const uint32_t ph = (it / SP_STAGES) & 1;
                    mbar_wait(smem_u32(&bars[SP_STAGES + s]), ph ^ 1);
                    const uint32_t full = smem_u32(&bars[s]);
                    const uint32_t dst = smem_u32(smem + s * SP_STAGE_BYTES);
                    mbar_expect_tx(full, (uint32_t)SP_STAGE_BYTES);
                    tma_load_4d(dst, &p.mmap[i], full, kb * TC_BK, w0, h0, n0);
                    tma_load_3d(dst + TC_A_BYTES, &p.gmap[i], full, kb * TC_BK, c0, wn);
                    tma_load_3d(dst + TC_A_BYTES + SP_CB * TC_BK * 4, &p.bmap[i], full, kb * TC_BK, c0, wn);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_BM, 2 * SP_CB);
            int it = 0;
            for (int i = 0; i < p.nmaps; ++i) {
                const int kbs = p.K[i] / TC_BK;
                for (int kb = 0; kb < kbs; ++kb, ++it) {
                    const int s = it % SP_STAGES;
                    const uint32_t ph = (it / SP_STAGES) & 1;
                    mbar_wait(smem_u32(&bars[s]), ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * SP_STAGE_BYTES);
                    const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                    const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + TC_A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k)
                        tc_mma_tf32(tmem_base + (uint32_t)(i * 2 * SP_CB), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                    (kb | k) != 0);
                    tc_commit(smem_u32(&bars[SP_STAGES + s]));
                }
            }
            tc_commit(smem_u32(&bars[2 * SP_STAGES]));
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int tw = row % p.TW, r2 = row / p.TW;
        const int th = r2 % p.TH, tn = r2 / p.TH;
        const int n = n0 + tn, h = h0 + th, w = w0 + tw;
        const bool valid = (n < p.N) && (h < p.H) && (w < p.W);
        const int Hs = p.H / p.up, Ws = p.W / p.up;
        if constexpr (!BWD) {
        // ===== forward epilogue: thread = pixel of the tile; while the MMAs run, fetch and normalise x
        float v[SP_CB];
        if (valid) {
            const float4* xr = reinterpret_cast<const float4*>(x + (((long long)n * Hs + h / p.up) * Ws + w / p.up) * p.C + c0);
            const float* mp = mean + (p.instance ? n * p.C : 0) + c0;
            const float* rp = rstd + (p.instance ? n * p.C : 0) + c0;
#pragma unroll
            for (int j = 0; j < SP_CB / 4; ++j) {
                float4 t = xr[j];
                float4 m = *reinterpret_cast<const float4*>(mp + 4 * j);
                float4 r = *reinterpret_cast<const float4*>(rp + 4 * j);
                v[4 * j + 0] = (t.x - m.x) * r.x; v[4 * j + 1] = (t.y - m.y) * r.y;
                v[4 * j + 2] = (t.z - m.z) * r.z; v[4 * j + 3] = (t.w - m.w) * r.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < SP_CB; ++j) v[j] = 0.f;
        }
        mbar_wait(smem_u32(&bars[2 * SP_STAGES]), 0);
        tc_fence_after();
        for (int i = 0; i < p.nmaps; ++i) {
            const float* bgp = p.bg[i] ? p.bg[i] + (long long)(p.per_sample[i] ? n : 0) * p.b_nstride[i] + c0 : nullptr;
            const float* bbp = p.bb[i] ? p.bb[i] + (long long)(p.per_sample[i] ? n : 0) * p.b_nstride[i] + c0 : nullptr;
#pragma unroll
            for (int c = 0; c < SP_CB; c += 32) {
                uint32_t g[32], b[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(i * 2 * SP_CB + c);
                tc_ld32(taddr, g);
                tc_ld32(taddr + SP_CB, b);
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float gv = __uint_as_float(g[j]) + (bgp ? bgp[c + j] : 0.f);
                        float bv = __uint_as_float(b[j]) + (bbp ? bbp[c + j] : 0.f);
                        v[c + j] = v[c + j] * (1.f + gv) + bv;
                    }
                }
            }
        }
        if (valid) {
            float4* orow = reinterpret_cast<float4*>(out + (((long long)n * p.H + h) * p.W + w) * p.C + c0);
#pragma unroll
            for (int j = 0; j < SP_CB / 4; ++j)
                orow[j] = make_float4(fsv_act(v[4 * j], p.act), fsv_act(v[4 * j + 1], p.act), fsv_act(v[4 * j + 2], p.act),
                                      fsv_act(v[4 * j + 3], p.act));
        }
        } else {
        // ===== backward epilogue: per 32-channel chunk, rebuild the modulation chain from gamma/beta in TMEM, then walk it
        // backwards: dbeta_i = g, dgamma_i = g * v_{i-1}, g *= (1 + gamma_i); finally dxhat = g.
        mbar_wait(smem_u32(&bars[2 * SP_STAGES]), 0);
        tc_fence_after();
        const long long pix = ((long long)n * p.H + h) * p.W + w;
#pragma unroll 1
        for (int c = 0; c < SP_CB; c += 32) {
            float v[32], vprev[FSV_SPADE_MAX_MAPS][32];
            if (valid) {
                const float4* xr = reinterpret_cast<const float4*>(x + (((long long)n * Hs + h / p.up) * Ws + w / p.up) * p.C + c0 + c);
                const float* mp = mean + (p.instance ? n * p.C : 0) + c0 + c;
                const float* rp = rstd + (p.instance ? n * p.C : 0) + c0 + c;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 t = xr[j];
                    float4 m = *reinterpret_cast<const float4*>(mp + 4 * j);
                    float4 r = *reinterpret_cast<const float4*>(rp + 4 * j);
                    v[4 * j + 0] = (t.x - m.x) * r.x; v[4 * j + 1] = (t.y - m.y) * r.y;
                    v[4 * j + 2] = (t.z - m.z) * r.z; v[4 * j + 3] = (t.w - m.w) * r.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < FSV_SPADE_MAX_MAPS; ++i) {
                if (i < p.nmaps) {
                    const float* bgp = p.bg[i] ? p.bg[i] + (long long)(p.per_sample[i] ? n : 0) * p.b_nstride[i] + c0 + c : nullptr;
                    const float* bbp = p.bb[i] ? p.bb[i] + (long long)(p.per_sample[i] ? n : 0) * p.b_nstride[i] + c0 + c : nullptr;
                    uint32_t g[32], b[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(i * 2 * SP_CB + c);
                    tc_ld32(taddr, g);
                    tc_ld32(taddr + SP_CB, b);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float gv = __uint_as_float(g[j]) + ((bgp && valid) ? bgp[j] : 0.f);
                        float bv = __uint_as_float(b[j]) + ((bbp && valid) ? bbp[j] : 0.f);
                        vprev[i][j] = v[j];
                        v[j] = v[j] * (1.f + gv) + bv;
                    }
                }
            }
            float gr[32];
            if (valid) {
                const float4* dr = reinterpret_cast<const float4*>(dout + pix * p.C + c0 + c);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 t = dr[j];
                    gr[4 * j + 0] = t.x; gr[4 * j + 1] = t.y; gr[4 * j + 2] = t.z; gr[4 * j + 3] = t.w;
                }
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (p.act == FSV_ACT_LRELU) gr[j] *= (v[j] > 0.f ? 1.f : FSV_LRELU_SLOPE);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) gr[j] = 0.f;
            }
#pragma unroll
            for (int i = FSV_SPADE_MAX_MAPS - 1; i >= 0; --i) {
                if (i < p.nmaps) {
                    const float* bgp = p.bg[i] ? p.bg[i] + (long long)(p.per_sample[i] ? n : 0) * p.b_nstride[i] + c0 + c : nullptr;
                    uint32_t g[32];
                    tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(i * 2 * SP_CB + c), g);
                    if (valid) {
                        float4* db = reinterpret_cast<float4*>(p.dbeta[i] + pix * p.dgb_ld[i] + c0 + c);
                        float4* dg = reinterpret_cast<float4*>(p.dgamma[i] + pix * p.dgb_ld[i] + c0 + c);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            db[j] = make_float4(gr[4 * j], gr[4 * j + 1], gr[4 * j + 2], gr[4 * j + 3]);
                            dg[j] = make_float4(gr[4 * j] * vprev[i][4 * j], gr[4 * j + 1] * vprev[i][4 * j + 1],
                                                gr[4 * j + 2] * vprev[i][4 * j + 2], gr[4 * j + 3] * vprev[i][4 * j + 3]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float gv = __uint_as_float(g[j]) + ((bgp && valid) ? bgp[j] : 0.f);
                        gr[j] *= (1.f + gv);
                    }
                }
            }
            if (valid) {
                float4* dxr = reinterpret_cast<float4*>(dxhat + pix * p.C + c0 + c);
#pragma unroll
                for (int j = 0; j < 8; ++j) dxr[j] = make_float4(gr[4 * j], gr[4 * j + 1], gr[4 * j + 2], gr[4 * j + 3]);
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
}

// ------------------------------------------------------------------ persistent variant (FSV_SPADE_PERSIST)
// The kernel above lives for ONE 128-pixel x SP_CB-channel tile: barrier init, TMEM allocation, one or two TMA round trips, a handful of
// MMAs, the epilogue -- about 10 us per CTA for ~73 KB of HBM traffic, two CTAs per SM (round-2 timeline: 145 - 152 us for the
// 512x512x64 layers = 2.0 TB/s).  Here one CTA per SM walks tiles (tile = blockIdx.x, + gridDim.x, ...) with the protocol of
// conv_tc.cu's k_conv_tc_p: the operand ring runs continuously across tiles and the gamma/beta accumulators are double-buffered in
// TMEM (2 x nmaps*2*SP_CB columns <= 512), so the TMA + MMA of tile i+1 overlap the epilogue of tile i.  Two epilogue groups of four
// warps (one per TMEM buffer) alternate tiles, so one group's x / dout loads are in flight while the other computes and stores.
//   full[s] / empty[s]  : TMA -> MMA ring, stage counter carried across tiles
//   acc_full[b]         : tcgen05.commit after the last MMA of a tile -> epilogue group b
//   acc_empty[b]        : one arrive per warp of group b after its last tcgen05.ld of the tile -> the MMA issuer may overwrite buffer b
// Tile order: channel blocks of one pixel tile are consecutive tile indices, i.e. they run on neighbouring CTAs at the same time and the
// map tile they share is fetched from HBM once.
// NGRP epilogue groups: 2 for the forward (168 registers per thread fit 320 threads); the backward epilogue needs ~250 registers, so it runs
// one group of four warps that alternates between the two accumulator buffers.
template <int SP_CB, bool BWD, int NGRP>
__global__ void __launch_bounds__(64 + 128 * NGRP, 1) k_spade_tc_p(const __grid_constant__ SpTcParams p, const float* __restrict__ x,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             float* __restrict__ out, const float* __restrict__ dout,
                                                             float* __restrict__ dxhat, int stages, int m_tiles) {
    constexpr int SP_STAGE_BYTES = SP_STAGE_BYTES_OF(SP_CB);
    constexpr int MAXM = SP_CB == 64 ? 2 : FSV_SPADE_MAX_MAPS;     // the host admits nmaps * 2 * SP_CB <= 256 only
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int STG = stages;
    // forward: per epilogue warp a 32-row x SP_CB staging area through which the output tile is transposed, so that a store instruction
    // writes whole pixels (SP_CB*4 contiguous bytes each) instead of 16 bytes of 32 different pixels (ncu, session 17: half of every
    // written sector unused, 32 sectors per store request)
    constexpr int SPP_STAGING = BWD ? 0 : NGRP * 4 * 32 * SP_CB * 4;
    uint8_t* stage_base = smem + STG * SP_STAGE_BYTES;
    uint64_t* bars = (uint64_t*)(stage_base + SPP_STAGING);         // full[STG], empty[STG], acc_full[2], acc_empty[2]
    uint64_t* acc_full = bars + 2 * STG;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncb = p.C / SP_CB;
    const int total = m_tiles * ncb;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STG; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[STG + s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&acc_full[b]), 1);
            mbar_init(smem_u32(&acc_empty[b]), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    uint32_t accw = 32;
    while (accw < (uint32_t)(p.nmaps * 2 * SP_CB)) accw <<= 1;
    const uint32_t tmem_cols = 2 * accw;                             // host guarantees <= 512
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                int mt = tile / ncb;
                const int c0 = (tile - mt * ncb) * SP_CB;
                const int tw_i = mt % p.tiles_w; mt /= p.tiles_w;
                const int th_i = mt % p.tiles_h; mt /= p.tiles_h;
                const int n0 = mt * p.TN, h0 = th_i * p.TH, w0 = tw_i * p.TW;
                for (int i = 0; i < p.nmaps; ++i) {
                    const int kbs = p.K[i] / TC_BK;
                    const int wn = p.per_sample[i] ? n0 : 0;
                    for (int kb = 0; kb < kbs; ++kb, ++it) {
                        const int s = it % STG;
                        const uint32_t ph = (it / STG) & 1;
                        mbar_wait(smem_u32(&bars[STG + s]), ph ^ 1);
                        const uint32_t full = smem_u32(&bars[s]);
                        const uint32_t dst = smem_u32(smem + s * SP_STAGE_BYTES);
                        mbar_expect_tx(full, (uint32_t)SP_STAGE_BYTES);
                        tma_load_4d(dst, &p.mmap[i], full, kb * TC_BK, w0, h0, n0);
                        tma_load_3d(dst + TC_A_BYTES, &p.gmap[i], full, kb * TC_BK, c0, wn);
                        tma_load_3d(dst + TC_A_BYTES + SP_CB * TC_BK * 4, &p.bmap[i], full, kb * TC_BK, c0, wn);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_BM, 2 * SP_CB);
            uint32_t it = 0, ti = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++ti) {
                const uint32_t buf = ti & 1;
                mbar_wait(smem_u32(&acc_empty[buf]), ((ti >> 1) & 1) ^ 1);      // group `buf` has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * accw;
                for (int i = 0; i < p.nmaps; ++i) {
                    const int kbs = p.K[i] / TC_BK;
                    for (int kb = 0; kb < kbs; ++kb, ++it) {
                        const int s = it % STG;
                        const uint32_t ph = (it / STG) & 1;
                        mbar_wait(smem_u32(&bars[s]), ph);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(smem + s * SP_STAGE_BYTES);
                        const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                        const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + TC_A_BYTES);
#pragma unroll
                        for (int k = 0; k < TC_BK / 8; ++k)
                            tc_mma_tf32(d_tmem + (uint32_t)(i * 2 * SP_CB), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                        (kb | k) != 0);
                        tc_commit(smem_u32(&bars[STG + s]));
                    }
                }
                tc_commit(smem_u32(&acc_full[buf]));
            }
        }
    } else {
        const int grp = (warp - 2) >> 2;              // epilogue group = TMEM buffer
        const int q = warp & 3;                       // TMEM lane quadrant this warp may read
        const int row = q * 32 + lane;
        const int tw = row % p.TW, r2 = row / p.TW;
        const int th = r2 % p.TH, tn = r2 / p.TH;
        const int Hs = p.H / p.up, Ws = p.W / p.up;
        uint32_t ti = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++ti) {
            const int buf = (int)(ti & 1);
            if (NGRP == 2 && buf != grp) continue;
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * accw;
            int mt = tile / ncb;
            const int c0 = (tile - mt * ncb) * SP_CB;
            const int tw_i = mt % p.tiles_w; mt /= p.tiles_w;
            const int th_i = mt % p.tiles_h; mt /= p.tiles_h;
            const int n = mt * p.TN + tn, h = th_i * p.TH + th, w = tw_i * p.TW + tw;
            const bool valid = (n < p.N) && (h < p.H) && (w < p.W);
            const uint32_t par = (ti >> 1) & 1;
            if constexpr (!BWD) {
                float v[SP_CB];
                if (valid) {
                    const float4* xr = reinterpret_cast<const float4*>(x + (((long long)n * Hs + h / p.up) * Ws + w / p.up) * p.C + c0);
                    const float* mp = mean + (p.instance ? n * p.C : 0) + c0;
                    const float* rp = rstd + (p.instance ? n * p.C : 0) + c0;
#pragma unroll
                    for (int j = 0; j < SP_CB / 4; ++j) {
                        float4 t = xr[j];
                        float4 m = *reinterpret_cast<const float4*>(mp + 4 * j);
                        float4 r = *reinterpret_cast<const float4*>(rp + 4 * j);
                        v[4 * j + 0] = (t.x - m.x) * r.x; v[4 * j + 1] = (t.y - m.y) * r.y;
                        v[4 * j + 2] = (t.z - m.z) * r.z; v[4 * j + 3] = (t.w - m.w) * r.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < SP_CB; ++j) v[j] = 0.f;
                }
                mbar_wait(smem_u32(&acc_full[buf]), par);
                tc_fence_after();
                for (int i = 0; i < p.nmaps; ++i) {
                    const float* bgp = p.bg[i] ? p.bg[i] + (long long)(p.per_sample[i] ? n : 0) * p.b_nstride[i] + c0 : nullptr;
                    const float* bbp = p.bb[i] ? p.bb[i] + (long long)(p.per_sample[i] ? n : 0) * p.b_nstride[i] + c0 : nullptr;
#pragma unroll
                    for (int c = 0; c < SP_CB; c += 32) {
                        uint32_t g[32], b[32];
                        const uint32_t taddr = tacc + (uint32_t)(i * 2 * SP_CB + c);
                        tc_ld32(taddr, g);
                        tc_ld32(taddr + SP_CB, b);
                        if (valid) {
                            // biases: per-channel constants, 16-byte loads (the scalar form issued 4 x 32 loads per map and chunk -- 3/4 of
                            // this kernel's load instructions in the ncu capture of session 17)
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 g4 = bgp ? *reinterpret_cast<const float4*>(bgp + c + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                                const float4 b4 = bbp ? *reinterpret_cast<const float4*>(bbp + c + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                                v[c + j + 0] = v[c + j + 0] * (1.f + (__uint_as_float(g[j + 0]) + g4.x)) + (__uint_as_float(b[j + 0]) + b4.x);
                                v[c + j + 1] = v[c + j + 1] * (1.f + (__uint_as_float(g[j + 1]) + g4.y)) + (__uint_as_float(b[j + 1]) + b4.y);
                                v[c + j + 2] = v[c + j + 2] * (1.f + (__uint_as_float(g[j + 2]) + g4.z)) + (__uint_as_float(b[j + 2]) + b4.z);
                                v[c + j + 3] = v[c + j + 3] * (1.f + (__uint_as_float(g[j + 3]) + g4.w)) + (__uint_as_float(b[j + 3]) + b4.w);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));      // accumulator drained: the MMAs of the next tile on this buffer may start
                {
                    constexpr int LPR = SP_CB / 4;          // lanes per pixel row (float4 each)
                    constexpr int RPI = 32 / LPR;           // pixel rows per store instruction
                    float4* stg = reinterpret_cast<float4*>(stage_base) + (warp - 2) * (32 * LPR);
                    const int sw = lane & 7;                // XOR swizzle of the 16-byte column: conflict-free writes and reads
#pragma unroll
                    for (int j = 0; j < LPR; ++j)
                        stg[lane * LPR + (j ^ sw)] = make_float4(fsv_act(v[4 * j], p.act), fsv_act(v[4 * j + 1], p.act), fsv_act(v[4 * j + 2], p.act),
                                                                 fsv_act(v[4 * j + 3], p.act));
                    const long long pixoff = valid ? (((long long)n * p.H + h) * p.W + w) * p.C + c0 : -1;
                    __syncwarp();
#pragma unroll 4
                    for (int it = 0; it < 32 / RPI; ++it) {
                        const int rr = it * RPI + lane / LPR, c4 = lane % LPR;
                        const long long off = __shfl_sync(0xffffffffu, pixoff, rr);
                        const float4 t = stg[rr * LPR + (c4 ^ (rr & 7))];
                        if (off >= 0) *reinterpret_cast<float4*>(out + off + c4 * 4) = t;
                    }
                    __syncwarp();
                }
            } else {
                // backward epilogue: identical to k_spade_tc<.., true>, per 32-channel chunk; gamma is read twice (forward walk, reverse walk)
                mbar_wait(smem_u32(&acc_full[buf]), par);
                tc_fence_after();
                const long long pix = ((long long)n * p.H + h) * p.W + w;
#pragma unroll 1
                for (int c = 0; c < SP_CB; c += 32) {
                    float v[32], vprev[MAXM][32];
                    if (valid) {
                        const float4* xr = reinterpret_cast<const float4*>(x + (((long long)n * Hs + h / p.up) * Ws + w / p.up) * p.C + c0 + c);
                        const float* mp = mean + (p.instance ? n * p.C : 0) + c0 + c;
                        const float* rp = rstd + (p.instance ? n * p.C : 0) + c0 + c;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 t = xr[j];
                            float4 m = *reinterpret_cast<const float4*>(mp + 4 * j);
                            float4 r = *reinterpret_cast<const float4*>(rp + 4 * j);
                            v[4 * j + 0] = (t.x - m.x) * r.x; v[4 * j + 1] = (t.y - m.y) * r.y;
                            v[4 * j + 2] = (t.z - m.z) * r.z; v[4 * j + 3] = (t.w - m.w) * r.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < MAXM; ++i) {
                        if (i < p.nmaps) {
                            const float* bgp = p.bg[i] ? p.bg[i] + (long long)(p.per_sample[i] ? n : 0) * p.b_nstride[i] + c0 + c : nullptr;
                            const float* bbp = p.bb[i] ? p.bb[i] + (long long)(p.per_sample[i] ? n : 0) * p.b_nstride[i] + c0 + c : nullptr;
                            uint32_t g[32], b[32];
                            const uint32_t taddr = tacc + (uint32_t)(i * 2 * SP_CB + c);
                            tc_ld32(taddr, g);
                            tc_ld32(taddr + SP_CB, b);
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 g4 = (bgp && valid) ? *reinterpret_cast<const float4*>(bgp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                                const float4 b4 = (bbp && valid) ? *reinterpret_cast<const float4*>(bbp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                                vprev[i][j + 0] = v[j + 0]; v[j + 0] = v[j + 0] * (1.f + (__uint_as_float(g[j + 0]) + g4.x)) + (__uint_as_float(b[j + 0]) + b4.x);
                                vprev[i][j + 1] = v[j + 1]; v[j + 1] = v[j + 1] * (1.f + (__uint_as_float(g[j + 1]) + g4.y)) + (__uint_as_float(b[j + 1]) + b4.y);
                                vprev[i][j + 2] = v[j + 2]; v[j + 2] = v[j + 2] * (1.f + (__uint_as_float(g[j + 2]) + g4.z)) + (__uint_as_float(b[j + 2]) + b4.z);
                                vprev[i][j + 3] = v[j + 3]; v[j + 3] = v[j + 3] * (1.f + (__uint_as_float(g[j + 3]) + g4.w)) + (__uint_as_float(b[j + 3]) + b4.w);
                            }
                        }
                    }
                    float gr[32];
                    if (valid) {
                        const float4* dr = reinterpret_cast<const float4*>(dout + pix * p.C + c0 + c);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 t = dr[j];
                            gr[4 * j + 0] = t.x; gr[4 * j + 1] = t.y; gr[4 * j + 2] = t.z; gr[4 * j + 3] = t.w;
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (p.act == FSV_ACT_LRELU) gr[j] *= (v[j] > 0.f ? 1.f : FSV_LRELU_SLOPE);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) gr[j] = 0.f;
                    }
#pragma unroll
                    for (int i = MAXM - 1; i >= 0; --i) {
                        if (i < p.nmaps) {
                            const float* bgp = p.bg[i] ? p.bg[i] + (long long)(p.per_sample[i] ? n : 0) * p.b_nstride[i] + c0 + c : nullptr;
                            uint32_t g[32];
                            tc_ld32(tacc + (uint32_t)(i * 2 * SP_CB + c), g);
                            if (valid) {
                                float4* db = reinterpret_cast<float4*>(p.dbeta[i] + pix * p.dgb_ld[i] + c0 + c);
                                float4* dg = reinterpret_cast<float4*>(p.dgamma[i] + pix * p.dgb_ld[i] + c0 + c);
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    db[j] = make_float4(gr[4 * j], gr[4 * j + 1], gr[4 * j + 2], gr[4 * j + 3]);
                                    dg[j] = make_float4(gr[4 * j] * vprev[i][4 * j], gr[4 * j + 1] * vprev[i][4 * j + 1],
                                                        gr[4 * j + 2] * vprev[i][4 * j + 2], gr[4 * j + 3] * vprev[i][4 * j + 3]);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 g4 = (bgp && valid) ? *reinterpret_cast<const float4*>(bgp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                                gr[j + 0] *= (1.f + (__uint_as_float(g[j + 0]) + g4.x));
                                gr[j + 1] *= (1.f + (__uint_as_float(g[j + 1]) + g4.y));
                                gr[j + 2] *= (1.f + (__uint_as_float(g[j + 2]) + g4.z));
                                gr[j + 3] *= (1.f + (__uint_as_float(g[j + 3]) + g4.w));
                            }
                        }
                    }
                    if (valid) {
                        float4* dxr = reinterpret_cast<float4*>(dxhat + pix * p.C + c0 + c);
#pragma unroll
                        for (int j = 0; j < 8; ++j) dxr[j] = make_float4(gr[4 * j], gr[4 * j + 1], gr[4 * j + 2], gr[4 * j + 3]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
}

// ------------------------------------------------------------------ host side
static void sp_pick_tile(int H, int W, int& TW, int& TH, int& TN) {
    TW = 16; TH = 8; TN = 1;
    if (W < 16) {
        TW = 1; while (TW * 2 <= W && TW < 16) TW *= 2;
        int rem = 128 / TW;
        TH = 1; while (TH * 2 <= H && TH * 2 <= rem) TH *= 2;
        TN = rem / TH;
    }
}

extern "C" int fsv_spade_fwd_tc_eligible(const fsv_spade_desc* d) {
    if (!d || d->nmaps < 1 || d->nmaps > FSV_SPADE_MAX_MAPS) return 0;
    if (d->C % 32 != 0) return 0;
    int TW, TH, TN;
    sp_pick_tile(d->H, d->W, TW, TH, TN);
    for (int i = 0; i < d->nmaps; ++i) {
        if (d->K[i] % TC_BK != 0 || d->m_ld[i] % 4 != 0 || d->m_coff[i] % 4 != 0) return 0;
        if (d->w_nstride[i] != 0 && (TN != 1 || d->w_nstride[i] % 4 != 0)) return 0;   // per-sample weights: tiles must not span samples
    }
    if ((long long)d->N * d->H * d->W < 128) return 0;
    return fsv_get_encode_tiled() != nullptr ? 1 : 0;
}

static int spade_tc_launch(const fsv_spade_desc* d, const float* x, const float* mean, const float* rstd,
                           const float* const* maps, const float* const* wg, const float* const* bg,
                           const float* const* wb, const float* const* bb, float* out, const float* dout, float* dxhat,
                           float* const* dgamma, float* const* dbeta, bool bwd, void* stream, const char* who) {
    FSV_REQUIRE(d != nullptr, "%s: null descriptor", who);
    if (!fsv_spade_fwd_tc_eligible(d)) {
        fsv_set_error("%s: shape not eligible (need C%%32==0, K%%32==0)", who);
        return FSV_ENOTSUP;
    }
    FSV_REQUIRE((((uintptr_t)x) & 15) == 0 && (((uintptr_t)mean) & 15) == 0 && (((uintptr_t)rstd) & 15) == 0, "%s: pointers must be 16-byte aligned", who);
    // channels per CTA: 64 where C allows (one fetch of the map tile per 64 channels), or 32 for every layer with FSV_SPADE_CB=32 /
    // FSV_SPADE_CB_BWD=32 (smaller CTAs: more of them resident per SM; A/B switch)
    static int force_f = -1, force_b = -1;
    if (force_f < 0) { const char* e = getenv("FSV_SPADE_CB"); force_f = e ? atoi(e) : 0; const char* e2 = getenv("FSV_SPADE_CB_BWD"); force_b = e2 ? atoi(e2) : 0; }
    const int CB = (d->C % 64 == 0 && (bwd ? force_b : force_f) != 32) ? 64 : 32;
    SpTcParams p;
    memset(&p, 0, sizeof(p));
    p.nmaps = d->nmaps; p.N = d->N; p.H = d->H; p.W = d->W; p.C = d->C; p.up = d->up; p.instance = d->mode == FSV_NORM_INSTANCE; p.act = d->act;
    sp_pick_tile(d->H, d->W, p.TW, p.TH, p.TN);
    p.tiles_w = fsv_cdiv(d->W, p.TW); p.tiles_h = fsv_cdiv(d->H, p.TH);
    const int tiles_n = fsv_cdiv(d->N, p.TN);
    PFN_encodeTiled enc = fsv_get_encode_tiled();
    for (int i = 0; i < d->nmaps; ++i) {
        FSV_REQUIRE(maps[i] && wg[i] && wb[i], "%s: map %d has null pointers", who, i);
        FSV_REQUIRE((((uintptr_t)(maps[i] + d->m_coff[i])) & 15) == 0 && (((uintptr_t)wg[i]) & 15) == 0 && (((uintptr_t)wb[i]) & 15) == 0,
                    "%s: map %d pointers must be 16-byte aligned", who, i);
        p.K[i] = d->K[i];
        p.per_sample[i] = d->w_nstride[i] != 0;
        p.bg[i] = bg[i]; p.bb[i] = bb[i]; p.b_nstride[i] = d->w_nstride[i];
        if (bwd) {
            FSV_REQUIRE(dgamma[i] && dbeta[i], "%s: null dgamma/dbeta for map %d", who, i);
            p.dgamma[i] = dgamma[i]; p.dbeta[i] = dbeta[i];
            p.dgb_ld[i] = d->dgb_ld[i] > 0 ? d->dgb_ld[i] : d->C;
            FSV_REQUIRE(p.dgb_ld[i] % 4 == 0 && (((uintptr_t)dgamma[i]) & 15) == 0 && (((uintptr_t)dbeta[i]) & 15) == 0, "%s: dgamma/dbeta alignment", who);
        }
        const long long ld = d->m_ld[i];
        {
            cuuint64_t dims[4] = {(cuuint64_t)d->K[i], (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
            cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)ld * d->W * 4, (cuuint64_t)ld * d->W * d->H * 4};
            cuuint32_t box[4] = {TC_BK, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TN};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = enc(&p.mmap[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)(maps[i] + d->m_coff[i]), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            FSV_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(map %d) failed with %d", who, i, (int)r);
        }
        for (int which = 0; which < 2; ++which) {
            const float* wp = which == 0 ? wg[i] : wb[i];
            const int nw = p.per_sample[i] ? d->N : 1;
            const long long ns = p.per_sample[i] ? d->w_nstride[i] : (long long)d->C * d->K[i];
            cuuint64_t dims[3] = {(cuuint64_t)d->K[i], (cuuint64_t)d->C, (cuuint64_t)nw};
            cuuint64_t strides[2] = {(cuuint64_t)d->K[i] * 4, (cuuint64_t)ns * 4};
            cuuint32_t box[3] = {TC_BK, (cuuint32_t)CB, 1};
            cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = enc(which == 0 ? &p.gmap[i] : &p.bmap[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)wp, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            FSV_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled(weights %d/%d) failed with %d", who, i, which, (int)r);
        }
    }
    const int smem_bytes = SP_STAGES * SP_STAGE_BYTES_OF(CB) + (2 * SP_STAGES + 1) * 8 + 16 + 1024;
    static unsigned long long configured = 0;
    if (fsv_first_on_device(&configured)) {
        FSV_CUDA(cudaFuncSetAttribute(k_spade_tc<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        FSV_CUDA(cudaFuncSetAttribute(k_spade_tc<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        FSV_CUDA(cudaFuncSetAttribute(k_spade_tc<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        FSV_CUDA(cudaFuncSetAttribute(k_spade_tc<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        FSV_CUDA(cudaFuncSetAttribute((k_spade_tc<32, false, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    }
    dim3 grid(p.tiles_w * p.tiles_h * tiles_n, d->C / CB);
    cudaStream_t st = (cudaStream_t)stream;
    // persistent variant: FSV_SPADE_PERSIST bit 0 = forward, bit 1 = backward; needs both accumulator buffers in the 512 TMEM columns
    static int persist = -1;
    if (persist < 0) { const char* e = getenv("FSV_SPADE_PERSIST"); persist = e ? atoi(e) : 3; }
    bool bias_al = true;            // the persistent kernels read the biases with 16-byte loads
    for (int i = 0; i < d->nmaps; ++i)
        bias_al = bias_al && (!bg[i] || (((uintptr_t)bg[i]) & 15) == 0) && (!bb[i] || (((uintptr_t)bb[i]) & 15) == 0);
    if ((persist & (bwd ? 2 : 1)) && d->nmaps * 2 * CB <= 256 && bias_al) {
        const int m_tiles = p.tiles_w * p.tiles_h * tiles_n;
        const long long total = (long long)m_tiles * (d->C / CB);
        int num_k = 0;
        for (int i = 0; i < d->nmaps; ++i) num_k += d->K[i] / TC_BK;
        const int staging = bwd ? 0 : 2 * 4 * 32 * CB * 4;          // forward: output transposition area of the eight epilogue warps
        int stages = (176 * 1024 - staging) / SP_STAGE_BYTES_OF(CB);   // one CTA per SM (it owns all of TMEM): a deep ring, several tiles ahead
        if (stages > 8) stages = 8;
        if (stages > 4 * num_k) stages = 4 * num_k;
        if (stages < 2) stages = 2;
        const int smem_p = stages * SP_STAGE_BYTES_OF(CB) + staging + (2 * stages + 4) * 8 + 16 + 1024;
        static unsigned long long configured_p = 0;
        if (fsv_first_on_device(&configured_p)) {
            FSV_CUDA(cudaFuncSetAttribute((k_spade_tc_p<64, false, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            FSV_CUDA(cudaFuncSetAttribute((k_spade_tc_p<32, false, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            FSV_CUDA(cudaFuncSetAttribute((k_spade_tc_p<64, true, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            FSV_CUDA(cudaFuncSetAttribute((k_spade_tc_p<32, true, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        }
        long long gx = fsv_sm_count();
        if (gx > total) gx = total;
        if (!bwd) {
            if (CB == 64) k_spade_tc_p<64, false, 2><<<(unsigned)gx, 320, smem_p, st>>>(p, x, mean, rstd, out, nullptr, nullptr, stages, m_tiles);
            else k_spade_tc_p<32, false, 2><<<(unsigned)gx, 320, smem_p, st>>>(p, x, mean, rstd, out, nullptr, nullptr, stages, m_tiles);
        } else {
            if (CB == 64) k_spade_tc_p<64, true, 1><<<(unsigned)gx, 192, smem_p, st>>>(p, x, mean, rstd, nullptr, dout, dxhat, stages, m_tiles);
            else k_spade_tc_p<32, true, 1><<<(unsigned)gx, 192, smem_p, st>>>(p, x, mean, rstd, nullptr, dout, dxhat, stages, m_tiles);
        }
        FSV_CHECK_LAUNCH(who);
        return FSV_OK;
    }
    if (!bwd) {
        if (CB == 64) k_spade_tc<64, false><<<grid, 192, smem_bytes, st>>>(p, x, mean, rstd, out, nullptr, nullptr);
        else if (force_f == 32) k_spade_tc<32, false, 3><<<grid, 192, smem_bytes, st>>>(p, x, mean, rstd, out, nullptr, nullptr);   // 3 CTAs / SM
        else k_spade_tc<32, false><<<grid, 192, smem_bytes, st>>>(p, x, mean, rstd, out, nullptr, nullptr);
    } else {
        if (CB == 64) k_spade_tc<64, true><<<grid, 192, smem_bytes, st>>>(p, x, mean, rstd, nullptr, dout, dxhat);
        else k_spade_tc<32, true><<<grid, 192, smem_bytes, st>>>(p, x, mean, rstd, nullptr, dout, dxhat);
    }
    FSV_CHECK_LAUNCH(who);
    return FSV_OK;
}

extern "C" int fsv_spade_fwd_tc(const fsv_spade_desc* d, const float* x, const float* mean, const float* rstd,
                                const float* const* maps, const float* const* wg, const float* const* bg,
                                const float* const* wb, const float* const* bb, float* out, void* stream) {
    FSV_REQUIRE((((uintptr_t)out) & 15) == 0, "spade_fwd_tc: out must be 16-byte aligned");
    return spade_tc_launch(d, x, mean, rstd, maps, wg, bg, wb, bb, out, nullptr, nullptr, nullptr, nullptr, false, stream, "spade_fwd_tc");
}

extern "C" int fsv_spade_bwd_tc(const fsv_spade_desc* d, const float* x, const float* mean, const float* rstd,
                                const float* const* maps, const float* const* wg, const float* const* bg,
                                const float* const* wb, const float* const* bb, const float* dout,
                                float* dxhat, float* const* dgamma, float* const* dbeta, void* stream) {
    FSV_REQUIRE((((uintptr_t)dout) & 15) == 0 && (((uintptr_t)dxhat) & 15) == 0, "spade_bwd_tc: dout/dxhat must be 16-byte aligned");
    return spade_tc_launch(d, x, mean, rstd, maps, wg, bg, wb, bb, nullptr, dout, dxhat, dgamma, dbeta, true, stream, "spade_bwd_tc");
}
