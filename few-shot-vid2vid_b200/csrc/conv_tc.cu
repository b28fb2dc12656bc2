// tcgen05 / TMA implicit-GEMM convolution (forward) for NHWC fp32 activations, TF32 tensor cores,
// fp32 accumulation in TMEM.  sm_100a only.
//
// Replaces cuDNN's conv2d for the FLOP-carrying layers of the hot path (architecture.py:60,81-84;
// generator.py:473-489,523-537; discriminator.py:69-88): 3x3 stride-1/2 (pad 1) and 4x4 stride-1/2
// (pad 2) with Cin % 32 == 0 and Cout % 16 == 0.  The data gradient of a stride-1 conv is the same
// kernel run on dy with spatially flipped, channel-transposed weights (host side, ops.py).
//
// GEMM view:  D[pixel, cout] = sum_{tap} sum_{ci}  X[pixel shifted by tap, ci] * W[cout, tap, ci]
//   M tile  = 128 output pixels = TN images x TH rows x TW cols (one TMA box of the NHWC input per tap,
//             out-of-bounds rows/cols zero-filled by TMA == the conv's zero padding; no im2col buffer)
//   N tile  = BN output channels (<= 128), accumulator = 128 lanes x BN fp32 columns of TMEM
//   K block = 32 input channels of one tap = 128-byte rows, SWIZZLE_128B, K-major for both operands
//   stride 2: the input is addressed through four "parity" tensor maps (even/odd rows x even/odd
//             cols), so every tap is again a dense box load.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected
// lane issues tcgen05.mma.kind::tf32, 4 per K block), warps 2..5 = epilogue (tcgen05.ld 32x32b, bias +
// residual + activation + scale, 16-byte global stores).  smem ring of STAGES x (16 KB A + BN*128 B B)
// guarded by full/empty mbarriers; tcgen05.commit releases ring slots and signals the epilogue.
//
// Roofline: tensor-bound in the limit; with 4-byte operands a 128x128x32 tile moves 32 KB per
// 1.05 MFLOP (32 FLOP/B) through L2, so this first version is L2-bandwidth bound below the TF32 peak
// (DESIGN.md section 5).
#include "tc_common.cuh"
#include <stdlib.h>

#define TC_MAX_TAPS 16
#define TC_MAX_STAGES 8

struct TcTap { int map, dh, dw, wk; };

struct __align__(64) TcParams {
    CUtensorMap amap[4];
    CUtensorMap bmap;
    TcTap taps[TC_MAX_TAPS];
    int ntaps, Cin, Cout, N, Ho, Wo;
    int TW, TH, TN, tiles_w, tiles_h;
    int y_ld, y_coff, res_ld, res_coff, act;
    float out_scale;
    int BN;
    int w_per_sample;           // B operand: 3-D tensor map (K, rows, sample); tiles never span samples (TN == 1)
    long long b_nstride;        // per-sample bias stride (floats), 0 = shared
    int stages;                 // smem ring depth (2..TC_MAX_STAGES)
    int m_tiles;                // pixel tiles (tiles_w * tiles_h * tiles_n), used by the persistent variant
    int OH, OW, os, oph, opw;   // output buffer dims and pixel stride/offset: pixel (ho,wo) of the GEMM lands at (ho*os+oph, wo*os+opw)
    // persistent variant only: ncls > 1 = several output-parity classes of one layer (stride-2 dgrad, collapsed up2 forward) in ONE
    // launch.  Class c owns taps[cls_t0[c] .. +cls_nt[c]) and writes at pixel offset (c >> 1, c & 1); the classes share the tile
    // geometry, so tile index = (class, channel block, pixel tile).
    int ncls;
    int cls_t0[4], cls_nt[4];
    // bring-up experiment (FSV_TC_SHIFT_EXP, one-tile-per-CTA kernel only): the A box is loaded one pixel to the left and the UMMA descriptor
    // starts one 128-byte row later -- does a start address that is not aligned to the 1024-byte swizzle pattern address the rows it should?
    // 1: base_offset 0, 2: base_offset = (start >> 7) & 7.  Pixels with tw == TW-1 are wrong by construction.  See scripts/shift_exp.py.
    int shift_exp;
};

// ------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(192, 1) k_conv_tc(const __grid_constant__ TcParams p, const float* __restrict__ bias,
                                                    const float* __restrict__ residual, float* __restrict__ y) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for SWIZZLE_128B
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int BN = p.BN;
    const int STG = p.stages;
    const int b_bytes = BN * TC_BK * 4;
    const int stage_bytes = TC_A_BYTES + b_bytes;
    uint64_t* bars = (uint64_t*)(smem + STG * stage_bytes);   // full[STAGES], empty[STAGES], tmem_full
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STG + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile coordinates
    int mt = blockIdx.x;
    const int tw_i = mt % p.tiles_w; mt /= p.tiles_w;
    const int th_i = mt % p.tiles_h; mt /= p.tiles_h;
    const int n0 = mt * p.TN;
    const int h0 = th_i * p.TH, w0 = tw_i * p.TW;
    const int co0 = blockIdx.y * BN;
    const int kblocks = p.Cin / TC_BK;
    const int num_k = p.ntaps * kblocks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STG; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[STG + s]), 1);
        }
        mbar_init(smem_u32(&bars[2 * STG]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < BN) tmem_cols <<= 1;
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
            // ===== TMA producer
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % STG;
                const uint32_t ph = (kb / STG) & 1;
                mbar_wait(smem_u32(&bars[STG + s]), ph ^ 1);          // slot free
                const int t = kb / kblocks, cb = kb - t * kblocks;
                const TcTap tap = p.taps[t];
                const uint32_t full = smem_u32(&bars[s]);
                const uint32_t a_dst = smem_u32(smem + s * stage_bytes);
                mbar_expect_tx(full, (uint32_t)stage_bytes);
                tma_load_4d(a_dst, &p.amap[tap.map], full, cb * TC_BK, w0 + tap.dw - (p.shift_exp ? 1 : 0), h0 + tap.dh, n0);
                tma_load_3d(a_dst + TC_A_BYTES, &p.bmap, full, tap.wk * p.Cin + cb * TC_BK, co0, p.w_per_sample ? n0 : 0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer
            const uint32_t idesc = make_idesc_tf32(TC_BM, BN);
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % STG;
                const uint32_t ph = (kb / STG) & 1;
                mbar_wait(smem_u32(&bars[s]), ph);                           // operands landed
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
                uint64_t adesc = make_kmajor_sw128_desc(a_addr + (p.shift_exp ? 128u : 0u));
                if (p.shift_exp == 2) adesc |= (uint64_t)(((a_addr + 128u) >> 7) & 7u) << 49;
                const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + TC_A_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {
                    // advance 8 tf32 = 32 bytes along K inside the 128-byte swizzle atom: +2 in (addr >> 4) units
                    tc_mma_tf32(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                }
                tc_commit(smem_u32(&bars[STG + s]));                   // frees the smem slot when the MMAs retire
            }
            tc_commit(smem_u32(&bars[2 * STG]));                       // accumulator complete
        }
    } else {
        // ===== epilogue: warps 2..5 -> TMEM lane quadrant (warp % 4)
        const int q = warp & 3;
        const int row = q * 32 + lane;                // GEMM row = pixel inside the tile
        mbar_wait(smem_u32(&bars[2 * STG]), 0);
        tc_fence_after();
        const int tw = row % p.TW;
        const int r2 = row / p.TW;
        const int th = r2 % p.TH;
        const int tn = r2 / p.TH;
        const int n = n0 + tn, ho = h0 + th, wo = w0 + tw;
        const bool valid = (n < p.N) && (ho < p.Ho) && (wo < p.Wo);
        const long long pix = ((long long)n * p.OH + (ho * p.os + p.oph)) * p.OW + (wo * p.os + p.opw);
        float* yrow = y + pix * p.y_ld + p.y_coff + co0;
        const float* rrow = residual ? residual + pix * p.res_ld + p.res_coff + co0 : nullptr;
        for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            if (valid) {
                const int ncol = min(32, BN - c);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (j < ncol) {
                        float4 o;
                        o.x = __uint_as_float(v[j]); o.y = __uint_as_float(v[j + 1]);
                        o.z = __uint_as_float(v[j + 2]); o.w = __uint_as_float(v[j + 3]);
                        if (bias) {
                            float4 b = *reinterpret_cast<const float4*>(bias + (long long)n * p.b_nstride + co0 + c + j);
                            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                        }
                        if (rrow) {
                            float4 r = *reinterpret_cast<const float4*>(rrow + c + j);
                            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                        }
                        o.x = fsv_act(o.x, p.act) * p.out_scale; o.y = fsv_act(o.y, p.act) * p.out_scale;
                        o.z = fsv_act(o.z, p.act) * p.out_scale; o.w = fsv_act(o.w, p.act) * p.out_scale;
                        *reinterpret_cast<float4*>(yrow + c + j) = o;
                    }
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

// ------------------------------------------------------------------ persistent variant (NOT the default; FSV_TC_PERSIST=1)
// Round-2 candidate, written after the round-1 GPU budget was spent: it compiles for sm_100a but has NOT run on hardware
// yet, so nothing selects it unless FSV_TC_PERSIST=1 is set.  Motivation: the
// 256x256- and 128x128-resolution layers have 9-36 K blocks per tile, so a CTA spends more time in its prologue
// (barrier init, TMEM allocation, first TMA round trip) and epilogue than in its main loop.  Here each CTA walks tiles
// tile = blockIdx.x, blockIdx.x + gridDim.x, ...; the operand ring runs continuously across tiles and the accumulator is
// double-buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps the main loop of tile i+1:
//   full[s]/empty[s]   : TMA -> MMA ring, exactly as in k_conv_tc, stage counter carried across tiles
//   acc_full[b]        : tcgen05.commit after the last MMA of a tile -> epilogue warps
//   acc_empty[b]       : one arrive per epilogue warp after its last tcgen05.ld of the tile -> MMA issuer may overwrite
__global__ void __launch_bounds__(192, 1) k_conv_tc_p(const __grid_constant__ TcParams p, const float* __restrict__ bias,
                                                      const float* __restrict__ residual, float* __restrict__ y) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int BN = p.BN;
    const int STG = p.stages;
    const int stage_bytes = TC_A_BYTES + BN * TC_BK * 4;
    uint64_t* bars = (uint64_t*)(smem + STG * stage_bytes);   // full[STG], empty[STG], acc_full[2], acc_empty[2]
    uint64_t* acc_full = bars + 2 * STG;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = p.Cin / TC_BK;
    const int m_tiles = p.m_tiles;
    const int per_cls = m_tiles * (p.Cout / BN);
    const int ncls = p.ncls > 1 ? p.ncls : 1;
    const int total = per_cls * ncls;

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
    while ((int)accw < BN) accw <<= 1;
    const uint32_t tmem_cols = 2 * accw;
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
                const int cls = tile / per_cls, tcl = tile - cls * per_cls;
                const int t0 = p.ncls > 1 ? p.cls_t0[cls] : 0;
                const int num_k = (p.ncls > 1 ? p.cls_nt[cls] : p.ntaps) * kblocks;
                int mt = tcl % m_tiles;
                const int co0 = (tcl / m_tiles) * BN;
                const int tw_i = mt % p.tiles_w; mt /= p.tiles_w;
                const int th_i = mt % p.tiles_h; mt /= p.tiles_h;
                const int n0 = mt * p.TN, h0 = th_i * p.TH, w0 = tw_i * p.TW;
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    const int s = it % STG;
                    const uint32_t ph = (it / STG) & 1;
                    mbar_wait(smem_u32(&bars[STG + s]), ph ^ 1);
                    const int t = kb / kblocks, cb = kb - t * kblocks;
                    const TcTap tap = p.taps[t0 + t];
                    const uint32_t full = smem_u32(&bars[s]);
                    const uint32_t a_dst = smem_u32(smem + s * stage_bytes);
                    mbar_expect_tx(full, (uint32_t)stage_bytes);
                    tma_load_4d(a_dst, &p.amap[tap.map], full, cb * TC_BK, w0 + tap.dw, h0 + tap.dh, n0);
                    tma_load_3d(a_dst + TC_A_BYTES, &p.bmap, full, tap.wk * p.Cin + cb * TC_BK, co0, p.w_per_sample ? n0 : 0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_BM, BN);
            uint32_t it = 0, i = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
                const uint32_t buf = i & 1;
                const int num_k = (p.ncls > 1 ? p.cls_nt[tile / per_cls] : p.ntaps) * kblocks;
                mbar_wait(smem_u32(&acc_empty[buf]), ((i >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * accw;
                for (int kb = 0; kb < num_k; ++kb, ++it) {
                    const int s = it % STG;
                    const uint32_t ph = (it / STG) & 1;
                    mbar_wait(smem_u32(&bars[s]), ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
                    const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                    const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + TC_A_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k)
                        tc_mma_tf32(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    tc_commit(smem_u32(&bars[STG + s]));
                }
                tc_commit(smem_u32(&acc_full[buf]));
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t i = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
            const uint32_t buf = i & 1;
            const int cls = tile / per_cls, tcl = tile - cls * per_cls;
            const int oph = p.ncls > 1 ? (cls >> 1) : p.oph, opw = p.ncls > 1 ? (cls & 1) : p.opw;
            int mt = tcl % m_tiles;
            const int co0 = (tcl / m_tiles) * BN;
            const int tw_i = mt % p.tiles_w; mt /= p.tiles_w;
            const int th_i = mt % p.tiles_h; mt /= p.tiles_h;
            const int n0 = mt * p.TN, h0 = th_i * p.TH, w0 = tw_i * p.TW;
            const int tw = row % p.TW;
            const int r2 = row / p.TW;
            const int th = r2 % p.TH;
            const int tn = r2 / p.TH;
            const int n = n0 + tn, ho = h0 + th, wo = w0 + tw;
            const bool valid = (n < p.N) && (ho < p.Ho) && (wo < p.Wo);
            const long long pix = ((long long)n * p.OH + (ho * p.os + oph)) * p.OW + (wo * p.os + opw);
            float* yrow = y + pix * p.y_ld + p.y_coff + co0;
            const float* rrow = residual ? residual + pix * p.res_ld + p.res_coff + co0 : nullptr;
            mbar_wait(smem_u32(&acc_full[buf]), (i >> 1) & 1);
            tc_fence_after();
            for (int c = 0; c < BN; c += 32) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * accw + (uint32_t)c, v);
                if (valid) {
                    const int ncol = min(32, BN - c);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        if (j < ncol) {
                            float4 o;
                            o.x = __uint_as_float(v[j]); o.y = __uint_as_float(v[j + 1]);
                            o.z = __uint_as_float(v[j + 2]); o.w = __uint_as_float(v[j + 3]);
                            if (bias) {
                                float4 b = *reinterpret_cast<const float4*>(bias + (long long)n * p.b_nstride + co0 + c + j);
                                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                            }
                            if (rrow) {
                                float4 r = *reinterpret_cast<const float4*>(rrow + c + j);
                                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                            }
                            o.x = fsv_act(o.x, p.act) * p.out_scale; o.y = fsv_act(o.y, p.act) * p.out_scale;
                            o.z = fsv_act(o.z, p.act) * p.out_scale; o.w = fsv_act(o.w, p.act) * p.out_scale;
                            *reinterpret_cast<float4*>(yrow + c + j) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
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
PFN_encodeTiled fsv_get_encode_tiled();
static PFN_encodeTiled get_encode() { return fsv_get_encode_tiled(); }
PFN_encodeTiled fsv_get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)f;
    }
    return fn;
}

static int pick_bn(int cout) {
    if (cout % 128 == 0) return 128;
    if (cout <= 128 && cout % 16 == 0) return cout;
    if (cout % 64 == 0) return 64;
    if (cout % 32 == 0) return 32;
    return 0;
}

// Few output-pixel tiles (low-resolution, wide layers: 1024 -> 1024 3x3 at 8x8 streams a 37.7 MB weight through 32 CTAs): narrow the
// N tile until the grid covers the SMs, so that the weight stream is spread over all of them (A tiles are tiny there; the extra
// A re-reads stay in L2).  FSV_TC_BN_OCC=0 disables.
static int occupancy_bn(int bn, int cout, long long m_tiles) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("FSV_TC_BN_OCC"); on = e ? atoi(e) : 1; }
    if (!on) return bn;
    const long long sms = fsv_sm_count();
    // level 1 (default): narrow while less than half the SMs would get a CTA; level 2 (experiment): while not every SM gets one
    // (two narrower CTAs per SM keep twice the operand bytes in flight of one wide CTA)
    const long long f = on >= 2 ? 1 : 2;
    while (bn >= 64 && f * m_tiles * (cout / bn) <= sms && cout % (bn / 2) == 0 && (bn / 2) % 16 == 0) bn /= 2;
    return bn;
}

extern "C" int fsv_conv2d_tc_eligible(const fsv_conv_desc* d) {
    if (!d) return 0;
    if (d->up != 1 || d->in_act != FSV_ACT_NONE) return 0;
    if (d->w_nstride != 0 || d->b_nstride != 0) return 0;
    if (d->Cin % TC_BK != 0 || d->x_coff % 4 != 0 || d->x_ld % 4 != 0) return 0;
    if (pick_bn(d->Cout) == 0) return 0;
    if (d->y_ld % 4 != 0 || d->y_coff % 4 != 0 || d->res_ld % 4 != 0 || d->res_coff % 4 != 0) return 0;
    if (d->kh * d->kw > TC_MAX_TAPS) return 0;
    if (d->stride != 1 && d->stride != 2) return 0;
    if ((long long)d->Ho * d->Wo * d->N < 64) return 0;          // tiny problems: the SIMT kernel is as good
    return get_encode() != nullptr ? 1 : 0;
}

static int encode_act_map(CUtensorMap* m, const float* base, int C, int ld, int Wd, int Hd, int N, long long sw, long long sh,
                          long long sn, int TW, int TH, int TN) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)sw * 4, (cuuint64_t)sh * 4, (cuuint64_t)sn * 4};
    cuuint32_t box[4] = {TC_BK, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TN};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    (void)ld;
    CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

static void pick_tile(int Ho, int Wo, int& TW, int& TH, int& TN) {
    TW = 16; TH = 8; TN = 1;
    if (Wo < 16) {
        TW = 1; while (TW * 2 <= Wo && TW < 16) TW *= 2;
        int rem = 128 / TW;
        TH = 1; while (TH * 2 <= Ho && TH * 2 <= rem) TH *= 2;
        TN = rem / TH;
    }
}

static int encode_weight_map(CUtensorMap* m, const float* w, long long kdim, int rows, int BN, int nsamples = 1, long long nstride = 0) {
    // (K, rows, sample): shared weights are the 1-sample case
    if (nsamples <= 1 || nstride == 0) { nsamples = 1; nstride = kdim * rows; }
    cuuint64_t dims[3] = {(cuuint64_t)kdim, (cuuint64_t)rows, (cuuint64_t)nsamples};
    cuuint64_t strides[2] = {(cuuint64_t)kdim * 4, (cuuint64_t)nstride * 4};
    cuuint32_t box[3] = {TC_BK, (cuuint32_t)BN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)w, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

// Ring depth.  The kernel is latency-bound on the L2 -> smem operand stream (ncu: lts throughput 10-40 %, tensor pipe
// 12-20 %), so bytes in flight per SM are what counts: either a deep ring with one CTA per SM or a shallower ring that
// lets two CTAs share the SM (the second CTA's main loop then also hides the first one's epilogue).
static int tc_env_stages() {
    static int v = -2;
    if (v == -2) {
        const char* e = getenv("FSV_TC_STAGES");
        v = e ? atoi(e) : -1;
    }
    return v;
}
// FSV_TC_DEEP=1: a launch whose tiles do not even fill the SMs once (one CTA per SM whatever the ring costs) takes a 200 KB ring, i.e.
// twice the bytes in flight per SM (off by default: in the multi-stream training graph the big CTA also keeps the other streams'
// kernels off its SM).
static int tc_env_deep() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FSV_TC_DEEP"); v = (e && atoi(e) != 0) ? 1 : 0; }
    return v;
}
static int pick_stages(int stage_bytes, int num_k, long long total_tiles) {
    int st = tc_env_stages();
    if (st <= 0) {
        const int budget = (tc_env_deep() && total_tiles <= fsv_sm_count()) ? 200 * 1024 : 96 * 1024;      // default: two CTAs per SM
        st = budget / stage_bytes;
        if (st < 3) st = 3;
    }
    const int max_fit = (200 * 1024) / stage_bytes;
    if (st > max_fit) st = max_fit;
    if (st > TC_MAX_STAGES) st = TC_MAX_STAGES;
    if (st > num_k) st = num_k;
    if (st < 2) st = 2;
    return st;
}

static int tc_persist() {
    static int persist = -1;
    if (persist < 0) { const char* e = getenv("FSV_TC_PERSIST"); persist = (e && atoi(e) == 0) ? 0 : 1; }
    return persist;
}
// the four output-parity classes of a stride-2 dgrad / collapsed up2 forward as ONE launch of the persistent kernel (each class alone
// fills a quarter of the SMs on the low-resolution layers, and the four launches were serialised on one stream).  FSV_TC_MERGE=0: four launches.
static int tc_merge_classes() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FSV_TC_MERGE"); v = (e && atoi(e) == 0) ? 0 : 1; }
    return v && tc_persist();
}

static int launch_tc(TcParams& p, int tiles_n, const float* bias, const float* residual, float* y, cudaStream_t st, const char* who) {
    const int stage_bytes = TC_A_BYTES + p.BN * TC_BK * 4;
    const int ncls = p.ncls > 1 ? p.ncls : 1;
    p.m_tiles = p.tiles_w * p.tiles_h * tiles_n;
    int max_taps = p.ntaps;
    if (ncls > 1) { max_taps = 0; for (int c = 0; c < ncls; ++c) max_taps = p.cls_nt[c] > max_taps ? p.cls_nt[c] : max_taps; }
    p.stages = pick_stages(stage_bytes, max_taps * (p.Cin / TC_BK), (long long)p.m_tiles * (p.Cout / p.BN) * ncls);
    const int smem_bytes = p.stages * stage_bytes + (2 * TC_MAX_STAGES + 1) * 8 + 16 + 1024;
    static unsigned long long configured = 0;
    if (fsv_first_on_device(&configured)) {
        FSV_CUDA(cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    }
    const int persist = tc_persist();
    FSV_REQUIRE(persist || ncls == 1, "%s: merged parity classes need the persistent kernel", who);
    if (persist) {      // default since round 2 (passes tests/test_gpu_tc.py on the B200, -1.2 ms per pose512 step); FSV_TC_PERSIST=0 = one tile per CTA
        static unsigned long long configured_p = 0;
        if (fsv_first_on_device(&configured_p)) {
            FSV_CUDA(cudaFuncSetAttribute(k_conv_tc_p, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        }
        const int smem_p = smem_bytes + 4 * 8;
        const long long total = (long long)p.m_tiles * (p.Cout / p.BN) * ncls;
        const int per_sm = smem_p <= 110 * 1024 ? 2 : 1;
        long long gx = (long long)fsv_sm_count() * per_sm;
        if (gx > total) gx = total;
        k_conv_tc_p<<<(unsigned)gx, 192, smem_p, st>>>(p, bias, residual, y);
        FSV_CHECK_LAUNCH(who);
        return FSV_OK;
    }
    {
        static int exp_mode = -1;
        if (exp_mode < 0) { const char* e = getenv("FSV_TC_SHIFT_EXP"); exp_mode = e ? atoi(e) : 0; }
        p.shift_exp = exp_mode;
    }
    dim3 grid(p.m_tiles, p.Cout / p.BN);
    k_conv_tc<<<grid, 192, smem_bytes, st>>>(p, bias, residual, y);
    FSV_CHECK_LAUNCH(who);
    return FSV_OK;
}

extern "C" int fsv_conv2d_fwd_tc(const fsv_conv_desc* d, const float* x, const float* w, const float* bias,
                                 const float* residual, float* y, void* stream) {
    FSV_REQUIRE(d != nullptr, "conv2d_fwd_tc: null descriptor");
    if (!fsv_conv2d_tc_eligible(d)) {
        fsv_set_error("conv2d_fwd_tc: shape not eligible for the tcgen05 path (need Cin%%32==0, Cout%%16==0, up==1, shared weights)");
        return FSV_ENOTSUP;
    }
    FSV_REQUIRE((((uintptr_t)x) & 15) == 0 && (((uintptr_t)w) & 15) == 0 && (((uintptr_t)y) & 15) == 0, "conv2d_fwd_tc: pointers must be 16-byte aligned");
    TcParams p;
    memset(&p, 0, sizeof(p));
    int TW, TH, TN;
    pick_tile(d->Ho, d->Wo, TW, TH, TN);
    p.TW = TW; p.TH = TH; p.TN = TN;
    p.tiles_w = fsv_cdiv(d->Wo, TW); p.tiles_h = fsv_cdiv(d->Ho, TH);
    const int tiles_n = fsv_cdiv(d->N, TN);
    const int BN = occupancy_bn(pick_bn(d->Cout), d->Cout, (long long)p.tiles_w * p.tiles_h * tiles_n);
    p.ntaps = d->kh * d->kw; p.Cin = d->Cin; p.Cout = d->Cout; p.N = d->N; p.Ho = d->Ho; p.Wo = d->Wo;
    p.OH = d->Ho; p.OW = d->Wo; p.os = 1; p.oph = 0; p.opw = 0;
    p.y_ld = d->y_ld; p.y_coff = d->y_coff; p.res_ld = d->res_ld; p.res_coff = d->res_coff; p.act = d->act; p.out_scale = d->out_scale;
    p.BN = BN;
    const long long ld = d->x_ld;
    const float* xb = x + d->x_coff;
    int rc = 0;
    if (d->stride == 1) {
        rc = encode_act_map(&p.amap[0], xb, d->Cin, d->x_ld, d->W, d->H, d->N, ld, ld * d->W, ld * d->W * d->H, TW, TH, TN);
        FSV_REQUIRE(rc == 0, "conv2d_fwd_tc: cuTensorMapEncodeTiled(A) failed with %d", rc);
        for (int r = 0; r < d->kh; ++r)
            for (int s = 0; s < d->kw; ++s) {
                TcTap& t = p.taps[r * d->kw + s];
                t.map = 0; t.dh = r - d->pad; t.dw = s - d->pad; t.wk = r * d->kw + s;
            }
    } else {
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                int Hd = (d->H - ph + 1) / 2, Wd = (d->W - pw + 1) / 2;
                if (Hd <= 0 || Wd <= 0) { Hd = Hd > 0 ? Hd : 1; Wd = Wd > 0 ? Wd : 1; }
                rc = encode_act_map(&p.amap[ph * 2 + pw], xb + ((long long)ph * d->W + pw) * ld, d->Cin, d->x_ld, Wd, Hd, d->N,
                                    2 * ld, 2 * ld * d->W, ld * d->W * d->H, TW, TH, TN);
                FSV_REQUIRE(rc == 0, "conv2d_fwd_tc: cuTensorMapEncodeTiled(A parity %d%d) failed with %d", ph, pw, rc);
            }
        for (int r = 0; r < d->kh; ++r)
            for (int s = 0; s < d->kw; ++s) {
                TcTap& t = p.taps[r * d->kw + s];
                int qh = r - d->pad, qw = s - d->pad;
                int ph = ((qh % 2) + 2) % 2, pw = ((qw % 2) + 2) % 2;
                t.map = ph * 2 + pw; t.dh = (qh - ph) / 2; t.dw = (qw - pw) / 2; t.wk = r * d->kw + s;
            }
    }
    rc = encode_weight_map(&p.bmap, w, (long long)p.ntaps * d->Cin, d->Cout, BN);
    FSV_REQUIRE(rc == 0, "conv2d_fwd_tc: cuTensorMapEncodeTiled(B) failed with %d", rc);
    return launch_tc(p, tiles_n, bias, residual, y, (cudaStream_t)stream, "conv2d_fwd_tc");
}

// ------------------------------------------------------------------ data gradient on the same kernel
// dx[n,h,w,ci] = sum_{r,s,co} dy[n,(h+pad-r)/stride,(w+pad-s)/stride,co] * w[co][r][s][ci]  (where divisible and in range).
// GEMM operands: A = dy (NHWC, K = Cout), B = wt[ci][r][s][co] (the OHWI weight with its channel axes swapped).
//   stride 1: one launch, tap (r,s) reads dy at (+pad-r, +pad-s).
//   stride 2: four launches, one per output parity (h%2, w%2); each uses the taps with r = (h+pad) mod 2 (same for s)
//             and writes its quarter of dx through the strided-output epilogue.
extern "C" int fsv_conv2d_dgrad_tc_eligible(const fsv_conv_desc* d) {
    if (!d) return 0;
    if (d->up != 1) return 0;
    if (d->w_nstride != 0) {   // per-sample weights: wt is (N, Cin, taps, Cout) and a tile must stay inside one sample
        int TW, TH, TN;
        pick_tile(d->stride == 1 ? d->H : (d->H + 1) / 2, d->stride == 1 ? d->W : (d->W + 1) / 2, TW, TH, TN);
        if (TN != 1 || d->stride != 1) return 0;
    }
    if (d->Cout % TC_BK != 0 || d->y_coff % 4 != 0 || d->y_ld % 4 != 0) return 0;
    if (pick_bn(d->Cin) == 0 || d->x_ld % 4 != 0 || d->x_coff % 4 != 0) return 0;
    if (d->kh * d->kw > TC_MAX_TAPS) return 0;
    if (d->stride != 1 && d->stride != 2) return 0;
    if ((long long)d->H * d->W * d->N < 256) return 0;
    return get_encode() != nullptr ? 1 : 0;
}

extern "C" int fsv_conv2d_dgrad_tc(const fsv_conv_desc* d, const float* dy, const float* wt, float* dx, void* stream) {
    FSV_REQUIRE(d != nullptr, "conv2d_dgrad_tc: null descriptor");
    if (!fsv_conv2d_dgrad_tc_eligible(d)) {
        fsv_set_error("conv2d_dgrad_tc: shape not eligible for the tcgen05 path");
        return FSV_ENOTSUP;
    }
    FSV_REQUIRE((((uintptr_t)dy) & 15) == 0 && (((uintptr_t)wt) & 15) == 0 && (((uintptr_t)dx) & 15) == 0, "conv2d_dgrad_tc: pointers must be 16-byte aligned");
    const int BN0 = pick_bn(d->Cin);
    const int taps_all = d->kh * d->kw;
    const long long ld = d->y_ld;
    const float* dyb = dy + d->y_coff;
    if (d->stride == 2 && d->H % 2 == 0 && d->W % 2 == 0 && tc_merge_classes()) {
        // all four parity classes have H/2 x W/2 pixels: one launch, tile index = (class, channel block, pixel tile)
        const int OHc = d->H / 2, OWc = d->W / 2;
        TcParams p;
        memset(&p, 0, sizeof(p));
        int TW, TH, TN;
        pick_tile(OHc, OWc, TW, TH, TN);
        p.TW = TW; p.TH = TH; p.TN = TN;
        p.tiles_w = fsv_cdiv(OWc, TW); p.tiles_h = fsv_cdiv(OHc, TH);
        const int tiles_n = fsv_cdiv(d->N, TN);
        p.Cin = d->Cout; p.Cout = d->Cin; p.N = d->N; p.Ho = OHc; p.Wo = OWc;
        p.OH = d->H; p.OW = d->W; p.os = 2;
        p.y_ld = d->x_ld; p.y_coff = d->x_coff; p.res_ld = d->x_ld; p.res_coff = 0; p.act = FSV_ACT_NONE; p.out_scale = 1.f;
        const int BN = occupancy_bn(BN0, d->Cin, 4LL * p.tiles_w * p.tiles_h * tiles_n);
        p.BN = BN;
        p.ncls = 4;
        int nt = 0;
        for (int cls = 0; cls < 4; ++cls) {
            const int ph = cls >> 1, pw = cls & 1;
            p.cls_t0[cls] = nt;
            for (int r = 0; r < d->kh; ++r)
                for (int s = 0; s < d->kw; ++s) {
                    const int th = ph + d->pad - r, tw = pw + d->pad - s;
                    if ((th & 1) || (tw & 1)) continue;
                    TcTap& t = p.taps[nt++];
                    t.map = 0; t.wk = r * d->kw + s;
                    t.dh = th / 2; t.dw = tw / 2;          // even (may be negative: exact division)
                }
            p.cls_nt[cls] = nt - p.cls_t0[cls];
            FSV_REQUIRE(p.cls_nt[cls] > 0, "conv2d_dgrad_tc: kernel %dx%d stride %d leaves a parity class without taps", d->kh, d->kw, d->stride);
        }
        p.ntaps = nt;
        int rc = encode_act_map(&p.amap[0], dyb, d->Cout, d->y_ld, d->Wo, d->Ho, d->N, ld, ld * d->Wo, ld * d->Wo * d->Ho, TW, TH, TN);
        FSV_REQUIRE(rc == 0, "conv2d_dgrad_tc: cuTensorMapEncodeTiled(A) failed with %d", rc);
        p.w_per_sample = d->w_nstride != 0;
        rc = encode_weight_map(&p.bmap, wt, (long long)taps_all * d->Cout, d->Cin, BN, p.w_per_sample ? d->N : 1,
                               (long long)taps_all * d->Cout * d->Cin);
        FSV_REQUIRE(rc == 0, "conv2d_dgrad_tc: cuTensorMapEncodeTiled(B) failed with %d", rc);
        return launch_tc(p, tiles_n, nullptr, nullptr, dx, (cudaStream_t)stream, "conv2d_dgrad_tc");
    }
    const int nclass = d->stride == 1 ? 1 : 4;
    for (int cls = 0; cls < nclass; ++cls) {
        const int ph = d->stride == 1 ? 0 : cls / 2, pw = d->stride == 1 ? 0 : cls % 2;
        const int OHc = d->stride == 1 ? d->H : (d->H - ph + 1) / 2;     // output rows of this parity class
        const int OWc = d->stride == 1 ? d->W : (d->W - pw + 1) / 2;
        if (OHc <= 0 || OWc <= 0) continue;
        TcParams p;
        memset(&p, 0, sizeof(p));
        int TW, TH, TN;
        pick_tile(OHc, OWc, TW, TH, TN);
        p.TW = TW; p.TH = TH; p.TN = TN;
        p.tiles_w = fsv_cdiv(OWc, TW); p.tiles_h = fsv_cdiv(OHc, TH);
        const int tiles_n = fsv_cdiv(d->N, TN);
        p.Cin = d->Cout; p.Cout = d->Cin; p.N = d->N; p.Ho = OHc; p.Wo = OWc;
        p.OH = d->H; p.OW = d->W; p.os = d->stride; p.oph = ph; p.opw = pw;
        p.y_ld = d->x_ld; p.y_coff = d->x_coff; p.res_ld = d->x_ld; p.res_coff = 0; p.act = FSV_ACT_NONE; p.out_scale = 1.f;
        const int BN = occupancy_bn(BN0, d->Cin, (long long)p.tiles_w * p.tiles_h * tiles_n);
        p.BN = BN;
        int nt = 0;
        for (int r = 0; r < d->kh; ++r)
            for (int s = 0; s < d->kw; ++s) {
                int th = ph + d->pad - r, tw = pw + d->pad - s;
                if (d->stride == 2 && ((th & 1) || (tw & 1))) continue;
                TcTap& t = p.taps[nt++];
                t.map = 0; t.wk = r * d->kw + s;
                t.dh = d->stride == 1 ? th : th / 2;       // th is even here for stride 2 (may be negative: exact division)
                t.dw = d->stride == 1 ? tw : tw / 2;
            }
        p.ntaps = nt;
        if (nt == 0) {   // no tap reaches this parity class (e.g. 1x1 stride-2): its outputs are zero
            FSV_REQUIRE(false, "conv2d_dgrad_tc: kernel %dx%d stride %d leaves a parity class without taps", d->kh, d->kw, d->stride);
        }
        int rc = encode_act_map(&p.amap[0], dyb, d->Cout, d->y_ld, d->Wo, d->Ho, d->N, ld, ld * d->Wo, ld * d->Wo * d->Ho, TW, TH, TN);
        FSV_REQUIRE(rc == 0, "conv2d_dgrad_tc: cuTensorMapEncodeTiled(A) failed with %d", rc);
        p.w_per_sample = d->w_nstride != 0;
        rc = encode_weight_map(&p.bmap, wt, (long long)taps_all * d->Cout, d->Cin, BN, p.w_per_sample ? d->N : 1,
                               (long long)taps_all * d->Cout * d->Cin);
        FSV_REQUIRE(rc == 0, "conv2d_dgrad_tc: cuTensorMapEncodeTiled(B) failed with %d", rc);
        rc = launch_tc(p, tiles_n, nullptr, nullptr, dx, (cudaStream_t)stream, "conv2d_dgrad_tc");
        if (rc) return rc;
    }
    return FSV_OK;
}

// ------------------------------------------------------------------ 3x3 conv over a nearest-x2-upsampled input, without the upsample
// y = conv3x3(up2(x)).  An output pixel of parity (ph, pw) only ever sees a 2x2 neighbourhood of SOURCE pixels: the
// three kernel rows collapse onto two source rows ({r0 | r1+r2} for even output rows, {r0+r1 | r2} for odd ones; same for
// columns), so the layer is four stride-1 2x2-tap convolutions of the source image -- one per output parity, written
// through the strided-output epilogue -- with pre-summed weights w4[co][ph][pw][a][b][ci] (built on the host side from
// the 3x3 kernel; 16 "taps").  That is 4/9 of the MACs of the reference's Upsample -> Conv2d (generator.py:484,537) and no
// 4x-size intermediate.
extern "C" int fsv_conv2d_fwd_tc_up2_eligible(const fsv_conv_desc* d) {
    if (!d || d->up != 2 || d->kh != 3 || d->kw != 3 || d->stride != 1 || d->pad != 1) return 0;
    fsv_conv_desc t = *d;
    t.up = 1;
    return fsv_conv2d_tc_eligible(&t);
}

extern "C" int fsv_conv2d_fwd_tc_up2(const fsv_conv_desc* d, const float* x, const float* w4, const float* bias,
                                     const float* residual, float* y, void* stream) {
    FSV_REQUIRE(d != nullptr, "conv2d_fwd_tc_up2: null descriptor");
    if (!fsv_conv2d_fwd_tc_up2_eligible(d)) {
        fsv_set_error("conv2d_fwd_tc_up2: not eligible (need up=2, 3x3 stride 1 pad 1, Cin%%32==0, Cout%%16==0)");
        return FSV_ENOTSUP;
    }
    FSV_REQUIRE((((uintptr_t)x) & 15) == 0 && (((uintptr_t)w4) & 15) == 0 && (((uintptr_t)y) & 15) == 0, "conv2d_fwd_tc_up2: pointers must be 16-byte aligned");
    const int Hs = d->H / 2, Ws = d->W / 2;
    const int BN0 = pick_bn(d->Cout);
    const long long ld = d->x_ld;
    if (tc_merge_classes()) {
        TcParams p;
        memset(&p, 0, sizeof(p));
        int TW, TH, TN;
        pick_tile(Hs, Ws, TW, TH, TN);
        p.TW = TW; p.TH = TH; p.TN = TN;
        p.tiles_w = fsv_cdiv(Ws, TW); p.tiles_h = fsv_cdiv(Hs, TH);
        const int tiles_n = fsv_cdiv(d->N, TN);
        p.Cin = d->Cin; p.Cout = d->Cout; p.N = d->N; p.Ho = Hs; p.Wo = Ws;
        p.OH = d->Ho; p.OW = d->Wo; p.os = 2;
        p.y_ld = d->y_ld; p.y_coff = d->y_coff; p.res_ld = d->res_ld; p.res_coff = d->res_coff; p.act = d->act; p.out_scale = d->out_scale;
        const int BN = occupancy_bn(BN0, d->Cout, 4LL * p.tiles_w * p.tiles_h * tiles_n);
        p.BN = BN;
        p.ncls = 4;
        p.ntaps = 16;
        for (int cls = 0; cls < 4; ++cls) {
            const int ph = cls >> 1, pw = cls & 1;
            p.cls_t0[cls] = 4 * cls;
            p.cls_nt[cls] = 4;
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) {
                    TcTap& t = p.taps[4 * cls + a * 2 + b];
                    t.map = 0; t.dh = a + ph - 1; t.dw = b + pw - 1; t.wk = ((ph * 2 + pw) * 2 + a) * 2 + b;
                }
        }
        int rc = encode_act_map(&p.amap[0], x + d->x_coff, d->Cin, d->x_ld, Ws, Hs, d->N, ld, ld * Ws, ld * Ws * Hs, TW, TH, TN);
        FSV_REQUIRE(rc == 0, "conv2d_fwd_tc_up2: cuTensorMapEncodeTiled(A) failed with %d", rc);
        rc = encode_weight_map(&p.bmap, w4, 16LL * d->Cin, d->Cout, BN);
        FSV_REQUIRE(rc == 0, "conv2d_fwd_tc_up2: cuTensorMapEncodeTiled(B) failed with %d", rc);
        return launch_tc(p, tiles_n, bias, residual, y, (cudaStream_t)stream, "conv2d_fwd_tc_up2");
    }
    for (int cls = 0; cls < 4; ++cls) {
        const int ph = cls >> 1, pw = cls & 1;
        TcParams p;
        memset(&p, 0, sizeof(p));
        int TW, TH, TN;
        pick_tile(Hs, Ws, TW, TH, TN);
        p.TW = TW; p.TH = TH; p.TN = TN;
        p.tiles_w = fsv_cdiv(Ws, TW); p.tiles_h = fsv_cdiv(Hs, TH);
        const int tiles_n = fsv_cdiv(d->N, TN);
        p.Cin = d->Cin; p.Cout = d->Cout; p.N = d->N; p.Ho = Hs; p.Wo = Ws;
        p.OH = d->Ho; p.OW = d->Wo; p.os = 2; p.oph = ph; p.opw = pw;
        p.y_ld = d->y_ld; p.y_coff = d->y_coff; p.res_ld = d->res_ld; p.res_coff = d->res_coff; p.act = d->act; p.out_scale = d->out_scale;
        const int BN = occupancy_bn(BN0, d->Cout, (long long)p.tiles_w * p.tiles_h * tiles_n);
        p.BN = BN;
        p.ntaps = 4;
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) {
                TcTap& t = p.taps[a * 2 + b];
                t.map = 0; t.dh = a + ph - 1; t.dw = b + pw - 1; t.wk = ((ph * 2 + pw) * 2 + a) * 2 + b;
            }
        int rc = encode_act_map(&p.amap[0], x + d->x_coff, d->Cin, d->x_ld, Ws, Hs, d->N, ld, ld * Ws, ld * Ws * Hs, TW, TH, TN);
        FSV_REQUIRE(rc == 0, "conv2d_fwd_tc_up2: cuTensorMapEncodeTiled(A) failed with %d", rc);
        rc = encode_weight_map(&p.bmap, w4, 16LL * d->Cin, d->Cout, BN);
        FSV_REQUIRE(rc == 0, "conv2d_fwd_tc_up2: cuTensorMapEncodeTiled(B) failed with %d", rc);
        rc = launch_tc(p, tiles_n, bias, residual, y, (cudaStream_t)stream, "conv2d_fwd_tc_up2");
        if (rc) return rc;
    }
    return FSV_OK;
}
