// tcgen05 / TMA weight gradient of the convolutions (TF32 operands, fp32 accumulation in TMEM, split-K with
// fp32 reductions into dW).  sm_100a only.
//
//   dW[co][r][s][ci] = sum over pixels  dy[px][co] * x[px shifted by (r,s)][ci]
//
// is, per tap, a GEMM whose reduction dimension is the PIXEL index.  tcgen05.mma wants both operands K-major
// in shared memory (K = pixels contiguous), while the activations live in HBM as NHWC (channels contiguous).
// So the operands are first re-laid out by a tiled transpose kernel into planar "channel-major" buffers
//   T[plane][n][c][h][w]   (row pitch padded to 16 bytes)
// -- dy as one plane; x as one plane PER KERNEL COLUMN s (and per row parity for stride 2):
//   X[s][ph][n][c][hq][wo] = x[n][hq*stride+ph][wo*stride + s - pad][c]   (0 outside the image; optionally read
//   through the nearest x2 upsample the conv reads x with)
// The column shift has to be baked into the copy because a TMA box must start on a 16-byte boundary of the
// innermost dimension -- a one-pixel (4-byte) shift along W is not addressable -- while row shifts are free (outer
// coordinate, out-of-bounds rows zero-filled by TMA == the conv's zero padding).  With that, x and dy share the
// same (wo) column grid and every tap is a dense, aligned box.  The copies are pure HBM streaming.
//
// GEMM tile: M = 128 channels of the wider operand, N = up to 128 channels of the narrower one, K block = 32
// pixels (a PW x PH x PN box, PW*PH*PN = 32, 128-byte K rows, SWIZZLE_128B).  grid = (split-K chunks, taps,
// M-tiles x N-tiles).  Warp roles and the smem ring are those of conv_tc.cu; the epilogue adds the tile into
// dW with fp32 atomics (red.global.add.f32).
#include "tc_common.cuh"
#include <stdlib.h>

#define WG_STAGES 4
#define WG_MAX_STAGES 8
#define WG_MAX_TAPS 16

struct WgTap { int plane, dh, dw; };

struct __align__(64) WgParams {
    CUtensorMap dymap;          // planar dy: dims (Wo, Ho, Cout, N)
    CUtensorMap xmap[8];        // planar x planes (kernel column s, row parity ph): dims (Wo, Hq, Cin, N)
    WgTap taps[WG_MAX_TAPS];
    int ntaps, Cin, Cout, N, Ho, Wo;
    int PW, PH, PN, nWB, nHB, nNB;     // K-block pixel box and block counts
    int role;                   // 0: M = Cout (A from dy, B from x)   1: M = Cin (A from x, B from dy)
    int BN, mtiles, ntiles;
    int kb_per_split;
};

__global__ void __launch_bounds__(192, 1) k_wgrad_tc(const __grid_constant__ WgParams p, float* __restrict__ dw) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int BN = p.BN;
    const int stage_bytes = TC_A_BYTES + BN * TC_BK * 4;
    uint64_t* bars = (uint64_t*)(smem + WG_STAGES * stage_bytes);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * WG_STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int tap_i = blockIdx.y;
    const WgTap tap = p.taps[tap_i];
    const int mt = blockIdx.z / p.ntiles, nt = blockIdx.z - mt * p.ntiles;
    const int m0 = mt * TC_BM, n0 = nt * BN;
    const int KB = p.nWB * p.nHB * p.nNB;
    const int kb0 = blockIdx.x * p.kb_per_split;
    const int kb1 = min(kb0 + p.kb_per_split, KB);
    const int num_k = kb1 - kb0;
    if (num_k <= 0) return;        // uniform for the whole CTA

    if (threadIdx.x == 0) {
        for (int s = 0; s < WG_STAGES; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&bars[WG_STAGES + s]), 1);
        }
        mbar_init(smem_u32(&bars[2 * WG_STAGES]), 1);
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
            const CUtensorMap* amap = p.role == 0 ? &p.dymap : &p.xmap[tap.plane];
            const CUtensorMap* bmap = p.role == 0 ? &p.xmap[tap.plane] : &p.dymap;
            const int a_dh = p.role == 0 ? 0 : tap.dh, a_dw = p.role == 0 ? 0 : tap.dw;
            const int b_dh = p.role == 0 ? tap.dh : 0, b_dw = p.role == 0 ? tap.dw : 0;
            for (int i = 0; i < num_k; ++i) {
                const int s = i % WG_STAGES;
                const uint32_t ph = (i / WG_STAGES) & 1;
                mbar_wait(smem_u32(&bars[WG_STAGES + s]), ph ^ 1);
                int kb = kb0 + i;
                const int wb = kb % p.nWB; kb /= p.nWB;
                const int hb = kb % p.nHB; kb /= p.nHB;
                const int w0 = wb * p.PW, h0 = hb * p.PH, nn0 = kb * p.PN;
                const uint32_t full = smem_u32(&bars[s]);
                const uint32_t a_dst = smem_u32(smem + s * stage_bytes);
                mbar_expect_tx(full, (uint32_t)stage_bytes);
                tma_load_4d(a_dst, amap, full, w0 + a_dw, h0 + a_dh, m0, nn0);
                tma_load_4d(a_dst + TC_A_BYTES, bmap, full, w0 + b_dw, h0 + b_dh, n0, nn0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_BM, BN);
            for (int i = 0; i < num_k; ++i) {
                const int s = i % WG_STAGES;
                const uint32_t ph = (i / WG_STAGES) & 1;
                mbar_wait(smem_u32(&bars[s]), ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
                const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + TC_A_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k)
                    tc_mma_tf32(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (i | k) != 0);
                tc_commit(smem_u32(&bars[WG_STAGES + s]));
            }
            tc_commit(smem_u32(&bars[2 * WG_STAGES]));
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        mbar_wait(smem_u32(&bars[2 * WG_STAGES]), 0);
        tc_fence_after();
        const int Mdim = p.role == 0 ? p.Cout : p.Cin;
        const int Ndim = p.role == 0 ? p.Cin : p.Cout;
        const int m = m0 + row;
        for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            if (m < Mdim) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = n0 + c + j;
                    if (c + j < BN && n < Ndim) {
                        const int co = p.role == 0 ? m : n, ci = p.role == 0 ? n : m;
                        atomicAdd(dw + ((long long)co * p.ntaps + tap_i) * p.Cin + ci, __uint_as_float(v[j]));
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

// ------------------------------------------------------------------ MN-major variant: operands straight from NHWC
// tcgen05 also accepts TF32 operands whose M/N index is the contiguous one ("MN-major"), provided the tile uses the
// SWIZZLE_128B_BASE32B shared-memory layout (cute: Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> over 4 K-rows x 128 B;
// TMA produces it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  With that, a K block of 32 pixels is one TMA box
// (32 channels, PW, PH, PN) of the NHWC tensor per 32-channel column block -- rows = pixels (K), 128 contiguous
// bytes = 32 channels (M or N) -- exactly how the activations already sit in HBM: no re-layout pass, no workspace,
// and (the box's inner dimension being channels) pixel shifts of +-1 are plain outer coordinates again.
//   UMMA descriptor: layout type 1, SBO = 512 B (next group of 4 K-rows), LBO = 4096 B (next 32-channel block),
//   K step of 8 pixels = +1024 B on the start address; instruction descriptor a_major = b_major = 1.
#define WGMN_BLOCK_BYTES (32 * 128)          // one 32-channel x 32-pixel block

struct __align__(64) WgMnParams {
    CUtensorMap dymap;          // NHWC dy: dims (Cout, Wo, Ho, N)
    CUtensorMap xmap[4];        // NHWC x (stride 1: map 0; stride 2: parity maps): dims (Cin, Wq, Hq, N)
    WgTap taps[WG_MAX_TAPS];
    int ntaps, Cin, Cout, N, Ho, Wo;
    int PW, PH, PN, nWB, nHB, nNB;
    int role, BN, mtiles, ntiles, kb_per_split;
    int per_sample, KBs, spn;          // per-sample weights: K blocks per sample, splits per sample (grid.x = N * spn)
    long long dw_nstride;
    int stages;                         // smem ring depth (2..WG_MAX_STAGES)
    int lbo16, sbo16, kstep16, ltype;   // descriptor fields in 16-byte units (tunable while bringing the layout up)
    // bring-up experiment (FSV_WG_SHIFT_EXP): the x box is loaded one pixel to the left and its descriptor starts one 128-byte K row later
    // (1: base_offset 0, 2: (start >> 7) & 3, 3: (start >> 7) & 7); K row 31 of every block is then garbage -- scripts/shift_exp.py zeroes
    // the dy pixels it meets.
    int shift_exp;
};

__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t saddr, int lbo16, int sbo16, int ltype) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(lbo16 & 0x3FFF) << 16;          // LBO: next 32-element block along M/N
    d |= (uint64_t)(sbo16 & 0x3FFF) << 32;          // SBO: next group of K rows
    d |= (uint64_t)1 << 46;                          // version (Blackwell)
    d |= (uint64_t)(ltype & 7) << 61;                // 1 = SWIZZLE_128B_BASE32B
    return d;
}

__global__ void __launch_bounds__(192, 1) k_wgrad_tc_mn(const __grid_constant__ WgMnParams p, float* __restrict__ dw) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int BN = p.BN;
    const int STG = p.stages;
    const int a_bytes = 4 * WGMN_BLOCK_BYTES, b_bytes = (BN / 32) * WGMN_BLOCK_BYTES;
    const int stage_bytes = a_bytes + b_bytes;
    uint64_t* bars = (uint64_t*)(smem + STG * stage_bytes);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STG + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tap_i = blockIdx.y;
    const WgTap tap = p.taps[tap_i];
    const int mt = blockIdx.z / p.ntiles, nt = blockIdx.z - mt * p.ntiles;
    const int m0 = mt * TC_BM, n0 = nt * BN;
    const int KB = p.nWB * p.nHB * p.nNB;
    int kb0 = blockIdx.x * p.kb_per_split;
    int kb1 = min(kb0 + p.kb_per_split, KB);
    if (p.per_sample) {            // split-K chunks never straddle a sample; each sample reduces into its own dW
        const int sn = blockIdx.x / p.spn, sj = blockIdx.x - sn * p.spn;
        kb0 = sn * p.KBs + sj * p.kb_per_split;
        kb1 = min(kb0 + p.kb_per_split, (sn + 1) * p.KBs);
        dw += (long long)sn * p.dw_nstride;
    }
    const int num_k = kb1 - kb0;
    if (num_k <= 0) return;

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
            const CUtensorMap* amap = p.role == 0 ? &p.dymap : &p.xmap[tap.plane];
            const CUtensorMap* bmap = p.role == 0 ? &p.xmap[tap.plane] : &p.dymap;
            const int a_dh = p.role == 0 ? 0 : tap.dh, a_dw = p.role == 0 ? 0 : tap.dw;
            const int b_dh = p.role == 0 ? tap.dh : 0, b_dw = p.role == 0 ? tap.dw : 0;
            for (int i = 0; i < num_k; ++i) {
                const int s = i % STG;
                const uint32_t ph = (i / STG) & 1;
                mbar_wait(smem_u32(&bars[STG + s]), ph ^ 1);
                int kb = kb0 + i;
                const int wb = kb % p.nWB; kb /= p.nWB;
                const int hb = kb % p.nHB; kb /= p.nHB;
                const int w0 = wb * p.PW, h0 = hb * p.PH, nn0 = kb * p.PN;
                const uint32_t full = smem_u32(&bars[s]);
                const uint32_t a_dst = smem_u32(smem + s * stage_bytes);
                mbar_expect_tx(full, (uint32_t)stage_bytes);
                const int xs = p.shift_exp ? 1 : 0;
                for (int j = 0; j < 4; ++j)
                    tma_load_4d(a_dst + j * WGMN_BLOCK_BYTES, amap, full, m0 + 32 * j, w0 + a_dw - (p.role == 1 ? xs : 0), h0 + a_dh, nn0);
                for (int j = 0; j < BN / 32; ++j)
                    tma_load_4d(a_dst + a_bytes + j * WGMN_BLOCK_BYTES, bmap, full, n0 + 32 * j, w0 + b_dw - (p.role == 0 ? xs : 0), h0 + b_dh, nn0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(TC_BM, BN) | (1u << 15) | (1u << 16);   // A and B MN-major
            for (int i = 0; i < num_k; ++i) {
                const int s = i % STG;
                const uint32_t ph = (i / STG) & 1;
                mbar_wait(smem_u32(&bars[s]), ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
                uint64_t adesc = make_mnmajor_desc(a_addr + ((p.shift_exp && p.role == 1) ? 128u : 0u), p.lbo16, p.sbo16, p.ltype);
                uint64_t bdesc = make_mnmajor_desc(a_addr + a_bytes + ((p.shift_exp && p.role == 0) ? 128u : 0u), p.lbo16, p.sbo16, p.ltype);
                if (p.shift_exp >= 2) {
                    const uint32_t xa = (p.role == 1 ? a_addr : a_addr + a_bytes) + 128u;
                    const uint64_t bo = (uint64_t)((xa >> 7) & (p.shift_exp == 2 ? 3u : 7u)) << 49;
                    if (p.role == 1) adesc |= bo; else bdesc |= bo;
                }
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k)     // 8 pixels = 8 rows of 128 B = +1024 B = +64 in (addr >> 4)
                    tc_mma_tf32(tmem_base, adesc + (uint64_t)(k * p.kstep16), bdesc + (uint64_t)(k * p.kstep16), idesc, (i | k) != 0);
                tc_commit(smem_u32(&bars[STG + s]));
            }
            tc_commit(smem_u32(&bars[2 * STG]));
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        mbar_wait(smem_u32(&bars[2 * STG]), 0);
        tc_fence_after();
        const int Mdim = p.role == 0 ? p.Cout : p.Cin;
        const int Ndim = p.role == 0 ? p.Cin : p.Cout;
        const int m = m0 + row;
        for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            if (m < Mdim) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = n0 + c + j;
                    if (c + j < BN && n < Ndim) {
                        const int co = p.role == 0 ? m : n, ci = p.role == 0 ? n : m;
                        atomicAdd(dw + ((long long)co * p.ntaps + tap_i) * p.Cin + ci, __uint_as_float(v[j]));
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

// ------------------------------------------------------------------ NHWC -> planar (channel-major) transpose
// dst[plane][n][c][hq][w] (row pitch `pitch`) <- src (N, Hs, Ws, C) NHWC slice (ld, coff), seen through an optional
// nearest x2 upsample (up) as an (H, W) = (Hs*up, Ws*up) image:
//   plane = s * nph + ph ;  row = hq * rstride + ph ;  col = w * cstride + s - pad     (zero outside the image)
// dy uses ns = nph = 1, rstride = cstride = 1, pad = 0.
// One block reads the (32*cstride + ns - 1) source pixels x 32 channels that feed a 32-column x 32-channel output tile
// ONCE (128-byte coalesced rows) and writes all ns column-shifted planes from shared memory (128-byte rows again).
#define TP_MAXSPAN 68
__global__ void __launch_bounds__(256) k_to_planar(const float* __restrict__ src, float* __restrict__ dst, int N, int Hs, int Ws, int C,
                                                   int ld, int coff, int up, int ns, int nph, int rstride, int cstride, int pad,
                                                   int Hq, int Wq, int pitch) {
    __shared__ float tile[TP_MAXSPAN][33];
    const int w0 = blockIdx.x * 32, c0 = blockIdx.z * 32;
    int y = blockIdx.y;                       // (ph * N + n) * Hq + hq
    const int hq = y % Hq; y /= Hq;
    const int n = y % N;
    const int ph = y / N;
    const int H = Hs * up, W = Ws * up;
    const int row = hq * rstride + ph;
    const int span = 31 * cstride + ns;       // source columns col0 .. col0+span-1 cover every (w, s) of this tile
    const int col0 = w0 * cstride - pad;
    for (int i = threadIdx.y; i < span; i += 8) {     // i = source column within the span, threadIdx.x = channel
        int col = col0 + i, c = c0 + threadIdx.x;
        float v = 0.f;
        if (row < H && col >= 0 && col < W && c < C)
            v = src[(((long long)n * Hs + row / up) * Ws + col / up) * ld + coff + c];
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int s = 0; s < ns; ++s) {
        const int plane = s * nph + ph;
        for (int i = threadIdx.y; i < 32; i += 8) {   // i = channel within the tile, threadIdx.x = output column
            int c = c0 + i, w = w0 + threadIdx.x;
            if (c < C && w < pitch) {
                float v = w < Wq ? tile[threadIdx.x * cstride + s][i] : 0.f;
                dst[((((long long)plane * N + n) * C + c) * Hq + hq) * pitch + w] = v;
            }
        }
    }
}

static int launch_to_planar(const float* src, float* dst, int N, int Hs, int Ws, int C, int ld, int coff, int up, int ns, int nph,
                            int rstride, int cstride, int pad, int Hq, int Wq, int pitch, cudaStream_t st) {
    FSV_REQUIRE(31 * cstride + ns <= TP_MAXSPAN, "to_planar: span too large");
    dim3 grid(fsv_cdiv(pitch, 32), nph * N * Hq, fsv_cdiv(C, 32));
    FSV_REQUIRE(grid.z <= 65535, "to_planar: grid too large");
    k_to_planar<<<grid, dim3(32, 8), 0, st>>>(src, dst, N, Hs, Ws, C, ld, coff, up, ns, nph, rstride, cstride, pad, Hq, Wq, pitch);
    FSV_CHECK_LAUNCH("to_planar");
    return FSV_OK;
}

// ------------------------------------------------------------------ host side
static bool wgrad_mn_enabled();
static int wg_bn(int nsmall) {
    if (nsmall >= 128) return 128;
    if (nsmall % 16 == 0) return nsmall;
    return 0;
}
static inline int pitch4(int w) { return (w + 3) & ~3; }

struct WgGeom { int Hq, nph, planes, pitch; long long x_floats, dy_floats; };
static WgGeom wg_geom(const fsv_conv_desc* d) {
    WgGeom g;
    g.nph = d->stride == 2 ? 2 : 1;
    g.Hq = d->stride == 2 ? (d->H + 1) / 2 : d->H;
    g.planes = d->kw * g.nph;
    g.pitch = pitch4(d->Wo);
    g.x_floats = (long long)g.planes * d->N * d->Cin * g.Hq * g.pitch;
    g.dy_floats = (long long)d->N * d->Cout * d->Ho * g.pitch;
    return g;
}

extern "C" int fsv_conv2d_wgrad_tc_eligible(const fsv_conv_desc* d) {
    if (!d) return 0;
    if (d->in_act != FSV_ACT_NONE) return 0;
    if (d->w_nstride != 0) {      // per-sample dW: MN-major path only, and a 32-pixel K block must fit inside one sample
        if (!wgrad_mn_enabled() || d->up != 1 || (long long)d->Ho * d->Wo < 32 || d->w_nstride % 4 != 0) return 0;
        int PW = 1; while (PW * 2 <= d->Wo && PW < 32) PW *= 2;
        int PH = 1; while (PH * 2 <= d->Ho && PW * PH * 2 <= 32) PH *= 2;
        if (PW * PH != 32) return 0;
    }
    if (d->stride != 1 && d->stride != 2) return 0;
    if (d->stride == 2 && d->up != 1) return 0;
    if (d->kh * d->kw > WG_MAX_TAPS || d->kw > 4) return 0;
    if (d->Cin % 16 != 0 || d->Cout % 16 != 0) return 0;
    if (d->Cin < 32 && d->Cout < 32) return 0;
    int nsmall = d->Cin < d->Cout ? d->Cin : d->Cout;
    if (wg_bn(nsmall) == 0) return 0;
    // K rows are 32 pixels of one image row (128-byte TMA inner box); rows narrower than 32 are zero-filled by TMA, which
    // wastes MMA issue slots but keeps the (small, low-resolution) layers off the FFMA path
    if (d->Wo < 4) return 0;
    if ((long long)d->N * d->Ho * d->Wo < 256) return 0;
    return fsv_get_encode_tiled() != nullptr ? 1 : 0;
}

extern "C" long long fsv_conv2d_wgrad_tc_workspace(const fsv_conv_desc* d) {
    if (!fsv_conv2d_wgrad_tc_eligible(d)) return 0;
    WgGeom g = wg_geom(d);
    return (g.x_floats + g.dy_floats) * 4 + 256;
}

static int encode_planar(CUtensorMap* m, const float* base, int Wd, int Hd, int N, int C, int pitch, int Hq, int CB) {
    // dims (W, H, C, N) in memory order; strides in bytes for H, C, N.  The K block is 32 pixels of one image row,
    // so the box is (32, 1, CB, 1): CB channel rows of 128 contiguous bytes, starting on a 128-byte boundary.
    cuuint64_t dims[4] = {(cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)C, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * Hq * 4, (cuuint64_t)pitch * Hq * C * 4};
    cuuint32_t box[4] = {32, 1, (cuuint32_t)CB, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fsv_get_encode_tiled()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

static int g_wgrad_mn = -1;
extern "C" void fsv_set_wgrad_mn(int on) { g_wgrad_mn = on; }
static bool wgrad_mn_enabled() {
    if (g_wgrad_mn < 0) {
        const char* e = getenv("FSV_WGRAD_MN");
        g_wgrad_mn = e ? atoi(e) : 1;      // default: operands in place (MN-major); FSV_WGRAD_MN=0 selects the planar re-layout path
    }
    return g_wgrad_mn != 0;
}

static int encode_nhwc_box(CUtensorMap* m, const float* base, int C, long long ld, int Wd, int Hd, int N, long long sw, long long sh,
                           long long sn, int PW, int PH, int PN) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)sw * 4, (cuuint64_t)sh * 4, (cuuint64_t)sn * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)PW, (cuuint32_t)PH, (cuuint32_t)PN};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    (void)ld;
    CUresult r = fsv_get_encode_tiled()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

// operands read in place (NHWC), no workspace use except a materialised x2 upsample
static int wgrad_tc_mn(const fsv_conv_desc* d, const float* x, const float* dy, float* dw, float* workspace, int accumulate, cudaStream_t st) {
    const int taps = d->kh * d->kw;
    if (!accumulate) {
        const size_t wsize = (size_t)d->Cout * taps * d->Cin;
        if (d->w_nstride == 0) FSV_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * wsize, st));
        else if ((size_t)d->w_nstride == wsize) FSV_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * wsize * d->N, st));
        else FSV_CUDA(cudaMemset2DAsync(dw, sizeof(float) * d->w_nstride, 0, sizeof(float) * wsize, d->N, st));
    }
    const float* xb = x + d->x_coff;
    long long xld = d->x_ld;
    if (d->up == 2) {     // dense boxes need the upsampled image: rebuild it in the workspace (HBM-cheap, not kept from forward)
        float* xu = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
        int rc = fsv_upsample2x_fwd(x + d->x_coff, xu, d->N, d->H / 2, d->W / 2, d->Cin, st);
        if (rc) return rc;
        FSV_REQUIRE(d->x_ld == d->Cin && d->x_coff == 0, "conv2d_wgrad_tc: up=2 needs a dense x");
        xb = xu; xld = d->Cin;
    }
    WgMnParams p;
    memset(&p, 0, sizeof(p));
    p.ntaps = taps; p.Cin = d->Cin; p.Cout = d->Cout; p.N = d->N; p.Ho = d->Ho; p.Wo = d->Wo;
    int PW = 1; while (PW * 2 <= d->Wo && PW < 32) PW *= 2;
    int PH = 1; while (PH * 2 <= d->Ho && PW * PH * 2 <= 32) PH *= 2;
    int PN = 32 / (PW * PH);
    p.PW = PW; p.PH = PH; p.PN = PN;
    p.nWB = fsv_cdiv(d->Wo, PW); p.nHB = fsv_cdiv(d->Ho, PH); p.nNB = fsv_cdiv(d->N, PN);
    p.role = d->Cout >= d->Cin ? 0 : 1;
    const int Mdim = p.role == 0 ? d->Cout : d->Cin, Ndim = p.role == 0 ? d->Cin : d->Cout;
    p.BN = Ndim >= 128 ? 128 : ((Ndim + 31) / 32) * 32;
    p.mtiles = fsv_cdiv(Mdim, TC_BM); p.ntiles = fsv_cdiv(Ndim, p.BN);
    const long long yld = d->y_ld;
    int rc = encode_nhwc_box(&p.dymap, dy + d->y_coff, d->Cout, yld, d->Wo, d->Ho, d->N, yld, yld * d->Wo, yld * d->Wo * d->Ho, PW, PH, PN);
    FSV_REQUIRE(rc == 0, "conv2d_wgrad_tc(mn): cuTensorMapEncodeTiled(dy) failed with %d", rc);
    if (d->stride == 1) {
        rc = encode_nhwc_box(&p.xmap[0], xb, d->Cin, xld, d->W, d->H, d->N, xld, xld * d->W, xld * d->W * d->H, PW, PH, PN);
        FSV_REQUIRE(rc == 0, "conv2d_wgrad_tc(mn): cuTensorMapEncodeTiled(x) failed with %d", rc);
        for (int r = 0; r < d->kh; ++r)
            for (int s2 = 0; s2 < d->kw; ++s2) {
                WgTap& t = p.taps[r * d->kw + s2];
                t.plane = 0; t.dh = r - d->pad; t.dw = s2 - d->pad;
            }
    } else {
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                int Hd = (d->H - ph + 1) / 2, Wd = (d->W - pw + 1) / 2;
                if (Hd < 1) Hd = 1;
                if (Wd < 1) Wd = 1;
                rc = encode_nhwc_box(&p.xmap[ph * 2 + pw], xb + ((long long)ph * d->W + pw) * xld, d->Cin, xld, Wd, Hd, d->N, 2 * xld,
                                     2 * xld * d->W, xld * d->W * d->H, PW, PH, PN);
                FSV_REQUIRE(rc == 0, "conv2d_wgrad_tc(mn): cuTensorMapEncodeTiled(x parity) failed with %d", rc);
            }
        for (int r = 0; r < d->kh; ++r)
            for (int s2 = 0; s2 < d->kw; ++s2) {
                WgTap& t = p.taps[r * d->kw + s2];
                int qh = r - d->pad, qw = s2 - d->pad;
                int ph = ((qh % 2) + 2) % 2, pw = ((qw % 2) + 2) % 2;
                t.plane = ph * 2 + pw; t.dh = (qh - ph) / 2; t.dw = (qw - pw) / 2;
            }
    }
    // validated on B200 by a parameter sweep (scripts/mn_sweep.sh, round 1): only this combination reproduces the reference
    p.lbo16 = WGMN_BLOCK_BYTES >> 4;   // 4096 B between 32-channel blocks
    p.sbo16 = 512 >> 4;                // 512 B between groups of 4 K rows
    p.kstep16 = 1024 >> 4;             // 8 pixels (one tf32 MMA K step) = 8 rows x 128 B
    p.ltype = 1;                       // SWIZZLE_128B_BASE32B
    {
        static int exp_mode = -1;
        if (exp_mode < 0) { const char* e = getenv("FSV_WG_SHIFT_EXP"); exp_mode = e ? atoi(e) : 0; }
        p.shift_exp = exp_mode;
    }
    const int KB = p.nWB * p.nHB * p.nNB;
    const long long base = (long long)taps * p.mtiles * p.ntiles;
    // split-K: enough CTAs for ~1.5 waves; every extra split costs Cout*taps*Cin fp32 reductions into dW through L2
    long long splits = ((long long)fsv_sm_count() * 3 / 2 + base - 1) / base;
    if (splits > KB / 8) splits = KB / 8;
    if (splits < 1) splits = 1;
    p.kb_per_split = (int)((KB + splits - 1) / splits);
    splits = (KB + p.kb_per_split - 1) / p.kb_per_split;
    if (d->w_nstride != 0) {
        FSV_REQUIRE(PN == 1, "conv2d_wgrad_tc: per-sample weights need K blocks inside one sample");
        p.per_sample = 1; p.dw_nstride = d->w_nstride; p.KBs = p.nWB * p.nHB;
        long long spn = ((long long)fsv_sm_count() * 3 / 2 + base * d->N - 1) / (base * d->N);
        if (spn > p.KBs / 8) spn = p.KBs / 8;
        if (spn < 1) spn = 1;
        p.kb_per_split = (int)((p.KBs + spn - 1) / spn);
        p.spn = (int)((p.KBs + p.kb_per_split - 1) / p.kb_per_split);
        splits = (long long)d->N * p.spn;
    }
    {
        // ring depth: see conv_tc.cu pick_stages (latency-bound operand stream: two CTAs per SM by default)
        static int env_st = -2;
        if (env_st == -2) { const char* e = getenv("FSV_WG_STAGES"); env_st = e ? atoi(e) : -1; }
        const int stage_bytes = 4 * WGMN_BLOCK_BYTES + (p.BN / 32) * WGMN_BLOCK_BYTES;
        int stg = env_st > 0 ? env_st : (96 * 1024) / stage_bytes;
        if (stg < 3 && env_st <= 0) stg = 3;
        if (stg > (200 * 1024) / stage_bytes) stg = (200 * 1024) / stage_bytes;
        if (stg > WG_MAX_STAGES) stg = WG_MAX_STAGES;
        if (stg < 2) stg = 2;
        p.stages = stg;
    }
    const int smem_bytes = p.stages * (4 * WGMN_BLOCK_BYTES + (p.BN / 32) * WGMN_BLOCK_BYTES) + (2 * WG_MAX_STAGES + 1) * 8 + 16 + 1024;
    static unsigned long long configured = 0;
    if (fsv_first_on_device(&configured)) {
        FSV_CUDA(cudaFuncSetAttribute(k_wgrad_tc_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    dim3 grid((unsigned)splits, taps, p.mtiles * p.ntiles);
    k_wgrad_tc_mn<<<grid, 192, smem_bytes, st>>>(p, dw);
    FSV_CHECK_LAUNCH("conv2d_wgrad_tc_mn");
    return FSV_OK;
}

extern "C" int fsv_conv2d_wgrad_tc(const fsv_conv_desc* d, const float* x, const float* dy, float* dw, float* workspace,
                                   int accumulate, void* stream) {
    FSV_REQUIRE(d != nullptr && workspace != nullptr, "conv2d_wgrad_tc: null descriptor / workspace");
    if (!fsv_conv2d_wgrad_tc_eligible(d)) {
        fsv_set_error("conv2d_wgrad_tc: shape not eligible for the tcgen05 path");
        return FSV_ENOTSUP;
    }
    if (wgrad_mn_enabled() && d->x_ld % 4 == 0 && d->x_coff % 4 == 0 && d->y_ld % 4 == 0 && d->y_coff % 4 == 0 &&
        (d->up == 1 || (d->x_ld == d->Cin && d->x_coff == 0)))
        return wgrad_tc_mn(d, x, dy, dw, workspace, accumulate, (cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    WgGeom g = wg_geom(d);
    float* xT = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float* dyT = xT + g.x_floats;
    const int taps = d->kh * d->kw;
    if (!accumulate) FSV_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->Cout * taps * d->Cin, st));
    int rc = launch_to_planar(x, xT, d->N, d->H / d->up, d->W / d->up, d->Cin, d->x_ld, d->x_coff, d->up, d->kw, g.nph, d->stride,
                              d->stride, d->pad, g.Hq, d->Wo, g.pitch, st);
    if (rc) return rc;
    rc = launch_to_planar(dy, dyT, d->N, d->Ho, d->Wo, d->Cout, d->y_ld, d->y_coff, 1, 1, 1, 1, 1, 0, d->Ho, d->Wo, g.pitch, st);
    if (rc) return rc;

    WgParams p;
    memset(&p, 0, sizeof(p));
    p.ntaps = taps; p.Cin = d->Cin; p.Cout = d->Cout; p.N = d->N; p.Ho = d->Ho; p.Wo = d->Wo;
    p.PW = 32; p.PH = 1; p.PN = 1;
    p.nWB = fsv_cdiv(d->Wo, p.PW); p.nHB = d->Ho; p.nNB = d->N;
    p.role = d->Cout >= d->Cin ? 0 : 1;
    const int Mdim = p.role == 0 ? d->Cout : d->Cin, Ndim = p.role == 0 ? d->Cin : d->Cout;
    p.BN = wg_bn(Ndim);
    p.mtiles = fsv_cdiv(Mdim, TC_BM); p.ntiles = fsv_cdiv(Ndim, p.BN);
    const int dyCB = p.role == 0 ? TC_BM : p.BN, xCB = p.role == 0 ? p.BN : TC_BM;
    rc = encode_planar(&p.dymap, dyT, d->Wo, d->Ho, d->N, d->Cout, g.pitch, d->Ho, dyCB);
    FSV_REQUIRE(rc == 0, "conv2d_wgrad_tc: cuTensorMapEncodeTiled(dy) failed with %d", rc);
    for (int pl = 0; pl < g.planes; ++pl) {
        int ph = pl % g.nph;
        int Hd = d->stride == 2 ? (d->H - ph + 1) / 2 : d->H;      // valid rows of this parity
        if (Hd < 1) Hd = 1;
        rc = encode_planar(&p.xmap[pl], xT + (long long)pl * d->N * d->Cin * g.Hq * g.pitch, d->Wo, Hd, d->N, d->Cin, g.pitch, g.Hq, xCB);
        FSV_REQUIRE(rc == 0, "conv2d_wgrad_tc: cuTensorMapEncodeTiled(x plane %d) failed with %d", pl, rc);
    }
    for (int r = 0; r < d->kh; ++r)
        for (int s2 = 0; s2 < d->kw; ++s2) {
            WgTap& t = p.taps[r * d->kw + s2];
            int qh = r - d->pad;
            int ph = d->stride == 2 ? ((qh % 2) + 2) % 2 : 0;
            t.plane = s2 * g.nph + ph;
            t.dh = d->stride == 2 ? (qh - ph) / 2 : qh;
            t.dw = 0;                                             // the column shift lives in the plane
        }
    const int KB = p.nWB * p.nHB * p.nNB;
    const long long base = (long long)taps * p.mtiles * p.ntiles;
    long long splits = ((long long)fsv_sm_count() * 3 + base - 1) / base;
    if (splits > KB / 8) splits = KB / 8;             // at least 8 K blocks per CTA
    if (splits < 1) splits = 1;
    p.kb_per_split = (int)((KB + splits - 1) / splits);
    splits = (KB + p.kb_per_split - 1) / p.kb_per_split;
    const int smem_bytes = WG_STAGES * (TC_A_BYTES + p.BN * TC_BK * 4) + (2 * WG_STAGES + 1) * 8 + 16 + 1024;
    static unsigned long long configured = 0;
    if (fsv_first_on_device(&configured)) {
        FSV_CUDA(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    dim3 grid((unsigned)splits, taps, p.mtiles * p.ntiles);
    k_wgrad_tc<<<grid, 192, smem_bytes, st>>>(p, dw);
    FSV_CHECK_LAUNCH("conv2d_wgrad_tc");
    return FSV_OK;
}
