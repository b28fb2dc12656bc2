// Spectral normalisation of a conv / linear weight, fused with the OIHW -> OHWI repack the conv kernels want.
// Replaces torch.nn.utils.spectral_norm's forward-pre-hook as the reference applies it
// (models/networks/architecture.py:60,81-84, generator.py:106-109, normalization.py:64-65,
//  discriminator.py:69-88): in training mode ONE power iteration per module call
//      v <- normalize(W^T u),  u <- normalize(W v)       (in place on the weight_u / weight_v buffers, eps 1e-12)
//  then sigma = u . (W v) with u, v held constant, and W_sn = W / sigma.  W is weight_orig viewed as (R, K = Cin*taps).
// Backward (u, v constants):  dW = (dW_sn - (sum dW_sn * W_sn) u v^T) / sigma.
//
// The torch hook costs ~20 tiny launches per module call (mv, norm, clamp, div, clone, dot, div + the permute copy and
// their backward); ~200 module calls per training step made that the largest launch-count item of the step.  Here a
// call is 3 launches forward and 2 backward, HBM/L2-bound: W is read three times and written once, all coalesced
// except the tap transposition (a 36-byte-stride gather out of L2).
//
// Cross-block reductions use the "last block done" pattern (partials in a workspace, a ticket counter that the last
// block resets), so no memset nodes are needed and the summation order -- hence the result -- is deterministic.
#include "common.cuh"

#define SN_THREADS 256

// Tickets of the per-weight kernels, one bank per reduction lane (norm.cu: fsv_set_reduction_lane): per-weight calls run wherever the
// module runs -- the generator's branch streams in a recording forward, the gathering weight-gradient lane in backward -- and two
// streams sharing one counter corrupt it for the rest of the process (round-2 session 21: garbage sigma in every later call).
#define SN_LANES 8
int fsv_current_reduction_lane();
__device__ unsigned int g_sn_ticket[SN_LANES][4];
#define SN_MAXCHUNKS 4096
__device__ unsigned int g_sn_colticket[SN_LANES][SN_MAXCHUNKS];

__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < SN_THREADS / 32; ++i) r += sh[i];
    return r;
}

// returns true in every thread of the block that took the last ticket
__device__ __forceinline__ bool last_block(unsigned int* ticket, unsigned int nblocks, int* sh_flag) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        *sh_flag = (t == nblocks - 1);
        if (t == nblocks - 1) *ticket = 0u;
    }
    __syncthreads();
    bool last = *sh_flag != 0;
    if (last) __threadfence();
    return last;
}

// phase 1: t = W^T u.  Grid (column chunks, row splits); the last row-split block of each column chunk (one ticket per
// chunk) folds that chunk's partials in a fixed order into t (kept in split 0's slot) and leaves the chunk's |t|^2.
__device__ __forceinline__ void sn_wtu_body(const float* __restrict__ W, const float* __restrict__ u, int R, int K, int rows_per_split,
                                            float* part, float* __restrict__ nrm_part, int chunk, int split, int nsplits,
                                            unsigned int* ticket) {
    __shared__ float sh[SN_THREADS / 32];
    __shared__ int flag;
    int c = chunk * SN_THREADS + threadIdx.x;
    int r0 = split * rows_per_split, r1 = min(R, r0 + rows_per_split);
    if (c < K) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int r = r0;
        for (; r + 3 < r1; r += 4) {
            a0 = fmaf(W[(size_t)r * K + c], u[r], a0);
            a1 = fmaf(W[(size_t)(r + 1) * K + c], u[r + 1], a1);
            a2 = fmaf(W[(size_t)(r + 2) * K + c], u[r + 2], a2);
            a3 = fmaf(W[(size_t)(r + 3) * K + c], u[r + 3], a3);
        }
        for (; r < r1; ++r) a0 = fmaf(W[(size_t)r * K + c], u[r], a0);
        part[(size_t)split * K + c] = (a0 + a1) + (a2 + a3);
    }
    if (!last_block(ticket, nsplits, &flag)) return;
    const volatile float* vp = part;
    float t = 0.f;
    if (c < K) {
        for (int s = 0; s < nsplits; ++s) t += vp[(size_t)s * K + c];
        part[c] = t;
    }
    float nrm = block_sum(t * t, sh);
    if (threadIdx.x == 0) nrm_part[chunk] = nrm;
}
__global__ void __launch_bounds__(SN_THREADS) k_sn_wtu(const float* __restrict__ W, const float* __restrict__ u, int R, int K,
                                                       int rows_per_split, float* part, float* __restrict__ nrm_part, int lane) {
    sn_wtu_body(W, u, R, K, rows_per_split, part, nrm_part, blockIdx.x, blockIdx.y, gridDim.y, &g_sn_colticket[lane][blockIdx.x]);
}

// phase 2: v = t / max(|t|, eps) (training), s = W v (warp per row); last block: sigma, u
__device__ __forceinline__ void sn_wv_body(const float* __restrict__ W, const float* __restrict__ t, const float* __restrict__ nrm_part,
                                           int nchunks, int R, int K, int power, float eps, float* s, float* u_buf,
                                           float* __restrict__ u_save, float* __restrict__ v_buf, float* __restrict__ v_save,
                                           float* __restrict__ sigma, int blk, int nblk, unsigned int* ticket) {
    __shared__ float sh[SN_THREADS / 32];
    __shared__ int flag;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float inv = 1.f;
    if (power) {
        float n = 0.f;
        for (int i = threadIdx.x; i < nchunks; i += SN_THREADS) n += nrm_part[i];
        n = block_sum(n, sh);
        inv = 1.f / fmaxf(sqrtf(n), eps);
    }
    int r = blk * (SN_THREADS / 32) + warp;
    if (r < R) {
        const float* wr = W + (size_t)r * K;
        float a = 0.f;
        if ((K & 3) == 0) {
            const float4* w4 = reinterpret_cast<const float4*>(wr);
            const float4* v4 = reinterpret_cast<const float4*>(t);
            float b = 0.f;
            int k = lane;
            for (; k + 32 < (K >> 2); k += 64) {
                float4 a4 = w4[k], b4 = v4[k], c4 = w4[k + 32], d4 = v4[k + 32];
                a = fmaf(a4.x, b4.x, a);
                a = fmaf(a4.y, b4.y, a);
                a = fmaf(a4.z, b4.z, a);
                a = fmaf(a4.w, b4.w, a);
                b = fmaf(c4.x, d4.x, b);
                b = fmaf(c4.y, d4.y, b);
                b = fmaf(c4.z, d4.z, b);
                b = fmaf(c4.w, d4.w, b);
            }
            for (; k < (K >> 2); k += 32) {
                float4 a4 = w4[k], b4 = v4[k];
                a = fmaf(a4.x, b4.x, a);
                a = fmaf(a4.y, b4.y, a);
                a = fmaf(a4.z, b4.z, a);
                a = fmaf(a4.w, b4.w, a);
            }
            a += b;
        } else {
            for (int k = lane; k < K; k += 32) a = fmaf(wr[k], t[k], a);
        }
        a = warp_sum(a) * inv;
        if (lane == 0) s[r] = a;
    }
    // the normalised v: every block writes its slice (training: also into the weight_v buffer)
    {
        int per = (K + nblk - 1) / nblk;
        int k0 = blk * per, k1 = min(K, k0 + per);
        for (int k = k0 + threadIdx.x; k < k1; k += SN_THREADS) {
            float vv = t[k] * inv;
            v_save[k] = vv;
            if (power) v_buf[k] = vv;
        }
    }
    if (!last_block(ticket, nblk, &flag)) return;
    const volatile float* sv = s;
    if (power) {
        float nrm = 0.f;
        for (int i = threadIdx.x; i < R; i += SN_THREADS) nrm += sv[i] * sv[i];
        nrm = block_sum(nrm, sh);
        float invs = 1.f / fmaxf(sqrtf(nrm), eps);
        float sg = 0.f;
        for (int i = threadIdx.x; i < R; i += SN_THREADS) {
            float un = sv[i] * invs;
            u_buf[i] = un;
            u_save[i] = un;
            sg += un * sv[i];     // sigma = u_new . (W v)
        }
        sg = block_sum(sg, sh);
        if (threadIdx.x == 0) *sigma = sg;
    } else {
        float sg = 0.f;
        for (int i = threadIdx.x; i < R; i += SN_THREADS) {
            float uo = u_buf[i];
            u_save[i] = uo;
            sg += uo * sv[i];
        }
        sg = block_sum(sg, sh);
        if (threadIdx.x == 0) *sigma = sg;
    }
}

__global__ void __launch_bounds__(SN_THREADS) k_sn_wv(const float* __restrict__ W, const float* __restrict__ t, const float* __restrict__ nrm_part,
                                                      int nchunks, int R, int K, int power, float eps, float* s, float* u_buf,
                                                      float* __restrict__ u_save, float* __restrict__ v_buf, float* __restrict__ v_save,
                                                      float* __restrict__ sigma, int lane) {
    sn_wv_body(W, t, nrm_part, nchunks, R, K, power, eps, s, u_buf, u_save, v_buf, v_save, sigma, blockIdx.x, gridDim.x, &g_sn_ticket[lane][1]);
}

// phase 3: W_sn (R, taps, Cin) = W (R, Cin, taps) / sigma.  Block = (SN_CI input channels of one row): a coalesced
// read of the SN_CI*taps contiguous floats, transposed through shared memory (odd pitch), coalesced writes per tap.
#define SN_CI 128
#define SN_MAXTAPS 16
__global__ void __launch_bounds__(SN_CI) k_sn_scale(const float* __restrict__ W, const float* __restrict__ sigma, int Cin, int taps,
                                                    float* __restrict__ out) {
    __shared__ float sm[SN_CI * (SN_MAXTAPS + 1)];
    float inv = 1.f / *sigma;
    int r = blockIdx.y, ci0 = blockIdx.x * SN_CI;
    int nci = min(SN_CI, Cin - ci0);
    size_t K = (size_t)Cin * taps;
    if (taps == 1) {
        if ((int)threadIdx.x < nci) out[r * K + ci0 + threadIdx.x] = W[r * K + ci0 + threadIdx.x] * inv;
        return;
    }
    int tp = taps | 1;
    const float* src = W + r * K + (size_t)ci0 * taps;
    for (int e = threadIdx.x; e < nci * taps; e += SN_CI) {
        int j = e / taps, t = e - j * taps;
        sm[j * tp + t] = src[e] * inv;
    }
    __syncthreads();
    if ((int)threadIdx.x < nci)
        for (int t = 0; t < taps; ++t) out[r * K + (size_t)t * Cin + ci0 + threadIdx.x] = sm[threadIdx.x * tp + t];
}

// phase 3, two-output variant: besides W_sn (R, taps, Cin) also emit wt (Cin, taps, R) = the same weight with its
// channel axes swapped, which is the B operand of the tcgen05 data gradient (conv_tc.cu) -- otherwise a strided torch
// copy per conv per step.  Block = (RB rows) x (32 input channels) x all taps, staged in shared memory with odd pitches
// so that both output orders are written in full 128-byte (64-byte for RB = 16) segments without bank conflicts.
__device__ __forceinline__ void sn_scale_t_body(const float* __restrict__ W, const float* __restrict__ sigma, int R, int Cin, int taps, int RB,
                                                float* __restrict__ out, float* __restrict__ wt, int bx, int by) {
    extern __shared__ float smt[];                 // [(row * 33 + j) * tp + t]
    const float inv = 1.f / *sigma;
    const int tp = taps | 1;
    const int r0 = by * RB, ci0 = bx * 32;
    const int nr = min(RB, R - r0), nj = min(32, Cin - ci0);
    const size_t K = (size_t)Cin * taps;
    const int run = nj * taps;                     // contiguous floats per row in the source
    for (int e = threadIdx.x; e < nr * run; e += SN_THREADS) {
        const int row = e / run, off = e - row * run;
        const int j = off / taps, t = off - j * taps;
        smt[(row * 33 + j) * tp + t] = W[(size_t)(r0 + row) * K + (size_t)ci0 * taps + off] * inv;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nr * taps * 32; e += SN_THREADS) {        // W_sn[r][t][ci]
        const int j = e & 31, rt = e >> 5;
        const int t = rt % taps, row = rt / taps;
        if (j < nj) out[(size_t)(r0 + row) * K + (size_t)t * Cin + ci0 + j] = smt[(row * 33 + j) * tp + t];
    }
    if (wt == nullptr) return;
    for (int e = threadIdx.x; e < nj * taps * RB; e += SN_THREADS) {        // wt[ci][t][r]
        const int row = e % RB, jt = e / RB;
        const int t = jt % taps, j = jt / taps;
        if (row < nr) wt[((size_t)(ci0 + j) * taps + t) * R + r0 + row] = smt[(row * 33 + j) * tp + t];
    }
}
__global__ void __launch_bounds__(SN_THREADS) k_sn_scale_t(const float* __restrict__ W, const float* __restrict__ sigma, int R, int Cin,
                                                           int taps, int RB, float* __restrict__ out, float* __restrict__ wt) {
    sn_scale_t_body(W, sigma, R, Cin, taps, RB, out, wt, blockIdx.x, blockIdx.y);
}

// backward phase 1: c = sum dW_sn * W_sn  (both OHWI, contiguous)
__device__ __forceinline__ void sn_dot_body(const float* __restrict__ a, const float* __restrict__ b, long long total, float* part,
                                            float* __restrict__ c_out, int blk, int nblk, unsigned int* ticket) {
    __shared__ float sh[SN_THREADS / 32];
    __shared__ int flag;
    float acc = 0.f;
    for (long long i = (long long)blk * SN_THREADS + threadIdx.x; i < total; i += (long long)nblk * SN_THREADS)
        acc = fmaf(a[i], b[i], acc);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) part[blk] = acc;
    if (!last_block(ticket, nblk, &flag)) return;
    const volatile float* pv = part;
    float t = 0.f;
    for (int i = threadIdx.x; i < nblk; i += SN_THREADS) t += pv[i];
    t = block_sum(t, sh);
    if (threadIdx.x == 0) *c_out = t;
}
__global__ void __launch_bounds__(SN_THREADS) k_sn_dot(const float* __restrict__ a, const float* __restrict__ b, long long total,
                                                       float* part, float* __restrict__ c_out, int lane) {
    sn_dot_body(a, b, total, part, c_out, blockIdx.x, gridDim.x, &g_sn_ticket[lane][2]);
}

// backward phase 2: dW (R, Cin, taps) = (dW_sn (R, taps, Cin) - c u v^T) / sigma; same tiling as k_sn_scale, transposing back
__device__ __forceinline__ void sn_bwd_body(const float* __restrict__ dws, const float* __restrict__ u, const float* __restrict__ v,
                                            const float* __restrict__ sigma, const float* __restrict__ c, int Cin, int taps,
                                            float* __restrict__ dw, int bx, int by) {
    __shared__ float sm[SN_CI * (SN_MAXTAPS + 1)];
    float inv = 1.f / *sigma;
    int r = by, ci0 = bx * SN_CI;
    int nci = min(SN_CI, Cin - ci0);
    size_t K = (size_t)Cin * taps;
    float cu = *c * u[r];
    if (taps == 1) {
        if ((int)threadIdx.x < nci) {
            size_t i = r * K + ci0 + threadIdx.x;
            dw[i] = (dws[i] - cu * v[ci0 + threadIdx.x]) * inv;
        }
        return;
    }
    int tp = taps | 1;
    if ((int)threadIdx.x < nci)
        for (int t = 0; t < taps; ++t) sm[threadIdx.x * tp + t] = dws[r * K + (size_t)t * Cin + ci0 + threadIdx.x];
    __syncthreads();
    float* dst = dw + r * K + (size_t)ci0 * taps;
    const float* vv = v + (size_t)ci0 * taps;
    for (int e = threadIdx.x; e < nci * taps; e += SN_CI) {
        int j = e / taps, t = e - j * taps;
        dst[e] = (sm[j * tp + t] - cu * vv[e]) * inv;
    }
}
__global__ void __launch_bounds__(SN_CI) k_sn_bwd(const float* __restrict__ dws, const float* __restrict__ u, const float* __restrict__ v,
                                                  const float* __restrict__ sigma, const float* __restrict__ c, int Cin, int taps,
                                                  float* __restrict__ dw) {
    sn_bwd_body(dws, u, v, sigma, c, Cin, taps, dw, blockIdx.x, blockIdx.y);
}


static int sn_splits(int R, int K, int* rows_per_split) {
    int gx = fsv_cdiv(K, SN_THREADS);
    int want = (2 * fsv_sm_count() + gx - 1) / gx;
    int rs = want < 1 ? 1 : want;
    int max_rs = R / 16 > 0 ? R / 16 : 1;
    if (rs > max_rs) rs = max_rs;
    if (rs > 32) rs = 32;
    *rows_per_split = fsv_cdiv(R, rs);
    return fsv_cdiv(R, *rows_per_split);
}

// ------------------------------------------------------------------------------------------------ grouped forward
// All spectral weights a network touches in one forward, in THREE launches (one per phase) instead of three per weight:
// ~110 weights x 3 phases x 2 generator forwards (+ the discriminators) per training step were ~660 dependent launches of
// pure latency.  The host builds a plan once (fsv_spectral_group_plan): per-weight descriptors with absolute parameter /
// buffer pointers (stable for the life of the modules), offsets into a per-call output arena and work arena (their base
// pointers are kernel arguments, so fresh allocations per call need no table update -> CUDA-graph friendly), and per-phase
// block maps (flattened grid -> (item, local block)).  Tickets of the last-block reductions live in a per-plan buffer,
// so groups on different streams do not share counters.  Arithmetic and summation order per weight are those of the
// single-weight kernels above (same bodies).
__global__ void __launch_bounds__(SN_THREADS) k_sng_wtu(const fsv_sn_item* __restrict__ items, const int2* __restrict__ map, float* work,
                                                        unsigned int* tickets) {
    const int2 m = map[blockIdx.x];
    const fsv_sn_item it = items[m.x];
    const int chunk = m.y % it.nchunks, split = m.y / it.nchunks;
    float* w = work + it.work_off;
    sn_wtu_body(it.w_orig, it.u, it.R, it.K, it.rps, w, w + (size_t)it.rs * it.K + it.R, chunk, split, it.rs, tickets + it.ticket_off + chunk);
}
__global__ void __launch_bounds__(SN_THREADS) k_sng_wv(const fsv_sn_item* __restrict__ items, const int2* __restrict__ map, float* work,
                                                       float* out, unsigned int* tickets, int power, float eps) {
    const int2 m = map[blockIdx.x];
    const fsv_sn_item it = items[m.x];
    float* w = work + it.work_off;
    float* s = w + (size_t)it.rs * it.K;
    float* uvs = out + it.uvs_off;
    sn_wv_body(it.w_orig, power ? w : it.v, s + it.R, it.nchunks, it.R, it.K, power, eps, s, it.u, uvs + it.K, it.v, uvs, uvs + it.K + it.R,
               m.y, it.nblk2, tickets + it.ticket_off + it.nchunks);
}
__global__ void __launch_bounds__(SN_THREADS) k_sng_scale(const fsv_sn_item* __restrict__ items, const int2* __restrict__ map, float* out,
                                                          int emit_wt) {
    const int2 m = map[blockIdx.x];
    const fsv_sn_item it = items[m.x];
    const int gx = (it.Cin + 31) / 32;
    const int RB = it.taps <= 9 ? 32 : 16;
    sn_scale_t_body(it.w_orig, out + it.uvs_off + it.K + it.R, it.R, it.Cin, it.taps, RB, out + it.out_off,
                    (emit_wt && it.want_wt) ? out + it.wt_off : nullptr, m.y % gx, m.y / gx);
}

extern "C" int fsv_spectral_group_plan(fsv_sn_item* items, int n, long long* totals) {
    FSV_REQUIRE(items && totals && n > 0, "spectral_group_plan: bad args");
    long long out = 0, work = 0, tick = 0, b1 = 0, b2 = 0, b3 = 0, bwork = 0, bb1 = 0, bb2 = 0;
    for (int i = 0; i < n; ++i) {
        fsv_sn_item& it = items[i];
        FSV_REQUIRE(it.R > 0 && it.R <= 65535 && it.Cin > 0 && it.taps > 0 && it.taps <= SN_MAXTAPS, "spectral_group_plan: bad dims in item %d", i);
        it.K = it.Cin * it.taps;
        it.nchunks = fsv_cdiv(it.K, SN_THREADS);
        it.rs = sn_splits(it.R, it.K, &it.rps);
        it.nblk2 = fsv_cdiv(it.R, SN_THREADS / 32);
        const int RB = it.taps <= 9 ? 32 : 16;
        it.nblk3 = fsv_cdiv(it.Cin, 32) * fsv_cdiv(it.R, RB);
        const long long rk = (long long)it.R * it.K;
        it.out_off = out; out += (rk + 3) / 4 * 4;
        it.wt_off = out; if (it.want_wt) out += (rk + 3) / 4 * 4;
        it.uvs_off = out; out += ((long long)it.K + it.R + 1 + 3) / 4 * 4;     // [v | u | sigma], 16-byte aligned slots
        it.work_off = work; work += ((long long)it.rs * it.K + it.R + it.nchunks + 3) / 4 * 4;
        it.ticket_off = (int)tick; tick += it.nchunks + 2;      // phase-1 tickets per chunk, one for phase 2, one for the backward dot
        it.blk1 = (int)b1; b1 += (long long)it.nchunks * it.rs;
        it.blk2 = (int)b2; b2 += it.nblk2;
        it.blk3 = (int)b3; b3 += it.nblk3;
        long long nb1 = (rk + SN_THREADS * 4 - 1) / (SN_THREADS * 4);
        const long long cap = 2LL * fsv_sm_count();
        if (nb1 > cap) nb1 = cap;
        it.nb1 = (int)nb1;
        it.nb2 = fsv_cdiv(it.Cin, SN_CI) * it.R;
        it.bwd_work_off = bwork; bwork += (nb1 + 1 + 3) / 4 * 4;
        bb1 += it.nb1; bb2 += it.nb2;
    }
    totals[0] = out; totals[1] = work; totals[2] = tick; totals[3] = b1; totals[4] = b2; totals[5] = b3;
    totals[6] = bwork; totals[7] = bb1; totals[8] = bb2;
    return FSV_OK;
}

// map layout: [b1 int2 | b2 int2 | b3 int2]; filled on the host by the caller from the plan (item index, local block)
extern "C" int fsv_spectral_group_fwd(const fsv_sn_item* items_dev, const int* map_dev, const long long* totals, int power, float eps,
                                      int emit_wt, float* out, float* work, unsigned int* tickets, void* stream) {
    FSV_REQUIRE(items_dev && map_dev && totals && out && work && tickets, "spectral_group_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int2* map = reinterpret_cast<const int2*>(map_dev);
    const long long b1 = totals[3], b2 = totals[4], b3 = totals[5];
    if (power) {
        k_sng_wtu<<<(unsigned)b1, SN_THREADS, 0, st>>>(items_dev, map, work, tickets);
        FSV_CHECK_LAUNCH("spectral_group_wtu");
    }
    k_sng_wv<<<(unsigned)b2, SN_THREADS, 0, st>>>(items_dev, map + b1, work, out, tickets, power, eps);
    FSV_CHECK_LAUNCH("spectral_group_wv");
    const size_t sm = (size_t)32 * 33 * 9 * sizeof(float) > (size_t)16 * 33 * 17 * sizeof(float) ? (size_t)32 * 33 * 9 * sizeof(float) : (size_t)16 * 33 * 17 * sizeof(float);
    k_sng_scale<<<(unsigned)b3, SN_THREADS, sm, st>>>(items_dev, map + b1 + b2, out, emit_wt);
    FSV_CHECK_LAUNCH("spectral_group_scale");
    return FSV_OK;
}

// ------------------------------------------------------------------------------------------------ grouped backward
// dW_orig of every spectral weight of a group in TWO launches.  The incoming dW_sn tensors are whatever the conv weight-gradient
// kernels allocated, so their addresses come in a per-call device table `dws` (NULL = this weight received no gradient: skipped);
// W_sn / [v|u|sigma] are read from the forward's arena, dW_orig goes to a per-call arena at the forward arena's out_off offsets.
__global__ void __launch_bounds__(SN_THREADS) k_sng_dot(const fsv_sn_item* __restrict__ items, const int2* __restrict__ map,
                                                        const float* const* __restrict__ dws, const float* __restrict__ out, float* work,
                                                        unsigned int* tickets) {
    const int2 m = map[blockIdx.x];
    const fsv_sn_item it = items[m.x];
    const float* g = dws[m.x];
    if (g == nullptr) return;
    float* w = work + it.bwd_work_off;
    sn_dot_body(g, out + it.out_off, (long long)it.R * it.K, w, w + it.nb1, m.y, it.nb1, tickets + it.ticket_off + it.nchunks + 1);
}
__global__ void __launch_bounds__(SN_CI) k_sng_bwd(const fsv_sn_item* __restrict__ items, const int2* __restrict__ map,
                                                   const float* const* __restrict__ dws, const float* __restrict__ out, const float* __restrict__ work,
                                                   float* __restrict__ dw_arena) {
    const int2 m = map[blockIdx.x];
    const fsv_sn_item it = items[m.x];
    const float* g = dws[m.x];
    if (g == nullptr) return;
    const float* uvs = out + it.uvs_off;
    const int gx = (it.Cin + SN_CI - 1) / SN_CI;
    sn_bwd_body(g, uvs + it.K, uvs, uvs + it.K + it.R, work + it.bwd_work_off + it.nb1, it.Cin, it.taps, dw_arena + it.out_off, m.y % gx, m.y / gx);
}

extern "C" int fsv_spectral_group_bwd(const fsv_sn_item* items_dev, const int* map_bwd_dev, const long long* totals_bwd, const float* const* dws_dev,
                                      const float* out, float* dw_arena, float* work, unsigned int* tickets, void* stream) {
    FSV_REQUIRE(items_dev && map_bwd_dev && totals_bwd && dws_dev && out && dw_arena && work && tickets, "spectral_group_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int2* map = reinterpret_cast<const int2*>(map_bwd_dev);
    k_sng_dot<<<(unsigned)totals_bwd[1], SN_THREADS, 0, st>>>(items_dev, map, dws_dev, out, work, tickets);
    FSV_CHECK_LAUNCH("spectral_group_dot");
    k_sng_bwd<<<(unsigned)totals_bwd[2], SN_CI, 0, st>>>(items_dev, map + totals_bwd[1], dws_dev, out, work, dw_arena);
    FSV_CHECK_LAUNCH("spectral_group_bwd");
    return FSV_OK;
}

extern "C" long long fsv_spectral_workspace(int R, int K) {
    int rps;
    int rs = sn_splits(R, K, &rps);
    long long a = (long long)rs * K + R + fsv_cdiv(K, SN_THREADS);   // forward: partials (t in split 0) + s + per-chunk |t|^2
    long long b = 4096 + 1;                                          // backward: block partials + c
    return (a > b ? a : b) * (long long)sizeof(float);
}

extern "C" int fsv_spectral_fwd(const float* w_orig, float* u, float* v, int R, int Cin, int taps, int power, float eps,
                                float* w_out, float* wt_out, float* uvs, float* work, void* stream) {
    FSV_REQUIRE(R > 0 && R <= 65535 && Cin > 0 && taps > 0 && taps <= SN_MAXTAPS, "spectral_fwd: bad dims (R %d Cin %d taps %d)", R, Cin, taps);
    FSV_REQUIRE(w_orig && u && v && w_out && uvs && work, "spectral_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int K = Cin * taps;
    int nchunks = fsv_cdiv(K, SN_THREADS);
    FSV_REQUIRE(nchunks <= SN_MAXCHUNKS, "spectral_fwd: K = %d too large", K);
    int rps;
    int rs = sn_splits(R, K, &rps);
    float* part = work;
    float* s = work + (size_t)rs * K;
    float* nrm_part = s + R;
    float* v_save = uvs;            // uvs = [v (K) | u (R) | sigma]: v first keeps it 16-byte aligned for the float4 loads
    float* u_save = uvs + K;
    float* sigma = uvs + K + R;
    const int lane = fsv_current_reduction_lane();
    FSV_REQUIRE(lane >= 0 && lane < SN_LANES, "spectral_fwd: reduction lane %d out of range", lane);
    if (power) {
        dim3 g(nchunks, rs);
        k_sn_wtu<<<g, SN_THREADS, 0, st>>>(w_orig, u, R, K, rps, part, nrm_part, lane);
        FSV_CHECK_LAUNCH("spectral_wtu");
    }
    k_sn_wv<<<fsv_cdiv(R, SN_THREADS / 32), SN_THREADS, 0, st>>>(w_orig, power ? part : v, nrm_part, nchunks, R, K, power, eps, s, u, u_save,
                                                                 v, v_save, sigma, lane);
    FSV_CHECK_LAUNCH("spectral_wv");
    if (wt_out) {
        const int RB = taps <= 9 ? 32 : 16;
        const size_t sm = (size_t)RB * 33 * (taps | 1) * sizeof(float);
        dim3 gt(fsv_cdiv(Cin, 32), fsv_cdiv(R, RB));
        FSV_REQUIRE(gt.y <= 65535, "spectral_fwd: too many rows");
        k_sn_scale_t<<<gt, SN_THREADS, sm, st>>>(w_orig, sigma, R, Cin, taps, RB, w_out, wt_out);
        FSV_CHECK_LAUNCH("spectral_scale_t");
        return FSV_OK;
    }
    dim3 g3(fsv_cdiv(Cin, SN_CI), R);
    k_sn_scale<<<g3, SN_CI, 0, st>>>(w_orig, sigma, Cin, taps, w_out);
    FSV_CHECK_LAUNCH("spectral_scale");
    return FSV_OK;
}

extern "C" int fsv_spectral_bwd(const float* dw_ohwi, const float* w_sn_ohwi, const float* uvs, int R, int Cin, int taps, float* dw_orig,
                                float* work, void* stream) {
    FSV_REQUIRE(R > 0 && R <= 65535 && Cin > 0 && taps > 0 && taps <= SN_MAXTAPS, "spectral_bwd: bad dims (R %d Cin %d taps %d)", R, Cin, taps);
    FSV_REQUIRE(dw_ohwi && w_sn_ohwi && uvs && dw_orig && work, "spectral_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int K = Cin * taps;
    long long total = (long long)R * K;
    int blocks = (int)((total + SN_THREADS * 4 - 1) / (SN_THREADS * 4));
    int cap = 2 * fsv_sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks > 4096) blocks = 4096;
    float* part = work;
    float* c = work + 4096;
    const int lane = fsv_current_reduction_lane();
    FSV_REQUIRE(lane >= 0 && lane < SN_LANES, "spectral_bwd: reduction lane %d out of range", lane);
    k_sn_dot<<<blocks, SN_THREADS, 0, st>>>(dw_ohwi, w_sn_ohwi, total, part, c, lane);
    FSV_CHECK_LAUNCH("spectral_dot");
    dim3 g2(fsv_cdiv(Cin, SN_CI), R);
    k_sn_bwd<<<g2, SN_CI, 0, st>>>(dw_ohwi, uvs + K, uvs, uvs + K + R, c, Cin, taps, dw_orig);
    FSV_CHECK_LAUNCH("spectral_bwd");
    return FSV_OK;
}
