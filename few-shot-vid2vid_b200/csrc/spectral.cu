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

__device__ unsigned int g_sn_ticket[4];

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

// phase 1: t = W^T u  (partials over row splits), last block: v = t / max(|t|, eps)
__global__ void __launch_bounds__(SN_THREADS) k_sn_wtu(const float* __restrict__ W, const float* __restrict__ u, int R, int K,
                                                       int rows_per_split, float eps, float* part,
                                                       float* __restrict__ v_buf, float* __restrict__ v_save) {
    __shared__ float sh[SN_THREADS / 32];
    __shared__ int flag;
    int c = blockIdx.x * SN_THREADS + threadIdx.x;
    int r0 = blockIdx.y * rows_per_split, r1 = min(R, r0 + rows_per_split);
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
        part[(size_t)blockIdx.y * K + c] = (a0 + a1) + (a2 + a3);
    }
    if (!last_block(&g_sn_ticket[0], gridDim.x * gridDim.y, &flag)) return;
    // this block alone: reduce the partials (fixed order), normalise
    const volatile float* vp = part;
    float nrm = 0.f;
    for (int k = threadIdx.x; k < K; k += SN_THREADS) {
        float t = 0.f;
        for (int s = 0; s < (int)gridDim.y; ++s) t += vp[(size_t)s * K + k];
        part[k] = t;       // split 0's slot now holds the full sum (each k touched by one thread only)
        nrm += t * t;
    }
    nrm = block_sum(nrm, sh);
    float inv = 1.f / fmaxf(sqrtf(nrm), eps);
    for (int k = threadIdx.x; k < K; k += SN_THREADS) {
        float t = part[k] * inv;
        v_buf[k] = t;
        v_save[k] = t;
    }
}

// phase 2: s = W v (warp per row); last block: sigma, u
__global__ void __launch_bounds__(SN_THREADS) k_sn_wv(const float* __restrict__ W, const float* __restrict__ v, int R, int K, int power,
                                                      float eps, float* s, float* u_buf,
                                                      float* __restrict__ u_save, float* __restrict__ sigma) {
    __shared__ float sh[SN_THREADS / 32];
    __shared__ int flag;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int r = blockIdx.x * (SN_THREADS / 32) + warp;
    if (r < R) {
        const float* wr = W + (size_t)r * K;
        float a = 0.f;
        if ((K & 3) == 0) {
            const float4* w4 = reinterpret_cast<const float4*>(wr);
            const float4* v4 = reinterpret_cast<const float4*>(v);
            for (int k = lane; k < (K >> 2); k += 32) {
                float4 a4 = w4[k], b4 = v4[k];
                a = fmaf(a4.x, b4.x, a);
                a = fmaf(a4.y, b4.y, a);
                a = fmaf(a4.z, b4.z, a);
                a = fmaf(a4.w, b4.w, a);
            }
        } else {
            for (int k = lane; k < K; k += 32) a = fmaf(wr[k], v[k], a);
        }
        a = warp_sum(a);
        if (lane == 0) s[r] = a;
    }
    if (!last_block(&g_sn_ticket[1], gridDim.x, &flag)) return;
    const volatile float* sv = s;
    if (power) {
        float nrm = 0.f;
        for (int i = threadIdx.x; i < R; i += SN_THREADS) nrm += sv[i] * sv[i];
        nrm = block_sum(nrm, sh);
        float inv = 1.f / fmaxf(sqrtf(nrm), eps);
        float sg = 0.f;
        for (int i = threadIdx.x; i < R; i += SN_THREADS) {
            float un = sv[i] * inv;
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

// phase 3: W_sn (R, taps, Cin) = W (R, Cin, taps) / sigma
__global__ void __launch_bounds__(SN_THREADS) k_sn_scale(const float* __restrict__ W, const float* __restrict__ sigma, int Cin, int taps,
                                                         long long total, float* __restrict__ out) {
    float inv = 1.f / *sigma;
    int K = Cin * taps;
    for (long long i = (long long)blockIdx.x * SN_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * SN_THREADS) {
        long long r = i / K;
        int k = (int)(i - r * K);
        int t = k / Cin, ci = k - t * Cin;
        out[i] = W[r * K + (long long)ci * taps + t] * inv;
    }
}

// backward phase 1: c = sum dW_sn * W_sn  (both OHWI, contiguous)
__global__ void __launch_bounds__(SN_THREADS) k_sn_dot(const float* __restrict__ a, const float* __restrict__ b, long long total,
                                                       float* part, float* __restrict__ c_out) {
    __shared__ float sh[SN_THREADS / 32];
    __shared__ int flag;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * SN_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * SN_THREADS)
        acc = fmaf(a[i], b[i], acc);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
    if (!last_block(&g_sn_ticket[2], gridDim.x, &flag)) return;
    const volatile float* pv = part;
    float t = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += SN_THREADS) t += pv[i];
    t = block_sum(t, sh);
    if (threadIdx.x == 0) *c_out = t;
}

// backward phase 2: dW (R, Cin, taps) = (dW_sn (R, taps, Cin) - c u v^T) / sigma
__global__ void __launch_bounds__(SN_THREADS) k_sn_bwd(const float* __restrict__ dws, const float* __restrict__ u, const float* __restrict__ v,
                                                       const float* __restrict__ sigma, const float* __restrict__ c, int Cin, int taps,
                                                       long long total, float* __restrict__ dw) {
    float inv = 1.f / *sigma, cc = *c;
    int K = Cin * taps;
    for (long long i = (long long)blockIdx.x * SN_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * SN_THREADS) {
        long long r = i / K;
        int k = (int)(i - r * K);         // OIHW column: ci * taps + t
        int ci = k / taps, t = k - ci * taps;
        float g = dws[r * K + (long long)t * Cin + ci];
        dw[i] = (g - cc * u[r] * v[k]) * inv;
    }
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

extern "C" long long fsv_spectral_workspace(int R, int K) {
    int rps;
    int rs = sn_splits(R, K, &rps);
    long long a = (long long)rs * K + R;   // forward: partials + s
    long long b = 4096 + 1;                // backward: block partials + c
    return (a > b ? a : b) * (long long)sizeof(float);
}

extern "C" int fsv_spectral_fwd(const float* w_orig, float* u, float* v, int R, int Cin, int taps, int power, float eps,
                                float* w_out, float* uvs, float* work, void* stream) {
    FSV_REQUIRE(R > 0 && Cin > 0 && taps > 0, "spectral_fwd: bad dims");
    FSV_REQUIRE(w_orig && u && v && w_out && uvs && work, "spectral_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int K = Cin * taps;
    int rps;
    int rs = sn_splits(R, K, &rps);
    float* part = work;
    float* s = work + (size_t)rs * K;
    float* v_save = uvs;            // uvs = [v (K) | u (R) | sigma]: v first keeps it 16-byte aligned for the float4 loads
    float* u_save = uvs + K;
    float* sigma = uvs + K + R;
    if (power) {
        dim3 g(fsv_cdiv(K, SN_THREADS), rs);
        k_sn_wtu<<<g, SN_THREADS, 0, st>>>(w_orig, u, R, K, rps, eps, part, v, v_save);
        FSV_CHECK_LAUNCH("spectral_wtu");
    } else {
        FSV_CUDA(cudaMemcpyAsync(v_save, v, (size_t)K * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    k_sn_wv<<<fsv_cdiv(R, SN_THREADS / 32), SN_THREADS, 0, st>>>(w_orig, v_save, R, K, power, eps, s, u, u_save, sigma);
    FSV_CHECK_LAUNCH("spectral_wv");
    long long total = (long long)R * K;
    int blocks = (int)((total + SN_THREADS * 4 - 1) / (SN_THREADS * 4));
    int cap = 8 * fsv_sm_count();
    if (blocks > cap) blocks = cap;
    k_sn_scale<<<blocks, SN_THREADS, 0, st>>>(w_orig, sigma, Cin, taps, total, w_out);
    FSV_CHECK_LAUNCH("spectral_scale");
    return FSV_OK;
}

extern "C" int fsv_spectral_bwd(const float* dw_ohwi, const float* w_sn_ohwi, const float* uvs, int R, int Cin, int taps, float* dw_orig,
                                float* work, void* stream) {
    FSV_REQUIRE(R > 0 && Cin > 0 && taps > 0, "spectral_bwd: bad dims");
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
    k_sn_dot<<<blocks, SN_THREADS, 0, st>>>(dw_ohwi, w_sn_ohwi, total, part, c);
    FSV_CHECK_LAUNCH("spectral_dot");
    int b2 = (int)((total + SN_THREADS * 4 - 1) / (SN_THREADS * 4));
    int cap2 = 8 * fsv_sm_count();
    if (b2 > cap2) b2 = cap2;
    k_sn_bwd<<<b2, SN_THREADS, 0, st>>>(dw_ohwi, uvs + K, uvs, uvs + K + R, c, Cin, taps, total, dw_orig);
    FSV_CHECK_LAUNCH("spectral_bwd");
    return FSV_OK;
}
