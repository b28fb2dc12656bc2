// Batch / instance normalisation on NHWC activations.
//
// Replaces the reference's (Sync)BatchNorm2d / InstanceNorm2d library calls
// (models/networks/normalization.py:32-35,78-82) and the LeakyReLU that always follows
// them (architecture.py:65-68, generator.py:473-486, discriminator.py:76-85).
//
// Roofline: all kernels here are HBM-streaming.  Training-mode forward is two passes
// (statistics, then normalise+affine+activation): algorithmic bytes = 4*(2|x| + |y|).
// Per-channel reductions walk NHWC rows with one warp per 32 consecutive channels
// (128-byte coalesced rows), fp32 partials per thread, fp64 combine.
#include "common.cuh"
#include <stdlib.h>

#define RED_TX 32
#define RED_TY 8
#define RED_MAX_TICKETS 8192

// Tickets of the last-block reductions, one bank per "reduction lane": kernels of one lane must be stream-ordered among themselves,
// different lanes may run concurrently (the drop-in generator runs its reference-encoder / hyper-network branch on a second stream
// next to the flow / warp branch; each stream selects its own lane with fsv_set_reduction_lane, a thread-local host setting).
// Lanes: 0 main stream, 1 / 3 generator branch streams, 2 discriminator-update stream, 4 .. 7 weight-gradient lanes (fsv/ops.py).
#define RED_LANES 8
__device__ unsigned int g_red_ticket[RED_LANES][RED_MAX_TICKETS];
static thread_local int t_red_lane = 0;
extern "C" int fsv_set_reduction_lane(int lane) {
    FSV_REQUIRE(lane >= 0 && lane < RED_LANES, "set_reduction_lane: lane must be in [0, %d)", RED_LANES);
    t_red_lane = lane;
    return FSV_OK;
}
int fsv_current_reduction_lane() { return t_red_lane; }       // spectral.cu keeps its per-weight tickets in the same lanes

// Generic two-value per-(group, channel) reduction over the rows of an NHWC slice: f(row, c) -> float2.
// grid (row blocks, 32-channel blocks, groups), block 32 x 8.  Every block leaves its fp64 partials in `part`; the last
// block to finish for a (group, channel block) -- one ticket each, reset by that block -- adds the partials in a fixed
// order and hands the totals to the epilogue functor e(g, c, a, b).  No memsets, no float atomics: one launch, and the
// result does not depend on block scheduling.
template <class F, class E>
__global__ void __launch_bounds__(RED_TX * RED_TY) k_chan_reduce2(F f, E e, long long rows_per_group, long long rows_per_block, int C,
                                                                  double* part, int lane) {
    __shared__ double s_a[RED_TY][RED_TX], s_b[RED_TY][RED_TX];
    __shared__ int s_last;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = blockIdx.y * RED_TX + tx;
    const int g = blockIdx.z;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > rows_per_group) r1 = rows_per_group;
    float sa0 = 0.f, sb0 = 0.f, sa1 = 0.f, sb1 = 0.f, sa2 = 0.f, sb2 = 0.f, sa3 = 0.f, sb3 = 0.f;
    if (c < C) {
        const long long base = g * rows_per_group;
        const typename F::Ctx ctx = f.begin(g, c);          // per-(group, channel) constants, loaded once per thread
        long long r = r0 + ty;
        for (; r + 3 * RED_TY < r1; r += 4 * RED_TY) {
            float2 v0 = f(ctx, base + r, c), v1 = f(ctx, base + r + RED_TY, c), v2 = f(ctx, base + r + 2 * RED_TY, c),
                   v3 = f(ctx, base + r + 3 * RED_TY, c);
            sa0 += v0.x; sb0 += v0.y; sa1 += v1.x; sb1 += v1.y; sa2 += v2.x; sb2 += v2.y; sa3 += v3.x; sb3 += v3.y;
        }
        for (; r < r1; r += RED_TY) {
            float2 v = f(ctx, base + r, c);
            sa0 += v.x; sb0 += v.y;
        }
    }
    s_a[ty][tx] = ((double)sa0 + (double)sa1) + ((double)sa2 + (double)sa3);
    s_b[ty][tx] = ((double)sb0 + (double)sb1) + ((double)sb2 + (double)sb3);
    __syncthreads();
    const long long slot = ((long long)g * gridDim.y + blockIdx.y) * gridDim.x;       // first row block of this (g, channel block)
    if (ty == 0) {
        double ta = 0.0, tb = 0.0;
#pragma unroll
        for (int i = 0; i < RED_TY; ++i) { ta += s_a[i][tx]; tb += s_b[i][tx]; }
        double* pp = part + ((slot + blockIdx.x) * RED_TX + tx) * 2;
        pp[0] = ta; pp[1] = tb;
    }
    __threadfence();
    __syncthreads();
    if (tx == 0 && ty == 0) {
        unsigned int* tk = &g_red_ticket[lane][g * gridDim.y + blockIdx.y];
        unsigned int t = atomicAdd(tk, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) *tk = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const volatile double* vp = part;
    double ta = 0.0, tb = 0.0;
    for (int rb = ty; rb < (int)gridDim.x; rb += RED_TY) {
        const long long o = ((slot + rb) * RED_TX + tx) * 2;
        ta += vp[o]; tb += vp[o + 1];
    }
    s_a[ty][tx] = ta; s_b[ty][tx] = tb;
    __syncthreads();
    if (ty == 0 && c < C) {
        ta = 0.0; tb = 0.0;
#pragma unroll
        for (int i = 0; i < RED_TY; ++i) { ta += s_a[i][tx]; tb += s_b[i][tx]; }
        e(g, c, ta, tb);
    }
}

static void reduce_geometry(int groups, long long rows_per_group, int C, int* row_blocks, long long* rows_per_block) {
    const int cb = fsv_cdiv(C, RED_TX);
    long long want = (4LL * fsv_sm_count() + (long long)cb * groups - 1) / ((long long)cb * groups);
    long long max_rb = (rows_per_group + 63) / 64;
    if (want > max_rb) want = max_rb;
    if (want < 1) want = 1;
    long long rpb = (rows_per_group + want - 1) / want;
    rpb = ((rpb + RED_TY - 1) / RED_TY) * RED_TY;
    *rows_per_block = rpb;
    *row_blocks = (int)((rows_per_group + rpb - 1) / rpb);
}
// doubles of workspace a reduction over (groups, rows_per_group, C) needs
extern "C" long long fsv_norm_work_doubles(int groups, int C, long long rows_per_group) {
    int rb; long long rpb;
    reduce_geometry(groups, rows_per_group, C, &rb, &rpb);
    return (long long)groups * fsv_cdiv(C, RED_TX) * rb * RED_TX * 2;
}

template <class F, class E>
static int launch_reduce2(F f, E e, int groups, long long rows_per_group, int C, double* work, cudaStream_t st, const char* name) {
    FSV_REQUIRE(work != nullptr, "%s: null workspace", name);
    FSV_REQUIRE((long long)groups * fsv_cdiv(C, RED_TX) <= RED_MAX_TICKETS && groups <= 65535, "%s: too many (group, channel-block) pairs", name);
    int rb; long long rpb;
    reduce_geometry(groups, rows_per_group, C, &rb, &rpb);
    dim3 grid(rb, fsv_cdiv(C, RED_TX), groups);
    dim3 block(RED_TX, RED_TY);
    k_chan_reduce2<<<grid, block, 0, st>>>(f, e, rows_per_group, rpb, C, work, t_red_lane);
    FSV_CHECK_LAUNCH(name);
    return FSV_OK;
}

struct StoreE {   // totals -> a[g*C+c], b[g*C+c]
    double *a, *b;
    int C;
    __device__ void operator()(int g, int c, double ta, double tb) const {
        a[(long long)g * C + c] = ta;
        b[(long long)g * C + c] = tb;
    }
};

struct StoreParamE {   // one group: totals -> a[c], b[c] and, as floats, the affine parameter gradients (dbias = a, dweight = b)
    double *a, *b;
    float *dweight, *dbias;
    __device__ void operator()(int, int c, double ta, double tb) const {
        a[c] = ta;
        b[c] = tb;
        dbias[c] = (float)ta;
        dweight[c] = (float)tb;
    }
};

struct StatsF {
    const float* x;
    int ld, coff;
    struct Ctx {};
    __device__ Ctx begin(int, int) const { return Ctx(); }
    __device__ float2 operator()(const Ctx&, long long row, int c) const {
        float v = x[row * ld + coff + c];
        return make_float2(v, v * v);
    }
};

extern "C" int fsv_norm_stats(const float* x, int N, int HW, int C, int ld, int coff, int mode, double* sum, double* sumsq, double* work,
                              void* stream) {
    FSV_REQUIRE(N > 0 && HW > 0 && C > 0 && ld >= coff + C, "norm_stats: bad dims");
    int groups = mode == FSV_NORM_INSTANCE ? N : 1;
    long long rpg = mode == FSV_NORM_INSTANCE ? HW : (long long)N * HW;
    StatsF f{x, ld, coff};
    StoreE e{sum, sumsq, C};
    return launch_reduce2(f, e, groups, rpg, C, work, (cudaStream_t)stream, "norm_stats");
}

// statistics and their finalisation in ONE launch (the training-mode forward of every BatchNorm / InstanceNorm / SPADE)
struct FinalizeE {
    double count, ucount;
    float eps, momentum;
    float *running_mean, *running_var;
    int update_running, C;
    float *mean, *rstd;
    __device__ void operator()(int g, int c, double sum, double sumsq) const {
        const long long i = (long long)g * C + c;
        double m = sum / count;
        double var = sumsq / count - m * m;
        if (var < 0.0) var = 0.0;
        mean[i] = (float)m;
        rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
        if (update_running) {   // batch mode: one group
            double unbiased = ucount > 1.0 ? var * (ucount / (ucount - 1.0)) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
    }
};
extern "C" int fsv_norm_stats_finalize(const float* x, int N, int HW, int C, int ld, int coff, int mode, double unbias_mul, float eps,
                                       float momentum, float* running_mean, float* running_var, int update_running, float* mean,
                                       float* rstd, double* work, void* stream) {
    FSV_REQUIRE(N > 0 && HW > 0 && C > 0 && ld >= coff + C && mean && rstd, "norm_stats_finalize: bad args");
    int groups = mode == FSV_NORM_INSTANCE ? N : 1;
    FSV_REQUIRE(!update_running || (groups == 1 && running_mean && running_var), "norm_stats_finalize: running update needs batch mode");
    long long rpg = mode == FSV_NORM_INSTANCE ? HW : (long long)N * HW;
    StatsF f{x, ld, coff};
    FinalizeE e{(double)rpg, (double)rpg * unbias_mul, eps, momentum, running_mean, running_var, update_running, C, mean, rstd};
    return launch_reduce2(f, e, groups, rpg, C, work, (cudaStream_t)stream, "norm_stats_finalize");
}

__global__ void k_norm_finalize(const double* __restrict__ sum, const double* __restrict__ sumsq, int total, int C, double count,
                                double ucount, float eps, float momentum, float* running_mean, float* running_var, int update_running,
                                float* __restrict__ mean, float* __restrict__ rstd) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    double m = sum[i] / count;
    double var = sumsq[i] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[i] = (float)m;
    rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
    if (update_running && i < C) {   // batch mode: one group
        double unbiased = ucount > 1.0 ? var * (ucount / (ucount - 1.0)) : var;
        running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * (float)m;
        running_var[i] = (1.f - momentum) * running_var[i] + momentum * (float)unbiased;
    }
}
extern "C" int fsv_norm_finalize(const double* sum, const double* sumsq, int groups, int C, double count, double unbias_count,
                                 float eps, float momentum, float* running_mean, float* running_var, int update_running,
                                 float* mean, float* rstd, void* stream) {
    FSV_REQUIRE(groups > 0 && C > 0 && count > 0, "norm_finalize: bad dims");
    FSV_REQUIRE(!update_running || (groups == 1 && running_mean && running_var), "norm_finalize: running update needs batch mode");
    int total = groups * C;
    k_norm_finalize<<<fsv_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(sum, sumsq, total, C, count, unbias_count, eps, momentum,
                                                                            running_mean, running_var, update_running, mean, rstd);
    FSV_CHECK_LAUNCH("norm_finalize");
    return FSV_OK;
}

__global__ void k_from_running(const float* __restrict__ rm, const float* __restrict__ rv, int C, float eps,
                               float* __restrict__ mean, float* __restrict__ rstd) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C) return;
    mean[i] = rm[i];
    rstd[i] = rsqrtf(rv[i] + eps);
}
extern "C" int fsv_norm_from_running(const float* running_mean, const float* running_var, int C, float eps,
                                     float* mean, float* rstd, void* stream) {
    FSV_REQUIRE(C > 0, "norm_from_running: bad dims");
    k_from_running<<<fsv_cdiv(C, 256), 256, 0, (cudaStream_t)stream>>>(running_mean, running_var, C, eps, mean, rstd);
    FSV_CHECK_LAUNCH("norm_from_running");
    return FSV_OK;
}

// ---------------------------------------------------------------- apply forward
__global__ void k_norm_apply_fwd(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                 const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ y,
                                 long long total, int C, long long HWC, int instance, int act) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        int s = instance ? (int)(i / HWC) * C + c : c;
        float v = (x[i] - mean[s]) * rstd[s];
        if (weight) v = v * weight[c] + bias[c];
        y[i] = fsv_act(v, act);
    }
}
// float4 variant (C % 4 == 0): 16-byte loads/stores
__global__ void k_norm_apply_fwd4(const float4* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                  const float* __restrict__ weight, const float* __restrict__ bias, float4* __restrict__ y,
                                  long long total4, int C, long long HWC, int instance, int act) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        long long e = i * 4;
        int c = (int)(e % C);
        int s = instance ? (int)(e / HWC) * C + c : c;
        float4 v = x[i];
        float4 m = *reinterpret_cast<const float4*>(mean + s);
        float4 r = *reinterpret_cast<const float4*>(rstd + s);
        v.x = (v.x - m.x) * r.x; v.y = (v.y - m.y) * r.y; v.z = (v.z - m.z) * r.z; v.w = (v.w - m.w) * r.w;
        if (weight) {
            float4 w = *reinterpret_cast<const float4*>(weight + c);
            float4 b = *reinterpret_cast<const float4*>(bias + c);
            v.x = v.x * w.x + b.x; v.y = v.y * w.y + b.y; v.z = v.z * w.z + b.z; v.w = v.w * w.w + b.w;
        }
        v.x = fsv_act(v.x, act); v.y = fsv_act(v.y, act); v.z = fsv_act(v.z, act); v.w = fsv_act(v.w, act);
        y[i] = v;
    }
}
static inline int ew_grid(long long items) {
    long long b = (items + 255) / 256, cap = (long long)fsv_sm_count() * 16;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}
extern "C" int fsv_norm_apply_fwd(const float* x, const float* mean, const float* rstd, const float* weight, const float* bias,
                                  float* y, int N, int HW, int C, int mode, int act, void* stream) {
    FSV_REQUIRE(N > 0 && HW > 0 && C > 0, "norm_apply_fwd: bad dims");
    FSV_REQUIRE((weight == nullptr) == (bias == nullptr), "norm_apply_fwd: weight and bias go together");
    long long total = (long long)N * HW * C;
    int inst = mode == FSV_NORM_INSTANCE;
    if (C % 4 == 0)
        k_norm_apply_fwd4<<<ew_grid(total / 4), 256, 0, (cudaStream_t)stream>>>((const float4*)x, mean, rstd, weight, bias, (float4*)y,
                                                                                 total / 4, C, (long long)HW * C, inst, act);
    else
        k_norm_apply_fwd<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(x, mean, rstd, weight, bias, y, total, C, (long long)HW * C, inst, act);
    FSV_CHECK_LAUNCH("norm_apply_fwd");
    return FSV_OK;
}

// ---------------------------------------------------------------- apply backward
// derivative of the activation that followed the norm.  With y == NULL (fsv_norm_apply_bwd2) the post-activation tensor is not read:
// for the sign-type activations (LeakyReLU / ReLU) its sign is recomputed from the normalised input with the forward's own formula
// (xhat * weight + bias), which removes one full-tensor read from each of the two backward passes.
__device__ __forceinline__ float norm_act_grad(const float* __restrict__ y, long long i, float xh, float w, float b, int act) {
    if (act == FSV_ACT_NONE) return 1.f;
    if (y) return fsv_act_grad(y[i], act);
    const float pre = xh * w + b;
    return act == FSV_ACT_LRELU ? (pre > 0.f ? 1.f : FSV_LRELU_SLOPE) : (pre > 0.f ? 1.f : 0.f);
}
struct NormBwdF {   // (sum dy', sum dy'*xhat), dy' = dy*act'(y)
    const float *x, *y, *dy, *mean, *rstd, *weight, *bias;
    int C, act, instance;
    long long HW;
    struct Ctx { float mean, rstd, w, b; };
    __device__ Ctx begin(int g, int c) const {          // instance mode: group g IS the sample
        const int s = instance ? g * C + c : c;
        Ctx k;
        k.mean = mean[s]; k.rstd = rstd[s]; k.w = weight ? weight[c] : 1.f; k.b = bias ? bias[c] : 0.f;
        return k;
    }
    __device__ float2 operator()(const Ctx& k, long long row, int c) const {
        long long i = row * C + c;
        float xh = (x[i] - k.mean) * k.rstd;
        float g = dy[i] * norm_act_grad(y, i, xh, k.w, k.b, act);
        return make_float2(g, g * xh);
    }
};
__global__ void k_norm_apply_bwd(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                                 const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ weight,
                                 const float* __restrict__ bias, const double* __restrict__ A, const double* __restrict__ B, float* __restrict__ dx,
                                 long long total, int C, long long HW, int instance, int act, int batch_stats, float inv_cnt) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        int s = instance ? (int)(i / (HW * C)) * C + c : c;
        float r = rstd[s];
        float w = weight ? weight[c] : 1.f;
        float xh = (x[i] - mean[s]) * r;
        float g = dy[i] * norm_act_grad(y, i, xh, w, bias ? bias[c] : 0.f, act);
        float v = g;
        if (batch_stats) v = g - (float)A[s] * inv_cnt - xh * (float)B[s] * inv_cnt;
        dx[i] = r * w * v;
    }
}
// float4 variant (C % 4 == 0): the scalar kernel above spends its time on 64-bit index arithmetic, not on memory
__global__ void k_norm_apply_bwd4(const float4* __restrict__ x, const float4* __restrict__ y, const float4* __restrict__ dy,
                                  const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ weight,
                                  const float* __restrict__ bias, const double* __restrict__ A, const double* __restrict__ B, float4* __restrict__ dx,
                                  long long total4, int C, long long HWC, int instance, int act, int batch_stats, float inv_cnt) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * 4;
        const int c = (int)(e % C);
        const int s = instance ? (int)(e / HWC) * C + c : c;
        const float4 r = *reinterpret_cast<const float4*>(rstd + s);
        const float4 d4 = dy[i];
        float w[4] = {1.f, 1.f, 1.f, 1.f}, bb[4] = {0.f, 0.f, 0.f, 0.f};
        if (weight) {
            const float4 w4 = *reinterpret_cast<const float4*>(weight + c);
            w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
        }
        if (bias) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias + c);
            bb[0] = b4.x; bb[1] = b4.y; bb[2] = b4.z; bb[3] = b4.w;
        }
        const float rr[4] = {r.x, r.y, r.z, r.w};
        const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
        const bool need_x = batch_stats || (y == nullptr && act != FSV_ACT_NONE);
        float xh[4] = {0.f, 0.f, 0.f, 0.f};
        if (need_x) {
            const float4 m = *reinterpret_cast<const float4*>(mean + s);
            const float4 x4 = x[i];
            xh[0] = (x4.x - m.x) * rr[0]; xh[1] = (x4.y - m.y) * rr[1]; xh[2] = (x4.z - m.z) * rr[2]; xh[3] = (x4.w - m.w) * rr[3];
        }
        float yv[4] = {0.f, 0.f, 0.f, 0.f};
        if (y != nullptr && act != FSV_ACT_NONE) {
            const float4 y4 = y[i];
            yv[0] = y4.x; yv[1] = y4.y; yv[2] = y4.z; yv[3] = y4.w;
        }
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float ag = 1.f;
            if (act != FSV_ACT_NONE) {
                if (y != nullptr) ag = fsv_act_grad(yv[j], act);
                else {
                    const float pre = xh[j] * w[j] + bb[j];
                    ag = act == FSV_ACT_LRELU ? (pre > 0.f ? 1.f : FSV_LRELU_SLOPE) : (pre > 0.f ? 1.f : 0.f);
                }
            }
            const float g = dv[j] * ag;
            o[j] = batch_stats ? rr[j] * w[j] * (g - (float)A[s + j] * inv_cnt - xh[j] * (float)B[s + j] * inv_cnt) : rr[j] * w[j] * g;
        }
        dx[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}
__global__ void k_norm_param_grads(const double* __restrict__ A, const double* __restrict__ B, int groups, int C,
                                   float* __restrict__ dweight, float* __restrict__ dbias) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double a = 0.0, b = 0.0;
    for (int g = 0; g < groups; ++g) {
        a += A[(long long)g * C + c];
        b += B[(long long)g * C + c];
    }
    dbias[c] = (float)a;
    dweight[c] = (float)b;
}
static int norm_apply_bwd_impl(const float* x, const float* y, const float* dy, const float* mean, const float* rstd,
                               const float* weight, const float* bias, float* dx, float* dweight, float* dbias, double* scratch,
                               int N, int HW, int C, int mode, int act, int batch_stats, void* stream) {
    FSV_REQUIRE(N > 0 && HW > 0 && C > 0 && scratch, "norm_apply_bwd: bad args");
    FSV_REQUIRE(y != nullptr || act == FSV_ACT_NONE || act == FSV_ACT_LRELU || act == FSV_ACT_RELU,
                "norm_apply_bwd: without y only the sign-type activations can be differentiated");
    cudaStream_t st = (cudaStream_t)stream;
    int inst = mode == FSV_NORM_INSTANCE;
    int groups = inst ? N : 1;
    long long rpg = inst ? HW : (long long)N * HW;
    double* A = scratch;
    double* B = scratch + (size_t)groups * C;
    NormBwdF f{x, y, dy, mean, rstd, weight, bias, C, act, inst, (long long)HW};
    if (weight) FSV_REQUIRE(dweight && dbias, "norm_apply_bwd: affine needs dweight/dbias");
    int rc;
    if (weight && groups == 1) {          // batch mode: the reduction's last block writes the parameter gradients itself
        StoreParamE e{A, B, dweight, dbias};
        rc = launch_reduce2(f, e, groups, rpg, C, B + (size_t)groups * C, st, "norm_bwd_reduce");
    } else {
        StoreE e{A, B, C};
        rc = launch_reduce2(f, e, groups, rpg, C, B + (size_t)groups * C, st, "norm_bwd_reduce");
    }
    if (rc) return rc;
    if (weight && groups != 1) {
        k_norm_param_grads<<<fsv_cdiv(C, 256), 256, 0, st>>>(A, B, groups, C, dweight, dbias);
        FSV_CHECK_LAUNCH("norm_param_grads");
    }
    long long total = (long long)N * HW * C;
    if (C % 4 == 0 && ((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)dy) | ((uintptr_t)dx)) & 15) == 0)
        k_norm_apply_bwd4<<<ew_grid(total / 4), 256, 0, st>>>((const float4*)x, (const float4*)y, (const float4*)dy, mean, rstd, weight, bias, A, B,
                                                              (float4*)dx, total / 4, C, (long long)HW * C, inst, act, batch_stats,
                                                              (float)(1.0 / (double)rpg));
    else
        k_norm_apply_bwd<<<ew_grid(total), 256, 0, st>>>(x, y, dy, mean, rstd, weight, bias, A, B, dx, total, C, HW, inst, act, batch_stats,
                                                          (float)(1.0 / (double)rpg));
    FSV_CHECK_LAUNCH("norm_apply_bwd");
    return FSV_OK;
}

extern "C" int fsv_norm_apply_bwd(const float* x, const float* y, const float* dy, const float* mean, const float* rstd,
                                  const float* weight, float* dx, float* dweight, float* dbias, double* scratch,
                                  int N, int HW, int C, int mode, int act, int batch_stats, void* stream) {
    FSV_REQUIRE(y != nullptr, "norm_apply_bwd: y is required (fsv_norm_apply_bwd2 differentiates the activation without it)");
    return norm_apply_bwd_impl(x, y, dy, mean, rstd, weight, nullptr, dx, dweight, dbias, scratch, N, HW, C, mode, act, batch_stats, stream);
}
extern "C" int fsv_norm_apply_bwd2(const float* x, const float* dy, const float* mean, const float* rstd, const float* weight,
                                   const float* bias, float* dx, float* dweight, float* dbias, double* scratch,
                                   int N, int HW, int C, int mode, int act, int batch_stats, void* stream) {
    return norm_apply_bwd_impl(x, nullptr, dy, mean, rstd, weight, bias, dx, dweight, dbias, scratch, N, HW, C, mode, act, batch_stats, stream);
}

// ---------------------------------------------------------------- SPADE norm backward (x may be read through a x2 upsample)
struct SpadeNormBwdF {   // rows are full-resolution pixels
    const float *x, *g, *mean, *rstd;
    int C, H, W, up, instance;
    struct Ctx { float mean, rstd; };
    __device__ Ctx begin(int grp, int c) const {
        const int s = instance ? grp * C + c : c;
        Ctx k;
        k.mean = mean[s]; k.rstd = rstd[s];
        return k;
    }
    __device__ float2 operator()(const Ctx& k, long long row, int c) const {
        float xv;
        if (up == 1) xv = x[row * C + c];            // same layout as g: no index arithmetic
        else {
            int w = (int)(row % W);
            long long q = row / W;
            int h = (int)(q % H);
            long long n = q / H;
            int Hs = H / up, Ws = W / up;
            xv = x[((n * Hs + h / up) * Ws + w / up) * C + c];
        }
        float xh = (xv - k.mean) * k.rstd;
        float gv = g[row * C + c];
        return make_float2(gv, gv * xh);
    }
};
// x2-upsampled input: the same two sums taken over the SOURCE pixels -- sum_hi g = sum_lo G, sum_hi g*xhat = sum_lo G*xhat with G = the
// 2x2 block sum of g (xhat is constant over the block).  One pixel decode (32-bit) and one x load per four g loads; the per-element
// form above spent most of its time in 64-bit divisions (round-2 timeline: 0.42 - 0.46 ms per 512x512x64 call, 11x the HBM time).
struct SpadeNormBwdUp2F {   // rows are source-resolution pixels
    const float *x, *g, *mean, *rstd;
    int C, Hs, Ws, instance;
    struct Ctx { float mean, rstd; };
    __device__ Ctx begin(int grp, int c) const {
        const int s = instance ? grp * C + c : c;
        Ctx k;
        k.mean = mean[s]; k.rstd = rstd[s];
        return k;
    }
    __device__ float2 operator()(const Ctx& k, long long row, int c) const {
        const unsigned r = (unsigned)row;
        const unsigned ws = r % (unsigned)Ws, q = r / (unsigned)Ws;
        const unsigned hs = q % (unsigned)Hs, n = q / (unsigned)Hs;
        const long long wc = (long long)Ws * 2 * C;                                   // one upsampled image row
        const float* gp = g + ((long long)(n * (unsigned)Hs + hs) * 2) * wc + (long long)ws * 2 * C + c;
        const float gs = (gp[0] + gp[C]) + (gp[wc] + gp[wc + C]);
        const float xh = (x[row * C + c] - k.mean) * k.rstd;
        return make_float2(gs, gs * xh);
    }
};
__global__ void k_spade_norm_bwd(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ mean,
                                 const float* __restrict__ rstd, const double* __restrict__ A, const double* __restrict__ B,
                                 float* __restrict__ dx, int N, int Hs, int Ws, int C, int up, int instance, int batch_stats, float inv_cnt) {
    long long total = (long long)N * Hs * Ws * C;
    int W = Ws * up;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long p = i / C;
        int ws = (int)(p % Ws);
        long long q = p / Ws;
        int hs = (int)(q % Hs);
        long long n = q / Hs;
        int s = instance ? (int)n * C + c : c;
        float gs = 0.f;
        for (int a = 0; a < up; ++a)
            for (int b = 0; b < up; ++b)
                gs += g[(((n * Hs * up + hs * up + a) * W) + ws * up + b) * C + c];
        float r = rstd[s];
        float v = gs;
        if (batch_stats) {
            float k = (float)(up * up);
            float xh = (x[i] - mean[s]) * r;
            v = gs - k * (float)A[s] * inv_cnt - k * xh * (float)B[s] * inv_cnt;
        }
        dx[i] = r * v;
    }
}
// float4 variant (C % 4 == 0)
__global__ void k_spade_norm_bwd4(const float4* __restrict__ x, const float4* __restrict__ g, const float* __restrict__ mean,
                                  const float* __restrict__ rstd, const double* __restrict__ A, const double* __restrict__ B,
                                  float4* __restrict__ dx, int N, int Hs, int Ws, int C, int up, int instance, int batch_stats, float inv_cnt) {
    const int C4 = C >> 2;
    const long long total4 = (long long)N * Hs * Ws * C4;
    const int W = Ws * up;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        const long long p = i / C4;
        const int ws = (int)(p % Ws);
        const long long q = p / Ws;
        const int hs = (int)(q % Hs);
        const long long n = q / Hs;
        const int c = c4 * 4;
        const int s = instance ? (int)n * C + c : c;
        float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int a = 0; a < up; ++a)
            for (int b = 0; b < up; ++b) {
                const float4 t = g[(((n * Hs * up + hs * up + a) * W) + ws * up + b) * C4 + c4];
                gs.x += t.x; gs.y += t.y; gs.z += t.z; gs.w += t.w;
            }
        const float4 r = *reinterpret_cast<const float4*>(rstd + s);
        float4 v = gs;
        if (batch_stats) {
            const float k = (float)(up * up) * inv_cnt;
            const float4 m = *reinterpret_cast<const float4*>(mean + s);
            const float4 xv = x[i];
            v.x = gs.x - k * (float)A[s + 0] - k * (xv.x - m.x) * r.x * (float)B[s + 0];
            v.y = gs.y - k * (float)A[s + 1] - k * (xv.y - m.y) * r.y * (float)B[s + 1];
            v.z = gs.z - k * (float)A[s + 2] - k * (xv.z - m.z) * r.z * (float)B[s + 2];
            v.w = gs.w - k * (float)A[s + 3] - k * (xv.w - m.w) * r.w * (float)B[s + 3];
        }
        dx[i] = make_float4(r.x * v.x, r.y * v.y, r.z * v.z, r.w * v.w);
    }
}
extern "C" int fsv_spade_norm_bwd(const float* x, const float* dxhat, const float* mean, const float* rstd, float* dx,
                                  double* scratch, int N, int H, int W, int C, int up, int mode, int batch_stats, void* stream) {
    FSV_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && (up == 1 || up == 2) && H % up == 0 && W % up == 0, "spade_norm_bwd: bad dims");
    cudaStream_t st = (cudaStream_t)stream;
    int inst = mode == FSV_NORM_INSTANCE;
    int groups = inst ? N : 1;
    long long rpg = inst ? (long long)H * W : (long long)N * H * W;
    double* A = scratch;
    double* B = scratch + (size_t)groups * C;
    if (batch_stats) {
        StoreE e{A, B, C};
        int rc;
        static int src_form = -1;       // FSV_SPADE_NORM_SRC=0: the per-upsampled-pixel form for up = 2 as well
        if (src_form < 0) { const char* ev = getenv("FSV_SPADE_NORM_SRC"); src_form = (ev && atoi(ev) == 0) ? 0 : 1; }
        if (up == 2 && src_form && (long long)N * H * W < (1LL << 31)) {
            SpadeNormBwdUp2F f{x, dxhat, mean, rstd, C, H / 2, W / 2, inst};
            rc = launch_reduce2(f, e, groups, rpg / 4, C, B + (size_t)groups * C, st, "spade_norm_bwd_reduce");
        } else {
            SpadeNormBwdF f{x, dxhat, mean, rstd, C, H, W, up, inst};
            rc = launch_reduce2(f, e, groups, rpg, C, B + (size_t)groups * C, st, "spade_norm_bwd_reduce");
        }
        if (rc) return rc;
    }
    long long total = (long long)N * (H / up) * (W / up) * C;
    if (C % 4 == 0 && ((((uintptr_t)x) | ((uintptr_t)dxhat) | ((uintptr_t)dx)) & 15) == 0)
        k_spade_norm_bwd4<<<ew_grid(total / 4), 256, 0, st>>>((const float4*)x, (const float4*)dxhat, mean, rstd, A, B, (float4*)dx, N, H / up,
                                                              W / up, C, up, inst, batch_stats, (float)(1.0 / (double)rpg));
    else
        k_spade_norm_bwd<<<ew_grid(total), 256, 0, st>>>(x, dxhat, mean, rstd, A, B, dx, N, H / up, W / up, C, up, inst, batch_stats,
                                                          (float)(1.0 / (double)rpg));
    FSV_CHECK_LAUNCH("spade_norm_bwd");
    return FSV_OK;
}
