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

#define RED_TX 32
#define RED_TY 8
#define RED_ROWS_PER_BLOCK 256

// Generic two-value per-(group, channel) reduction over the rows of an NHWC slice.
// f(row, c) -> float2; results are atomically added (fp64) into a[g*C+c], b[g*C+c].
template <class F>
__global__ void k_chan_reduce2(F f, long long rows_per_group, int C, double* __restrict__ a, double* __restrict__ b) {
    int c = blockIdx.y * RED_TX + threadIdx.x;
    int g = blockIdx.z;
    long long r0 = (long long)blockIdx.x * RED_ROWS_PER_BLOCK;
    long long r1 = r0 + RED_ROWS_PER_BLOCK;
    if (r1 > rows_per_group) r1 = rows_per_group;
    float sa = 0.f, sb = 0.f;
    if (c < C) {
        for (long long r = r0 + threadIdx.y; r < r1; r += RED_TY) {
            float2 v = f(g * rows_per_group + r, c);
            sa += v.x;
            sb += v.y;
        }
    }
    __shared__ double s_a[RED_TY][RED_TX], s_b[RED_TY][RED_TX];
    s_a[threadIdx.y][threadIdx.x] = (double)sa;
    s_b[threadIdx.y][threadIdx.x] = (double)sb;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double ta = 0.0, tb = 0.0;
#pragma unroll
        for (int i = 0; i < RED_TY; ++i) {
            ta += s_a[i][threadIdx.x];
            tb += s_b[i][threadIdx.x];
        }
        atomicAdd(&a[(long long)g * C + c], ta);
        atomicAdd(&b[(long long)g * C + c], tb);
    }
}

template <class F>
static int launch_reduce2(F f, int groups, long long rows_per_group, int C, double* a, double* b, cudaStream_t st, const char* name) {
    FSV_CUDA(cudaMemsetAsync(a, 0, sizeof(double) * (size_t)groups * C, st));
    FSV_CUDA(cudaMemsetAsync(b, 0, sizeof(double) * (size_t)groups * C, st));
    dim3 grid(fsv_cdiv(rows_per_group, RED_ROWS_PER_BLOCK), fsv_cdiv(C, RED_TX), groups);
    dim3 block(RED_TX, RED_TY);
    k_chan_reduce2<<<grid, block, 0, st>>>(f, rows_per_group, C, a, b);
    FSV_CHECK_LAUNCH(name);
    return FSV_OK;
}

struct StatsF {
    const float* x;
    int ld, coff;
    __device__ float2 operator()(long long row, int c) const {
        float v = x[row * ld + coff + c];
        return make_float2(v, v * v);
    }
};

extern "C" int fsv_norm_stats(const float* x, int N, int HW, int C, int ld, int coff, int mode, double* sum, double* sumsq, void* stream) {
    FSV_REQUIRE(N > 0 && HW > 0 && C > 0 && ld >= coff + C, "norm_stats: bad dims");
    int groups = mode == FSV_NORM_INSTANCE ? N : 1;
    long long rpg = mode == FSV_NORM_INSTANCE ? HW : (long long)N * HW;
    StatsF f{x, ld, coff};
    return launch_reduce2(f, groups, rpg, C, sum, sumsq, (cudaStream_t)stream, "norm_stats");
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
struct NormBwdF {   // (sum dy', sum dy'*xhat), dy' = dy*act'(y)
    const float *x, *y, *dy, *mean, *rstd;
    int C, act, instance;
    long long HW;
    __device__ float2 operator()(long long row, int c) const {
        long long i = row * C + c;
        int s = instance ? (int)(row / HW) * C + c : c;
        float g = dy[i] * fsv_act_grad(y[i], act);
        float xh = (x[i] - mean[s]) * rstd[s];
        return make_float2(g, g * xh);
    }
};
__global__ void k_norm_apply_bwd(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                                 const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ weight,
                                 const double* __restrict__ A, const double* __restrict__ B, float* __restrict__ dx,
                                 long long total, int C, long long HW, int instance, int act, int batch_stats, float inv_cnt) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        int s = instance ? (int)(i / (HW * C)) * C + c : c;
        float r = rstd[s];
        float g = dy[i] * fsv_act_grad(y[i], act);
        float w = weight ? weight[c] : 1.f;
        float v = g;
        if (batch_stats) {
            float xh = (x[i] - mean[s]) * r;
            v = g - (float)A[s] * inv_cnt - xh * (float)B[s] * inv_cnt;
        }
        dx[i] = r * w * v;
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
extern "C" int fsv_norm_apply_bwd(const float* x, const float* y, const float* dy, const float* mean, const float* rstd,
                                  const float* weight, float* dx, float* dweight, float* dbias, double* scratch,
                                  int N, int HW, int C, int mode, int act, int batch_stats, void* stream) {
    FSV_REQUIRE(N > 0 && HW > 0 && C > 0 && scratch, "norm_apply_bwd: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    int inst = mode == FSV_NORM_INSTANCE;
    int groups = inst ? N : 1;
    long long rpg = inst ? HW : (long long)N * HW;
    double* A = scratch;
    double* B = scratch + (size_t)groups * C;
    NormBwdF f{x, y, dy, mean, rstd, C, act, inst, (long long)HW};
    int rc = launch_reduce2(f, groups, rpg, C, A, B, st, "norm_bwd_reduce");
    if (rc) return rc;
    if (weight) {
        FSV_REQUIRE(dweight && dbias, "norm_apply_bwd: affine needs dweight/dbias");
        k_norm_param_grads<<<fsv_cdiv(C, 256), 256, 0, st>>>(A, B, groups, C, dweight, dbias);
        FSV_CHECK_LAUNCH("norm_param_grads");
    }
    long long total = (long long)N * HW * C;
    k_norm_apply_bwd<<<ew_grid(total), 256, 0, st>>>(x, y, dy, mean, rstd, weight, A, B, dx, total, C, HW, inst, act, batch_stats,
                                                      (float)(1.0 / (double)rpg));
    FSV_CHECK_LAUNCH("norm_apply_bwd");
    return FSV_OK;
}

// ---------------------------------------------------------------- SPADE norm backward (x may be read through a x2 upsample)
struct SpadeNormBwdF {   // rows are full-resolution pixels
    const float *x, *g, *mean, *rstd;
    int C, H, W, up, instance;
    __device__ float2 operator()(long long row, int c) const {
        int w = (int)(row % W);
        long long q = row / W;
        int h = (int)(q % H);
        long long n = q / H;
        int Hs = H / up, Ws = W / up;
        int s = instance ? (int)n * C + c : c;
        float xv = x[((n * Hs + h / up) * Ws + w / up) * C + c];
        float xh = (xv - mean[s]) * rstd[s];
        float gv = g[row * C + c];
        return make_float2(gv, gv * xh);
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
        SpadeNormBwdF f{x, dxhat, mean, rstd, C, H, W, up, inst};
        int rc = launch_reduce2(f, groups, rpg, C, A, B, st, "spade_norm_bwd_reduce");
        if (rc) return rc;
    }
    long long total = (long long)N * (H / up) * (W / up) * C;
    k_spade_norm_bwd<<<ew_grid(total), 256, 0, st>>>(x, dxhat, mean, rstd, A, B, dx, N, H / up, W / up, C, up, inst, batch_stats,
                                                      (float)(1.0 / (double)rpg));
    FSV_CHECK_LAUNCH("spade_norm_bwd");
    return FSV_OK;
}
