// Label preprocessing and face-region kernels of the pose pipeline.
//
// Reference: models/input_process.py:52-93 (get_fg_mask: MaxPool2d(15,1,7) + threshold; get_part_mask / get_face_mask:
// DensePose part-id tests, built with a python loop over 25 ids into a ByteTensor), loss_collector.py:178-179
// (AvgPool2d(15,1,7) of the face mask), models/face_refiner.py:24-83 (get_face_region: nonzero() + four .item() host
// round trips PER SAMPLE, then a python loop of slice + F.interpolate + torch.cat).  Here the box is computed on the
// device (one block per sample, min/max reduction) and stays there; the crop kernels read it from device memory, so the
// whole training step has no host synchronisation and can be captured into one CUDA graph.
//
// All kernels are HBM/latency-bound streaming kernels over label planes (a few MB); 15x15 windows are evaluated
// separably through a shared-memory tile (row pass, then column pass).
#include "common.cuh"

#define PP_TW 32
#define PP_TH 8
#define PP_R 7   // window radius (15x15)

// ------------------------------------------------------------------------------------------------ 15x15 window kernels
// mode 0: out = any(src > thr) over the window clipped to the image   (== (MaxPool2d(15,1,7)(src) > thr).float(): the pool
//          pads with -inf, input_process.py:58-60)
// mode 1: out = sum(test(src)) / 225 with zero padding               (== AvgPool2d(15,1,7)(mask), count_include_pad=True)
//          where test(v) = 1 if (v/2+0.5)*24 is within 0.1 of 23 or 24  (get_face_mask, input_process.py:81-93)
__device__ __forceinline__ float pp_face_test(float v) {
    float part = (v / 2.f + 0.5f) * 24.f;
    return ((part > (float)(23.0 - 0.1) && part < (float)(23.0 + 0.1)) || (part > (float)(24.0 - 0.1) && part < (float)(24.0 + 0.1))) ? 1.f : 0.f;
}

__global__ void k_window15(const float* __restrict__ src, long long s_n, float* __restrict__ dst, int H, int W, int mode, float thr) {
    __shared__ float tile[PP_TH + 2 * PP_R][PP_TW + 2 * PP_R + 1];
    __shared__ float rowp[PP_TH + 2 * PP_R][PP_TW];
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * PP_TW, y0 = blockIdx.y * PP_TH;
    const float* s = src + (long long)n * s_n;
    const int tid = threadIdx.y * PP_TW + threadIdx.x;
    for (int i = tid; i < (PP_TH + 2 * PP_R) * (PP_TW + 2 * PP_R); i += PP_TW * PP_TH) {
        int ty = i / (PP_TW + 2 * PP_R), tx = i % (PP_TW + 2 * PP_R);
        int y = y0 + ty - PP_R, x = x0 + tx - PP_R;
        float v = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            float r = s[(long long)y * W + x];
            v = mode == 0 ? (r > thr ? 1.f : 0.f) : pp_face_test(r);
        }
        tile[ty][tx] = v;
    }
    __syncthreads();
    for (int i = tid; i < (PP_TH + 2 * PP_R) * PP_TW; i += PP_TW * PP_TH) {
        int ty = i / PP_TW, tx = i % PP_TW;
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * PP_R + 1; ++k) a += tile[ty][tx + k];
        rowp[ty][tx] = a;
    }
    __syncthreads();
    int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x < W && y < H) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * PP_R + 1; ++k) a += rowp[threadIdx.y + k][threadIdx.x];
        dst[((long long)n * H + y) * W + x] = mode == 0 ? (a > 0.f ? 1.f : 0.f) : a / 225.f;   // integer-valued sums: exact
    }
}

extern "C" int fsv_fg_mask(const float* label, long long n_stride, float* out, int N, int H, int W, float thr, void* stream) {
    FSV_REQUIRE(label && out && N > 0 && H > 0 && W > 0, "fg_mask: bad args");
    dim3 grid(fsv_cdiv(W, PP_TW), fsv_cdiv(H, PP_TH), N), block(PP_TW, PP_TH);
    k_window15<<<grid, block, 0, (cudaStream_t)stream>>>(label, n_stride, out, H, W, 0, thr);
    FSV_CHECK_LAUNCH("fg_mask");
    return FSV_OK;
}

extern "C" int fsv_face_mask_avg15(const float* part, long long n_stride, float* out, int N, int H, int W, void* stream) {
    FSV_REQUIRE(part && out && N > 0 && H > 0 && W > 0, "face_mask_avg15: bad args");
    dim3 grid(fsv_cdiv(W, PP_TW), fsv_cdiv(H, PP_TH), N), block(PP_TW, PP_TH);
    k_window15<<<grid, block, 0, (cudaStream_t)stream>>>(part, n_stride, out, H, W, 1, 0.f);
    FSV_CHECK_LAUNCH("face_mask_avg15");
    return FSV_OK;
}

// ------------------------------------------------------------------------------------------------ DensePose part masks
// get_part_mask (input_process.py:63-79): 9 body-part groups of the 25 DensePose ids; out is NHWC (N, H, W, 9) in {0,1}
__global__ void k_part_masks(const float* __restrict__ part, long long s_n, float* __restrict__ out, int N, long long HW) {
    // group of id j (0..24): [[0],[1,2],[3,4],[5,6],[7,9,8,10],[11,13,12,14],[15,17,16,18],[19,21,20,22],[23,24]]
    const int grp[25] = {0, 1, 1, 2, 2, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8};
    long long total = (long long)N * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long n = i / HW, p = i - n * HW;
        float v = (part[n * s_n + p] / 2.f + 0.5f) * 24.f;
        float m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        int j = (int)floorf(v + 0.5f);              // the only id whose +-0.1 window can contain v
        if (j >= 0 && j <= 24 && v > (float)((double)j - 0.1) && v < (float)((double)j + 0.1)) m[grp[j]] = 1.f;   // python j-0.1 is a double, cast to fp32 for the compare
        float* o = out + i * 9;
#pragma unroll
        for (int g = 0; g < 9; ++g) o[g] = m[g];
    }
}

extern "C" int fsv_part_masks(const float* part, long long n_stride, float* out, int N, int H, int W, void* stream) {
    FSV_REQUIRE(part && out && N > 0 && H > 0 && W > 0, "part_masks: bad args");
    long long total = (long long)N * H * W;
    long long blocks = (total + 255) / 256, cap = (long long)fsv_sm_count() * 16;
    k_part_masks<<<(int)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream>>>(part, n_stride, out, N, (long long)H * W);
    FSV_CHECK_LAUNCH("part_masks");
    return FSV_OK;
}

// ------------------------------------------------------------------------------------------------ face box on the device
// face_refiner.py:52-83 get_face_region.  The "face" pixels are those where up to three planes all exceed a threshold
// (OpenPose: the three keypoint-image channels > 0; DensePose: the part channel > 0.9); one block per sample reduces
// min/max row/column, thread 0 finishes the integer arithmetic of the reference and writes box[n] = {ys, ye, xs, xe}.
#define BB_SPLITS 32
__device__ unsigned int g_bb_ticket[1024];
__device__ int g_bb_part[1024][BB_SPLITS][4];

__global__ void k_face_bbox(const float* __restrict__ p0, const float* __restrict__ p1, const float* __restrict__ p2,
                            long long s0, long long s1, long long s2, float thr, int H, int W, int openpose, int crop_smaller,
                            int* __restrict__ box) {
    // grid (sample, split): every block reduces a slice of the plane; the last block of a sample (ticket) combines the partial
    // boxes and finishes the integer arithmetic of the reference.  2 blocks over a 512x512 plane took 150 us; this takes ~10.
    const int n = blockIdx.x, sp = blockIdx.y;
    const float* a = p0 + (long long)n * s0;
    const float* b = p1 ? p1 + (long long)n * s1 : nullptr;
    const float* c = p2 ? p2 + (long long)n * s2 : nullptr;
    int ymin = 1 << 30, ymax = -1, xmin = 1 << 30, xmax = -1;
    const int total = H * W;
    const int per = (total + BB_SPLITS - 1) / BB_SPLITS;
    const int i1 = min(total, (sp + 1) * per);
    for (int i = sp * per + threadIdx.x; i < i1; i += blockDim.x) {
        bool f = a[i] > thr && (!b || b[i] > thr) && (!c || c[i] > thr);
        if (f) {
            int y = i / W, x = i - y * W;
            ymin = min(ymin, y); ymax = max(ymax, y); xmin = min(xmin, x); xmax = max(xmax, x);
        }
    }
    __shared__ int sm[4][32];
    __shared__ int s_last;
    for (int o = 16; o > 0; o >>= 1) {
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    }
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane == 0) { sm[0][warp] = ymin; sm[1][warp] = ymax; sm[2][warp] = xmin; sm[3][warp] = xmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < nw; ++k) {
            ymin = min(ymin, sm[0][k]); ymax = max(ymax, sm[1][k]); xmin = min(xmin, sm[2][k]); xmax = max(xmax, sm[3][k]);
        }
        volatile int* part = g_bb_part[n][sp];
        part[0] = ymin; part[1] = ymax; part[2] = xmin; part[3] = xmax;
        __threadfence();
        unsigned int t = atomicAdd(&g_bb_ticket[n], 1u);
        s_last = (t == BB_SPLITS - 1);
        if (s_last) g_bb_ticket[n] = 0u;
    }
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    ymin = 1 << 30; ymax = -1; xmin = 1 << 30; xmax = -1;
    for (int k = 0; k < BB_SPLITS; ++k) {
        volatile int* part = g_bb_part[n][k];
        ymin = min(ymin, part[0]); ymax = max(ymax, part[1]); xmin = min(xmin, part[2]); xmax = max(xmax, part[3]);
    }
    {
        int yc, xc, len;
        if (ymax >= 0) {
            int ys = ymin, ye = ymax, xs = xmin, xe = xmax;
            if (openpose) {
                xc = (xs + xe) / 2; yc = (ys * 3 + ye * 2) / 5;
                len = (int)((double)(xe - xs) * 2.5);
            } else {
                xc = (xs + xe) / 2; yc = (ys + ye) / 2;
                len = (int)((double)(ye - ys) * 1.25);
            }
            len = min(W, max(32, len));
            yc = max(len / 2, min(H - 1 - len / 2, yc));
            xc = max(len / 2, min(W - 1 - len / 2, xc));
        } else {
            yc = H / 4; xc = W / 2; len = H / 32 * 8;
        }
        box[n * 4 + 0] = yc - len / 2 + crop_smaller;
        box[n * 4 + 1] = yc + len / 2 - crop_smaller;
        box[n * 4 + 2] = xc - len / 2 + crop_smaller;
        box[n * 4 + 3] = xc + len / 2 - crop_smaller;
    }
}

extern "C" int fsv_face_bbox(const float* p0, const float* p1, const float* p2, long long s0, long long s1, long long s2, float thr,
                             int N, int H, int W, int openpose, int crop_smaller, int* box, void* stream) {
    FSV_REQUIRE(p0 && box && N > 0 && N <= 1024 && H > 0 && W > 0, "face_bbox: bad args (N <= 1024)");
    k_face_bbox<<<dim3(N, BB_SPLITS), 256, 0, (cudaStream_t)stream>>>(p0, p1, p2, s0, s1, s2, thr, H, W, openpose, crop_smaller, box);
    FSV_CHECK_LAUNCH("face_bbox");
    return FSV_OK;
}

// ------------------------------------------------------------------------------------------------ crop + nearest resize
// face_refiner.py:34-38: output_i = F.interpolate(image[i, -3:, ys:ye, xs:xe], size=(S, S))  (mode 'nearest').
// ATen's nearest index: src = min(floor(dst * scale), in - 1) with scale = (float)in / out evaluated in fp32.
__device__ __forceinline__ int pp_nearest(int d, int in, int out) {
    float scale = (float)in / (float)out;
    return min((int)floorf((float)d * scale), in - 1);
}

// src addressed by generic element strides (sn, sc, sh, sw) so NCHW tensors and NCHW-shaped views of NHWC buffers both
// work; dst is NHWC (N, S, S, dst_ld) at channel offset dst_coff and batch offset handled by the caller's pointer.
__global__ void k_crop_resize_fwd(const float* __restrict__ src, long long sn, long long sc, long long sh, long long sw,
                                  const int* __restrict__ box, float* __restrict__ dst, int N, int C, int S, int dst_ld, int dst_coff) {
    long long total = (long long)N * S * S;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ox = (int)(i % S);
        long long q = i / S;
        int oy = (int)(q % S);
        int n = (int)(q / S);
        int ys = box[n * 4], ye = box[n * 4 + 1], xs = box[n * 4 + 2], xe = box[n * 4 + 3];
        int y = ys + pp_nearest(oy, ye - ys, S), x = xs + pp_nearest(ox, xe - xs, S);
        const float* s = src + n * sn + y * sh + x * sw;
        float* d = dst + i * dst_ld + dst_coff;
        for (int c = 0; c < C; ++c) d[c] = s[c * sc];
    }
}

// adjoint: one thread per source pixel of the image; pixels outside the box get 0, pixels inside gather the (few) output
// pixels that map onto them -- deterministic, no atomics, dsrc fully written (NHWC (N, H, W, C) contiguous).
__global__ void k_crop_resize_bwd(const float* __restrict__ ddst, int dst_ld, int dst_coff, const int* __restrict__ box,
                                  float* __restrict__ dsrc, int N, int C, int H, int W, int S) {
    long long total = (long long)N * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int x = (int)(i % W);
        long long q = i / W;
        int y = (int)(q % H);
        int n = (int)(q / H);
        int ys = box[n * 4], ye = box[n * 4 + 1], xs = box[n * 4 + 2], xe = box[n * 4 + 3];
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (y >= ys && y < ye && x >= xs && x < xe) {
            int inh = ye - ys, inw = xe - xs, ry = y - ys, rx = x - xs;
            int oy0 = max(0, (int)floorf((float)ry * (float)S / (float)inh) - 1), oy1 = min(S - 1, (int)ceilf((float)(ry + 1) * (float)S / (float)inh) + 1);
            int ox0 = max(0, (int)floorf((float)rx * (float)S / (float)inw) - 1), ox1 = min(S - 1, (int)ceilf((float)(rx + 1) * (float)S / (float)inw) + 1);
            for (int oy = oy0; oy <= oy1; ++oy) {
                if (pp_nearest(oy, inh, S) != ry) continue;
                for (int ox = ox0; ox <= ox1; ++ox) {
                    if (pp_nearest(ox, inw, S) != rx) continue;
                    const float* g = ddst + (((long long)n * S + oy) * S + ox) * dst_ld + dst_coff;
                    for (int c = 0; c < C; ++c) acc[c] += g[c];
                }
            }
        }
        for (int c = 0; c < C; ++c) dsrc[i * C + c] = acc[c];
    }
}

extern "C" int fsv_crop_resize_fwd(const float* src, long long sn, long long sc, long long sh, long long sw, const int* box,
                                   float* dst, int N, int C, int S, int dst_ld, int dst_coff, void* stream) {
    FSV_REQUIRE(src && box && dst && N > 0 && C > 0 && S > 0 && dst_ld >= dst_coff + C, "crop_resize_fwd: bad args");
    long long total = (long long)N * S * S, blocks = (total + 255) / 256, cap = (long long)fsv_sm_count() * 16;
    k_crop_resize_fwd<<<(int)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream>>>(src, sn, sc, sh, sw, box, dst, N, C, S, dst_ld, dst_coff);
    FSV_CHECK_LAUNCH("crop_resize_fwd");
    return FSV_OK;
}

extern "C" int fsv_crop_resize_bwd(const float* ddst, int dst_ld, int dst_coff, const int* box, float* dsrc, int N, int C, int H, int W,
                                   int S, void* stream) {
    FSV_REQUIRE(ddst && box && dsrc && N > 0 && C > 0 && C <= 4 && S > 0 && H > 0 && W > 0 && dst_ld >= dst_coff + C, "crop_resize_bwd: bad args (C <= 4)");
    long long total = (long long)N * H * W, blocks = (total + 255) / 256, cap = (long long)fsv_sm_count() * 16;
    k_crop_resize_bwd<<<(int)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream>>>(ddst, dst_ld, dst_coff, box, dsrc, N, C, H, W, S);
    FSV_CHECK_LAUNCH("crop_resize_bwd");
    return FSV_OK;
}
