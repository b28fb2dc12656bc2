/*
 * fsv_b200.h -- C ABI of the B200-native (sm_100a) few-shot vid2vid hot path.
 *
 * This is the drop-in boundary of SURVEY.md section 8(b).  The reference has no FFI
 * of its own on this path: every hot op is an ATen/cuDNN library call made from the
 * nn.Modules returned by models/networks/__init__.py:29-55 (define_G / define_D).
 * Each entry point below therefore cites the reference call site(s) whose library
 * call it replaces.  The Python host side (few-shot-vid2vid_b200/fsv) binds these
 * with ctypes from torch.autograd.Function wrappers; see INTEGRATION.md.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no C++ / torch types cross the boundary.
 *  - every function returns 0 on success, a negative FSV_E* code on error;
 *    fsv_last_error() returns a thread-local message for the last failure.
 *  - all pointers are DEVICE pointers on the current CUDA device; work is enqueued
 *    on the cudaStream_t passed last (as void*); nothing synchronises the host.
 *  - activations are fp32 NHWC: element (n,h,w,c) of a buffer with channel stride
 *    `ld` and channel offset `coff` lives at ((n*H + h)*W + w)*ld + coff + c.  A
 *    consumer/producer can thus address a channel slice of a wider (concatenated)
 *    buffer without a copy.
 *  - conv weights are "OHWI": w[co][r][s][ci] (Cout, kh, kw, Cin) contiguous.
 *    A per-sample weight (the hyper-network's output, base_network.py:56-71
 *    batch_conv) is addressed as w + n*w_nstride.
 *  - caller allocates all outputs and workspaces.
 */
#ifndef FSV_B200_H_
#define FSV_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define FSV_OK 0
#define FSV_EINVAL (-1)   /* bad argument / unsupported shape */
#define FSV_ECUDA (-2)    /* CUDA runtime or launch error */
#define FSV_ENOTSUP (-3)  /* valid request this build cannot serve (e.g. no tcgen05 path) */

/* activation codes fused into producers */
#define FSV_ACT_NONE 0
#define FSV_ACT_LRELU 1    /* LeakyReLU(0.2): architecture.py:15-17 */
#define FSV_ACT_TANH 2     /* generator.py:211 */
#define FSV_ACT_SIGMOID 3  /* generator.py:487 conv_mask */
#define FSV_ACT_RELU 4     /* VGG19 perceptual features: models/networks/vgg.py:45-59 */

/* normalisation grouping */
#define FSV_NORM_BATCH 0     /* statistics over (N,H,W): SyncBatchNorm local stats, normalization.py:33,80 */
#define FSV_NORM_INSTANCE 1  /* statistics over (H,W) per sample: InstanceNorm2d, normalization.py:35,82 */

const char* fsv_last_error(void);
int fsv_version(void);
/* sm count and compute capability of the current device */
int fsv_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------ layout plumbing */
/* NCHW (contiguous) -> NHWC slice.  Replaces the implicit layout of every reference tensor
 * at the module boundary (generator.py:181 inputs, discriminator.py:49 input). */
int fsv_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, int dst_ld, int dst_coff, void* stream);
/* NHWC slice -> NCHW (contiguous); with accumulate!=0 adds into dst (gradient of the above). */
int fsv_nhwc_to_nchw(const float* src, float* dst, int N, int C, int H, int W, int src_ld, int src_coff,
                     int accumulate, void* stream);
/* copy a channel slice: dst[row, dcoff+c] = src[row, scoff+c]  (torch.cat on dim 1: generator.py:441-443,562-563) */
int fsv_copy_channels(const float* src, int src_ld, int src_coff, float* dst, int dst_ld, int dst_coff,
                      long long rows, int C, int accumulate, void* stream);
/* nearest x2 upsample (generator.py:124,207; nn.Upsample(scale_factor=2) generator.py:484,537) and its adjoint */
int fsv_upsample2x_fwd(const float* x, float* y, int N, int Hs, int Ws, int C, void* stream);
int fsv_upsample2x_bwd(const float* dy, float* dx, int N, int Hs, int Ws, int C, void* stream);
/* MaxPool2d(2, 2) of the VGG19 feature stack (vgg.py:48-59; H, W even or odd: floor); bwd routes dy to the arg-max of each window
 * (first maximum in row-major window order, as ATen) */
int fsv_maxpool2_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream);
int fsv_maxpool2_bwd(const float* x, const float* dy, float* dx, int N, int H, int W, int C, void* stream);
/* AvgPool2d(3, stride 2, pad 1, count_include_pad=False): discriminator.py:28 */
int fsv_avgpool3s2_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream);
int fsv_avgpool3s2_bwd(const float* dy, float* dx, int N, int H, int W, int C, void* stream);
/* g = dy * act'(y) given the post-activation output y (out_scale is applied after act) */
int fsv_act_bwd(const float* y, const float* dy, float* g, long long n, int act, float out_scale, void* stream);

/* ------------------------------------------------------------------ convolution / linear */
/* One descriptor for F.conv2d (architecture.py:60,81-84; generator.py:473-489,523-537;
 * discriminator.py:69-88), F.linear (generator.py:260-270: N=1,H=rows,W=1,k=1) and the
 * per-sample batch_conv (base_network.py:56-71: w_nstride != 0). */
typedef struct fsv_conv_desc {
    int N, H, W, Cin;      /* conv input dims (H,W are AFTER the optional upsample-on-load) */
    int x_ld, x_coff;      /* input buffer channel stride / offset */
    int up;                /* 1, or 2 = nearest x2 upsample on load (buffer is H/2 x W/2) */
    int Cout, kh, kw, stride, pad;
    int Ho, Wo;            /* output dims */
    int y_ld, y_coff;      /* output buffer channel stride / offset */
    int act;               /* FSV_ACT_* fused after bias (+residual) */
    float out_scale;       /* multiplies the result after act (generator.py:502 flow_multiplier) */
    long long w_nstride;   /* 0 = shared weight; else floats between per-sample weights */
    long long b_nstride;   /* same for the bias */
    int res_ld, res_coff;  /* residual buffer (added before act), used when residual != NULL */
    int in_act;            /* FSV_ACT_NONE or FSV_ACT_LRELU applied to x on load (generator.py:210 conv_img(actvn(x))) */
    int use_tc;            /* 0 = SIMT path, 1 = tcgen05/TMA path (FSV_ENOTSUP if the shape is not eligible), -1 = auto */
} fsv_conv_desc;

/* y = act(conv(x, w) + bias + residual) * out_scale */
int fsv_conv2d_fwd(const fsv_conv_desc* d, const float* x, const float* w, const float* bias,
                   const float* residual, float* y, void* stream);
/* dx (at conv-input resolution H x W, channel stride x_ld/x_coff) = conv_transpose(dy, w); accumulate!=0 adds */
int fsv_conv2d_dgrad(const fsv_conv_desc* d, const float* dy, const float* w, float* dx, int accumulate, void* stream);
/* dw[co][r][s][ci] (+= if accumulate) = sum_pixels dy * x ; per-sample when w_nstride != 0.
 * dbias (may be NULL) = sum_pixels dy (per sample when b_nstride != 0). */
int fsv_conv2d_wgrad(const fsv_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias,
                     int accumulate, void* stream);
/* 1 if the tcgen05/TMA implicit-GEMM path can serve this descriptor (fwd) */
int fsv_conv2d_tc_eligible(const fsv_conv_desc* d);
/* y = conv3x3(nearest_up2(x)) without materialising the upsample: four 2x2-tap stride-1 convolutions of the source image (one
 * per output parity) with host-pre-summed weights w4[co][ph][pw][a][b][ci] (Cout,16,Cin); d describes the conv at the
 * upsampled resolution with up = 2.  4/9 of the MACs of Upsample -> Conv2d (generator.py:484,537). */
int fsv_conv2d_fwd_tc_up2_eligible(const fsv_conv_desc* d);
/* w (Cout, 3, 3, Cin) OHWI -> the pre-summed w4 (Cout, 16, Cin) fsv_conv2d_fwd_tc_up2 takes */
int fsv_up2_weights(const float* w, float* w4, int Cout, int Cin, void* stream);
int fsv_conv2d_fwd_tc_up2(const fsv_conv_desc* d, const float* x, const float* w4, const float* bias,
                          const float* residual, float* y, void* stream);
/* Backward of y = conv3x3(nearest_up2(x)) at source resolution (autograd of generator.py:484,537; csrc/layout.cu).
 * Data gradient: dx = conv 4x4 / stride 2 / pad 1 of dy (fsv_conv2d_fwd with that descriptor) with the folded weights
 * wf (Cin, 4, 4, Cout) built here from wt (Cin, 3, 3, Cout), the weight with its channel axes swapped.
 * Weight gradient: fsv_conv2d_wgrad_tc of that 4x4 / stride-2 descriptor with dy in the role of x and x in the role of dy
 * gives dw16 (Cin, 4, 4, Cout); fsv_up2_wgrad_fold sums it into dw (Cout, 3, 3, Cin) (+= if accumulate). */
int fsv_up2_dgrad_weights(const float* wt, float* wf, int Cin, int Cout, void* stream);
int fsv_up2_wgrad_fold(const float* dw16, float* dw, int Cout, int Cin, int accumulate, void* stream);
/* Data gradient on the tcgen05/TMA kernel.  wt is the weight with its channel axes swapped: wt[ci][r][s][co]
 * (Cin, kh, kw, Cout).  Stride 1: one launch; stride 2: one launch per output parity class with a strided-output
 * epilogue.  dx is fully overwritten.  FSV_ENOTSUP when fsv_conv2d_dgrad_tc_eligible() is 0. */
int fsv_conv2d_dgrad_tc_eligible(const fsv_conv_desc* d);
int fsv_conv2d_dgrad_tc(const fsv_conv_desc* d, const float* dy, const float* wt, float* dx, void* stream);
/* Weight gradient on tcgen05: dw[co][r][s][ci] (+= if accumulate) = sum_pixels dy * x.  The operands are first
 * re-laid out into planar channel-major buffers inside `workspace` (fsv_conv2d_wgrad_tc_workspace() bytes, caller
 * allocated) so that the pixel axis is the contiguous GEMM-K axis.  Bias gradients stay with fsv_conv2d_wgrad. */
int fsv_conv2d_wgrad_tc_eligible(const fsv_conv_desc* d);
long long fsv_conv2d_wgrad_tc_workspace(const fsv_conv_desc* d);
int fsv_conv2d_wgrad_tc(const fsv_conv_desc* d, const float* x, const float* dy, float* dw, float* workspace,
                        int accumulate, void* stream);

/* ------------------------------------------------------------------ normalisation */
/* Per-channel reductions run as ONE launch: blocks leave fp64 partials in a workspace, the last block of each
 * (group, 32-channel block) adds them in a fixed order (deterministic; no memsets, no float atomics).
 * fsv_norm_work_doubles: doubles of workspace such a reduction needs (rows_per_group = N*HW for batch, HW for instance). */
long long fsv_norm_work_doubles(int groups, int C, long long rows_per_group);
/* The reductions' tickets come in 4 independent banks ("lanes"): calls issued under different lanes may run concurrently on
 * different streams; calls of one lane must be stream-ordered.  Thread-local host setting, default lane 0. */
int fsv_set_reduction_lane(int lane);
/* per-(group, channel) sum and sum of squares of an NHWC slice; groups = 1 (batch) or N (instance).
 * Outputs are DOUBLE [groups*C]; work: fsv_norm_work_doubles(groups, C, rows_per_group) doubles. */
int fsv_norm_stats(const float* x, int N, int HW, int C, int ld, int coff, int mode, double* sum, double* sumsq, double* work,
                   void* stream);
/* fsv_norm_stats + fsv_norm_finalize fused (the training-mode forward of every BatchNorm / InstanceNorm / SPADE,
 * normalization.py:32-35,42,78-82): count = rows per group, unbias_count = count * unbias_mul. */
int fsv_norm_stats_finalize(const float* x, int N, int HW, int C, int ld, int coff, int mode, double unbias_mul, float eps,
                            float momentum, float* running_mean, float* running_var, int update_running, float* mean,
                            float* rstd, double* work, void* stream);
/* mean/rstd from the sums; for mode batch + training also the running-stat update of
 * F.batch_norm (momentum, unbiased variance).  count = elements per channel that were summed;
 * unbias_count = elements per channel as seen by the reference (4x count when the reference
 * normalises a nearest-x2-upsampled copy of the tensor the sums were taken over). */
int fsv_norm_finalize(const double* sum, const double* sumsq, int groups, int C, double count, double unbias_count,
                      float eps, float momentum, float* running_mean, float* running_var, int update_running,
                      float* mean, float* rstd, void* stream);
/* eval-mode stats from running buffers: mean = running_mean, rstd = 1/sqrt(running_var+eps) */
int fsv_norm_from_running(const float* running_mean, const float* running_var, int C, float eps,
                          float* mean, float* rstd, void* stream);
/* y = act((x-mean)*rstd*weight + bias); weight/bias may be NULL.
 * Replaces BatchNorm/InstanceNorm + LeakyReLU pairs (architecture.py:65-68; generator.py:473-486; discriminator.py:76-85). */
int fsv_norm_apply_fwd(const float* x, const float* mean, const float* rstd, const float* weight, const float* bias,
                       float* y, int N, int HW, int C, int mode, int act, void* stream);
/* backward of the above given y (for the activation mask) and dy:
 *   g = dy*act'(y)*weight;  dx = rstd*(g - S1/cnt - xhat*S2/cnt)  (batch_stats!=0)  or  rstd*g  (eval)
 *   dweight = sum dy'*xhat, dbias = sum dy'  (when weight != NULL)
 * scratch: double[2*groups*C + fsv_norm_work_doubles(groups, C, rows_per_group)]. */
int fsv_norm_apply_bwd(const float* x, const float* y, const float* dy, const float* mean, const float* rstd,
                       const float* weight, float* dx, float* dweight, float* dbias, double* scratch,
                       int N, int HW, int C, int mode, int act, int batch_stats, void* stream);

/* Same backward WITHOUT the post-activation tensor: for act in {none, LeakyReLU, ReLU} the activation's derivative is recomputed from
 * sign(xhat*weight + bias) (the forward's own expression), saving one full-tensor read in each of the two passes. */
int fsv_norm_apply_bwd2(const float* x, const float* dy, const float* mean, const float* rstd, const float* weight,
                        const float* bias, float* dx, float* dweight, float* dbias, double* scratch,
                        int N, int HW, int C, int mode, int act, int batch_stats, void* stream);

/* ------------------------------------------------------------------ fused SPADE (normalization.py:37-52 + architecture.py:96-97) */
#define FSV_SPADE_MAX_MAPS 3
typedef struct fsv_spade_desc {
    int N, H, W, C;            /* output dims; x buffer is (N, H/up, W/up, C) contiguous */
    int up;                    /* 1 or 2: nearest x2 upsample of x on load (generator.py:207 feeding bn_0/bn_s) */
    int mode;                  /* FSV_NORM_BATCH / FSV_NORM_INSTANCE: layout of mean/rstd ([C] or [N*C]) */
    int act;                   /* FSV_ACT_NONE (shortcut, architecture.py:103) or FSV_ACT_LRELU */
    int nmaps;                 /* 1..3 label maps (None maps are dropped by the caller) */
    int K[FSV_SPADE_MAX_MAPS];         /* hidden channels of map i */
    int m_ld[FSV_SPADE_MAX_MAPS];      /* channel stride of map i's buffer (map is N,H,W,K at offset m_coff) */
    int m_coff[FSV_SPADE_MAX_MAPS];
    long long w_nstride[FSV_SPADE_MAX_MAPS];  /* 0 = fixed mlp_gamma/mlp_beta weights; else per-sample stride (hyper-weights) */
    int dgb_ld[FSV_SPADE_MAX_MAPS];    /* backward only: pixel stride of the dgamma[i]/dbeta[i] buffers (0 = C); 2C lets the
                                          caller interleave them as one (N,H,W,2C) tensor for a single 1x1 dgrad/wgrad */
} fsv_spade_desc;
/* out = act( (((x-mean)*rstd) * (1+g_0) + b_0) * (1+g_1) + b_1 ... ),  g_i = Wg_i . map_i + bg_i  (1x1).
 * bg[i] / bb[i] may be NULL (no bias): the reference's adaptive path applies the hyper-weights without their
 * bias slots (normalization.py:48-50 indexes weights[0][j] -> the weight tensor only). */
int fsv_spade_fwd(const fsv_spade_desc* d, const float* x, const float* mean, const float* rstd,
                  const float* const* maps, const float* const* wg, const float* const* bg,
                  const float* const* wb, const float* const* bb, float* out, void* stream);
/* Same contract on the tcgen05 path (TF32 gamma/beta GEMM in TMEM, fp32 modulation epilogue) for C % 64 == 0 and
 * K_i % 32 == 0; FSV_ENOTSUP otherwise. */
int fsv_spade_fwd_tc_eligible(const fsv_spade_desc* d);
int fsv_spade_fwd_tc(const fsv_spade_desc* d, const float* x, const float* mean, const float* rstd,
                     const float* const* maps, const float* const* wg, const float* const* bg,
                     const float* const* wb, const float* const* bb, float* out, void* stream);
int fsv_spade_bwd_tc(const fsv_spade_desc* d, const float* x, const float* mean, const float* rstd,
                     const float* const* maps, const float* const* wg, const float* const* bg,
                     const float* const* wb, const float* const* bb, const float* dout,
                     float* dxhat, float* const* dgamma, float* const* dbeta, void* stream);
/* Backward.  Recomputes gamma/beta and the pre-activation value (so `out` is not needed), walks the modulation chain backwards and emits
 *   dxhat  (N,H,W,C): gradient w.r.t. the normalised activation (feed to fsv_spade_norm_bwd)
 *   dgamma[i], dbeta[i] (N,H,W,C each): gradients of the 1x1 conv outputs (feed to fsv_conv2d_dgrad/wgrad) */
int fsv_spade_bwd(const fsv_spade_desc* d, const float* x, const float* mean, const float* rstd,
                  const float* const* maps, const float* const* wg, const float* const* bg,
                  const float* const* wb, const float* const* bb, const float* dout,
                  float* dxhat, float* const* dgamma, float* const* dbeta, void* stream);
/* dx (N,H/up,W/up,C) from dxhat: batch_stats!=0: rstd*(sum_children(g) - k*S1/cnt - k*xhat*S2/cnt), else rstd*sum(g).
 * scratch: double[2*groups*C + fsv_norm_work_doubles(groups, C, rows_per_group)] (rows at full resolution H*W). */
int fsv_spade_norm_bwd(const float* x, const float* dxhat, const float* mean, const float* rstd, float* dx,
                       double* scratch, int N, int H, int W, int C, int up, int mode, int batch_stats, void* stream);

/* ------------------------------------------------------------------ warp + composite (base_network.py:13-37; generator.py:431-443,214-224) */
/* img (N,H,W,Ci) NHWC, flow (N,H,W,2) in pixels (x,y), mask (N,H,W,1) or NULL, raw (N,H,W,Ci) or NULL.
 * out (N,H,W,out_ld) at out_coff:  blend==0: out[..,0:Ci] = warp, out[..,Ci] = mask (the [warp,mask] concat, when mask!=NULL)
 *                                   blend==1: out[..,0:Ci] = raw*mask + warp*(1-mask)  (generator.py:217,224) */
int fsv_warp_fwd(const float* img, const float* flow, const float* mask, const float* raw, float* out,
                 int N, int H, int W, int Ci, int out_ld, int out_coff, int blend, void* stream);
/* gradients: dflow (N,H,W,2), dmask (N,H,W,1) (may be NULL), draw (may be NULL), dimg (may be NULL; atomically accumulated, caller zeroes) */
int fsv_warp_bwd(const float* img, const float* flow, const float* mask, const float* raw, const float* dout,
                 float* dflow, float* dmask, float* draw, float* dimg,
                 int N, int H, int W, int Ci, int out_ld, int out_coff, int blend, void* stream);

/* ------------------------------------------------------------------ reference-feature softmax (generator.py:385) */
/* row-wise softmax over C for rows x C contiguous; and its backward dx = y*(dy - sum(dy*y)) */
int fsv_softmax_rows_fwd(const float* x, float* y, long long rows, int C, void* stream);
int fsv_softmax_rows_bwd(const float* y, const float* dy, float* dx, long long rows, int C, void* stream);

/* ------------------------------------------------------------------ spectral normalisation of a weight */
/* Replaces the torch.nn.utils.spectral_norm forward-pre-hook the reference wraps its convs / linears in
 * (architecture.py:60,81-84, generator.py:106-109, normalization.py:64-65, discriminator.py:69-88).
 * w_orig: weight_orig, (R, Cin, taps) contiguous (OIHW, taps = kh*kw; 1 for nn.Linear); u: weight_u (R), v: weight_v
 * (K = Cin*taps).  power=1 (training): one power iteration, u and v advance IN PLACE; power=0 (eval): none.
 * w_out: (R, taps, Cin) = w_orig / sigma repacked OHWI for fsv_conv2d_*; wt_out: NULL or (Cin, taps, R), the same
 * weight with the channel axes swapped (the B operand fsv_conv2d_dgrad_tc wants); uvs: (K + R + 1) floats = [v | u | sigma]
 * as used for this call (the backward needs them after the buffers moved on).  work: fsv_spectral_workspace(R, K) bytes.
 * Backward, u and v constants:  dw_orig (R, Cin, taps) = (dw_ohwi - (sum dw_ohwi * w_sn_ohwi) u v^T) / sigma. */
long long fsv_spectral_workspace(int R, int K);
int fsv_spectral_fwd(const float* w_orig, float* u, float* v, int R, int Cin, int taps, int power, float eps, float* w_out,
                     float* wt_out, float* uvs, float* work, void* stream);
int fsv_spectral_bwd(const float* dw_ohwi, const float* w_sn_ohwi, const float* uvs, int R, int Cin, int taps, float* dw_orig,
                     float* work, void* stream);

/* Grouped forward: every spectral weight a network touches in one forward pass in THREE launches (one per phase).
 * Fill w_orig/u/v/R/Cin/taps/want_wt of each item, call fsv_spectral_group_plan (host; fills the derived fields and
 * totals[9], see below), upload the items and a
 * block map ([(item, local block)] for the three phases back to back, int32 pairs) once, and keep a zero-initialised ticket
 * buffer per plan.  Per call pass freshly allocated `out` / `work` arenas: item i's W_sn lives at out + out_off (OHWI), its
 * channel-swapped copy at out + wt_off (when want_wt and emit_wt), [v | u | sigma] at out + uvs_off. */
typedef struct fsv_sn_item {
    const float* w_orig; float* u; float* v;
    int R, Cin, taps, want_wt;
    /* derived by fsv_spectral_group_plan */
    int K, nchunks, rs, rps, nblk2, nblk3, ticket_off, blk1, blk2, blk3;
    long long out_off, wt_off, uvs_off, work_off;
    int nb1, nb2;               /* backward: blocks of the dot phase / the transpose phase */
    long long bwd_work_off;     /* float offset into the backward work arena (nb1 partials + the scalar c) */
} fsv_sn_item;
/* totals[9] = {out floats, work floats, ticket words, fwd phase-1/2/3 blocks, bwd work floats, bwd phase-1 / phase-2 blocks} */
int fsv_spectral_group_plan(fsv_sn_item* items, int n, long long* totals);
/* Grouped backward (u, v constants): dw_arena + out_off(i) <- (dW_sn_i - (sum dW_sn_i * W_sn_i) u v^T) / sigma in weight_orig layout, for every
 * item whose entry in the device pointer table dws_dev is non-NULL; `out` is the forward's arena; map_bwd_dev = [(item, local block)] of
 * the two phases back to back; totals_bwd = {work floats, phase-1 blocks, phase-2 blocks} (= totals + 6). */
int fsv_spectral_group_bwd(const fsv_sn_item* items_dev, const int* map_bwd_dev, const long long* totals_bwd, const float* const* dws_dev,
                           const float* out, float* dw_arena, float* work, unsigned int* tickets, void* stream);
int fsv_spectral_group_fwd(const fsv_sn_item* items_dev, const int* map_dev, const long long* totals, int power, float eps,
                           int emit_wt, float* out, float* work, unsigned int* tickets, void* stream);

/* ------------------------------------------------------------------ multi-tensor Adam (base_model.py:39-48, loss_collector.py:227) */
/* One launch for all parameters of an optimizer; same update rule as torch.optim.Adam(lr, betas, eps) without amsgrad / weight
 * decay.  items: device array describing each parameter; chunks: device array of (item, chunk) int32 pairs covering every item
 * in pieces of fsv_adam_chunks(1) elements; step_dev: device float holding the number of steps taken so far (advanced here). */
typedef struct fsv_adam_item {
    float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
    long long numel;
} fsv_adam_item;
long long fsv_adam_chunks(long long numel);
int fsv_adam_step(const fsv_adam_item* items_dev, const int* chunks_dev, long long nchunks, float* step_dev, float lr,
                  float beta1, float beta2, float eps, void* stream);

/* ------------------------------------------------------------------ fused generator-step losses (loss_collector.py:131-215) */
/* Flow / mask losses in one pass (see csrc/losses.cu for the formulas).  All frame tensors NHWC contiguous except tgt (NCHW);
 * absent inputs are NULL: warp1/mask1 (no previous-frame branch), fake + face_avg + fg_diff (non-pose), ref_body_warp/body,
 * ref_fg_warp/fg.  out2 = {F_Warp / lambda_flow, F_Mask / lambda_mask}; work: fsv_flow_mask_loss_work_doubles() doubles. */
typedef struct fsv_flow_mask_desc {
    const float *warp0, *mask0, *warp1, *mask1, *tgt, *fake, *ref_body_warp, *body, *ref_fg_warp, *fg, *face_avg, *fg_diff;
    int B, H, W;
} fsv_flow_mask_desc;
long long fsv_flow_mask_loss_work_doubles(void);
int fsv_flow_mask_loss_fwd(const fsv_flow_mask_desc* d, float* out2, double* work, void* stream);
/* g_warp / g_mask: device scalars = d(total loss)/d(out2[0]), d(total loss)/d(out2[1]) */
int fsv_flow_mask_loss_bwd(const fsv_flow_mask_desc* d, const float* g_warp, const float* g_mask, float* dwarp0, float* dwarp1,
                           float* dmask0, float* dmask1, float* dfake, float* dref_body_warp, float* dref_fg_warp, void* stream);
/* Feature matching (loss_collector.py:206-215): x = one discriminator feature for the batch [fake ; real] (contiguous, `half`
 * elements per half); out[0] = mean |x[:half] - x[half:]|; backward writes g * sign / half into dx[:half] and zeros into dx[half:]. */
int fsv_halves_l1_fwd(const float* x, long long half, float* out, double* work, void* stream);
int fsv_halves_l1_bwd(const float* x, long long half, const float* g, float* dx, void* stream);

/* ------------------------------------------------------------------ pose label preprocessing + face region */
/* (MaxPool2d(15, stride 1, pad 7)(plane) > thr).float(): get_fg_mask, models/input_process.py:52-61.  plane n starts at
 * label + n*n_stride (pass the address of channel 2 of an NCHW label and n_stride = C*H*W); out (N, H, W). */
int fsv_fg_mask(const float* label, long long n_stride, float* out, int N, int H, int W, float thr, void* stream);
/* AvgPool2d(15, stride 1, pad 7)(get_face_mask(part)): input_process.py:81-93 + loss_collector.py:178-179; out (N, H, W). */
int fsv_face_mask_avg15(const float* part, long long n_stride, float* out, int N, int H, int W, void* stream);
/* get_part_mask (input_process.py:63-79): the 9 body-part group masks of a DensePose part-id plane; out NHWC (N, H, W, 9). */
int fsv_part_masks(const float* part, long long n_stride, float* out, int N, int H, int W, void* stream);
/* get_face_region (models/face_refiner.py:52-83) without nonzero()/.item(): face pixels = all given planes (p1, p2 may be
 * NULL) > thr; plane k of sample n starts at pk + n*sk.  openpose != 0 selects the keypoint-box arithmetic (:64-66), else the
 * DensePose one (:68-69).  box[n] = {ys, ye, xs, xe} (int32, device memory), shrunk by crop_smaller on every side. */
int fsv_face_bbox(const float* p0, const float* p1, const float* p2, long long s0, long long s1, long long s2, float thr,
                  int N, int H, int W, int openpose, int crop_smaller, int* box, void* stream);
/* crop_face_region (face_refiner.py:34-38): dst[n] = F.interpolate(src[n, :, ys:ye, xs:xe], size=(S, S), mode='nearest') for the
 * device-resident boxes; src element (n,c,y,x) lives at n*sn + c*sc + y*sh + x*sw (NCHW tensors and NCHW-shaped views of NHWC
 * buffers alike), dst is NHWC (N, S, S, dst_ld) at channel offset dst_coff.  bwd: dsrc (N, H, W, C) NHWC, fully overwritten
 * (zero outside the box), deterministic (gather form, no atomics); C <= 4. */
int fsv_crop_resize_fwd(const float* src, long long sn, long long sc, long long sh, long long sw, const int* box,
                        float* dst, int N, int C, int S, int dst_ld, int dst_coff, void* stream);
int fsv_crop_resize_bwd(const float* ddst, int dst_ld, int dst_coff, const int* box, float* dsrc, int N, int C, int H, int W,
                        int S, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FSV_B200_H_ */
