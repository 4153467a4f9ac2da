/*
 * b200unet.h — C ABI of libb200unet.so: the sm_100a CUDA implementation of the
 * 3D U-Net segmentation hot path of fransiskusbudi/multimodal_segmentation_project.
 *
 * The reference has no FFI layer: its boundary is the Python API of
 *   models/unet.py:6-90        (DoubleConv, UNet3D)
 *   models/unet_dann.py:65-98  (UNet3D.forward(x, return_features))
 *   utils/metrics.py:6-190     (losses, metrics)
 *   train_dann.py:22-49        (GradientReversal, DomainDiscriminator)
 * Every export below replaces the torch library op(s) the cited reference line
 * dispatches to.  The host-side mirror (multimodal_segmentation_project_b200/)
 * binds these with ctypes; INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *  - every pointer is a raw DEVICE pointer borrowed for the duration of the call
 *    (torch owns all memory); nothing is retained, nothing synchronises;
 *  - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *  - every export returns 0 on success or a negative B200_ERR_* code, with a
 *    thread-local message readable through b200_last_error();
 *  - activations inside the network are channels-last ("NDHWC"): [N, D, H, W, C],
 *    dtype B200_F32 or B200_BF16; logits / targets at the loss+metric boundary are
 *    the reference's NCDHW fp32 / int64 tensors;
 *  - no CPU fallback exists: on a non-sm_100 device the kernels fail with
 *    B200_ERR_ARCH / B200_ERR_CUDA.
 */
#ifndef B200UNET_H
#define B200UNET_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  B200_OK = 0,
  B200_ERR_SHAPE = -1,
  B200_ERR_ALIGN = -2,
  B200_ERR_ARCH = -3,
  B200_ERR_CUDA = -4,
  B200_ERR_UNSUPPORTED = -5
};

enum { B200_F32 = 0, B200_BF16 = 1 };

/* weight packing modes for b200_pack_conv3_weights */
enum {
  B200_PACK_FPROP = 0, /* [tap][ci][co]                      (direct kernels, fprop)   */
  B200_PACK_DGRAD = 1, /* [tap'][co][ci], tap' = 26 - tap     (direct kernels, dgrad)   */
  B200_PACK_FPROP_TC = 2, /* tcgen05 core-matrix layout, fprop (bf16 only)              */
  B200_PACK_DGRAD_TC = 3  /* tcgen05 core-matrix layout, dgrad (bf16 only)              */
};

/* segmentation loss modes (utils/metrics.py) */
enum {
  B200_LOSS_DICE_CE = 0,    /* combined_loss            utils/metrics.py:14-40   */
  B200_LOSS_TVERSKY = 1,    /* tversky_loss             utils/metrics.py:137-156 */
  B200_LOSS_CE_TVERSKY = 2, /* combined_ce_tversky_loss utils/metrics.py:158-167 */
  B200_LOSS_DICE = 3,       /* 'dice' closure           train_unet.py:185-199    */
  B200_LOSS_CE = 4          /* nn.CrossEntropyLoss mean (the CE half on its own)  */
};

const char* b200_last_error(void);
int b200_version(void);
/* 0 if device `dev` is sm_100 (B200); B200_ERR_ARCH otherwise. */
int b200_check_device(int dev);
/* number of kernels launched by this library in this process (bench.py's gpu_launches). */
int64_t b200_launch_count(void);

/* ---------------------------------------------------------------- layout / packing */

/* NCDHW fp32 -> NDHWC (dtype). Replaces nothing in the reference; it is the entry
 * conversion for in_channels > 1 (for in_channels == 1 the two layouts coincide). */
int b200_ncdhw_to_ndhwc(int dtype, const float* x, void* y, int64_t N, int64_t C, int64_t S, void* stream);
int b200_ndhwc_to_ncdhw(int dtype, const void* x, float* y, int64_t N, int64_t C, int64_t S, void* stream);
/* elementwise dtype cast fp32 -> dtype / dtype -> fp32, n elements */
int b200_cast_from_f32(int dtype, const float* x, void* y, int64_t n, void* stream);
int b200_cast_to_f32(int dtype, const void* x, float* y, int64_t n, void* stream);

/* nn.Conv3d weight [Cout, Cin, 3,3,3] fp32 (models/unet.py:11,15) -> kernel layout. */
int b200_pack_conv3_weights(int mode, int dtype, const float* w, void* out, int Cout, int Cin, void* stream);
int64_t b200_pack_conv3_bytes(int mode, int dtype, int Cout, int Cin);

/* ---------------------------------------------------------------- 3x3x3 convolution
 * nn.Conv3d(k=3, p=1) forward (models/unet.py:11,15) and, with DGRAD-packed weights,
 * its data gradient.  Input is the virtual channel concat [x0 | x1] (x1 may be NULL,
 * c1 = 0) which replaces torch.cat((skip, up), 1) at models/unet.py:84; output
 * channels [0,co0) go to y0 and [co0, co0+co1) to y1 (y1 may be NULL) which is the
 * un-concat of the decoder's data gradient.  bias may be NULL.  impl: 0 = auto,
 * 1 = CUDA-core implicit GEMM, 2 = tcgen05/TMEM implicit GEMM (bf16, TC-packed weights).
 */
/* resolves impl=0 (auto) for a given problem: returns 1 (CUDA-core, B200_PACK_FPROP/DGRAD weights) or
 * 2 (tcgen05, B200_PACK_*_TC weights); the caller packs the weights accordingly. */
/* every tcgen05 weight layout of a model in ONE launch (once per optimiser step): `jobs` = device array of njobs + 1 records
 * { const float* w; void* out; int32 Cout, Cin, dgrad, pad; int64 group_begin } (40 bytes; group_begin counts 16-byte output
 * groups, 27*Cin*Cout/8 per job; record njobs carries the total).  out = what b200_pack_conv3_weights(mode FPROP_TC / DGRAD_TC)
 * writes for that layer. */
int b200_pack_conv3_batched(const void* jobs, int njobs, int64_t total_groups, void* stream);
int b200_conv3d_k3_select(int dtype, int impl, int c0, int c1, int co0, int co1, int N, int D, int H, int W);
/* selects the persistent tcgen05 convolution (one CTA per SM looping over tiles, double-buffered TMEM):
 * 0 never, 1 auto (default: every layer with >= 2 tiles per SM and <= 64 output channels per tile), 2 same as 1,
 * 3 only 16->16 layers — for tests and benchmarks. */
int b200_set_conv_persistent(int mode);
/* the row-streaming tcgen05 convolution (one 128-voxel row per M tile, A operand reused across the three kh taps) serves
 * full-resolution layers with 16 / 32 output channels: 1 = wherever it applies (default), 0 = never — tests and A/B timing. */
int b200_set_conv_rowstream(int on);
int b200_conv3d_k3(int dtype, int impl, const void* x0, int c0, const void* x1, int c1,
                   const void* wpack, const float* bias, void* y0, int co0, void* y1, int co1,
                   int N, int D, int H, int W, void* stream);

/* Convolution + BatchNorm3d batch statistics in one kernel (reference models/unet.py:11-12, :15-16: nn.Conv3d -> nn.BatchNorm3d
 * in training mode).  b200_conv3d_k3_bnstats_blocks() returns how many partial rows the fused kernel writes for a problem, or 0
 * when it does not apply (then call b200_conv3d_k3 + b200_bn_stats).  b200_conv3d_k3_bnstats() writes y0 and
 * partials[rows][2][co0] fp32 = per-CTA (sum, sum of squares) of (y - bias) over the bf16 values stored; finish with
 * b200_bn_finalize_ex(partials, rows, bias, ...).  partials must hold b200_bn_partials_bytes(co0). */
int b200_conv3d_k3_bnstats_blocks(int dtype, int impl, int c0, int c1, int co0, int co1, int N, int D, int H, int W);
int b200_conv3d_k3_bnstats(int dtype, int impl, const void* x0, int c0, const void* x1, int c1, const void* wpack,
                           const float* bias, void* y0, int co0, int N, int D, int H, int W, float* partials, void* stream);

/* Data gradient of a convolution whose INPUT was relu(bn(xprev)) (the second conv of a DoubleConv, models/unet.py:15) + the
 * BatchNorm-backward reduction of that previous layer in one kernel: y0 = gy and partials[rows][2][co0] = per-CTA
 * (sum g, invstd * sum g * (xprev - mean)) with g = gy * [bn(xprev) > 0] (= b200_bn_act_bwd_reduce's partials, no dropout).
 * rows = b200_conv3d_k3_bnbwd_blocks() (0: not served, run b200_conv3d_k3 + b200_bn_act_bwd_reduce); finish with
 * b200_bn_bwd_finalize_ex(partials, rows, ...) and b200_bn_act_bwd_apply. */
int b200_conv3d_k3_bnbwd_blocks(int dtype, int impl, int c0, int co0, int N, int D, int H, int W);
int b200_conv3d_k3_bnbwd(int dtype, int impl, const void* x0, int c0, const void* wpack, void* y0, int co0, int N, int D, int H,
                         int W, const void* xprev, const float* scale, const float* shift, const float* mean, const float* invstd,
                         float* partials, void* stream);

/* weight gradient of nn.Conv3d(k=3,p=1): dw[Cout, Cin, 3,3,3] fp32 (torch layout, overwritten),
 * dbias[Cout] fp32 (may be NULL).  workspace: b200_conv3d_wgrad_workspace() bytes. */
int64_t b200_conv3d_wgrad_workspace(int c0, int c1, int Cout, int N, int D, int H, int W);
/* process-wide selection of the weight-gradient kernel: 0 auto (tcgen05 for bf16 when the channel
 * counts allow), 1 CUDA-core split-K, 2 tcgen05 (error if unsupported) — for tests and benchmarks. */
int b200_set_wgrad_impl(int impl);
/* voxel-pair tcgen05 weight gradient for the 16-channel layers (wgrad_tc4.cu): on = 0 falls back to the M = 64 kernel
 * (wgrad_tc2.cu); dseg > 0 forces the d-run per CTA, 0 = planned — for tests and A/B timing.  Workspace sizes depend on
 * these settings: change them only between calls of b200_conv3d_wgrad_workspace / b200_conv3d_wgrad pairs. */
int b200_set_wgrad_pair(int on, int dseg);
/* test hook: the next wide-row tcgen05 weight gradient returns B200_ERR_UNSUPPORTED without launching (checks that a
 * failing kernel selection surfaces as an error instead of an uninitialised dw). */
int b200_debug_fail_next_wgrad(int on);
int b200_conv3d_wgrad(int dtype, const void* x0, int c0, const void* x1, int c1, const void* dy, int Cout,
                      float* dw, float* dbias, void* workspace, int64_t workspace_bytes,
                      int N, int D, int H, int W, void* stream);

/* ---------------------------------------------------------------- BatchNorm3d + ReLU + Dropout3d
 * models/unet.py:12-14,16-18.  Statistics are per channel over M = N*D*H*W rows.
 * bn_stats writes per-block partial sums of (x - x[row 0]) and (x - x[row 0])^2 to `partials`
 * (b200_bn_partials_bytes(C) bytes, fp32; the shift avoids cancellation when |mean| >> std),
 * bn_finalize (given the same x) reduces them in fixed order in fp64 -> deterministic.
 */
int64_t b200_bn_partials_bytes(int C);
int b200_bn_stats(int dtype, const void* x, int64_t M, int C, float* partials, void* stream);
/* training != 0: batch statistics (biased var for normalisation, unbiased for running_var,
 * momentum update, num_batches_tracked += 1); training == 0: running statistics.
 * Outputs scale[c] = gamma*invstd, shift[c] = beta, mean[c], invstd[c]; the apply kernels evaluate
 * (x - mean)*scale + shift (the reference's order of operations: no cancellation when |mean| >> std). */
int b200_bn_finalize(int dtype, const void* x, const float* partials, int64_t M, int C, const float* gamma, const float* beta,
                     float eps, float momentum, int training, float* running_mean, float* running_var,
                     int64_t* num_batches_tracked, float* scale, float* shift, float* mean, float* invstd,
                     void* stream);
/* training-mode finalize over `nblocks` partial rows whose sums were taken around shift_vec[c] (NULL = 0): for the partials of
 * b200_conv3d_k3_bnstats pass its row count and the convolution bias. */
int b200_bn_finalize_ex(const float* partials, int nblocks, const float* shift_vec, int64_t M, int C, const float* gamma,
                        const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                        int64_t* num_batches_tracked, float* scale, float* shift, float* mean, float* invstd, void* stream);
/* y = dropmask[n,c] * relu((x - mean[c])*scale[c] + shift[c]);  dropmask NULL = no dropout;
 * x: [N, S, C] rows, y same. relu: 0/1. */
int b200_bn_act_fwd(int dtype, const void* x, void* y, const float* scale, const float* shift, const float* mean,
                    const float* dropmask, int relu, int64_t N, int64_t S, int C, void* stream);
/* backward: g = gy * dropmask * [relu active]; partial sums of g and g*xhat */
int b200_bn_act_bwd_reduce(int dtype, const void* gy, const void* x, const float* scale, const float* shift,
                           const float* mean, const float* invstd, const float* dropmask, int relu,
                           int64_t N, int64_t S, int C, float* partials, void* stream);
/* reduces partials -> dgamma, dbeta (fp32, overwritten) and coefficient vectors for bwd_apply */
int b200_bn_bwd_finalize(const float* partials, int64_t M, int C, float* dgamma, float* dbeta, float* sums, void* stream);
/* same over an explicit number of partial rows (b200_head_bwd writes b200_head_blocks() rows) */
int b200_bn_bwd_finalize_ex(const float* partials, int nblocks, int64_t M, int C, float* dgamma, float* dbeta, float* sums,
                            void* stream);
/* dx = scale * (g - [training] (sum_g/M + xhat * sum_gxhat/M)) */
int b200_bn_act_bwd_apply(int dtype, const void* gy, const void* x, void* dx, const float* scale, const float* shift,
                          const float* mean, const float* invstd, const float* dropmask, int relu,
                          const float* sums, int training, int64_t N, int64_t S, int C, void* stream);
/* column sums of a [M, C] matrix -> out[C] fp32 (conv bias gradients) */
int b200_channel_sum(int dtype, const void* x, int64_t M, int C, float* partials, float* out, void* stream);

/* ---------------------------------------------------------------- MaxPool3d(2,2)  models/unet.py:40,71 */
int b200_maxpool2_fwd(int dtype, const void* x, void* y, int N, int D, int H, int W, int C, void* stream);
/* gradient goes to the first maximum in (d,h,w) scan order, as ATen's max_pool3d_with_indices */
/* BatchNorm3d + ReLU apply and MaxPool3d(2,2) of the result in ONE pass (models/unet.py:16-18 feeding :69-71: the encoder
 * hand-off): y [N,D,H,W,C] and pooled [N,D/2,H/2,W/2,C] from the pre-BN tensor x; element arithmetic of b200_bn_act_fwd
 * (relu = 1, no Dropout3d mask) and b200_maxpool2_fwd.  Needs even D/H/W and C/8 dividing 256 (b200_bn_act_pool_fwd_supported). */
int b200_bn_act_pool_fwd_supported(int D, int H, int W, int C);
int b200_bn_act_pool_fwd(int dtype, const void* x, void* y, void* pooled, const float* scale, const float* shift, const float* mean,
                         int N, int D, int H, int W, int C, void* stream);
int b200_maxpool2_bwd(int dtype, const void* x, const void* gy, void* gx, int N, int D, int H, int W, int C, void* stream);
/* the encoder output feeds both the pool and the decoder's skip connection (models/unet.py:69-71,84): gx = gskip + scatter(gy)
 * in one pass instead of the pool backward followed by autograd's accumulation kernel */
int b200_maxpool2_bwd_add(int dtype, const void* x, const void* gy, const void* gskip, void* gx, int N, int D, int H, int W, int C,
                          void* stream);

/* ---------------------------------------------------------------- ConvTranspose3d(k=2,s=2)  models/unet.py:56-58,79
 * x [N,D,H,W,Cin] -> y [N,2D,2H,2W,Cout]; w is the torch layout [Cin, Cout, 2,2,2] fp32. */
int b200_convt2_fwd(int dtype, const void* x, const float* w, const float* bias, void* y,
                    int N, int D, int H, int W, int Cin, int Cout, void* stream);
int b200_convt2_bwd_data(int dtype, const void* gy, const float* w, void* gx,
                         int N, int D, int H, int W, int Cin, int Cout, void* stream);
int64_t b200_convt2_wgrad_workspace(int Cin, int Cout, int N, int D, int H, int W);
int b200_convt2_bwd_weight(int dtype, const void* x, const void* gy, float* dw, float* dbias,
                           void* workspace, int64_t workspace_bytes,
                           int N, int D, int H, int W, int Cin, int Cout, void* stream);
/* nearest-neighbour resize of an NDHWC tensor (F.interpolate, models/unet.py:81-83) and its adjoint */
int b200_nearest_resize_fwd(int dtype, const void* x, void* y, int N, int D, int H, int W, int OD, int OH, int OW, int C, void* stream);
int b200_nearest_resize_bwd(int dtype, const void* gy, void* gx, int N, int D, int H, int W, int OD, int OH, int OW, int C, void* stream);

/* ---------------------------------------------------------------- fused head (training, bf16, 16 channels, 2..4 classes)
 * models/unet.py:16-18 (last BatchNorm3d + ReLU), :62,87 (final 1x1x1 conv) and utils/metrics.py:14-40, 65-167 in ONE pass each way.
 * b200_head_fwd: x = PRE-BatchNorm activation [N*S, 16] bf16; scale/shift/mean from b200_bn_finalize(_ex); target uint8 or int64
 * (label_bytes 1 / 8); writes fp32 NCDHW logits, the loss sums in b200_kd_loss_fwd's layout (feed b200_seg_loss_finalize) and, when
 * conf != NULL, the C x C confusion counts conf[target][argmax].  The normalised activation is never materialised.
 * b200_head_bwd: recomputes it, writes gy = d loss / d (normalised activation) as bf16, the 1x1 conv's dw[C][16] / db[C] and the
 * BatchNorm-backward partial sums bnpart[b200_head_blocks()][2][16] (finish with b200_bn_bwd_finalize_ex + b200_bn_act_bwd_apply).
 * Workspaces: wpart b200_head_blocks() * 68 floats, bnpart b200_head_blocks() * 32 floats. */
int b200_head_blocks(int64_t N, int64_t S);
int b200_head_fwd(const void* x, const float* scale, const float* shift, const float* mean, const float* w, const float* bias,
                  int round_bf16, const void* target, int label_bytes, int64_t N, int64_t S, int Cin, int C, float* logits,
                  double* sums, unsigned long long* conf, void* stream);
int b200_head_bwd(const float* logits, const void* target, int label_bytes, const float* coef, const float* gout, const void* x,
                  const float* scale, const float* shift, const float* mean, const float* invstd, const float* w, int64_t N,
                  int64_t S, int Cin, int C, void* gy, float* wpart, float* bnpart, float* dw, float* db, void* stream);

/* ---------------------------------------------------------------- final 1x1x1 conv  models/unet.py:62,87
 * x NDHWC (dtype) -> logits NCDHW fp32 [N, Cout, S]; w [Cout, Cin] fp32.
 * round_bf16 != 0 rounds the logits to bf16 first (what autocast + Accelerate's fp32 cast produce). */
int b200_conv1x1_fwd(int dtype, const void* x, const float* w, const float* bias, float* y,
                     int64_t N, int64_t S, int Cin, int Cout, int round_bf16, void* stream);
/* gy NCDHW fp32 -> gx NDHWC (dtype, may be NULL), dw[Cout,Cin], db[Cout] (fp32, overwritten; may be NULL) */
int b200_conv1x1_bwd(int dtype, const void* x, const float* w, const float* gy, void* gx, float* dw, float* db,
                     float* partials, int64_t N, int64_t S, int Cin, int Cout, void* stream);
int64_t b200_conv1x1_partials_bytes(int Cin, int Cout);

/* ---------------------------------------------------------------- losses  utils/metrics.py:14-40,137-190
 * logits [N, C, S] fp32 (NCDHW), target [N, S] int64 (the squeezed [N,1,...] tensor).
 * sums (double[4 + 4*C]) = {CE_sum, KL_sum, 0, 0, then per class I_k, P_k, T_k, 0}: batch-global.
 * seg_loss_fwd zeroes and fills sums; seg_loss_finalize writes the scalar loss (fp32) and the
 * per-class backward coefficients coef[2*C + 2].
 */
int b200_seg_loss_fwd(const float* logits, const int64_t* target, int64_t N, int C, int64_t S, double* sums, void* stream);
/* teacher != NULL adds the KD term sum_v sum_c q_c (log q_c - log_softmax(s/T)_c) into sums[1] */
int b200_kd_loss_fwd(const float* student, const float* teacher, const int64_t* target, float temperature,
                     int64_t N, int C, int64_t S, double* sums, void* stream);
/* loss = w_ce*CE + w_reg*region(mode: dice eps=1e-5 | tversky alpha,beta,eps) + w_kd*T^2*KL_mean */
int b200_seg_loss_finalize(const double* sums, int mode, float alpha, float beta, float kd_alpha, float temperature,
                           int has_kd, int64_t N, int C, int64_t S, float* loss, float* coef, void* stream);
/* dlogits = gout[0] * dL/dlogits; teacher may be NULL (no KD term) */
int b200_seg_loss_bwd(const float* logits, const float* teacher, const int64_t* target, const float* coef,
                      const float* gout, float temperature, int64_t N, int C, int64_t S, float* dlogits, void* stream);

/* ---------------------------------------------------------------- metrics  utils/metrics.py:65-129
 * conf[t*C + p] (int64, zeroed by the call) += 1 for target class t, argmax class p
 * (ties -> lowest index, NaN counts as maximum; torch.argmax semantics). Targets outside [0,C) are
 * counted in no cell. */
int b200_confusion(const float* logits, const int64_t* target, int64_t N, int C, int64_t S, int64_t* conf, void* stream);
/* argmax mask only: out [N, S] uint8 */
int b200_argmax(const float* logits, int64_t N, int C, int64_t S, uint8_t* out, void* stream);

/* sliding-window inference (BASELINE config #5; the evaluator test_model.py:248 runs whole volumes):
 * acc[c, d0+z, h0+y, w0+x] += logits[c, z, y, x], cnt[d0+z, h0+y, w0+x] += 1 for one window;
 * finalize divides acc by cnt (voxels never covered stay 0). acc [C,D,H,W] fp32, cnt [D,H,W] fp32. */
int b200_window_accumulate(float* acc, float* cnt, const float* logits, int C, int D, int H, int W,
                           int d0, int h0, int w0, int wd, int wh, int ww, void* stream);
int b200_window_finalize(float* acc, const float* cnt, int C, int64_t S, void* stream);

/* ---------------------------------------------------------------- DANN  models/unet_dann.py:79, train_dann.py:22-49 */
/* torch.mean(bottleneck, dim=[2,3,4]): x NDHWC [N,S,C] -> out fp32 [N,C]; and its adjoint */
int b200_gap_fwd(int dtype, const void* x, float* out, int64_t N, int64_t S, int C, void* stream);
/* gx[n,s,c] (+)= gout[n,c]/S; accumulate != 0 adds into gx */
int b200_gap_bwd(int dtype, const float* gout, void* gx, int accumulate, int64_t N, int64_t S, int C, void* stream);
/* out = alpha * x (gradient reversal backward: alpha = -lambda), n fp32 elements */
int b200_scale_f32(const float* x, float* out, float alpha, int64_t n, void* stream);
/* nn.Linear forward y[B,O] = act(x[B,I] @ w[O,I]^T + b) * dropmask ; relu 0/1 ; dropmask [B,O] or NULL */
int b200_linear_fwd(const float* x, const float* w, const float* b, const float* dropmask, int relu,
                    float* y, int B, int I, int O, void* stream);
/* backward of the above: g = gy * dropmask * [y>0 if relu]; gx = g @ w; dw = g^T @ x; db = sum g */
int b200_linear_bwd(const float* x, const float* w, const float* y, const float* gy, const float* dropmask, int relu,
                    float* gx, float* dw, float* db, int B, int I, int O, void* stream);
/* mean cross-entropy over B rows, labels int64; loss fp32[1]; dlogits = (softmax - onehot)/B */
int b200_ce_rows(const float* logits, const int64_t* labels, int B, int C, float* loss, float* dlogits, void* stream);

/* ---------------------------------------------------------------- optimiser (train_unet.py:226,378)
 * Fused AdamW over one flat fp32 buffer.  All per-step scalars live on the device so that a
 * captured CUDA graph stays valid: hyper = float[4] {lr, 1-beta1^t, 1-beta2^t, unused};
 * adamw_prepare increments *step (int64, device) and refreshes the bias corrections; the host
 * changes the learning rate by writing hyper[0].  Gradients are multiplied by grad_scale first
 * (1/world_size after a sum all-reduce). */
int b200_adamw_prepare(int64_t* step, float beta1, float beta2, float* hyper, void* stream);
int b200_adamw_flat(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper, float beta1,
                    float beta2, float eps, float weight_decay, float grad_scale, void* stream);

/* ---- device-side input pipeline: utils/dataloader.py:111-117 (CT window), :128-145 (MRI z-score, 1st-99th
 * percentile clip, min-max), :162-181 (AMOS / CHAOS label remaps) on tensors already in HBM --------------------------------- */
int b200_ct_window(const float* x, float* y, int64_t n, float lo, float hi, void* stream);
int64_t b200_preprocess_workspace_bytes(void);
/* out[0] = mean, out[1] = population variance of x (fp64, device memory); fixed-order fp64 accumulation */
int b200_moments_f32(const float* x, int64_t n, void* workspace, double* out, void* stream);
/* values[r] = element of 0-based rank ranks[r] in ascending order (exact order statistic, 3-pass radix select, no host sync) */
int b200_select_ranks_f32(const float* x, int64_t n, const int64_t* ranks, int nranks, float* values, void* workspace, void* stream);
/* params (fp64, device): mean, std + 1e-8 (float32 values), low, high, high - low + 1e-8: numpy's dtype flow of preprocess_mri */
int b200_mri_normalize(const float* x, float* y, int64_t n, const double* params, void* stream);
/* out = 0, then for each range in order: lo <= in <= hi -> val (host arrays, <= 8 ranges); out_u8: write uint8 instead of int64 */
int b200_label_remap(const int64_t* in, void* out, int64_t n, const int64_t* lo, const int64_t* hi, const int64_t* val, int nranges,
                     int out_u8, void* stream);

/* ---------------------------------------------------------------- intensity augmentations (utils/dataloader.py:252-260)
 * Device versions of the five MONAI transforms combined_transform() applies; MONAI (monai>=1.2.0) is not part of the reference
 * tree: csrc/augment_kernels.cu restates its published array transforms.  Random draws are made by the caller. float32 [C, D, H, W]. */
int b200_aug_bias_field(const float* x, float* y, int C, int D, int H, int W, int degree, const double* coeff, void* stream);
int b200_aug_gaussian_noise(const float* x, const float* z, float* y, int64_t n, float mean, float std, void* stream);
int b200_minmax_f32(const float* x, int64_t n, float* minmax, void* workspace, void* stream);
int b200_aug_adjust_contrast(const float* x, float* y, int64_t n, const float* minmax, float gamma, void* stream);
int b200_aug_histogram_shift(const float* x, float* y, int64_t n, const float* minmax, const double* ref, const double* floating, int ncp,
                             void* stream);
int b200_aug_coarse_dropout(float* img, int64_t* label, int C, int D, int H, int W, const int* holes, int nholes, float fill, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200UNET_H */
