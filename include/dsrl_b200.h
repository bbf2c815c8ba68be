/* dsrl_b200.h -- C-ABI of libdsrl_b200.so: the B200 (sm_100a) implementation of the DSRL hot path.
 *
 * The reference (sanje2v/DualSuperResLearningForSemSeg) is pure Python and has no FFI layer; the boundary
 * this library sits behind is the pair of Python import surfaces the training/benchmark scripts use:
 *
 *   from models.losses import FALoss            (models/losses/FALoss.py:5-34; called at
 *                                                command_handlers/train_or_resume.py:118,437,444)
 *   from metrices import mIoU, Accuracy         (metrices/mIoU.py:5-41, metrices/Accuracy.py:4-30; called at
 *                                                command_handlers/train_or_resume.py:389-390,476-481 and
 *                                                command_handlers/benchmark.py:56-57,76-77)
 *
 * Each entry point below names the reference computation it replaces.  Conventions:
 *   - plain C: pointers, sizes, ints.  No torch / C++ types cross the boundary.
 *   - every data pointer is a DEVICE pointer to contiguous memory; `stream` is a cudaStream_t.  All work is
 *     enqueued on that stream; no entry point synchronises, allocates or frees caller memory.
 *   - return value: DSRL_OK (0) or a negative DSRL_ERR_*; dsrl_last_error() gives a thread-local message.
 *   - re-entrant: the only global state is one-time, mutex-guarded kernel attribute / descriptor setup.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns DSRL_ERR_CUDA.
 */
#ifndef DSRL_B200_H
#define DSRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSRL_B200_VERSION 200 /* major*10000 + minor*100 + patch */

typedef void *dsrl_stream_t; /* cudaStream_t */

enum dsrl_status {
    DSRL_OK = 0,
    DSRL_ERR_BAD_SHAPE = -1,   /* FALoss.py:19-20 / mIoU.py:16-17 shape contracts violated */
    DSRL_ERR_BAD_DTYPE = -2,
    DSRL_ERR_UNSUPPORTED = -3, /* valid request this build cannot serve (size limits, mode/reduction combos) */
    DSRL_ERR_CUDA = -4,
    DSRL_ERR_BAD_ARG = -5      /* null pointer, workspace too small, unknown enum value */
};

enum dsrl_fa_mode {
    DSRL_FA_REFERENCE = 0, /* what FALoss.py:8-34 computes: spectral norm, per-(b,c) w x w Gram, all-pairs L1 */
    DSRL_FA_POSITION = 1   /* paper's N x N position affinity (opt-in; not in the reference) */
};

enum dsrl_reduction { DSRL_REDUCE_NONE = 0, DSRL_REDUCE_MEAN = 1, DSRL_REDUCE_SUM = 2 }; /* FALoss.py:32-34 */

enum dsrl_precision {
    DSRL_PREC_FP32 = 0,   /* reference mode: CUDA-core FP32 FMA (always); position mode: 3xTF32 split on tcgen05 */
    DSRL_PREC_TF32 = 1,   /* position mode default: one tcgen05 kind::tf32 pass, FP32 accumulate in TMEM */
    DSRL_PREC_BF16 = 2,   /* reserved: rejected with DSRL_ERR_UNSUPPORTED (a single BF16 pass misses the loss tolerance) */
    DSRL_PREC_F16 = 3,    /* position mode: FP16 operands (the 11-bit significand of TF32; unit-norm features need no more
                             exponent range), tcgen05 kind::f16 at twice the TF32 rate, FP32 accumulate */
    /* Flag, OR-ed onto one of the three position-mode precisions above.  The gradient of the position loss is a sum of
     * sign(S1 - S2) terms; tensor-core operand rounding flips the sign of the ~1e-4 of the entries that lie within its error of
     * zero (0.5-1 % relative-norm on the gradient for densely distributed inputs).  With this flag every entry whose
     * tensor-core value is below ~3.5 sigma of that error is re-decided in FP64 from the unrounded features (a second,
     * HBM/L2-bound kernel over ~1e-3 of the entries), and the gradient is the one the exact signs give. */
    DSRL_PREC_EXACT_SIGNS = 16
};

enum dsrl_dtype { DSRL_U8 = 0, DSRL_I32 = 1, DSRL_I64 = 2 };

/* ---- diagnostics ------------------------------------------------------------------------------------- */
int dsrl_version(void);
const char *dsrl_last_error(void);
/* Number of kernel launches (and memset nodes) this library has enqueued from the calling process since
 * load; bench.py reports the delta over its timed region as `gpu_launches`. */
uint64_t dsrl_launch_count(void);

/* ---- FA loss (replaces FALoss.forward, FALoss.py:18-34, and its autograd backward) --------------------- */

/* Bytes of the opaque `saved` blob forward() fills and backward() consumes, and of the scratch workspace
 * forward needs (backward needs none except reference mode with DSRL_REDUCE_NONE).  `precision` as passed to the
 * calls (it selects which operand copies the position-mode workspace holds; ignored in reference mode).  C2 is only
 * meaningful in position mode (reference mode requires C1 == C2, FALoss.py:20).
 * Return 0 for an invalid / unsupported geometry. */
size_t dsrl_fa_saved_bytes(int mode, int precision, int B, int C1, int C2, int H, int W, int k);
size_t dsrl_fa_workspace_bytes(int mode, int precision, int B, int C1, int C2, int H, int W, int k);

/* Position mode with DSRL_PREC_EXACT_SIGNS: copies the statistics the last forward left in `saved` (device) to
 * out[0..4) on the host: entries listed as near ties, signs corrected, ties dropped because a row's list was full (those
 * keep the tensor-core sign), and 1e6 * the largest |D_exact| / threshold among the corrected entries (well below 1e6 when
 * the threshold is wide enough).  Synchronises `stream`. */
int dsrl_fa_sign_stats(const void *saved, uint64_t *out, dsrl_stream_t stream);

/* Forward.  x1: (B, C1, H, W), x2: (B, C2, H, W) fp32.  k = subsample_factor (FALoss.py:14,23-24).
 * `workspace` and `saved` must be 16-byte aligned; position mode may launch a cluster kernel (CTA pairs).
 * loss_out: 1 float for mean/sum; (B, C, n*n) floats, n = (W/k)^2, for DSRL_REDUCE_NONE (FALoss.py:27-34).
 * need_grad != 0 also prepares the gradient in `saved` (fused forward+backward work, one pass over the
 * pairs).  For mean/sum the local (this rank's) sum of |.| terms is left as a double at saved[0..8) so that
 * a multi-GPU caller can all-reduce it. */
int dsrl_fa_forward(int mode, int precision, const float *x1, const float *x2, int B, int C1, int C2, int H,
                    int W, int k, int reduction, int need_grad, float *loss_out, void *saved,
                    size_t saved_bytes, void *workspace, size_t workspace_bytes, dsrl_stream_t stream);

/* Backward.  grad_out: device pointer -- 1 float (mean/sum) or (B, C, n*n) floats (none).
 * dx1 / dx2: (B, C, H, W) fp32 outputs, either may be NULL (input does not require grad).
 * workspace may be NULL except in reference mode with DSRL_REDUCE_NONE; x1 / x2 are not read in position mode. */
int dsrl_fa_backward(int mode, int precision, const float *x1, const float *x2, const void *saved,
                     size_t saved_bytes, const float *grad_out, float *dx1, float *dx2, int B, int C1, int C2,
                     int H, int W, int k, int reduction, void *workspace, size_t workspace_bytes,
                     dsrl_stream_t stream);

/* Forward + backward in one call for callers that already know the upstream gradient (a captured training step with a
 * constant loss weight, cf. `w2 * FALoss()(..)` at train_or_resume.py:437): same results as dsrl_fa_forward(need_grad=1)
 * followed by dsrl_fa_backward.  Reference mode at the training shapes runs it as ONE kernel launch; position mode
 * without pooling (k == 1, both dx given) writes dX from the gradient kernel itself (no backward launch); every other
 * case is the two calls back to back.  reduction must be mean or sum; grad_out points to 1 float on the device.
 * `saved` is scratch here: do not pass it to dsrl_fa_backward afterwards. */
int dsrl_fa_forward_backward(int mode, int precision, const float *x1, const float *x2, int B, int C1, int C2, int H,
                             int W, int k, int reduction, const float *grad_out, float *loss_out, float *dx1,
                             float *dx2, void *saved, size_t saved_bytes, void *workspace, size_t workspace_bytes,
                             dsrl_stream_t stream);

/* In-place `d1[i] *= *grad_out`, `d2[i] *= *grad_out` (either pointer may be NULL); returns at once on the device when
 * *grad_out == 1.  Lets an autograd wrapper run dsrl_fa_forward_backward with a unit upstream gradient in its forward (one
 * pass, no `saved` gradient round trip) and apply the real upstream gradient -- `w2` of `w2 * FALoss()(..)`,
 * train_or_resume.py:437 -- when backward() delivers it.  grad_out: 1 device float. */
int dsrl_scale_grads(const float *grad_out, float *d1, int64_t n1, float *d2, int64_t n2, dsrl_stream_t stream);

/* ---- segmentation counts (replaces the three np.histogram passes of mIoU.update, mIoU.py:21-29, and the
 *      two reductions of Accuracy.update, Accuracy.py:19-20) ---------------------------------------------- */

/* Row layout of `counts`, one row per update:  [0,NC) area_pred  [NC,2NC) area_inter  [2NC,3NC) area_target
 * [3NC] correct  [3NC+1] valid.  int64. */
#define DSRL_SEG_ROW_LEN(num_classes) (3 * (num_classes) + 2)

/* pred/target: num_updates consecutive maps of npix_per_update elements each (an `update()` call of the
 * reference sees (B,H,W) = one map here).  mask: uint8/bool per pixel, or NULL to derive
 * `target != ignore_label` in-kernel (what every reference call site passes: train_or_resume.py:479,
 * benchmark.py:73).  counts: [num_updates][DSRL_SEG_ROW_LEN] int64, OVERWRITTEN. */
int dsrl_seg_counts(const void *pred, int pred_dtype, const void *target, int target_dtype, const uint8_t *mask,
                    int64_t num_updates, int64_t npix_per_update, int num_classes, int ignore_label,
                    int64_t *counts, dsrl_stream_t stream);

/* Fused argmax + counts (replaces `np.argmax(SSSR_output, axis=1)` at benchmark.py:69 /
 * `t.argmax(SSSR_output, dim=1)` at train_or_resume.py:477 followed by the counts above).
 * logits: (num_updates*batch, num_classes, H*W) fp32, NCHW; each update covers `batch` images of `hw` pixels.
 * Tie-break = first maximum, NaN counts as maximum (numpy/torch argmax).  pred_out (int64, may be NULL)
 * optionally receives the argmax map. */
int dsrl_seg_counts_from_logits(const float *logits, const void *target, int target_dtype, const uint8_t *mask,
                                int64_t num_updates, int64_t batch, int64_t hw, int num_classes,
                                int ignore_label, int64_t *counts, int64_t *pred_out, dsrl_stream_t stream);

/* ---- cross-entropy with an ignore label (the caller-side loss next to FA; replaces
 *      `t.nn.CrossEntropyLoss(ignore_index=IGNORE_CLASS_LABEL)`, train_or_resume.py:116,435, i.e. torch's
 *      log_softmax + nll_loss forward and their two backward kernels, by one pass each) ----------------------- */

/* logits: (B, C, HW) fp32 NCHW; target: (B, HW) of `target_dtype` (uint8 as the dataset delivers it, or the
 * int64 of `target.long()`).  Pixels whose target equals ignore_index contribute neither loss nor gradient.  A target
 * outside [0, C) that is not ignore_index (torch raises a device assert there) makes the loss NaN; its pixel gets no
 * gradient.
 * reduction: DSRL_REDUCE_MEAN (sum / number of valid pixels; NaN when there is none, like torch) or DSRL_REDUCE_SUM.
 * saved: dsrl_ce_saved_bytes(B, HW) bytes, 16-byte aligned; holds the per-pixel log-sum-exp and the valid count
 * for dsrl_ce_backward, which writes dlogits = grad_out * dloss/dlogits (grad_out: 1 device float). */
size_t dsrl_ce_saved_bytes(int B, int64_t HW);
int dsrl_ce_forward(const float *logits, const void *target, int target_dtype, int B, int C, int64_t HW,
                    int64_t ignore_index, int reduction, float *loss_out, void *saved, size_t saved_bytes,
                    dsrl_stream_t stream);
int dsrl_ce_backward(const float *logits, const void *target, int target_dtype, int B, int C, int64_t HW,
                     int64_t ignore_index, int reduction, const void *saved, size_t saved_bytes,
                     const float *grad_out, float *dlogits, dsrl_stream_t stream);

/* ---- the stage-3 losses in shared passes ----------------------------------------------------
 * The reference's stage-3 step evaluates   CE(SSSR, target) + w1 * MSE(SISR, image) + w2 * FA(T1(SSSR), T2(SISR))
 * (train_or_resume.py:435-438) where T1 / T2 are the "feature transformers" Conv2d(C -> 1, kernel 1, stride 8, no bias) +
 * BatchNorm2d(1) + ReLU (models/DSRL.py:86-95,181,184).  The strided 1x1 convolution reads 1/64 of the pixels the CE / MSE
 * passes stream anyway, so those passes emit it ("tap"); BatchNorm + ReLU are folded into the FA kernel's pooling read; and
 * the backward passes of CE / MSE add the transformer path's input gradient dz * w[c] at the strided pixels in place and
 * collect dw.  Forward: 4 launches for the three losses, backward: 3.
 *
 * tap_w: [C] convolution weight; tap_z: (B, Hf, Wf) convolution output, Hf = (H-1)/stride + 1; tap_dz: gradient w.r.t. it;
 * tap_dw_part: [dsrl_tap_dw_blocks(B, H, W, stride)][C] per-block partial sums of dw (every row written; the rows actually
 * used by a call are returned in *dw_blocks_used; sum them over axis 0).  The vector path needs W % 4 == 0, stride % 4 == 0. */
int dsrl_ce_forward_tap(const float *logits, const void *target, int target_dtype, int B, int C, int H, int W,
                        int64_t ignore_index, int reduction, float *loss_out, void *saved, size_t saved_bytes,
                        const float *tap_w, float *tap_z, int tap_stride, dsrl_stream_t stream);
int64_t dsrl_tap_dw_blocks(int B, int H, int W, int tap_stride);
int dsrl_ce_backward_tap(const float *logits, const void *target, int target_dtype, int B, int C, int H, int W,
                         int64_t ignore_index, int reduction, const void *saved, size_t saved_bytes, const float *grad_out,
                         float *dlogits, const float *tap_w, const float *tap_dz, float *tap_dw_part, int tap_stride,
                         int64_t *dw_blocks_used, dsrl_stream_t stream);

/* Mean squared error over (B, C, H, W) fp32 tensors (`t.nn.MSELoss()`, train_or_resume.py:117,436) in one pass forward
 * (deterministic two-level sum) and one pass backward (dx = grad_out * 2 (x - y) / count), with the same optional taps
 * (tap_w == NULL: none).  H * W divisible by 4, 16-byte aligned tensors. */
size_t dsrl_mse_workspace_bytes(int B, int C, int H, int W);
int dsrl_mse_forward(const float *x, const float *y, int B, int C, int H, int W, float *loss_out, void *workspace,
                     size_t workspace_bytes, const float *tap_w, float *tap_z, int tap_stride, dsrl_stream_t stream);
int dsrl_mse_backward(const float *x, const float *y, int B, int C, int H, int W, const float *grad_out, float *dx,
                      const float *tap_w, const float *tap_dz, float *tap_dw_part, int tap_stride, dsrl_stream_t stream);

/* BatchNorm2d(1) + ReLU of the two transformers (models/DSRL.py:93-95).  Forward: batch statistics (training) or the running
 * ones (eval) -> bn_out[t] = {a, b, mean, invstd}, the transformer output being relu(a z + b); in training mode the running
 * statistics are updated as torch does (momentum, unbiased variance).  Backward: from dF (gradient w.r.t. the transformer
 * output for a unit upstream gradient) and grad_out (1 device float) -> dz (gradient w.r.t. the convolution output) and
 * dgb = {dgamma, dbeta}.  count = B * Hf * Wf.  One launch each for both transformers. */
int dsrl_ft_bn_forward(const float *z1, const float *z2, int64_t count, const float *gamma1, const float *beta1,
                       float *run_mean1, float *run_var1, const float *gamma2, const float *beta2, float *run_mean2,
                       float *run_var2, float eps, float momentum, int training, float *bn_out, dsrl_stream_t stream);
int dsrl_ft_bn_backward(const float *z1, const float *dF1, float *dz1, float *dgb1, const float *z2, const float *dF2,
                        float *dz2, float *dgb2, int64_t count, const float *bn, const float *grad_out, int training,
                        dsrl_stream_t stream);

/* FA loss (reference semantics, C = 1) of the two transformer outputs relu(a z + b), read straight from the convolution outputs
 * z1, z2 (B, 1, H, W) with bn = dsrl_ft_bn_forward's bn_out; forward + backward in ONE launch: loss and df1 / df2 = gradient
 * w.r.t. the transformer outputs, scaled by *grad_out.  saved / workspace as dsrl_fa_forward_backward (reference mode). */
int dsrl_fa_forward_backward_transformed(const float *z1, const float *z2, const float *bn, int B, int H, int W, int k,
                                         int reduction, const float *grad_out, float *loss_out, float *df1, float *df2,
                                         void *saved, size_t saved_bytes, void *workspace, size_t workspace_bytes,
                                         dsrl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DSRL_B200_H */
