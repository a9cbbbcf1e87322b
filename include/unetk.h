/*
 * unetk — C ABI of the B200 (sm_100a) kernel library behind the U-Net / AnomalyUNet hot path of
 * ukeSJTU/tiaozhanbei-unet.
 *
 * The reference has no FFI / plugin interface of its own: its hot path is the Python import
 * surface of src/model.py and src/train_utils.py, and below that stock PyTorch operators.  This
 * header is the operator-level boundary the drop-in Python modules (tiaozhanbei-unet_b200/model.py,
 * train_utils.py) bind with ctypes; every entry point names the reference operator it replaces
 * (paths relative to /root/reference).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  All pointers are BORROWED device pointers (the
 *    library never allocates or frees device memory); `stream` is a cudaStream_t passed as void*.
 *  - Asynchronous on `stream`.  Returns 0 on success, a negative UNETK_ERR_* code otherwise; the
 *    message is available from unetk_last_error() (thread-local).
 *  - Built for sm_100a only.  Any other device -> UNETK_ERR_ARCH.  There is no CPU or library
 *    fallback.
 *  - Activations: NHWC bf16, channel count a multiple of 64 (8 for the pure element-wise ops).
 *    Images / losses / network outputs: NCHW fp32, as the reference's callers hand them over.
 *  - Packed weights (bf16), produced by unetk_permute_batched from the fp32 reference layouts:
 *      conv3x3 forward  Wf[co][tap][ci]      = W[co][ci][ky][kx],           tap = 3*ky + kx
 *      conv3x3 dgrad    Wd[ci][tap][co]      = W[co][ci][2-ky][2-kx]        (taps flipped)
 *      convT   forward  Wt[tap*Cout+co][ci]  = W[ci][co][ky][kx],           tap = 2*ky + kx
 *      convT   dgrad    Wtd[ci][tap*Cout+co] = W[ci][co][ky][kx]
 *    Weight gradients come back fp32 as dW[co][tap][ci] and are re-laid-out to the reference
 *    layout with unetk_permute_batched.
 */
#ifndef UNETK_H_
#define UNETK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef UNETK_OK
#define UNETK_OK 0
#define UNETK_ERR_ARG (-1)       /* bad shape / alignment / null pointer */
#define UNETK_ERR_ARCH (-2)      /* device is not sm_100 */
#define UNETK_ERR_CUDA (-3)      /* a CUDA runtime / driver call failed */
#define UNETK_ERR_WORKSPACE (-4) /* workspace too small */
#endif

#define UNETK_VERSION 100 /* 0.1.0 */

/* Optional fused first pass of a BatchNorm+ReLU backward (HOST struct, read at launch; NULL = off).
 * A kernel that completes a gradient dA of an activation a = relu(y*scale+shift) also accumulates
 *   sums[0..C) += sum g,  sums[C..2C) += sum g*xhat,   g = dA*[y*scale+shift > 0], xhat = (y-mean)*invstd
 * (what unetk_bn_bwd_reduce computes), saving one pass over dA and y.  The caller zeroes sums. */
typedef struct {
  const void* y;          /* raw conv output of the producing layer, NHWC bf16, same shape as dA */
  const float* scale;
  const float* shift;
  const float* mean;
  const float* invstd;
  float* sums;            /* [2][C] */
} unetk_bn_ctx;

/* Training-mode BatchNorm finalize folded into the kernel that applies it (HOST struct, read at launch):
 * (scale, shift) are derived from the batch sums by every thread exactly as unetk_bn_finalize does; one
 * block also stores scale / shift / mean / invstd and updates running_mean / running_var (momentum,
 * unbiased variance) and num_batches_tracked (all three nullable). */
typedef struct {
  const float* stats;        /* [2][C] sum, sum of squares from the conv epilogue */
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  long long* num_batches_tracked;
  float* scale;              /* outputs, [C] each */
  float* shift;
  float* mean;
  float* invstd;
  double count;              /* N*H*W (x world size with synchronised statistics): exact beyond 2^24 */
  double inv_count;          /* 1 / count: written by the library at launch, callers leave it 0 */
  float unbias;              /* count / (count - 1) for running_var's unbiased variance: likewise written at launch */
  float eps;
  float momentum;
} unetk_bn_finalize_args;

int unetk_version(void);
const char* unetk_last_error(void);
/* 0 when the current CUDA device is compute capability 10.x, UNETK_ERR_ARCH otherwise. */
int unetk_check_device(void);
/* cudaMemsetAsync on the caller's stream (zeroing statistics / gradient accumulators). */
int unetk_memset_zero(void* ptr, size_t bytes, void* stream);
/* Diagnostics (tools/timeline.py): a one-thread kernel that stores the GPU's %globaltimer (ns) into *slot when
 * everything enqueued before it on `stream` has finished.  Works inside CUDA-graph capture, so a replayed
 * step can be laid out kernel by kernel and stream by stream without a profiler. */
int unetk_debug_stamp(uint64_t* slot, void* stream);

/* ---- nn.Conv2d(Cin, Cout, 3, padding=1, bias=False): src/model.py:14,17 -------------------- */
/* y[n,h,w,co] = sum_{tap,ci} cat(src1,src2)[n,h+dy,w+dx,ci] * Wf[co][tap][ci]   (tcgen05 implicit
 * GEMM).  src2/c2 fold torch.cat([x2, x1], dim=1) of src/model.py:65 into the K loop (c2 = 0: none).
 * stats != NULL: adds per-channel sum / sum of squares of the fp32 accumulators into stats[2][Cout]
 * (training BatchNorm, src/model.py:15,18).  scale/shift != NULL: y = relu?(acc*scale+shift)
 * (eval-mode BatchNorm + ReLU folded in).  stats and scale are mutually exclusive.
 * h2, w2 (0 = h, w): extent of src2 [n,h2,w2,c2], aligned at the origin of the h x w grid and read as zero
 * beyond -- F.pad of the up-sampled tensor to the skip's size (src/model.py:57-61) for inputs whose
 * size is not a multiple of 16; the same pair on dgrad bounds dx2, on wgrad it is src2's extent. */
int unetk_conv3x3_fwd(const void* src1, int c1, const void* src2, int c2, const void* wf, void* y,
                      float* stats, const float* scale, const float* shift, int relu, int n, int h,
                      int w, int cout, int h2, int w2, void* stream);
/* autograd of the above w.r.t. its input: dx = conv(dy, flipped W^T).  Columns [0,c1) go to dx1,
 * [c1,c1+c2) to dx2 (the two concat sources); accK != 0 accumulates into dxK (skip connections
 * receive several gradients).  stats != NULL: stats[c2] += sum over pixels of dx2 (the bias
 * gradient of the ConvTranspose2d that produced src2; the caller zeroes it). */
int unetk_conv3x3_dgrad(const void* dy, int cout, const void* wd, void* dx1, int c1, int acc1,
                        void* dx2, int c2, int acc2, float* stats, const unetk_bn_ctx* bn, int n, int h,
                        int w, int h2, int w2, void* stream);
/* autograd w.r.t. the weight: dw[co][tap][ci] += sum_pix dy * x  (split-K, fp32 atomics; the
 * caller zeroes dw). */
int unetk_conv3x3_wgrad(const void* dy, const void* src1, int c1, const void* src2, int c2,
                        float* dw, int n, int h, int w, int cout, int h2, int w2, void* stream);

/* First encoder conv, Conv2d(n_channels<=4, 64, 3, padding=1, bias=False) on the caller's NCHW
 * fp32 image (src/model.py:86 / :168 -> :14); w is the reference OIHW fp32 weight; dw is written
 * in OIHW directly. */
int unetk_first_conv_fwd(const float* x, const float* w, void* y, float* stats, int n, int cin,
                         int h, int w_, int cout, void* stream);
int unetk_first_conv_wgrad(const float* x, const void* dy, float* dw, int n, int cin, int h, int w_,
                           int cout, void* stream);
/* Tensor-core route for the same gradient.  unetk_image_to_nhwc64 writes channels [0,8) of the NHWC
 * bf16 image x64[n,h,w,64] (the caller zeroes x64 ONCE: the padding channels are never written);
 * unetk_first_conv_wgrad_tc runs the 64->64 conv3x3 weight gradient on it (scratch: 64*9*64 floats,
 * zeroed by the call) and gathers dw in OIHW [64][cin][3][3] (overwritten). */
int unetk_image_to_nhwc64(const float* x, void* x64, int n, int cin, int h, int w_, void* stream);
/* The forward pass can take the same route: unetk_first_conv_pack writes the OIHW fp32 weight as
 * Wf[64][9][64] bf16 (zero beyond cin), then unetk_conv3x3_fwd(x64, 64, NULL, 0, wf, ...) is the layer. */
int unetk_first_conv_pack(const float* w, void* wf, int cin, void* stream);
int unetk_first_conv_wgrad_tc(const void* x64, const void* dy, float* scratch, float* dw, int n,
                              int cin, int h, int w_, int cout, void* stream);

/* ---- nn.ConvTranspose2d(Cin, Cin/2, kernel_size=2, stride=2): src/model.py:51 -------------- */
/* up[n,2h+ky,2w+kx,co] = bias[co] + sum_ci a[n,h,w,ci] * W[ci][co][ky][kx]  (GEMM + pixel-shuffle
 * store); n,h,w describe the INPUT grid. */
int unetk_convt2x2_fwd(const void* a, int cin, const void* wt, const float* bias, void* up, int n,
                       int h, int w, int cout, void* stream);
int unetk_convt2x2_dgrad(const void* dup, int cout, const void* wtd, void* da, int cin, int acc,
                         const unetk_bn_ctx* bn, int n, int h, int w, void* stream);
/* dw[co][tap][ci] += sum_pix dup[2*pix+tap, co] * a[pix, ci] (caller zeroes dw). */
int unetk_convt2x2_wgrad(const void* a, int cin, const void* dup, int cout, float* dw, int n, int h,
                         int w, void* stream);

/* ---- nn.BatchNorm2d + nn.ReLU(inplace=True): src/model.py:15-16,18-19 ---------------------- */
/* Training: batch statistics -> scale = gamma*invstd, shift = beta - mean*scale; updates
 * running_mean / running_var (momentum, unbiased variance) and num_batches_tracked (int64). */
int unetk_bn_finalize(const float* stats, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, int64_t* num_batches_tracked,
                      float* scale, float* shift, float* mean, float* invstd, int c, double count,
                      float eps, float momentum, void* stream);
/* Eval: scale/shift from the running statistics. */
int unetk_bn_fold(const float* gamma, const float* beta, const float* running_mean,
                  const float* running_var, float* scale, float* shift, int c, float eps,
                  void* stream);
/* a = relu(y*scale + shift) */
int unetk_bnrelu_apply(const void* y, const float* scale, const float* shift, void* a, int64_t npix,
                       int c, void* stream);
/* ---- nn.MaxPool2d(2): src/model.py:32, fused with the BatchNorm+ReLU that feeds it --------- */
/* apply != 0: a = relu(y*scale+shift) is written too.  pooled = max over 2x2, idx (uint8 per
 * element, may be NULL) = position 0..3 of the first maximum. */
/* The same two kernels with unetk_bn_finalize folded in (training forward: conv -> this). */
int unetk_bnrelu_apply_fin(const void* y, const unetk_bn_finalize_args* fin, void* a, int64_t npix, int c,
                           void* stream);
int unetk_bnrelu_pool_fin(const void* y, const unetk_bn_finalize_args* fin, void* a, void* pooled, void* idx,
                          int n, int h, int w, int c, void* stream);
int unetk_bnrelu_pool(const void* y, const float* scale, const float* shift, void* a, void* pooled,
                      void* idx, int n, int h, int w, int c, int apply, void* stream);
/* da[2x2 window] (+)= scatter of dpooled by idx; accumulate != 0 adds to the existing da; extra (may be
 * NULL, same shape as da) is a further gradient of the same activation added on the fly. */
int unetk_maxpool_bwd(const void* dpooled, const void* idx, void* da, int n, int h, int w, int c,
                      int accumulate, const void* extra, const unetk_bn_ctx* bn, void* stream);
/* BatchNorm+ReLU backward in two passes: sums[2][C] (caller zeroes) then dy, dgamma, dbeta. */
int unetk_bn_bwd_reduce(const void* da, const void* y, const float* scale, const float* shift,
                        const float* mean, const float* invstd, float* sums, int64_t npix, int c,
                        void* stream);
/* param_sums (NULL = sums): the sums that become dgamma / dbeta -- under synchronised BatchNorm `sums` holds
 * the cross-rank MEAN of the per-rank sums (dy then uses the global statistics) while the parameter
 * gradients stay this rank's own sums, to be averaged with all other gradients. */
int unetk_bn_bwd_apply(const void* da, const void* y, const float* scale, const float* shift,
                       const float* mean, const float* invstd, const float* sums, const float* param_sums,
                       void* dy, float* dgamma, float* dbeta, int64_t npix, int c, void* stream);

/* ---- OutConv = nn.Conv2d(64, n_out, 1) (+ torch.sigmoid): src/model.py:69-75,201,208 ------- */
int unetk_head_fwd(const void* a, const float* w, const float* b, float* out, int n, int h, int w_,
                   int c, int n_out, int sigmoid, void* stream);
/* The same on the RAW output y of the last conv3x3: the training-mode BatchNorm of that layer (batch statistics
 * finalized as in unetk_bnrelu_apply_fin, running statistics updated) + ReLU are applied on the fly, the activation
 * rounded to bf16 as the stand-alone pass would have stored it; the activated tensor is never written
 * (src/model.py:18-19 feeding :200-201 / :207-208). */
int unetk_head_fwd_bn(const void* y, const unetk_bn_finalize_args* fin, const float* w, const float* b, float* out,
                      int n, int h, int w_, int c, int n_out, int sigmoid, void* stream);
/* Output stage of one AnomalyUNet decoder for a training step, in ONE pass over y (csrc/head_fused.cu):
 * unetk_head_fwd_bn (+sigmoid) + this decoder's term of CombinedLoss (src/train_utils.py:30-44) + unetk_head_bwd
 * with the BatchNorm-backward sums.  which = 0: reconstruction decoder, squared error against target = the image
 * [n, n_out, h, w]; which = 1: segmentation decoder, focal loss against target = the mask [n, 1, h, w].
 * out (nullable): the sigmoid outputs.  da: gradient w.r.t. the activated tensor (bf16 NHWC); dw / db / bn_sums are
 * ACCUMULATED into (zero them first).  out3 = {total, recon, seg}: each launch writes its own term, the launch of a
 * step that finishes second also writes total = recon_weight * recon + seg_weight * seg.  work: zeroed once by the
 * caller (unetk_head_loss_workspace_bytes), shared by the two launches of a step, self-cleaning. */
size_t unetk_head_loss_workspace_bytes(void);
int unetk_head_loss_bwd(const void* y, const unetk_bn_finalize_args* fin, const float* w, const float* b,
                        const float* target, float* out, void* da, float* dw, float* db, float* bn_sums, void* work,
                        size_t work_bytes, float* out3, int which, int n, int h, int w_, int c, int n_out,
                        float recon_weight, float seg_weight, float alpha, float gamma, void* stream);
/* a == NULL (forward went through unetk_head_fwd_bn): the activation is recomputed from bn->y / scale / shift.
 * bn->sums != NULL: the first pass of that BatchNorm's backward (see unetk_bn_ctx) is fused in. */
int unetk_head_bwd(const void* a, const float* w, const float* dout, const float* out, void* da,
                   float* dw, float* db, int n, int h, int w_, int c, int n_out, int sigmoid,
                   const unetk_bn_ctx* bn, void* stream);

/* ---- CombinedLoss.forward / focal_loss: src/train_utils.py:10-44 --------------------------- */
/* out3 = {total, recon, seg}; d_recon / d_amap (may be NULL) = gradient of `total`. */
size_t unetk_combined_loss_workspace_bytes(void);
int unetk_combined_loss(const float* recon, const float* amap, const float* image,
                        const float* mask, float* d_recon, float* d_amap, float* out3, void* work,
                        size_t work_bytes, int64_t n_recon, int64_t n_seg, float recon_weight,
                        float seg_weight, float focal_alpha, float focal_gamma, void* stream);
int unetk_scale_by_scalar(float* y, const float* scalar, int64_t n, void* stream);

/* ---- SSIMLoss: src/train_utils.py:47-104 --------------------------------------------------- */
/* sums[n] = per-image sum of the SSIM map; maps (may be NULL; 5 planes of N*C*H*W floats, 3 if
 * need2 == 0) hold the derivative planes consumed by unetk_ssim_bwd. */
int unetk_ssim_fwd(const float* img1, const float* img2, float* sums, float* maps, int n, int c,
                   int h, int w, int window_size, int need2, void* stream);
/* out = 1 - mean(ssim_map): out[0] when size_average, else out[n] (one value per image). */
int unetk_ssim_finalize(const float* sums, float* out, int n, int c, int h, int w, int size_average,
                        void* stream);
int unetk_ssim_bwd(const float* img1, const float* img2, const float* maps, const float* gout,
                   float* grad, int n, int c, int h, int w, int window_size, int which,
                   int per_image, void* stream);

/* ---- SegmentationUNet pipeline ---------------------------------------- */
/* CombinedSegmentationLoss.forward (src/metrics.py:285-335): logits NCHW fp32 [n,c,h,w] (c <= 8), target
 * int64 [n,h,w], class_w [c] or NULL.  out4 = {total, ce, dice, focal} with
 *   ce    = F.cross_entropy(logits, target, weight=class_w, ignore_index)        (weighted mean)
 *   dice  = 1 - mean_{n,c} (2 I + 1e-8) / (U + 1e-8) on softmax(logits) vs one-hot(target)  (:233-261)
 *   focal = mean over ALL pixels of alpha (1 - pt)^gamma ce                       (:264-282)
 *   total = sum of the terms whose weight is > 0, like the reference.
 * Targets outside [0, c) belong to no class (dice) and are skipped by ce / focal like ignore_index.
 * is_prob != 0: `logits` already holds probabilities (stand-alone dice_loss); only dice is meaningful.
 * work: unetk_seg_loss_workspace_bytes(n, c) bytes, written by _fwd and read by _bwd.
 * _bwd: dlogits = (*gout, or 1 when gout is NULL) * d total / d logits. */
size_t unetk_seg_loss_workspace_bytes(int n, int c);
int unetk_seg_loss_fwd(const float* logits, const int64_t* target, const float* class_w, float* out4,
                       void* work, size_t work_bytes, int n, int c, int h, int w, int64_t ignore_index,
                       float ce_w, float dice_w, float focal_w, float focal_alpha, float focal_gamma,
                       int is_prob, void* stream);
int unetk_seg_loss_bwd(const float* logits, const int64_t* target, const float* class_w, const void* work,
                       const float* gout, float* dlogits, int n, int c, int h, int w, int64_t ignore_index,
                       float ce_w, float dice_w, float focal_w, float focal_alpha, float focal_gamma,
                       int is_prob, void* stream);
/* nn.Dropout2d applied in place to an NHWC bf16 activation: x[n,h,w,c] *= mask[n,c] (src/model.py:130,146;
 * the caller draws mask = bernoulli(1-p)/(1-p)). */
int unetk_channel_scale(void* x, const float* mask, int n, int h, int w, int c, void* stream);
/* nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) of the bilinear=True decoders
 * (src/model.py:47-49), NHWC bf16: out[n,2h,2w,c] from x[n,h,w,c]; _bwd: dx (+)= gradient w.r.t. x. */
int unetk_upsample2x_fwd(const void* x, void* out, int n, int h, int w, int c, void* stream);
int unetk_upsample2x_bwd(const void* dout, void* dx, int n, int h, int w, int c, int accumulate, void* stream);
/* torch.argmax(logits, dim=1): NCHW fp32 -> int64 [n,h,w], first maximum wins (test.py:98). */
int unetk_argmax_channels(const float* logits, int64_t* out, int n, int c, int h, int w, void* stream);

/* ---- compute_anomaly_score: src/utils.py:205-215 (method 0 = 'mse'/'ssim', 1 = 'l1') ------- */
int unetk_anomaly_score(const float* recon, const float* orig, float* out, int n, int c, int h,
                        int w, int method, void* stream);

/* ---- evaluation reductions of validate_epoch / calculate_pixel_metrics on the device ----------
 * src/train_utils.py:232-246 and src/utils.py:97-108: for every threshold t (1..8 of them, a HOST array),
 * counts[t*4 + {0 tn, 1 fp, 2 fn, 3 tp}] += number of pixels by (truth > truth_threshold, pred > thresholds[t]),
 * sklearn's confusion_matrix(y_true, y_pred).ravel() order, over the images i with include[i] != 0 (all images
 * when include is NULL; the reference keeps the anomalous images only).  pred / truth: fp32 [n, elems_per_image];
 * counts: int64 [n_thresholds, 4] on the device, ACCUMULATED into (zero it before the first batch). */
int unetk_pixel_confusion(const float* pred, const float* truth, const uint8_t* include, const float* thresholds,
                          int n_thresholds, float truth_threshold, int64_t* counts, int n, int64_t elems_per_image,
                          void* stream);
/* SegmentationMetrics.update (src/metrics.py:22-45): cm[t * c + p] += 1 (int64 [c, c] on the device, ACCUMULATED)
 * over the pixels with label t and prediction p both in [0, c) — sklearn's
 * confusion_matrix(target, pred, labels=range(c)) — where p = argmax over the class axis of NCHW fp32 `logits`
 * (first maximum wins), or `preds` int64 [n, elems] when logits is NULL (exactly one of the two).  With
 * has_ignore the pixels whose label equals ignore_index are dropped first; *kept (device int64, accumulated)
 * counts the remaining pixels (total_samples).  c <= 8. */
int unetk_class_confusion(const float* logits, const int64_t* preds, const int64_t* target, int n, int c,
                          int64_t elems_per_image, int64_t ignore_index, int has_ignore, int64_t* cm, int64_t* kept,
                          void* stream);
/* Exact order statistics of an fp32 device array (NaN orders last, as in numpy): out2[0] = k-th smallest
 * (0-based), out2[1] = (k+1)-th smallest (= out2[0] when k = n-1) — the two neighbours
 * np.percentile(all_scores, 95) interpolates between (src/train_utils.py:213).  x is not modified. */
size_t unetk_select_workspace_bytes(void);
int unetk_select_kth(const float* x, int64_t n, int64_t k, float* out2, void* workspace, size_t workspace_bytes,
                     void* stream);

/* ---- evaluation-time input transform on the device: src/dataset.py:141-152 --------------------
 * transforms.Resize((S, S)) -> ToTensor() -> Normalize(mean, std) of a PIL image, i.e. Pillow's separable,
 * antialiased bilinear resampling in 8-bit fixed point (22 fractional bits, the horizontal pass rounded to a
 * byte before the vertical one), then (v / 255 - mean[c]) / std[c] in fp32.  Bit-exact against
 * Pillow + torchvision.
 * unetk_resize_ksize / unetk_resize_coeffs are HOST functions (no device needed): the tap table of one axis,
 * bounds int32 [out, 2] = (first source index, tap count), coeffs int32 [out, ksize].
 * unetk_resize_u8: src u8 [n, h, w, c] (c = 3 RGB or 1 mask) -> tmp u8 [n, h, ow, c] -> out_u8 [n, oh, ow, c]
 * and / or out_f32 [n, c, oh, ow] (either may be NULL); the tap tables are DEVICE copies; mean / std are HOST
 * arrays of c floats (NULL: 0 / 1, i.e. ToTensor only, the mask transform). */
int unetk_resize_ksize(int in_size, int out_size);
int unetk_resize_coeffs(int in_size, int out_size, int32_t* bounds, int32_t* coeffs);
int unetk_resize_u8(const uint8_t* src, uint8_t* tmp, uint8_t* out_u8, float* out_f32, int n, int h, int w, int c,
                    int oh, int ow, const int32_t* hbounds, const int32_t* hcoeffs, int hksize,
                    const int32_t* vbounds, const int32_t* vcoeffs, int vksize, const float* mean, const float* std,
                    void* stream);

/* ---- training augmentations on the device: src/dataset.py:131-138 -----------------------------
 * RandomHorizontalFlip -> RandomRotation -> ColorJitter of the resized uint8 image with HOST-drawn parameters
 * (torchvision's draws from torch's generator), restated from torchvision 0.26 + Pillow 12.2; bit-exact.
 * geom: DEVICE int32 [n, 8] = (mode, flip, a0, a1, a2, a3, a4, a5): mode 0 identity, 1 Pillow's affine_fixed
 * nearest-neighbour rotation (16.16 fixed point: source x = (a2 + a1 y + a0 x) >> 16, source y =
 * (a5 + a4 y + a3 x) >> 16, outside the image -> 0), 2 / 3 / 4 the exact 180 / 90 / 270 degree transposes;
 * the flip is applied to the source before the rotation.  src != dst, u8 [n, h, w, c]. */
int unetk_flip_rotate_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, int c, const int32_t* geom,
                         void* stream);
/* ColorJitter in place on u8 [n, h, w, 3]: four stages, ops DEVICE int32 [n, 4] per stage 0 brightness,
 * 1 contrast, 2 saturation (Image.blend of the black / mean-luma / luma image with the image, factors[n, stage]
 * = alpha), 3 hue (RGB -> HSV -> RGB, factors = the uint8 amount added to H), other values skip the stage.
 * gray_sums: DEVICE workspace of 4 * n uint64. */
int unetk_color_jitter_u8(uint8_t* img, int n, int h, int w, const int32_t* ops, const float* factors,
                          uint64_t* gray_sums, void* stream);
/* ToTensor + Normalize: u8 [n, h, w, c] -> fp32 [n, c, h, w] = (v / 255 - mean[c]) / std[c] (HOST arrays or NULL). */
int unetk_u8_to_tensor(const uint8_t* hwc, float* chw, int n, int h, int w, int c, const float* mean,
                       const float* std, void* stream);

/* Label masks of the segmentation datasets: transforms.Resize(size, interpolation=NEAREST) -> MaskToTensor
 * (src/gear_dataset.py:255-258, src/kolektorsdd_dataset.py:156-159), i.e. Pillow's nearest-neighbour resize.
 * unetk_resize_nearest_table is a HOST function: the source index of every output index of one axis (int32
 * [out]).  unetk_resize_nearest_u8: src u8 [n, h, w] -> out_i64 [n, oh, ow] (the .long() label map) and / or
 * out_u8 (either may be NULL); xtab [ow] / ytab [oh] are DEVICE copies of the tables. */
int unetk_resize_nearest_table(int in_size, int out_size, int32_t* table);
int unetk_resize_nearest_u8(const uint8_t* src, int64_t* out_i64, uint8_t* out_u8, int n, int h, int w, int oh, int ow,
                            const int32_t* xtab, const int32_t* ytab, void* stream);

/* ---- get_optimizer -> torch.optim.{Adam,AdamW,SGD}: src/train_utils.py:263-272 ------------- */
/* mode 0 Adam (L2 decay in the gradient), 1 AdamW, 2 SGD+momentum (beta1 = momentum, v unused). */
int unetk_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                    int mode, void* stream);

/* ---- layout plumbing ------------------------------------------------------------------------ */
/* One launch over a table of 64-byte descriptors (see tiaozhanbei-unet_b200/packing.py):
 * out[d0][d1][d2] = scale * in[i0*s0 + f(i1)*s1 + i2*s2]; out is bf16 (out_bf16 != 0) or fp32. */
int unetk_permute_batched(const void* descs, int n_desc, int n_chunks, const float* in, void* out,
                          int out_bf16, float scale, void* stream);
int unetk_permute_chunk_elems(void);
/* Tiled fast paths for the conv3x3 layers (64-byte descriptors {src_off, dst_off, dst2_off, cout, cin,
 * block_begin}, see tiaozhanbei-unet_b200/packing.py): OIHW fp32 -> Wf and Wd bf16 in one launch
 * (one block per 32x32 (co, ci) tile), and staged dW[co][tap][ci] -> OIHW fp32 * scale (one block per co). */
int unetk_pack_conv_weights(const void* descs, int n_desc, int n_blocks, const float* flat_p, void* wpack,
                            void* stream);
int unetk_unpack_conv_grads(const void* descs, int n_desc, int n_blocks, int max_cin, const float* gstage,
                            float* flat_g, float scale, void* stream);
/* The tail of the training step (src/train_utils.py:131-135: `optimizer.zero_grad()`, `backward()` leaving `.grad`,
 * `optimizer.step()`) for a group of tensors in ONE pass: staged gradient * hyper[8] -> g (`.grad`, reference
 * layout), the optimizer update of unetk_adam_step (same arithmetic, bit for bit) on p / m / v, the new weights ->
 * their bf16 kernel layouts in wpack, and zeroes over the staging range it has read.  descs: 64-byte records
 * {int64 off, wf_off, wd_off; int32 kind (0 plain, 1 conv3x3 OIHW, 2 ConvTranspose2d [ci][co][2][2]), cout, cin,
 * block_begin, pad[6]} (tiaozhanbei-unet_b200/packing.py: UpdateTable); conv / transposed-conv tensors need
 * channel counts that are multiples of 32.  hyper: DEVICE float[9] = {lr, beta1, beta2, eps, weight_decay,
 * 1 - beta1^step, sqrt(1 - beta2^step), first step (0 / 1), gradient scale} -- device memory so that the launch can
 * be replayed from a CUDA graph with a new step count; unetk_set_floats (values passed by value in the launch)
 * fills it.  n_blocks = work units of the table (32 x 32 channel tiles, 2048-element chunks); max_blocks > 0 bounds
 * the grid (the blocks then walk the units), for launches that run beside other streams' kernels. */
int unetk_fused_update(const void* descs, int n_desc, int n_blocks, int max_blocks, float* p, float* g, float* m, float* v,
                       float* gstage, void* wpack, const float* hyper, int mode, void* stream);
int unetk_update_plain_chunk_elems(void);
int unetk_set_floats(float* dst, const float* host_vals, int n, void* stream);
int unetk_nchw_to_nhwc(const float* in, void* out, int n, int c, int h, int w, void* stream);
int unetk_nhwc_to_nchw(const void* in, float* out, int n, int c, int h, int w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNETK_H_ */
