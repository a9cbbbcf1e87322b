// extern "C" surface declared in include/unetk.h; forwards 1:1 to the C++ launchers.
#include "../../include/unetk.h"

#include "common.cuh"
#include "kernels.h"

namespace unetk {
const char* last_error();
}

#define S(stream) static_cast<cudaStream_t>(stream)

extern "C" {

int unetk_version(void) { return UNETK_VERSION; }
const char* unetk_last_error(void) { return unetk::last_error(); }
int unetk_check_device(void) { return unetk::check_device_sm100(); }

int unetk_memset_zero(void* ptr, size_t bytes, void* stream) {
  if (ptr == nullptr && bytes != 0) {
    unetk::set_last_error("memset_zero: null pointer");
    return UNETK_ERR_ARG;
  }
  cudaError_t e = cudaMemsetAsync(ptr, 0, bytes, S(stream));
  if (e != cudaSuccess) {
    unetk::set_last_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e));
    return UNETK_ERR_CUDA;
  }
  return UNETK_OK;
}

__global__ void debug_stamp_kernel(uint64_t* slot) {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}
int unetk_debug_stamp(uint64_t* slot, void* stream) {
  if (slot == nullptr) {
    unetk::set_last_error("debug_stamp: null pointer");
    return UNETK_ERR_ARG;
  }
  debug_stamp_kernel<<<1, 1, 0, S(stream)>>>(slot);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int unetk_conv3x3_fwd(const void* src1, int c1, const void* src2, int c2, const void* wf, void* y,
                      float* stats, const float* scale, const float* shift, int relu, int n, int h,
                      int w, int cout, int h2, int w2, void* stream) {
  return unetk::conv3x3_fwd(src1, c1, src2, c2, wf, y, stats, scale, shift, relu, n, h, w, cout, h2, w2, S(stream));
}
static const unetk::BnReduceCtx* B(const unetk_bn_ctx* bn) { return reinterpret_cast<const unetk::BnReduceCtx*>(bn); }
static_assert(sizeof(unetk_bn_ctx) == sizeof(unetk::BnReduceCtx), "unetk_bn_ctx mirrors BnReduceCtx");

int unetk_conv3x3_dgrad(const void* dy, int cout, const void* wd, void* dx1, int c1, int acc1,
                        void* dx2, int c2, int acc2, float* stats, const unetk_bn_ctx* bn, int n, int h, int w,
                        int h2, int w2, void* stream) {
  return unetk::conv3x3_dgrad(dy, cout, wd, dx1, c1, acc1, dx2, c2, acc2, stats, B(bn), n, h, w, h2, w2, S(stream));
}
int unetk_conv3x3_wgrad(const void* dy, const void* src1, int c1, const void* src2, int c2,
                        float* dw, int n, int h, int w, int cout, int h2, int w2, void* stream) {
  return unetk::conv3x3_wgrad(dy, src1, c1, src2, c2, dw, n, h, w, cout, h2, w2, S(stream));
}
int unetk_first_conv_fwd(const float* x, const float* w, void* y, float* stats, int n, int cin,
                         int h, int w_, int cout, void* stream) {
  return unetk::first_conv_fwd(x, w, y, stats, n, cin, h, w_, cout, S(stream));
}
int unetk_image_to_nhwc64(const float* x, void* x64, int n, int cin, int h, int w_, void* stream) {
  return unetk::image_to_nhwc64(x, x64, n, cin, h, w_, S(stream));
}
int unetk_first_conv_pack(const float* w, void* wf, int cin, void* stream) {
  return unetk::first_conv_pack(w, wf, cin, S(stream));
}
int unetk_first_conv_wgrad_tc(const void* x64, const void* dy, float* scratch, float* dw, int n,
                              int cin, int h, int w_, int cout, void* stream) {
  return unetk::first_conv_wgrad_tc(x64, dy, scratch, dw, n, cin, h, w_, cout, S(stream));
}
int unetk_first_conv_wgrad(const float* x, const void* dy, float* dw, int n, int cin, int h, int w_,
                           int cout, void* stream) {
  return unetk::first_conv_wgrad(x, dy, dw, n, cin, h, w_, cout, S(stream));
}
int unetk_convt2x2_fwd(const void* a, int cin, const void* wt, const float* bias, void* up, int n,
                       int h, int w, int cout, void* stream) {
  return unetk::convt2x2_fwd(a, cin, wt, bias, up, n, h, w, cout, S(stream));
}
int unetk_convt2x2_dgrad(const void* dup, int cout, const void* wtd, void* da, int cin, int acc,
                         const unetk_bn_ctx* bn, int n, int h, int w, void* stream) {
  return unetk::convt2x2_dgrad(dup, cout, wtd, da, cin, acc, B(bn), n, h, w, S(stream));
}
int unetk_convt2x2_wgrad(const void* a, int cin, const void* dup, int cout, float* dw, int n, int h,
                         int w, void* stream) {
  return unetk::convt2x2_wgrad(a, cin, dup, cout, dw, n, h, w, S(stream));
}
int unetk_bn_finalize(const float* stats, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, int64_t* num_batches_tracked,
                      float* scale, float* shift, float* mean, float* invstd, int c, double count,
                      float eps, float momentum, void* stream) {
  return unetk::bn_finalize(stats, gamma, beta, running_mean, running_var,
                            reinterpret_cast<long long*>(num_batches_tracked), scale, shift, mean, invstd, c, count,
                            eps, momentum, S(stream));
}
int unetk_bn_fold(const float* gamma, const float* beta, const float* running_mean,
                  const float* running_var, float* scale, float* shift, int c, float eps,
                  void* stream) {
  return unetk::bn_fold(gamma, beta, running_mean, running_var, scale, shift, c, eps, S(stream));
}
int unetk_bnrelu_apply(const void* y, const float* scale, const float* shift, void* a, int64_t npix,
                       int c, void* stream) {
  return unetk::bnrelu_apply(y, scale, shift, a, static_cast<long>(npix), c, S(stream));
}
static const unetk::BnFinalizeArgs* F(const unetk_bn_finalize_args* f) { return reinterpret_cast<const unetk::BnFinalizeArgs*>(f); }
static_assert(sizeof(unetk_bn_finalize_args) == sizeof(unetk::BnFinalizeArgs), "unetk_bn_finalize_args mirrors BnFinalizeArgs");
int unetk_bnrelu_apply_fin(const void* y, const unetk_bn_finalize_args* fin, void* a, int64_t npix, int c, void* stream) {
  return unetk::bnrelu_apply_fin(y, F(fin), a, static_cast<long>(npix), c, S(stream));
}
int unetk_bnrelu_pool_fin(const void* y, const unetk_bn_finalize_args* fin, void* a, void* pooled, void* idx, int n,
                          int h, int w, int c, void* stream) {
  return unetk::bnrelu_pool_fin(y, F(fin), a, pooled, idx, n, h, w, c, S(stream));
}
int unetk_bnrelu_pool(const void* y, const float* scale, const float* shift, void* a, void* pooled,
                      void* idx, int n, int h, int w, int c, int apply, void* stream) {
  return unetk::bnrelu_pool(y, scale, shift, a, pooled, idx, n, h, w, c, apply, S(stream));
}
int unetk_maxpool_bwd(const void* dpooled, const void* idx, void* da, int n, int h, int w, int c,
                      int accumulate, const void* extra, const unetk_bn_ctx* bn, void* stream) {
  return unetk::maxpool_bwd(dpooled, idx, da, n, h, w, c, accumulate, extra, B(bn), S(stream));
}
int unetk_bn_bwd_reduce(const void* da, const void* y, const float* scale, const float* shift,
                        const float* mean, const float* invstd, float* sums, int64_t npix, int c,
                        void* stream) {
  return unetk::bn_bwd_reduce(da, y, scale, shift, mean, invstd, sums, static_cast<long>(npix), c, S(stream));
}
int unetk_bn_bwd_apply(const void* da, const void* y, const float* scale, const float* shift,
                       const float* mean, const float* invstd, const float* sums, const float* param_sums,
                       void* dy, float* dgamma, float* dbeta, int64_t npix, int c, void* stream) {
  return unetk::bn_bwd_apply(da, y, scale, shift, mean, invstd, sums, param_sums, dy, dgamma, dbeta,
                             static_cast<long>(npix), c, S(stream));
}
int unetk_head_fwd(const void* a, const float* w, const float* b, float* out, int n, int h, int w_,
                   int c, int n_out, int sigmoid, void* stream) {
  return unetk::head_fwd(a, w, b, out, n, h, w_, c, n_out, sigmoid, S(stream));
}
int unetk_head_fwd_bn(const void* y, const unetk_bn_finalize_args* fin, const float* w, const float* b, float* out,
                      int n, int h, int w_, int c, int n_out, int sigmoid, void* stream) {
  return unetk::head_fwd_bn(y, reinterpret_cast<const unetk::BnFinalizeArgs*>(fin), w, b, out, n, h, w_, c, n_out, sigmoid, S(stream));
}
size_t unetk_head_loss_workspace_bytes(void) { return unetk::head_loss_workspace_bytes(); }
int unetk_head_loss_bwd(const void* y, const unetk_bn_finalize_args* fin, const float* w, const float* b,
                        const float* target, float* out, void* da, float* dw, float* db, float* bn_sums, void* work,
                        size_t work_bytes, float* out3, int which, int n, int h, int w_, int c, int n_out,
                        float recon_weight, float seg_weight, float alpha, float gamma, void* stream) {
  return unetk::head_loss_bwd(y, reinterpret_cast<const unetk::BnFinalizeArgs*>(fin), w, b, target, out, da, dw, db, bn_sums,
                              work, work_bytes, out3, which, n, h, w_, c, n_out, recon_weight, seg_weight, alpha, gamma, S(stream));
}
int unetk_head_bwd(const void* a, const float* w, const float* dout, const float* out, void* da,
                   float* dw, float* db, int n, int h, int w_, int c, int n_out, int sigmoid,
                   const unetk_bn_ctx* bn, void* stream) {
  return unetk::head_bwd(a, w, dout, out, da, dw, db, n, h, w_, c, n_out, sigmoid, B(bn), S(stream));
}
size_t unetk_combined_loss_workspace_bytes(void) { return unetk::combined_loss_workspace_bytes(); }
int unetk_combined_loss(const float* recon, const float* amap, const float* image,
                        const float* mask, float* d_recon, float* d_amap, float* out3, void* work,
                        size_t work_bytes, int64_t n_recon, int64_t n_seg, float recon_weight,
                        float seg_weight, float focal_alpha, float focal_gamma, void* stream) {
  return unetk::combined_loss(recon, amap, image, mask, d_recon, d_amap, out3, work, work_bytes,
                              static_cast<long>(n_recon), static_cast<long>(n_seg), recon_weight, seg_weight, focal_alpha,
                              focal_gamma, S(stream));
}
int unetk_scale_by_scalar(float* y, const float* scalar, int64_t n, void* stream) {
  return unetk::scale_by_scalar(y, scalar, static_cast<long>(n), S(stream));
}
int unetk_ssim_fwd(const float* img1, const float* img2, float* sums, float* maps, int n, int c,
                   int h, int w, int window_size, int need2, void* stream) {
  return unetk::ssim_fwd(img1, img2, sums, maps, n, c, h, w, window_size, need2, S(stream));
}
int unetk_ssim_finalize(const float* sums, float* out, int n, int c, int h, int w, int size_average,
                        void* stream) {
  return unetk::ssim_finalize(sums, out, n, c, h, w, size_average, S(stream));
}
int unetk_ssim_bwd(const float* img1, const float* img2, const float* maps, const float* gout,
                   float* grad, int n, int c, int h, int w, int window_size, int which,
                   int per_image, void* stream) {
  return unetk::ssim_bwd(img1, img2, maps, gout, grad, n, c, h, w, window_size, which, per_image, S(stream));
}
size_t unetk_seg_loss_workspace_bytes(int n, int c) { return unetk::seg_loss_workspace_bytes(n, c); }
int unetk_seg_loss_fwd(const float* logits, const int64_t* target, const float* class_w, float* out4,
                       void* work, size_t work_bytes, int n, int c, int h, int w, int64_t ignore_index,
                       float ce_w, float dice_w, float focal_w, float focal_alpha, float focal_gamma,
                       int is_prob, void* stream) {
  return unetk::seg_loss_fwd(logits, reinterpret_cast<const long long*>(target), class_w, out4, work, work_bytes, n, c, h, w,
                             ignore_index, ce_w, dice_w, focal_w, focal_alpha, focal_gamma, is_prob, S(stream));
}
int unetk_seg_loss_bwd(const float* logits, const int64_t* target, const float* class_w, const void* work,
                       const float* gout, float* dlogits, int n, int c, int h, int w, int64_t ignore_index,
                       float ce_w, float dice_w, float focal_w, float focal_alpha, float focal_gamma,
                       int is_prob, void* stream) {
  return unetk::seg_loss_bwd(logits, reinterpret_cast<const long long*>(target), class_w, work, gout, dlogits, n, c, h, w,
                             ignore_index, ce_w, dice_w, focal_w, focal_alpha, focal_gamma, is_prob, S(stream));
}
int unetk_channel_scale(void* x, const float* mask, int n, int h, int w, int c, void* stream) {
  return unetk::channel_scale(x, mask, n, h, w, c, S(stream));
}
int unetk_upsample2x_fwd(const void* x, void* out, int n, int h, int w, int c, void* stream) {
  return unetk::upsample2x_fwd(x, out, n, h, w, c, S(stream));
}
int unetk_upsample2x_bwd(const void* dout, void* dx, int n, int h, int w, int c, int accumulate, void* stream) {
  return unetk::upsample2x_bwd(dout, dx, n, h, w, c, accumulate, S(stream));
}
int unetk_argmax_channels(const float* logits, int64_t* out, int n, int c, int h, int w, void* stream) {
  return unetk::argmax_channels(logits, reinterpret_cast<long long*>(out), n, c, h, w, S(stream));
}
int unetk_anomaly_score(const float* recon, const float* orig, float* out, int n, int c, int h,
                        int w, int method, void* stream) {
  return unetk::anomaly_score(recon, orig, out, n, c, h, w, method, S(stream));
}
int unetk_pixel_confusion(const float* pred, const float* truth, const uint8_t* include, const float* thresholds,
                          int n_thresholds, float truth_threshold, int64_t* counts, int n, int64_t elems_per_image,
                          void* stream) {
  return unetk::pixel_confusion(pred, truth, include, thresholds, n_thresholds, truth_threshold,
                                reinterpret_cast<long long*>(counts), n, static_cast<long>(elems_per_image), S(stream));
}
int unetk_class_confusion(const float* logits, const int64_t* preds, const int64_t* target, int n, int c,
                          int64_t elems_per_image, int64_t ignore_index, int has_ignore, int64_t* cm, int64_t* kept,
                          void* stream) {
  return unetk::class_confusion(logits, reinterpret_cast<const long long*>(preds),
                                reinterpret_cast<const long long*>(target), n, c, static_cast<long>(elems_per_image),
                                static_cast<long long>(ignore_index), has_ignore, reinterpret_cast<long long*>(cm),
                                reinterpret_cast<long long*>(kept), S(stream));
}
size_t unetk_select_workspace_bytes(void) { return unetk::select_workspace_bytes(); }
int unetk_select_kth(const float* x, int64_t n, int64_t k, float* out2, void* workspace, size_t workspace_bytes,
                     void* stream) {
  return unetk::select_kth(x, static_cast<long>(n), static_cast<long>(k), out2, workspace, workspace_bytes, S(stream));
}
int unetk_resize_ksize(int in_size, int out_size) { return unetk::resize_ksize(in_size, out_size); }
int unetk_resize_coeffs(int in_size, int out_size, int32_t* bounds, int32_t* coeffs) {
  return unetk::resize_coeffs(in_size, out_size, bounds, coeffs);
}
int unetk_resize_u8(const uint8_t* src, uint8_t* tmp, uint8_t* out_u8, float* out_f32, int n, int h, int w, int c,
                    int oh, int ow, const int32_t* hbounds, const int32_t* hcoeffs, int hksize,
                    const int32_t* vbounds, const int32_t* vcoeffs, int vksize, const float* mean, const float* std,
                    void* stream) {
  return unetk::resize_u8(src, tmp, out_u8, out_f32, n, h, w, c, oh, ow, hbounds, hcoeffs, hksize, vbounds, vcoeffs,
                          vksize, mean, std, S(stream));
}
int unetk_flip_rotate_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, int c, const int32_t* geom,
                         void* stream) {
  return unetk::flip_rotate_u8(src, dst, n, h, w, c, geom, S(stream));
}
int unetk_color_jitter_u8(uint8_t* img, int n, int h, int w, const int32_t* ops, const float* factors,
                          uint64_t* gray_sums, void* stream) {
  return unetk::color_jitter_u8(img, n, h, w, ops, factors, reinterpret_cast<unsigned long long*>(gray_sums),
                                S(stream));
}
int unetk_u8_to_tensor(const uint8_t* hwc, float* chw, int n, int h, int w, int c, const float* mean,
                       const float* std, void* stream) {
  return unetk::u8_to_tensor(hwc, chw, n, h, w, c, mean, std, S(stream));
}
int unetk_resize_nearest_table(int in_size, int out_size, int32_t* table) {
  return unetk::resize_nearest_table(in_size, out_size, table);
}
int unetk_resize_nearest_u8(const uint8_t* src, int64_t* out_i64, uint8_t* out_u8, int n, int h, int w, int oh, int ow,
                            const int32_t* xtab, const int32_t* ytab, void* stream) {
  return unetk::resize_nearest_u8(src, reinterpret_cast<long long*>(out_i64), out_u8, n, h, w, oh, ow, xtab, ytab,
                                  S(stream));
}
int unetk_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                    int mode, void* stream) {
  return unetk::adam_step(p, g, m, v, static_cast<long>(n), lr, beta1, beta2, eps, weight_decay,
                          static_cast<long>(step), grad_scale, mode, S(stream));
}
int unetk_permute_batched(const void* descs, int n_desc, int n_chunks, const float* in, void* out,
                          int out_bf16, float scale, void* stream) {
  return unetk::permute_batched(descs, n_desc, n_chunks, in, out, out_bf16, scale, S(stream));
}
int unetk_pack_conv_weights(const void* descs, int n_desc, int n_blocks, const float* flat_p, void* wpack,
                            void* stream) {
  return unetk::pack_conv_weights(descs, n_desc, n_blocks, flat_p, wpack, S(stream));
}
int unetk_unpack_conv_grads(const void* descs, int n_desc, int n_blocks, int max_cin, const float* gstage,
                            float* flat_g, float scale, void* stream) {
  return unetk::unpack_conv_grads(descs, n_desc, n_blocks, max_cin, gstage, flat_g, scale, S(stream));
}
int unetk_permute_chunk_elems(void) { return unetk::permute_chunk_elems(); }
int unetk_fused_update(const void* descs, int n_desc, int n_blocks, int max_blocks, float* p, float* g, float* m, float* v,
                       float* gstage, void* wpack, const float* hyper, int mode, void* stream) {
  return unetk::fused_update(descs, n_desc, n_blocks, max_blocks, p, g, m, v, gstage, wpack, hyper, mode, S(stream));
}
int unetk_update_plain_chunk_elems(void) { return unetk::update_plain_chunk_elems(); }
int unetk_set_floats(float* dst, const float* host_vals, int n, void* stream) {
  return unetk::set_floats(dst, host_vals, n, S(stream));
}
int unetk_nchw_to_nhwc(const float* in, void* out, int n, int c, int h, int w, void* stream) {
  return unetk::nchw_to_nhwc(in, out, n, c, h, w, S(stream));
}
int unetk_nhwc_to_nchw(const void* in, float* out, int n, int c, int h, int w, void* stream) {
  return unetk::nhwc_to_nchw(in, out, n, c, h, w, S(stream));
}

}  // extern "C"
