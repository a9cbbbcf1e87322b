// Internal C++ launch interface of the kernel library (one function per fused op).
// The public C ABI in include/unetk.h forwards to these 1:1.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace unetk {

// Optional fused first pass of a BatchNorm+ReLU backward (same layout as unetk_bn_ctx in unetk.h):
// the kernel that finishes a gradient dA also accumulates sums[0..C) += sum g, sums[C..2C) += sum g*xhat
// with g = dA * [y*scale+shift > 0], xhat = (y-mean)*invstd, for the layer whose raw output is y.
struct BnReduceCtx {
  const void* y;
  const float* scale;
  const float* shift;
  const float* mean;
  const float* invstd;
  float* sums;
};

// Training-mode BatchNorm finalize folded into the kernel that applies it (same layout as
// unetk_bn_finalize_args in unetk.h).
struct BnFinalizeArgs {
  const float* stats;        // [2][C] sum, sum of squares (conv epilogue)
  const float* gamma;
  const float* beta;
  float* running_mean;       // nullable pair
  float* running_var;
  long long* num_batches_tracked;   // nullable
  float* scale;              // outputs [C] each, kept for the backward pass
  float* shift;
  float* mean;
  float* invstd;
  double count;              // N*H*W (x world size with synchronised statistics): exact beyond 2^24
  double inv_count;          // 1 / count: written by the library at launch, callers leave it 0
  float unbias;              // count / (count - 1) for running_var's unbiased variance: likewise written at launch
  float eps;
  float momentum;
};

// inv_count / unbias are derived at launch from count: callers only fill count
inline BnFinalizeArgs with_reciprocals(BnFinalizeArgs f) {
  f.inv_count = 1.0 / f.count;
  f.unbias = f.count > 1.0 ? static_cast<float>(f.count / (f.count - 1.0)) : 1.f;
  return f;
}

#ifdef __CUDACC__
// (scale, shift) of channel c from the batch sums, exactly as bn_finalize_kernel (elementwise.cu) computes
// them: E[x^2] - E[x]^2 in double from the fp32 sums with the host's reciprocal of the count (two DMULs and a
// DFMA instead of two double divisions: every thread of the consuming kernel runs this).  `writer` threads
// also store scale / shift / mean / invstd for the backward pass and update the running statistics
// (num_batches_tracked is the caller's: one thread only).
__device__ __forceinline__ void bn_finalize1(const BnFinalizeArgs& f, int C, int c, bool writer, float& sc, float& sh,
                                             float* mean_out = nullptr, float* invstd_out = nullptr) {
  const double mean_d = static_cast<double>(f.stats[c]) * f.inv_count;
  const float mean = static_cast<float>(mean_d);
  float var = static_cast<float>(static_cast<double>(f.stats[C + c]) * f.inv_count - mean_d * mean_d);
  var = fmaxf(var, 0.f);
  const float invstd = rsqrtf(var + f.eps);
  sc = f.gamma[c] * invstd;
  sh = f.beta[c] - mean * sc;
  if (mean_out != nullptr) *mean_out = mean;
  if (invstd_out != nullptr) *invstd_out = invstd;
  if (writer) {
    f.scale[c] = sc;
    f.shift[c] = sh;
    f.mean[c] = mean;
    f.invstd[c] = invstd;
    if (f.running_mean != nullptr) {
      const float unbiased = var * f.unbias;
      f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * mean;
      f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * unbiased;
    }
  }
}
#endif

// igemm.cu — K-major implicit GEMMs (tcgen05)
int conv3x3_fwd(const void* src1, int c1, const void* src2, int c2, const void* wpk, void* y,
                float* stats, const float* scale, const float* shift, int relu, int n, int h, int w,
                int cout, int h2, int w2, cudaStream_t stream);
int conv3x3_dgrad(const void* dy, int cout, const void* wdpk, void* dx1, int c1, int acc1, void* dx2,
                  int c2, int acc2, float* stats, const BnReduceCtx* bn, int n, int h, int w, int h2, int w2,
                  cudaStream_t stream);
int convt2x2_fwd(const void* a, int cin, const void* wpk, const float* bias, void* up, int n, int h,
                 int w, int cout, cudaStream_t stream);
int convt2x2_dgrad(const void* dup, int cout, const void* wdpk, void* da, int cin, int acc,
                   const BnReduceCtx* bn, int n, int h, int w, cudaStream_t stream);

// wgrad.cu — MN-major split-K GEMMs (tcgen05)
int conv3x3_wgrad(const void* dy, const void* src1, int c1, const void* src2, int c2, float* dw,
                  int n, int h, int w, int cout, int h2, int w2, cudaStream_t stream);
int convt2x2_wgrad(const void* a, int cin, const void* dup, int cout, float* dw, int n, int h, int w,
                   cudaStream_t stream);

// first_conv.cu
int first_conv_fwd(const float* x, const float* w, void* y, float* stats, int N, int Cin, int H, int W,
                   int cout, cudaStream_t s);
int image_to_nhwc64(const float* x, void* x64, int N, int Cin, int H, int W, cudaStream_t s);
int first_conv_pack(const float* w, void* wf, int Cin, cudaStream_t s);
int first_conv_wgrad_tc(const void* x64, const void* dy, float* scratch, float* dw, int N, int Cin, int H, int W,
                        int cout, cudaStream_t s);
int first_conv_wgrad(const float* x, const void* dy, float* dw, int N, int Cin, int H, int W, int cout,
                     cudaStream_t s);

// elementwise.cu
int bn_finalize(const float* stats, const float* gamma, const float* beta, float* running_mean,
                float* running_var, long long* nbt, float* scale, float* shift, float* mean,
                float* invstd, int C, double count, float eps, float momentum, cudaStream_t s);
int bn_fold(const float* gamma, const float* beta, const float* rm, const float* rv, float* scale,
            float* shift, int C, float eps, cudaStream_t s);
int bnrelu_apply(const void* y, const float* scale, const float* shift, void* a, long npix, int C,
                 cudaStream_t s);
int bnrelu_apply_fin(const void* y, const BnFinalizeArgs* fin, void* a, long npix, int C, cudaStream_t s);
int bnrelu_pool_fin(const void* y, const BnFinalizeArgs* fin, void* a, void* pooled, void* idx, int N, int H, int W,
                    int C, cudaStream_t s);
int bnrelu_pool(const void* y, const float* scale, const float* shift, void* a, void* pooled,
                void* idx, int N, int H, int W, int C, int apply, cudaStream_t s);
int maxpool_bwd(const void* dp, const void* idx, void* da, int N, int H, int W, int C, int accumulate,
                const void* extra, const BnReduceCtx* bn, cudaStream_t s);
int bn_bwd_reduce(const void* da, const void* y, const float* scale, const float* shift,
                  const float* mean, const float* invstd, float* sums, long npix, int C,
                  cudaStream_t s);
int bn_bwd_apply(const void* da, const void* y, const float* scale, const float* shift,
                 const float* mean, const float* invstd, const float* sums, const float* param_sums, void* dy,
                 float* dgamma, float* dbeta, long npix, int C, cudaStream_t s);
int nchw_to_nhwc(const float* in, void* out, int N, int C, int H, int W, cudaStream_t s);
int nhwc_to_nchw(const void* in, float* out, int N, int C, int H, int W, cudaStream_t s);

// head_loss.cu
size_t head_loss_workspace_bytes();
int head_loss_bwd(const void* y, const BnFinalizeArgs* fin, const float* w, const float* b, const float* target,
                  float* out, void* da, float* dw, float* db, float* bn_sums, void* work, size_t work_bytes,
                  float* out3, int which, int N, int H, int W, int C, int n_out, float recon_weight, float seg_weight,
                  float alpha, float gamma, cudaStream_t s);
int head_fwd_bn(const void* y, const BnFinalizeArgs* fin, const float* w, const float* b, float* out, int N, int H,
                int W, int C, int n_out, int sigmoid, cudaStream_t s);
int head_fwd(const void* a, const float* w, const float* b, float* out, int N, int H, int W, int C,
             int n_out, int sigmoid, cudaStream_t s);
int head_bwd(const void* a, const float* w, const float* dout, const float* out, void* da, float* dw,
             float* db, int N, int H, int W, int C, int n_out, int sigmoid, const BnReduceCtx* bn, cudaStream_t s);
size_t combined_loss_workspace_bytes();
int combined_loss(const float* recon, const float* amap, const float* image, const float* mask,
                  float* d_recon, float* d_amap, float* out3, void* work, size_t work_bytes,
                  long n_recon, long n_seg, float rw, float sw, float alpha, float gamma,
                  cudaStream_t s);
int scale_by_scalar(float* y, const float* s, long n, cudaStream_t st);
// seg_loss.cu — SegmentationUNet pipeline extras
size_t seg_loss_workspace_bytes(int N, int C);
int seg_loss_fwd(const float* logits, const long long* target, const float* class_w, float* out4, void* work,
                 size_t work_bytes, int N, int C, int H, int W, long long ignore_index, float ce_w, float dice_w,
                 float focal_w, float focal_alpha, float focal_gamma, int is_prob, cudaStream_t s);
int seg_loss_bwd(const float* logits, const long long* target, const float* class_w, const void* work, const float* gout,
                 float* dlogits, int N, int C, int H, int W, long long ignore_index, float ce_w, float dice_w,
                 float focal_w, float focal_alpha, float focal_gamma, int is_prob, cudaStream_t s);
int channel_scale(void* x, const float* mask, int N, int H, int W, int C, cudaStream_t s);
int argmax_channels(const float* logits, long long* out, int N, int C, int H, int W, cudaStream_t s);
int upsample2x_fwd(const void* x, void* out, int N, int H, int W, int C, cudaStream_t s);
int upsample2x_bwd(const void* dout, void* dx, int N, int H, int W, int C, int accumulate, cudaStream_t s);
int pixel_confusion(const float* pred, const float* truth, const unsigned char* include, const float* thresholds,
                    int n_thresholds, float truth_threshold, long long* counts, int n, long elems_per_image,
                    cudaStream_t s);
int class_confusion(const float* logits, const long long* preds, const long long* target, int n, int c, long hw,
                    long long ignore_index, int has_ignore, long long* cm, long long* kept, cudaStream_t s);
size_t select_workspace_bytes();
int select_kth(const float* x, long n, long k, float* out2, void* workspace, size_t workspace_bytes, cudaStream_t s);
int resize_ksize(int in_size, int out_size);
int resize_coeffs(int in_size, int out_size, int* bounds, int* coeffs);
int resize_u8(const unsigned char* src, unsigned char* tmp, unsigned char* out_u8, float* out_f32, int n, int h,
              int w, int c, int oh, int ow, const int* hbounds, const int* hcoeffs, int hksize, const int* vbounds,
              const int* vcoeffs, int vksize, const float* mean, const float* stdv, cudaStream_t s);
int flip_rotate_u8(const unsigned char* src, unsigned char* dst, int n, int h, int w, int c, const int* geom,
                   cudaStream_t s);
int color_jitter_u8(unsigned char* img, int n, int h, int w, const int* ops, const float* factors,
                    unsigned long long* gray_sums, cudaStream_t s);
int u8_to_tensor(const unsigned char* hwc, float* chw, int n, int h, int w, int c, const float* mean,
                 const float* stdv, cudaStream_t s);
int resize_nearest_table(int in_size, int out_size, int* table);
int resize_nearest_u8(const unsigned char* src, long long* out_i64, unsigned char* out_u8, int n, int h, int w,
                      int oh, int ow, const int* xtab, const int* ytab, cudaStream_t s);
int anomaly_score(const float* recon, const float* orig, float* out, int N, int C, int H, int W,
                  int method, cudaStream_t s);
int ssim_fwd(const float* img1, const float* img2, float* sums, float* maps, int N, int C, int H,
             int W, int window_size, int need2, cudaStream_t s);
int ssim_finalize(const float* sums, float* out, int N, int C, int H, int W, int size_average,
                  cudaStream_t s);
int ssim_bwd(const float* img1, const float* img2, const float* maps, const float* gout, float* grad,
             int N, int C, int H, int W, int window_size, int which, int per_image, cudaStream_t s);

// optim.cu
int adam_step(float* p, const float* g, float* m, float* v, long n, float lr, float beta1,
              float beta2, float eps, float wd, long step, float grad_scale, int mode,
              cudaStream_t s);
int permute_batched(const void* descs, int n_desc, int n_chunks, const float* in, void* out,
                    int out_bf16, float scale, cudaStream_t s);
int permute_chunk_elems();
int pack_conv_weights(const void* descs, int n_desc, int n_blocks, const float* flat_p, void* wpack,
                      cudaStream_t s);
int unpack_conv_grads(const void* descs, int n_desc, int n_blocks, int max_cin, const float* gstage,
                      float* flat_g, float scale, cudaStream_t s);
int fused_update(const void* descs, int n_desc, int n_blocks, int max_blocks, float* p, float* g, float* m, float* v,
                 float* gstage, void* wpack, const float* hyper, int mode, cudaStream_t s);
int update_plain_chunk_elems();
int set_floats(float* dst, const float* host_vals, int n, cudaStream_t s);

}  // namespace unetk
