// Multi-class segmentation losses, channel dropout and argmax (sm_100a) — the operators of the reference's
// SegmentationUNet pipeline that the U-Net hot path does not already cover.
//
// Reference operators being replaced:
//   CombinedSegmentationLoss.forward  = ce_w * F.cross_entropy(pred, target, weight, ignore_index)
//                                     + dice_w * dice_loss(softmax(pred), target)
//                                     + focal_w * focal_loss(pred, target)          /root/reference/src/metrics.py:285-335
//   dice_loss (one-hot target, per (n, c) dice, 1 - mean)                           src/metrics.py:233-261
//   focal_loss (alpha (1 - pt)^gamma ce, mean over all pixels)                      src/metrics.py:264-282
//   nn.Dropout2d on the bottleneck of SegmentationUNet                              src/model.py:130,146
//   torch.argmax(outputs, dim=1)                                                    test.py:98, src/metrics.py:31
//
// Layouts as the callers pass them: logits NCHW fp32 [N, C, H, W] (C <= 8), targets int64 [N, H, W].
// One thread per pixel reads its C logits from C coalesced planes; the softmax is recomputed in the
// backward pass instead of being stored (the kernels are bandwidth bound: (C + 2) * 4 bytes per pixel).
// Pass 1 accumulates the cross-entropy numerator / denominator, the focal sum and the per-(n, c) dice
// sums; a one-block finalize turns them into the four scalars and the dice coefficients of the backward.
#include "common.cuh"
#include "kernels.h"

namespace unetk {

constexpr int kSegMaxC = 8;
constexpr int kSegThreads = 256;
// workspace (floats): [0] sum w_t*ce, [1] sum w_t, [2] sum focal, [3] pad, then per (n, c): I, P, T,
// then per (n, c): A, B (backward coefficients of the dice term)
__host__ __device__ inline size_t seg_ws_floats(int N, int C) { return 4 + static_cast<size_t>(N) * C * 5; }

// (1 - pt)^gamma and gamma (1 - pt)^(gamma - 1): gamma = 2 (the reference's default everywhere) avoids powf
__device__ __forceinline__ float seg_pow(float om, float gamma) { return gamma == 2.f ? om * om : powf(om, gamma); }
__device__ __forceinline__ float seg_dpow(float om, float gamma) {
  return gamma == 2.f ? 2.f * om : (om > 0.f ? gamma * powf(om, gamma - 1.f) : 0.f);
}

struct SegPixel {
  float p[kSegMaxC];   // probabilities
  float ce;            // -log p_t (0 when the target is not a class)
  int t;               // target class or -1
  bool valid;          // counts for cross entropy / focal
};

__device__ __forceinline__ SegPixel seg_load(const float* __restrict__ z, const long long* __restrict__ tgt, long img_off,
                                             long hw, long pix, int C, long long ignore_index, int is_prob) {
  SegPixel s;
  float v[kSegMaxC];
  float m = -INFINITY;
#pragma unroll
  for (int c = 0; c < kSegMaxC; ++c) {
    v[c] = (c < C) ? __ldg(z + img_off + c * hw + pix) : -INFINITY;
    m = fmaxf(m, v[c]);
  }
  const long long t = tgt[pix];
  s.t = (t >= 0 && t < C) ? static_cast<int>(t) : -1;
  s.valid = (t != ignore_index) && s.t >= 0;
  if (is_prob) {       // dice_loss on probabilities handed in by the caller
#pragma unroll
    for (int c = 0; c < kSegMaxC; ++c) s.p[c] = (c < C) ? v[c] : 0.f;
    s.ce = 0.f;
    return s;
  }
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < kSegMaxC; ++c) {
    s.p[c] = (c < C) ? expf(v[c] - m) : 0.f;
    sum += s.p[c];
  }
  const float inv = 1.f / sum;
  float zt = 0.f;
#pragma unroll
  for (int c = 0; c < kSegMaxC; ++c) {
    s.p[c] *= inv;
    if (c == s.t) zt = v[c];
  }
  s.ce = s.valid ? -(zt - m - logf(sum)) : 0.f;
  return s;
}

__global__ void __launch_bounds__(kSegThreads)
seg_loss_fwd_kernel(const float* __restrict__ z, const long long* __restrict__ tgt, const float* __restrict__ class_w,
                    float* __restrict__ ws, int C, long hw, long long ignore_index, float focal_alpha,
                    float focal_gamma, int is_prob) {
  const int n = blockIdx.y;
  const long img_off = static_cast<long>(n) * C * hw;
  float acc[3 + 3 * kSegMaxC];
#pragma unroll
  for (int i = 0; i < 3 + 3 * kSegMaxC; ++i) acc[i] = 0.f;
  for (long pix = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; pix < hw;
       pix += static_cast<long>(gridDim.x) * blockDim.x) {
    const SegPixel s = seg_load(z, tgt + static_cast<long>(n) * hw, img_off, hw, pix, C, ignore_index, is_prob);
    if (s.valid) {
      const float w = class_w != nullptr ? __ldg(class_w + s.t) : 1.f;
      acc[0] = fmaf(w, s.ce, acc[0]);
      acc[1] += w;
      const float pt = expf(-s.ce);
      acc[2] += focal_alpha * seg_pow(fmaxf(1.f - pt, 0.f), focal_gamma) * s.ce;
    }
#pragma unroll
    for (int c = 0; c < kSegMaxC; ++c) {
      const float y = (c == s.t) ? 1.f : 0.f;
      acc[3 + 3 * c] = fmaf(s.p[c], y, acc[3 + 3 * c]);
      acc[4 + 3 * c] += s.p[c];
      acc[5 + 3 * c] += y;
    }
  }
  __shared__ float red[kSegThreads / 32][3 + 3 * kSegMaxC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 3 + 3 * kSegMaxC; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3 + 3 * C) {
    float v = 0.f;
    for (int w = 0; w < kSegThreads / 32; ++w) v += red[w][threadIdx.x];
    const int i = threadIdx.x;
    if (i < 3) atomicAdd(ws + i, v);
    else atomicAdd(ws + 4 + (static_cast<size_t>(n) * C + (i - 3) / 3) * 3 + (i - 3) % 3, v);
  }
}

// out4 = {total, ce, dice, focal}; also the dice backward coefficients A = 2/(U+s)/(N*C), B = (2I+s)/(U+s)^2/(N*C)
__global__ void seg_loss_finalize_kernel(float* __restrict__ ws, float* __restrict__ out4, int N, int C, float npix_total,
                                         float ce_w, float dice_w, float focal_w, float smooth) {
  __shared__ float sdice;
  if (threadIdx.x == 0) sdice = 0.f;
  __syncthreads();
  const int nc = N * C;
  float* coef = ws + 4 + static_cast<size_t>(nc) * 3;
  float local = 0.f;
  for (int i = threadIdx.x; i < nc; i += blockDim.x) {
    const float I = ws[4 + 3 * i], U = ws[5 + 3 * i] + ws[6 + 3 * i];
    const float d = (2.f * I + smooth) / (U + smooth);
    local += d;
    coef[2 * i] = 2.f / (U + smooth) / nc;
    coef[2 * i + 1] = (2.f * I + smooth) / ((U + smooth) * (U + smooth)) / nc;
  }
  atomicAdd(&sdice, local);
  __syncthreads();
  if (threadIdx.x == 0) {
    const float ce = ws[1] > 0.f ? ws[0] / ws[1] : 0.f;
    const float dice = 1.f - sdice / nc;
    const float focal = ws[2] / npix_total;
    out4[1] = ce;
    out4[2] = dice;
    out4[3] = focal;
    out4[0] = (ce_w > 0.f ? ce_w * ce : 0.f) + (dice_w > 0.f ? dice_w * dice : 0.f) + (focal_w > 0.f ? focal_w * focal : 0.f);
  }
}

__global__ void __launch_bounds__(kSegThreads)
seg_loss_bwd_kernel(const float* __restrict__ z, const long long* __restrict__ tgt, const float* __restrict__ class_w,
                    const float* __restrict__ ws, const float* __restrict__ gout, float* __restrict__ dz, int N, int C,
                    long hw, long long ignore_index, float ce_w, float dice_w, float focal_w, float focal_alpha,
                    float focal_gamma, int is_prob) {
  const int n = blockIdx.y;
  const long img_off = static_cast<long>(n) * C * hw;
  const float g = gout != nullptr ? __ldg(gout) : 1.f;
  const float wsum = ws[1];
  const float k_ce = (ce_w > 0.f && wsum > 0.f) ? g * ce_w / wsum : 0.f;
  const float k_focal = focal_w > 0.f ? g * focal_w / (static_cast<float>(N) * hw) : 0.f;
  const float k_dice = dice_w > 0.f ? g * dice_w : 0.f;
  const float* coef = ws + 4 + static_cast<size_t>(N) * C * 3 + static_cast<size_t>(n) * C * 2;
  float A[kSegMaxC], B[kSegMaxC];
#pragma unroll
  for (int c = 0; c < kSegMaxC; ++c) {
    A[c] = (c < C) ? coef[2 * c] : 0.f;
    B[c] = (c < C) ? coef[2 * c + 1] : 0.f;
  }
  for (long pix = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; pix < hw;
       pix += static_cast<long>(gridDim.x) * blockDim.x) {
    const SegPixel s = seg_load(z, tgt + static_cast<long>(n) * hw, img_off, hw, pix, C, ignore_index, is_prob);
    // cross entropy + focal share d(ce)/dz_k = p_k - y_k
    float k_lin = 0.f;
    if (s.valid) {
      const float w = class_w != nullptr ? __ldg(class_w + s.t) : 1.f;
      const float pt = expf(-s.ce);
      const float om = fmaxf(1.f - pt, 0.f);
      const float dfdce = focal_alpha * (seg_pow(om, focal_gamma) + seg_dpow(om, focal_gamma) * pt * s.ce);
      k_lin = k_ce * w + k_focal * dfdce;
    }
    // dice: G_c = dL/dp_c = B - y_c A; through the softmax: dz_k = p_k (G_k - sum_c G_c p_c)
    float G[kSegMaxC];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < kSegMaxC; ++c) {
      G[c] = B[c] - ((c == s.t) ? A[c] : 0.f);
      dot = fmaf(G[c], s.p[c], dot);
    }
#pragma unroll
    for (int c = 0; c < kSegMaxC; ++c) {
      if (c < C) {
        const float y = (c == s.t) ? 1.f : 0.f;
        float d;
        if (is_prob) d = k_dice * G[c];
        else d = k_lin * (s.p[c] - y) + k_dice * s.p[c] * (G[c] - dot);
        dz[img_off + c * hw + pix] = d;
      }
    }
  }
}

// x[n, h, w, c] *= mask[n, c]  (Dropout2d: mask = bernoulli(1 - p) / (1 - p) per (n, c))
__global__ void __launch_bounds__(256)
channel_scale_kernel(uint4* __restrict__ x, const float* __restrict__ mask, long hw, int cgroups, long nvec) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % cgroups);
    const long n = i / (cgroups * hw);
    const float* m = mask + n * cgroups * 8 + cg * 8;
    uint4 u = x[i];
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    u.x = pack_bf16x2(a.x * m[0], a.y * m[1]);
    u.y = pack_bf16x2(b.x * m[2], b.y * m[3]);
    u.z = pack_bf16x2(c.x * m[4], c.y * m[5]);
    u.w = pack_bf16x2(d.x * m[6], d.y * m[7]);
    x[i] = u;
  }
}

// out[n, y, x] = index of the first maximum over the channel axis (torch.argmax(dim=1))
__global__ void __launch_bounds__(256)
argmax_channels_kernel(const float* __restrict__ z, long long* __restrict__ out, int C, long hw, long total) {
  for (long p = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; p < total;
       p += static_cast<long>(gridDim.x) * blockDim.x) {
    const long n = p / hw, r = p - n * hw;
    const float* b = z + n * C * hw + r;
    float best = __ldg(b);
    int bi = 0;
    for (int c = 1; c < C; ++c) {
      const float v = __ldg(b + c * hw);
      if (v > best || (v != v && best == best)) {   // NaN counts as the maximum, like ATen
        best = v;
        bi = c;
      }
    }
    out[p] = bi;
  }
}

// ----------------------------------------------------------------------------------------------
// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) of the bilinear=True decoders
// (src/model.py:47-49) on NHWC bf16: out[n, 2h, 2w, c].  Source coordinate o * (in - 1) / (out - 1), the
// four taps blended in fp32 exactly as ATen's upsample_bilinear2d does.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void up2_src(int o, int in, int out, int& i0, int& i1, float& l1) {
  const float r = out > 1 ? static_cast<float>(in - 1) / static_cast<float>(out - 1) : 0.f;
  const float f = r * o;
  i0 = static_cast<int>(f);
  i1 = i0 + ((i0 < in - 1) ? 1 : 0);
  l1 = f - i0;
}

__global__ void __launch_bounds__(256)
upsample2x_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int H, int W, int cgroups, long nvec) {
  const int OH = 2 * H, OW = 2 * W;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % cgroups);
    long pp = i / cgroups;
    const int ox = static_cast<int>(pp % OW);
    pp /= OW;
    const int oy = static_cast<int>(pp % OH);
    const long n = pp / OH;
    int y0, y1, x0, x1;
    float ly, lx;
    up2_src(oy, H, OH, y0, y1, ly);
    up2_src(ox, W, OW, x0, x1, lx);
    const uint4* b = x + n * H * W * cgroups + cg;
    const uint4 q00 = ldg_nc_v4(b + (static_cast<long>(y0) * W + x0) * cgroups), q01 = ldg_nc_v4(b + (static_cast<long>(y0) * W + x1) * cgroups);
    const uint4 q10 = ldg_nc_v4(b + (static_cast<long>(y1) * W + x0) * cgroups), q11 = ldg_nc_v4(b + (static_cast<long>(y1) * W + x1) * cgroups);
    const uint32_t* a = &q00.x; const uint32_t* bb = &q01.x; const uint32_t* c = &q10.x; const uint32_t* d = &q11.x;
    uint4 o;
    uint32_t* po = &o.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 v00 = unpack_bf16x2(a[k]), v01 = unpack_bf16x2(bb[k]), v10 = unpack_bf16x2(c[k]), v11 = unpack_bf16x2(d[k]);
      const float r0 = (1.f - ly) * ((1.f - lx) * v00.x + lx * v01.x) + ly * ((1.f - lx) * v10.x + lx * v11.x);
      const float r1 = (1.f - ly) * ((1.f - lx) * v00.y + lx * v01.y) + ly * ((1.f - lx) * v10.y + lx * v11.y);
      po[k] = pack_bf16x2(r0, r1);
    }
    out[i] = o;
  }
}

// gradient w.r.t. the input, gathered per input pixel from the (at most 4 x 4) output pixels that read it
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const uint4* __restrict__ dout, uint4* dx, int H, int W, int cgroups, long nvec, int accumulate) {
  const int OH = 2 * H, OW = 2 * W;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % cgroups);
    long pp = i / cgroups;
    const int ix = static_cast<int>(pp % W);
    pp /= W;
    const int iy = static_cast<int>(pp % H);
    const long n = pp / H;
    float acc[8];
    if (accumulate) {
      const uint4 o = dx[i];
      const float2 a = unpack_bf16x2(o.x), b = unpack_bf16x2(o.y), c = unpack_bf16x2(o.z), d = unpack_bf16x2(o.w);
      acc[0] = a.x; acc[1] = a.y; acc[2] = b.x; acc[3] = b.y; acc[4] = c.x; acc[5] = c.y; acc[6] = d.x; acc[7] = d.y;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    }
    const uint4* b = dout + n * OH * OW * cgroups + cg;
    const int oy_lo = max(0, 2 * iy - 3), oy_hi = min(OH - 1, 2 * iy + 3);
    const int ox_lo = max(0, 2 * ix - 3), ox_hi = min(OW - 1, 2 * ix + 3);
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      int y0, y1;
      float ly;
      up2_src(oy, H, OH, y0, y1, ly);
      const float wy = (y0 == iy ? 1.f - ly : 0.f) + (y1 == iy ? ly : 0.f) - ((y0 == iy && y1 == iy) ? 0.f : 0.f);
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        int x0, x1;
        float lx;
        up2_src(ox, W, OW, x0, x1, lx);
        const float wx = (x0 == ix ? 1.f - lx : 0.f) + (x1 == ix ? lx : 0.f);
        if (wx == 0.f) continue;
        const float wgt = wy * wx;
        const uint4 g = ldg_nc_v4(b + (static_cast<long>(oy) * OW + ox) * cgroups);
        const float2 a = unpack_bf16x2(g.x), bb = unpack_bf16x2(g.y), c = unpack_bf16x2(g.z), d = unpack_bf16x2(g.w);
        acc[0] = fmaf(wgt, a.x, acc[0]); acc[1] = fmaf(wgt, a.y, acc[1]); acc[2] = fmaf(wgt, bb.x, acc[2]); acc[3] = fmaf(wgt, bb.y, acc[3]);
        acc[4] = fmaf(wgt, c.x, acc[4]); acc[5] = fmaf(wgt, c.y, acc[5]); acc[6] = fmaf(wgt, d.x, acc[6]); acc[7] = fmaf(wgt, d.y, acc[7]);
      }
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    dx[i] = o;
  }
}

static int ew_blocks256(long nvec) {
  long blocks = (nvec + 255) / 256;
  const long cap = static_cast<long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

int upsample2x_fwd(const void* x, void* out, int N, int H, int W, int C, cudaStream_t s) {
  UNETK_CHECK_ARG(x && out, "upsample2x_fwd: null pointer");
  UNETK_CHECK_ARG(C > 0 && C % 8 == 0, "upsample2x_fwd: channel count must be a multiple of 8 (got %d)", C);
  const long nvec = static_cast<long>(N) * 4 * H * W * (C / 8);
  upsample2x_fwd_kernel<<<ew_blocks256(nvec), 256, 0, s>>>(static_cast<const uint4*>(x), static_cast<uint4*>(out), H, W, C / 8, nvec);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int upsample2x_bwd(const void* dout, void* dx, int N, int H, int W, int C, int accumulate, cudaStream_t s) {
  UNETK_CHECK_ARG(dout && dx, "upsample2x_bwd: null pointer");
  UNETK_CHECK_ARG(C > 0 && C % 8 == 0, "upsample2x_bwd: channel count must be a multiple of 8 (got %d)", C);
  const long nvec = static_cast<long>(N) * H * W * (C / 8);
  upsample2x_bwd_kernel<<<ew_blocks256(nvec), 256, 0, s>>>(static_cast<const uint4*>(dout), static_cast<uint4*>(dx), H, W, C / 8, nvec, accumulate);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

static int seg_grid_x(long hw, int N) {
  long b = (hw + kSegThreads - 1) / kSegThreads;
  const long cap = (static_cast<long>(num_sms()) * 8 + N - 1) / N;
  if (b > cap) b = cap;
  return static_cast<int>(b < 1 ? 1 : b);
}

size_t seg_loss_workspace_bytes(int N, int C) { return seg_ws_floats(N, C) * sizeof(float); }

int seg_loss_fwd(const float* logits, const long long* target, const float* class_w, float* out4, void* work,
                 size_t work_bytes, int N, int C, int H, int W, long long ignore_index, float ce_w, float dice_w,
                 float focal_w, float focal_alpha, float focal_gamma, int is_prob, cudaStream_t s) {
  UNETK_CHECK_ARG(logits && target && out4 && work, "seg_loss_fwd: null pointer");
  UNETK_CHECK_ARG(C >= 1 && C <= kSegMaxC, "seg_loss_fwd: 1..%d classes supported (got %d)", kSegMaxC, C);
  UNETK_CHECK_ARG(N >= 1 && H >= 1 && W >= 1, "seg_loss_fwd: empty input");
  if (work_bytes < seg_loss_workspace_bytes(N, C)) {
    set_last_error("seg_loss_fwd: workspace of %zu bytes needed, %zu given", seg_loss_workspace_bytes(N, C), work_bytes);
    return UNETK_ERR_WORKSPACE;
  }
  if (int e = check_device_sm100()) return e;
  const long hw = static_cast<long>(H) * W;
  float* ws = static_cast<float*>(work);
  UNETK_CUDA(cudaMemsetAsync(ws, 0, seg_loss_workspace_bytes(N, C), s));
  seg_loss_fwd_kernel<<<dim3(seg_grid_x(hw, N), N), kSegThreads, 0, s>>>(logits, target, class_w, ws, C, hw, ignore_index,
                                                                        focal_alpha, focal_gamma, is_prob);
  UNETK_CUDA(cudaGetLastError());
  seg_loss_finalize_kernel<<<1, 128, 0, s>>>(ws, out4, N, C, static_cast<float>(N) * hw, ce_w, dice_w, focal_w, 1e-8f);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int seg_loss_bwd(const float* logits, const long long* target, const float* class_w, const void* work, const float* gout,
                 float* dlogits, int N, int C, int H, int W, long long ignore_index, float ce_w, float dice_w,
                 float focal_w, float focal_alpha, float focal_gamma, int is_prob, cudaStream_t s) {
  UNETK_CHECK_ARG(logits && target && work && dlogits, "seg_loss_bwd: null pointer");
  UNETK_CHECK_ARG(C >= 1 && C <= kSegMaxC, "seg_loss_bwd: 1..%d classes supported (got %d)", kSegMaxC, C);
  if (int e = check_device_sm100()) return e;
  const long hw = static_cast<long>(H) * W;
  seg_loss_bwd_kernel<<<dim3(seg_grid_x(hw, N), N), kSegThreads, 0, s>>>(logits, target, class_w, static_cast<const float*>(work),
                                                                        gout, dlogits, N, C, hw, ignore_index, ce_w, dice_w,
                                                                        focal_w, focal_alpha, focal_gamma, is_prob);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int channel_scale(void* x, const float* mask, int N, int H, int W, int C, cudaStream_t s) {
  UNETK_CHECK_ARG(x && mask, "channel_scale: null pointer");
  UNETK_CHECK_ARG(C > 0 && C % 8 == 0, "channel_scale: channel count must be a multiple of 8 (got %d)", C);
  const long hw = static_cast<long>(H) * W, nvec = hw * N * (C / 8);
  long blocks = (nvec + 255) / 256;
  const long cap = static_cast<long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  channel_scale_kernel<<<static_cast<int>(blocks < 1 ? 1 : blocks), 256, 0, s>>>(static_cast<uint4*>(x), mask, hw, C / 8, nvec);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int argmax_channels(const float* logits, long long* out, int N, int C, int H, int W, cudaStream_t s) {
  UNETK_CHECK_ARG(logits && out, "argmax_channels: null pointer");
  UNETK_CHECK_ARG(C >= 1, "argmax_channels: no channels");
  const long hw = static_cast<long>(H) * W, total = hw * N;
  long blocks = (total + 255) / 256;
  const long cap = static_cast<long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  argmax_channels_kernel<<<static_cast<int>(blocks < 1 ? 1 : blocks), 256, 0, s>>>(logits, out, C, hw, total);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

}  // namespace unetk
