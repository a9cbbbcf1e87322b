// Evaluation-time input transform on the device (deterministic half):
//
//   transforms.Resize((S, S)) -> ToTensor() -> Normalize(mean, std)            src/dataset.py:141-152
//
// torchvision resizes a PIL image with Pillow's separable, antialiased bilinear resampling in 8-bit fixed
// point (Pillow 12.2 src/libImaging/Resample.c, a third-party dependency of the reference: not under
// /root/reference; restated from its published algorithm): per output index a window of source taps
// [first, first + count) with triangle weights normalised to 1 and rounded to 22 fractional bits; a pass
// accumulates 2^21 + sum(pixel * weight), shifts by 22 and clamps to a byte.  The horizontal pass is rounded
// to uint8 before the vertical one.  ToTensor divides by 255 in fp32, Normalize subtracts the mean and divides
// by the std in fp32: three correctly rounded operations, reproduced with __fdiv_rn / __fsub_rn.
// Integer / byte work: bit-exact against Pillow + torchvision.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace unetk {

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr int kMaxChannels = 4;

__device__ __forceinline__ unsigned char clip8(int v) {
  v >>= kPrecisionBits;
  return static_cast<unsigned char>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

struct NormArgs {
  float mean[kMaxChannels];
  float stdv[kMaxChannels];
};

// src u8 [n, h, w, c] -> dst u8 [n, h, ow, c]; one thread per (row, output column)
template <int C>
__global__ void __launch_bounds__(256)
resize_h_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, long rows, int w, int ow,
                const int* __restrict__ bounds, const int* __restrict__ coeffs, int ksize) {
  const long total = rows * ow;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long row = i / ow;
    const int ox = static_cast<int>(i - row * ow);
    const int first = __ldg(bounds + 2 * ox), count = __ldg(bounds + 2 * ox + 1);
    const int* k = coeffs + static_cast<long>(ox) * ksize;
    const unsigned char* s = src + (row * w + first) * C;
    int acc[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) acc[ch] = 1 << (kPrecisionBits - 1);
    for (int t = 0; t < count; ++t) {
      const int kv = __ldg(k + t);
#pragma unroll
      for (int ch = 0; ch < C; ++ch) acc[ch] += static_cast<int>(__ldg(s + t * C + ch)) * kv;
    }
#pragma unroll
    for (int ch = 0; ch < C; ++ch) dst[i * C + ch] = clip8(acc[ch]);
  }
}

// src u8 [n, h, ow, c] -> out_u8 [n, oh, ow, c] and / or out_f32 [n, c, oh, ow] = (v / 255 - mean) / std
template <int C>
__global__ void __launch_bounds__(256)
resize_v_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ out_u8,
                float* __restrict__ out_f32, int n, int h, int oh, int ow, const int* __restrict__ bounds,
                const int* __restrict__ coeffs, int ksize, NormArgs norm) {
  const long plane = static_cast<long>(oh) * ow, total = plane * n;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long img = i / plane;
    const long r = i - img * plane;
    const int oy = static_cast<int>(r / ow), ox = static_cast<int>(r - static_cast<long>(oy) * ow);
    const int first = __ldg(bounds + 2 * oy), count = __ldg(bounds + 2 * oy + 1);
    const int* k = coeffs + static_cast<long>(oy) * ksize;
    const unsigned char* s = src + ((img * h + first) * ow + ox) * C;
    int acc[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) acc[ch] = 1 << (kPrecisionBits - 1);
    for (int t = 0; t < count; ++t) {
      const int kv = __ldg(k + t);
      const unsigned char* p = s + static_cast<long>(t) * ow * C;
#pragma unroll
      for (int ch = 0; ch < C; ++ch) acc[ch] += static_cast<int>(__ldg(p + ch)) * kv;
    }
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
      const unsigned char v = clip8(acc[ch]);
      if (out_u8 != nullptr) out_u8[i * C + ch] = v;
      if (out_f32 != nullptr) {
        const float x = __fdiv_rn(static_cast<float>(v), 255.0f);                      // ToTensor
        out_f32[(img * C + ch) * plane + r] = __fdiv_rn(__fsub_rn(x, norm.mean[ch]), norm.stdv[ch]);   // Normalize
      }
    }
  }
}

inline double triangle(double x) {
  if (x < 0.0) x = -x;
  return x < 1.0 ? 1.0 - x : 0.0;
}

inline int grid_for(long items) {
  long blocks = (items + 255) / 256;
  const long cap = static_cast<long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

}  // namespace

// taps per output index of one axis (the window of the widest one)
int resize_ksize(int in_size, int out_size) {
  if (in_size < 1 || out_size < 1) return 0;
  double filterscale = static_cast<double>(in_size) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  return static_cast<int>(ceil(1.0 * filterscale)) * 2 + 1;
}

// HOST arrays: bounds int32 [out, 2] = (first source index, tap count), coeffs int32 [out, ksize]
int resize_coeffs(int in_size, int out_size, int* bounds, int* coeffs) {
  UNETK_CHECK_ARG(in_size >= 1 && out_size >= 1 && bounds && coeffs, "resize_coeffs: bad arguments");
  const double scale = static_cast<double>(in_size) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;     // bilinear: support 1
  const int ksize = static_cast<int>(ceil(support)) * 2 + 1;
  const double inv = 1.0 / filterscale;
  std::vector<double> w(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = triangle((x + xmin - center + 0.5) * inv);
      ww += w[x];
    }
    int* k = coeffs + static_cast<long>(xx) * ksize;
    for (int x = 0; x < xmax; ++x) {
      const double v = ww != 0.0 ? w[x] / ww : w[x];
      k[x] = v < 0 ? static_cast<int>(-0.5 + v * (1 << kPrecisionBits)) : static_cast<int>(0.5 + v * (1 << kPrecisionBits));
    }
    for (int x = xmax; x < ksize; ++x) k[x] = 0;
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return UNETK_OK;
}

int resize_u8(const unsigned char* src, unsigned char* tmp, unsigned char* out_u8, float* out_f32, int n, int h,
              int w, int c, int oh, int ow, const int* hbounds, const int* hcoeffs, int hksize, const int* vbounds,
              const int* vcoeffs, int vksize, const float* mean, const float* stdv, cudaStream_t s) {
  UNETK_CHECK_ARG(src && tmp && (out_u8 || out_f32) && hbounds && hcoeffs && vbounds && vcoeffs,
                  "resize_u8: null pointer");
  UNETK_CHECK_ARG(c == 1 || c == 3, "resize_u8: 1 (mask) or 3 (RGB) channels, got %d", c);
  UNETK_CHECK_ARG(n >= 1 && h >= 1 && w >= 1 && oh >= 1 && ow >= 1, "resize_u8: bad sizes");
  UNETK_CHECK_ARG(hksize == resize_ksize(w, ow) && vksize == resize_ksize(h, oh),
                  "resize_u8: tap tables do not belong to %dx%d -> %dx%d", h, w, oh, ow);
  NormArgs norm{};
  for (int ch = 0; ch < kMaxChannels; ++ch) {
    norm.mean[ch] = (mean != nullptr && ch < c) ? mean[ch] : 0.f;
    norm.stdv[ch] = (stdv != nullptr && ch < c) ? stdv[ch] : 1.f;
  }
  const long rows = static_cast<long>(n) * h;
  if (c == 3) {
    resize_h_kernel<3><<<grid_for(rows * ow), 256, 0, s>>>(src, tmp, rows, w, ow, hbounds, hcoeffs, hksize);
    resize_v_kernel<3><<<grid_for(static_cast<long>(n) * oh * ow), 256, 0, s>>>(tmp, out_u8, out_f32, n, h, oh, ow,
                                                                               vbounds, vcoeffs, vksize, norm);
  } else {
    resize_h_kernel<1><<<grid_for(rows * ow), 256, 0, s>>>(src, tmp, rows, w, ow, hbounds, hcoeffs, hksize);
    resize_v_kernel<1><<<grid_for(static_cast<long>(n) * oh * ow), 256, 0, s>>>(tmp, out_u8, out_f32, n, h, oh, ow,
                                                                               vbounds, vcoeffs, vksize, norm);
  }
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

}  // namespace unetk
