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

// ---------------------------------------------------------------------------------------------
// The random training augmentations of src/dataset.py:131-138 with host-drawn parameters
// (RandomHorizontalFlip -> RandomRotation -> ColorJitter on the resized uint8 image), restated from
// torchvision 0.26 + Pillow 12.2: Geometry.c affine_fixed (nearest neighbour in 16.16 fixed point), Blend.c
// (float32, truncating), Convert.c rgb2l / rgb2hsv / hsv2rgb (float32 ratios, double hue arithmetic).
// Every floating-point step uses the *_rn intrinsics so that no multiply-add is contracted.
// ---------------------------------------------------------------------------------------------
// geom: 8 ints per image = (mode, flip, a0, a1, a2, a3, a4, a5); mode 0 identity, 1 affine
// (source x = (a2 + a1 y + a0 x) >> 16, source y = (a5 + a4 y + a3 x) >> 16, outside -> 0), 2 / 3 / 4 the exact
// 180 / 90 / 270 degree transposes Image.rotate takes for those angles (90 / 270: square images)
template <int C>
__global__ void __launch_bounds__(256)
flip_rotate_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, int h, int w,
                   const int* __restrict__ geom) {
  const int img = blockIdx.y;
  const int* g = geom + img * 8;
  const int mode = g[0], flip = g[1];
  const long long a0 = g[2], a1 = g[3], a2 = g[4], a3 = g[5], a4 = g[6], a5 = g[7];
  const long plane = static_cast<long>(h) * w;
  const unsigned char* s = src + img * plane * C;
  unsigned char* d = dst + img * plane * C;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < plane;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int y = static_cast<int>(i / w), x = static_cast<int>(i - static_cast<long>(y) * w);
    long long xin, yin;
    if (mode == 1) {
      xin = (a2 + a1 * y + a0 * x) >> 16;
      yin = (a5 + a4 * y + a3 * x) >> 16;
    } else if (mode == 2) {
      xin = w - 1 - x;
      yin = h - 1 - y;
    } else if (mode == 3) {
      xin = w - 1 - y;
      yin = x;
    } else if (mode == 4) {
      xin = y;
      yin = h - 1 - x;
    } else {
      xin = x;
      yin = y;
    }
    const bool ok = xin >= 0 && xin < w && yin >= 0 && yin < h;
    if (flip) xin = w - 1 - xin;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) d[i * C + ch] = ok ? __ldg(s + (yin * w + xin) * C + ch) : 0;
  }
}

__device__ __forceinline__ int luma601(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// Image.blend(degenerate, image, alpha)
__device__ __forceinline__ unsigned char blend8(int in1, int in2, float alpha) {
  if (alpha == 0.f) return static_cast<unsigned char>(in1);
  if (alpha == 1.f) return static_cast<unsigned char>(in2);
  const float t = __fadd_rn(static_cast<float>(in1), __fmul_rn(alpha, static_cast<float>(in2 - in1)));
  if (alpha >= 0.f && alpha <= 1.f) return static_cast<unsigned char>(static_cast<int>(t));
  if (t <= 0.f) return 0;
  if (t >= 255.f) return 255;
  return static_cast<unsigned char>(static_cast<int>(t));
}

__device__ __forceinline__ int clampi(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

__device__ __forceinline__ void rgb2hsv8(int r, int g, int b, int& uh, int& us, int& uv) {
  const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
  uv = maxc;
  if (minc == maxc) {
    uh = 0;
    us = 0;
    return;
  }
  const float cr = static_cast<float>(maxc - minc);
  const float s = __fdiv_rn(cr, static_cast<float>(maxc));
  const float rc = __fdiv_rn(static_cast<float>(maxc - r), cr);
  const float gc = __fdiv_rn(static_cast<float>(maxc - g), cr);
  const float bc = __fdiv_rn(static_cast<float>(maxc - b), cr);
  float hf;
  if (r == maxc)
    hf = __fsub_rn(bc, gc);
  else if (g == maxc)
    hf = static_cast<float>(__dsub_rn(__dadd_rn(2.0, static_cast<double>(rc)), static_cast<double>(bc)));
  else
    hf = static_cast<float>(__dsub_rn(__dadd_rn(4.0, static_cast<double>(gc)), static_cast<double>(rc)));
  hf = static_cast<float>(fmod(__dadd_rn(__ddiv_rn(static_cast<double>(hf), 6.0), 1.0), 1.0));
  uh = clampi(static_cast<int>(__dmul_rn(static_cast<double>(hf), 255.0)));
  us = clampi(static_cast<int>(__dmul_rn(static_cast<double>(s), 255.0)));
}

__device__ __forceinline__ int round8(double x) { return clampi(static_cast<int>(floor(__dadd_rn(x, 0.5)))); }

__device__ __forceinline__ void hsv2rgb8(int h, int s, int v, int& r, int& g, int& b) {
  if (s == 0) {
    r = g = b = v;
    return;
  }
  const double hh = __ddiv_rn(__dmul_rn(static_cast<double>(h), 6.0), 255.0);
  const int i = static_cast<int>(floor(hh));
  const double f = static_cast<double>(static_cast<float>(__dsub_rn(hh, static_cast<double>(i))));
  const double fs = static_cast<double>(static_cast<float>(__ddiv_rn(static_cast<double>(s), 255.0)));
  const double vv = static_cast<double>(v);
  const int p = round8(__dmul_rn(vv, __dsub_rn(1.0, fs)));
  const int q = round8(__dmul_rn(vv, __dsub_rn(1.0, __dmul_rn(fs, f))));
  const int t = round8(__dmul_rn(vv, __dsub_rn(1.0, __dmul_rn(fs, __dsub_rn(1.0, f)))));
  switch (i % 6) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

// sum of the luma of the images whose adjustment at this stage is the contrast (ImageEnhance.Contrast's mean)
__global__ void __launch_bounds__(256)
gray_sum_kernel(const unsigned char* __restrict__ img, long plane, const int* __restrict__ ops, int stage,
                unsigned long long* __restrict__ sums) {
  const int n = blockIdx.y;
  if (ops[n * 4 + stage] != 1) return;
  const unsigned char* s = img + n * plane * 3;
  unsigned long long acc = 0;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < plane;
       i += static_cast<long>(gridDim.x) * blockDim.x)
    acc += luma601(__ldg(s + i * 3), __ldg(s + i * 3 + 1), __ldg(s + i * 3 + 2));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(sums + stage * gridDim.y + n, acc);
}

// one ColorJitter adjustment per image, in place: ops[n, stage] = 0 brightness, 1 contrast, 2 saturation, 3 hue
// (factor = the uint8 amount added to H), anything else: skip
__global__ void __launch_bounds__(256)
jitter_kernel(unsigned char* __restrict__ img, long plane, const int* __restrict__ ops,
              const float* __restrict__ factors, int stage, const unsigned long long* __restrict__ sums) {
  const int n = blockIdx.y;
  const int op = ops[n * 4 + stage];
  if (op < 0 || op > 3) return;
  const float factor = factors[n * 4 + stage];
  int mean = 0;
  if (op == 1)
    mean = static_cast<int>(__dadd_rn(__ddiv_rn(static_cast<double>(sums[stage * gridDim.y + n]), static_cast<double>(plane)), 0.5));
  const int dh = op == 3 ? static_cast<int>(factor) : 0;
  unsigned char* d = img + n * plane * 3;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < plane;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    int r = d[i * 3], g = d[i * 3 + 1], b = d[i * 3 + 2];
    if (op == 3) {
      int uh, us, uv;
      rgb2hsv8(r, g, b, uh, us, uv);
      uh = (uh + dh) & 255;
      hsv2rgb8(uh, us, uv, r, g, b);
    } else {
      const int y = luma601(r, g, b);
      const int d0 = op == 0 ? 0 : (op == 1 ? mean : y);
      r = blend8(d0, r, factor);
      g = blend8(d0, g, factor);
      b = blend8(d0, b, factor);
    }
    d[i * 3] = static_cast<unsigned char>(r);
    d[i * 3 + 1] = static_cast<unsigned char>(g);
    d[i * 3 + 2] = static_cast<unsigned char>(b);
  }
}

// ToTensor + Normalize: u8 [n, h, w, c] -> fp32 [n, c, h, w]
template <int C>
__global__ void __launch_bounds__(256)
u8_to_tensor_kernel(const unsigned char* __restrict__ src, float* __restrict__ out, long plane, long total, NormArgs norm) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long img = i / plane, r = i - img * plane;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
      const float x = __fdiv_rn(static_cast<float>(__ldg(src + i * C + ch)), 255.0f);
      out[(img * C + ch) * plane + r] = __fdiv_rn(__fsub_rn(x, norm.mean[ch]), norm.stdv[ch]);
    }
  }
}

// Image.resize(size, NEAREST) of a label mask (src/gear_dataset.py:255-258): a gather through the per-axis source
// index tables of Pillow's ImagingScaleAffine (resize_nearest_table); u8 [n, h, w] -> int64 [n, oh, ow]
// (MaskToTensor's .long()) and / or u8.
__global__ void __launch_bounds__(256)
resize_nearest_kernel(const unsigned char* __restrict__ src, long long* __restrict__ out_i64,
                      unsigned char* __restrict__ out_u8, int n, int h, int w, int oh, int ow,
                      const int* __restrict__ xtab, const int* __restrict__ ytab) {
  const long plane = static_cast<long>(oh) * ow, total = plane * n;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long img = i / plane, r = i - img * plane;
    const int oy = static_cast<int>(r / ow), ox = static_cast<int>(r - static_cast<long>(oy) * ow);
    const int sx = __ldg(xtab + ox), sy = __ldg(ytab + oy);
    const unsigned char v = (sx >= 0 && sy >= 0) ? __ldg(src + (img * h + sy) * w + sx) : 0;
    if (out_i64 != nullptr) out_i64[i] = v;
    if (out_u8 != nullptr) out_u8[i] = v;
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

static NormArgs make_norm(const float* mean, const float* stdv, int c) {
  NormArgs norm{};
  for (int ch = 0; ch < kMaxChannels; ++ch) {
    norm.mean[ch] = (mean != nullptr && ch < c) ? mean[ch] : 0.f;
    norm.stdv[ch] = (stdv != nullptr && ch < c) ? stdv[ch] : 1.f;
  }
  return norm;
}

static dim3 per_image_grid(long plane, int n) {
  long bx = (plane + 256 * 4 - 1) / (256 * 4);
  long cap = static_cast<long>(num_sms()) * 8 / n;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  return dim3(static_cast<unsigned>(bx < 1 ? 1 : bx), static_cast<unsigned>(n));
}

int flip_rotate_u8(const unsigned char* src, unsigned char* dst, int n, int h, int w, int c, const int* geom,
                   cudaStream_t s) {
  UNETK_CHECK_ARG(src && dst && geom && src != dst, "flip_rotate_u8: null or aliased pointers");
  UNETK_CHECK_ARG(c == 1 || c == 3, "flip_rotate_u8: 1 or 3 channels, got %d", c);
  UNETK_CHECK_ARG(n >= 1 && n <= 65535 && h >= 1 && w >= 1 && h < 32768 && w < 32768, "flip_rotate_u8: bad sizes");
  const dim3 grid = per_image_grid(static_cast<long>(h) * w, n);
  if (c == 3)
    flip_rotate_kernel<3><<<grid, 256, 0, s>>>(src, dst, h, w, geom);
  else
    flip_rotate_kernel<1><<<grid, 256, 0, s>>>(src, dst, h, w, geom);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int color_jitter_u8(unsigned char* img, int n, int h, int w, const int* ops, const float* factors,
                    unsigned long long* gray_sums, cudaStream_t s) {
  UNETK_CHECK_ARG(img && ops && factors && gray_sums, "color_jitter_u8: null pointer");
  UNETK_CHECK_ARG(n >= 1 && n <= 65535 && h >= 1 && w >= 1, "color_jitter_u8: bad sizes");
  const long plane = static_cast<long>(h) * w;
  const dim3 grid = per_image_grid(plane, n);
  UNETK_CUDA(cudaMemsetAsync(gray_sums, 0, sizeof(unsigned long long) * 4 * n, s));
  for (int stage = 0; stage < 4; ++stage) {
    gray_sum_kernel<<<grid, 256, 0, s>>>(img, plane, ops, stage, gray_sums);
    jitter_kernel<<<grid, 256, 0, s>>>(img, plane, ops, factors, stage, gray_sums);
  }
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int u8_to_tensor(const unsigned char* hwc, float* chw, int n, int h, int w, int c, const float* mean,
                 const float* stdv, cudaStream_t s) {
  UNETK_CHECK_ARG(hwc && chw, "u8_to_tensor: null pointer");
  UNETK_CHECK_ARG(c == 1 || c == 3, "u8_to_tensor: 1 or 3 channels, got %d", c);
  UNETK_CHECK_ARG(n >= 1 && h >= 1 && w >= 1, "u8_to_tensor: bad sizes");
  const long plane = static_cast<long>(h) * w, total = plane * n;
  const NormArgs norm = make_norm(mean, stdv, c);
  if (c == 3)
    u8_to_tensor_kernel<3><<<grid_for(total), 256, 0, s>>>(hwc, chw, plane, total, norm);
  else
    u8_to_tensor_kernel<1><<<grid_for(total), 256, 0, s>>>(hwc, chw, plane, total, norm);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

// HOST: source index per output index of one axis for Image.resize(NEAREST) (Geometry.c ImagingScaleAffine:
// the sampling position starts at scale / 2 and is ACCUMULATED in double, one addition per output index)
int resize_nearest_table(int in_size, int out_size, int* table) {
  UNETK_CHECK_ARG(in_size >= 1 && out_size >= 1 && table, "resize_nearest_table: bad arguments");
  const double step = static_cast<double>(in_size) / out_size;
  double pos = 0.0 + step * 0.5;
  for (int x = 0; x < out_size; ++x) {
    const int xin = pos < 0.0 ? -1 : static_cast<int>(pos);
    table[x] = (xin >= 0 && xin < in_size) ? xin : -1;
    pos += step;
  }
  return UNETK_OK;
}

int resize_nearest_u8(const unsigned char* src, long long* out_i64, unsigned char* out_u8, int n, int h, int w,
                      int oh, int ow, const int* xtab, const int* ytab, cudaStream_t s) {
  UNETK_CHECK_ARG(src && (out_i64 || out_u8) && xtab && ytab, "resize_nearest_u8: null pointer");
  UNETK_CHECK_ARG(n >= 1 && h >= 1 && w >= 1 && oh >= 1 && ow >= 1, "resize_nearest_u8: bad sizes");
  resize_nearest_kernel<<<grid_for(static_cast<long>(n) * oh * ow), 256, 0, s>>>(src, out_i64, out_u8, n, h, w, oh,
                                                                                ow, xtab, ytab);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

}  // namespace unetk
