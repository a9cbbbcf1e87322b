// First layer of the encoder: Conv2d(n_channels <= 4 -> 64, 3x3, pad 1, bias=False) reading the
// caller's NCHW fp32 image directly (/root/reference/src/model.py:86,168 `inc` = DoubleConv(3, 64),
// first conv at src/model.py:14).  K = 9 * Cin = 27 is far too short for a tcgen05 tile and the
// layer is bound by its 128 B/pixel bf16 output, so it is a CUDA-core direct convolution:
// one thread per output pixel, 64 fp32 accumulators, weights broadcast from shared memory.
// Forward also emits the BatchNorm statistics (sum, sum of squares per output channel).
// wgrad: dW[co, ci, tap] = sum_pix dY[pix, co] * X[pix + tap, ci], staged through shared memory.
#include "common.cuh"
#include "kernels.h"

namespace unetk {

constexpr int kFcCout = 64;
constexpr int kFcMaxK = 36;  // 9 taps x up to 4 input channels

__global__ void __launch_bounds__(128)
first_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                      uint4* __restrict__ y, float* __restrict__ stats, int N, int Cin, int H,
                      int W) {
  __shared__ __align__(16) float sw[kFcMaxK][kFcCout];  // [tap * Cin + ci][co]
  __shared__ float sstat[2][kFcCout];
  const int K = 9 * Cin;
  for (int i = threadIdx.x; i < K * kFcCout; i += blockDim.x) {
    const int k = i / kFcCout, co = i - k * kFcCout;
    const int tap = k / Cin, ci = k - tap * Cin;
    sw[k][co] = w[(co * Cin + ci) * 9 + tap];  // OIHW
  }
  if (threadIdx.x < kFcCout) sstat[0][threadIdx.x] = sstat[1][threadIdx.x] = 0.f;
  __syncthreads();
  const long hw = static_cast<long>(H) * W;
  const long total = hw * N;
  const long nblocks = (total + blockDim.x - 1) / blockDim.x;
  for (long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const long p = blk * blockDim.x + threadIdx.x;
    const bool valid = p < total;
    float acc[kFcCout];
#pragma unroll
    for (int j = 0; j < kFcCout; ++j) acc[j] = 0.f;
    if (valid) {
      const long n = p / hw;
      const int r = static_cast<int>(p - n * hw);
      const int yy = r / W, xx = r - yy * W;
      for (int tap = 0; tap < 9; ++tap) {
        const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
        if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
        for (int ci = 0; ci < Cin; ++ci) {
          const float v = __ldg(x + (n * Cin + ci) * hw + static_cast<long>(sy) * W + sx);
          const float4* wr = reinterpret_cast<const float4*>(sw[tap * Cin + ci]);
#pragma unroll
          for (int j = 0; j < kFcCout / 4; ++j) {
            const float4 ww = wr[j];
            acc[4 * j + 0] = fmaf(v, ww.x, acc[4 * j + 0]);
            acc[4 * j + 1] = fmaf(v, ww.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(v, ww.z, acc[4 * j + 2]);
            acc[4 * j + 3] = fmaf(v, ww.w, acc[4 * j + 3]);
          }
        }
      }
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 u;
        u.x = pack_bf16x2(acc[8 * g + 0], acc[8 * g + 1]);
        u.y = pack_bf16x2(acc[8 * g + 2], acc[8 * g + 3]);
        u.z = pack_bf16x2(acc[8 * g + 4], acc[8 * g + 5]);
        u.w = pack_bf16x2(acc[8 * g + 6], acc[8 * g + 7]);
        y[p * 8 + g] = u;
      }
    }
    if (stats != nullptr) {
      const int lane = threadIdx.x & 31;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float s[32], s2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          s[j] = acc[half * 32 + j];
          s2[j] = s[j] * s[j];
        }
        const float cs = warp_colsum32(s);
        const float cs2 = warp_colsum32(s2);
        atomicAdd(&sstat[0][half * 32 + lane], cs);
        atomicAdd(&sstat[1][half * 32 + lane], cs2);
      }
    }
  }
  if (stats != nullptr) {
    __syncthreads();
    if (threadIdx.x < kFcCout) {
      atomicAdd(stats + threadIdx.x, sstat[0][threadIdx.x]);
      atomicAdd(stats + kFcCout + threadIdx.x, sstat[1][threadIdx.x]);
    }
  }
}

// wgrad: block = 256 threads = 4 pixel slices x 64 threads; thread (slice, t) owns a 4 co x 8 k
// register tile (co = 4*(t%16).., k = 8*(t/16)..; k = tap*Cin + ci, 32 slots >= 27) and walks every
// fourth pixel of a 128-pixel chunk staged in shared memory: 3 vector LDS per 32 FMAs.
constexpr int kFwPix = 128;
constexpr int kFwK = 32;

__global__ void __launch_bounds__(256)
first_conv_wgrad_kernel(const float* __restrict__ x, const uint4* __restrict__ dy,
                        float* __restrict__ dw, int N, int Cin, int H, int W) {
  __shared__ __align__(16) float sdy[kFwPix][kFcCout];      // 32 KB
  __shared__ __align__(16) float sx[kFwPix][kFwK];          // 16 KB
  const int K = 9 * Cin;
  const int t = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const int co0 = (t & 15) * 4, k0 = (t >> 4) * 8;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const long hw = static_cast<long>(H) * W;
  const long total = hw * N;
  const long nchunks = (total + kFwPix - 1) / kFwPix;
  // staging role of this thread: pixel pp = tid % 128 (fixed), k = tid / 128, +2, +4, ...
  const int spp = threadIdx.x & (kFwPix - 1), sk0 = threadIdx.x >> 7;
  for (long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long p0 = ch * kFwPix;
    __syncthreads();
    for (int i = threadIdx.x; i < kFwPix * 8; i += blockDim.x) {
      const int pp = i >> 3, g = i & 7;
      const long p = p0 + pp;
      uint4 u = make_uint4(0, 0, 0, 0);
      if (p < total) u = ldg_nc_v4(dy + p * 8 + g);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      float4* row = reinterpret_cast<float4*>(&sdy[pp][g * 8]);
      row[0] = make_float4(a.x, a.y, b.x, b.y);
      row[1] = make_float4(c.x, c.y, d.x, d.y);
    }
    {
      // one pixel per thread: decode (n, y, x) once, then walk this thread's k slots
      const long p = p0 + spp;
      const bool pv = p < total;
      const long n = pv ? p / hw : 0;
      const int r = static_cast<int>(p - n * hw);
      const int yy = r / W, xx = r - yy * W;
      const float* xb = x + n * Cin * hw;
      int tap = 0, ci = sk0;          // k = tap * Cin + ci, advanced by 2 each step
      while (ci >= Cin) { ci -= Cin; ++tap; }
      for (int k = sk0; k < kFwK; k += 2) {
        float v = 0.f;
        if (pv && k < K) {
          const int sy = yy + tap / 3 - 1, sxx = xx + tap % 3 - 1;
          if (sy >= 0 && sy < H && sxx >= 0 && sxx < W) v = __ldg(xb + ci * hw + static_cast<long>(sy) * W + sxx);
        }
        sx[spp][k] = v;
        ci += 2;
        while (ci >= Cin) { ci -= Cin; ++tap; }
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int pp = slice; pp < kFwPix; pp += 4) {
      const float4 g = *reinterpret_cast<const float4*>(&sdy[pp][co0]);
      const float4 x0 = *reinterpret_cast<const float4*>(&sx[pp][k0]);
      const float4 x1 = *reinterpret_cast<const float4*>(&sx[pp][k0 + 4]);
      const float gv[4] = {g.x, g.y, g.z, g.w};
      const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(gv[i], xv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      if (k < K) {
        const int tap = k / Cin, ci = k - tap * Cin;
        atomicAdd(dw + ((co0 + i) * Cin + ci) * 9 + tap, acc[i][j]);  // OIHW
      }
    }
}

// ----------------------------------------------------------------------------------------------
// Tensor-core route for the weight gradient: the image is re-laid-out as NHWC bf16 with the channels
// zero-padded to 64 (only the first 16-byte group of every 128-byte pixel row is ever written; the
// caller zeroes the buffer once) and the gradient is the ordinary conv3x3 weight gradient of a
// 64 -> 64 layer (nine-tap tcgen05 kernel, wgrad.cu MODE 2) followed by a 64*Cin*9-element gather into
// OIHW.  0.36 ms -> ~0.1 ms at 16 x 256 x 256 (the CUDA-core kernel above is LDS bound).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
image_to_nhwc64_kernel(const float* __restrict__ x, uint4* __restrict__ x64, int Cin, long hw, long total) {
  for (long p = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; p < total;
       p += static_cast<long>(gridDim.x) * blockDim.x) {
    const long n = p / hw, r = p - n * hw;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < Cin) v[c] = __ldg(x + (n * Cin + c) * hw + r);
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    x64[p * 8] = u;
  }
}

// scratch [64 co][9 taps][64 ci] -> dw OIHW [64][Cin][3][3]
__global__ void first_conv_gather_kernel(const float* __restrict__ scratch, float* __restrict__ dw, int Cin) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kFcCout * Cin * 9) return;
  const int tap = i % 9, ci = (i / 9) % Cin, co = i / (9 * Cin);
  dw[i] = scratch[(co * 9 + tap) * 64 + ci];
}

// OIHW fp32 [64][Cin][3][3] -> Wf[co][tap][64 ci] bf16, zero beyond Cin (B operand of the 64 -> 64 kernel)
__global__ void first_conv_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, int Cin) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kFcCout * 9 * 64) return;
  const int ci = i & 63, tap = (i >> 6) % 9, co = i / (9 * 64);
  wf[i] = __float2bfloat16(ci < Cin ? w[(co * Cin + ci) * 9 + tap] : 0.f);
}

int first_conv_pack(const float* w, void* wf, int Cin, cudaStream_t s) {
  UNETK_CHECK_ARG(w && wf, "first_conv_pack: null pointer");
  UNETK_CHECK_ARG(Cin >= 1 && Cin <= 8, "first_conv_pack: 1..8 input channels supported (got %d)", Cin);
  first_conv_pack_kernel<<<(kFcCout * 9 * 64 + 255) / 256, 256, 0, s>>>(w, static_cast<__nv_bfloat16*>(wf), Cin);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int image_to_nhwc64(const float* x, void* x64, int N, int Cin, int H, int W, cudaStream_t s) {
  UNETK_CHECK_ARG(x && x64, "image_to_nhwc64: null pointer");
  UNETK_CHECK_ARG(Cin >= 1 && Cin <= 8, "image_to_nhwc64: 1..8 input channels supported (got %d)", Cin);
  const long hw = static_cast<long>(H) * W, total = hw * N;
  long blocks = (total + 255) / 256;
  const long cap = static_cast<long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  image_to_nhwc64_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(x, static_cast<uint4*>(x64), Cin, hw, total);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int first_conv_wgrad_tc(const void* x64, const void* dy, float* scratch, float* dw, int N, int Cin, int H, int W,
                        int cout, cudaStream_t s) {
  UNETK_CHECK_ARG(x64 && dy && scratch && dw, "first_conv_wgrad_tc: null pointer");
  UNETK_CHECK_ARG(Cin >= 1 && Cin <= 8, "first_conv_wgrad_tc: 1..8 input channels supported (got %d)", Cin);
  UNETK_CHECK_ARG(cout == kFcCout, "first_conv_wgrad_tc: 64 output channels required (got %d)", cout);
  UNETK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * kFcCout * 9 * 64, s));
  if (int e = conv3x3_wgrad(dy, x64, 64, nullptr, 0, scratch, N, H, W, cout, 0, 0, s)) return e;
  const int n = kFcCout * Cin * 9;
  first_conv_gather_kernel<<<(n + 255) / 256, 256, 0, s>>>(scratch, dw, Cin);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int first_conv_fwd(const float* x, const float* w, void* y, float* stats, int N, int Cin, int H, int W,
                   int cout, cudaStream_t s) {
  UNETK_CHECK_ARG(x && w && y, "first_conv_fwd: null pointer");
  UNETK_CHECK_ARG(Cin >= 1 && Cin <= 4, "first_conv_fwd: 1..4 input channels supported (got %d)", Cin);
  UNETK_CHECK_ARG(cout == kFcCout, "first_conv_fwd: 64 output channels required (got %d)", cout);
  const long total = static_cast<long>(N) * H * W;
  long blocks = (total + 127) / 128;
  const long cap = static_cast<long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  first_conv_fwd_kernel<<<static_cast<int>(blocks), 128, 0, s>>>(x, w, static_cast<uint4*>(y), stats, N, Cin, H, W);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int first_conv_wgrad(const float* x, const void* dy, float* dw, int N, int Cin, int H, int W, int cout,
                     cudaStream_t s) {
  UNETK_CHECK_ARG(x && dy && dw, "first_conv_wgrad: null pointer");
  UNETK_CHECK_ARG(Cin >= 1 && Cin <= 4, "first_conv_wgrad: 1..4 input channels supported (got %d)", Cin);
  UNETK_CHECK_ARG(cout == kFcCout, "first_conv_wgrad: 64 output channels required (got %d)", cout);
  const long total = static_cast<long>(N) * H * W;
  long blocks = (total + kFwPix - 1) / kFwPix;
  const long cap = static_cast<long>(num_sms()) * 2;
  if (blocks > cap) blocks = cap;
  UNETK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * cout * Cin * 9, s));
  first_conv_wgrad_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(x, static_cast<const uint4*>(dy), dw, N, Cin, H, W);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

}  // namespace unetk
