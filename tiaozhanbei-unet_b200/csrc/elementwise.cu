// Bandwidth-bound kernels around the convolutions (sm_100a): BatchNorm finalize / apply / backward,
// ReLU, 2x2 max-pool with index save, NCHW<->NHWC layout conversion.
//
// Reference operators being replaced: nn.BatchNorm2d + nn.ReLU(inplace=True) at
// /root/reference/src/model.py:15-16,18-19 and nn.MaxPool2d(2) at src/model.py:32 (and their
// autograd).  All tensors are NHWC bf16 with C a multiple of 8; every thread moves 16-byte
// vectors (8 channels) and keeps its per-channel constants in registers: the grid-stride is a
// multiple of C/8, so a thread always sees the same 8 channels.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace unetk {

constexpr int kEwThreads = 256;

static inline int ew_grid(long nvec, int cgroups) {
  // total threads must be a multiple of the channel-group count (<= 256, divides 256 or is a
  // multiple of it handled by the caller's check)
  long blocks = (nvec + kEwThreads - 1) / kEwThreads;
  // blocks of 256 threads per SM: half occupancy.  Stand-alone (tools/power_probe_ew.py, 64 ch @256^2) the passes run at
  // 5.7 / 6.0 / 6.0 TB/s with 4 per SM against 5.5 / 5.7 / 5.9 with 8, at the same ~166 pJ per byte (the memory system, not
  // the SMs, is what their energy goes to); inside the step (100 steps, five runs each way on two boxes) 4 per SM
  // is 0.3-1 % faster: fewer resident blocks in the way of the tensor-core kernels they run beside.
  static const int per_sm = [] { const char* e = getenv("UNETK_EW_BLOCKS_PER_SM"); const int v = e ? atoi(e) : 4; return v >= 1 && v <= 8 ? v : 4; }();
  const long cap = static_cast<long>(num_sms()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  // (blocks * 256) % cgroups == 0 must hold: cgroups is a power of two <= 128 in this network, or
  // otherwise we round blocks up to a multiple of cgroups.
  if ((blocks * kEwThreads) % cgroups != 0) blocks = ((blocks + cgroups - 1) / cgroups) * cgroups;
  return static_cast<int>(blocks);
}

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  return u;
}

// ----------------------------------------------------------------------------------------------
// BatchNorm finalize: batch statistics -> (scale, shift, mean, invstd), running-stat update.
// torch semantics (nn.BatchNorm2d, training): biased variance for normalisation, unbiased for
// running_var, momentum 0.1, num_batches_tracked += 1.
// ----------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const BnFinalizeArgs f, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && f.num_batches_tracked != nullptr) *f.num_batches_tracked += 1;
  if (c >= C) return;
  float sc, sh;
  bn_finalize1(f, C, c, true, sc, sh);      // the arithmetic every fused consumer uses: bit-identical constants
}

// Eval mode: fold running statistics into (scale, shift).
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ running_mean,
                               const float* __restrict__ running_var, float* __restrict__ scale,
                               float* __restrict__ shift, int C, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] * rsqrtf(running_var[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - running_mean[c] * sc;
}

// Training mode, finalize fused into the consumer (saves 26 dependent ~2 us launches per step): every
// thread derives (scale, shift) of its 8 channels from the batch sums, exactly as bn_finalize_kernel does;
// the first `cgroups` threads of block 0 also store scale / shift / mean / invstd for the backward pass
// and update the running statistics.
__device__ __forceinline__ void bn_finalize8(const BnFinalizeArgs& f, int C, int cg, bool writer,
                                             float (&sc)[8], float (&sh)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) bn_finalize1(f, C, cg * 8 + j, writer, sc[j], sh[j]);
  if (writer && cg == 0 && f.num_batches_tracked != nullptr) *f.num_batches_tracked += 1;
}

// ----------------------------------------------------------------------------------------------
// a = relu(y * scale + shift)
// ----------------------------------------------------------------------------------------------
template <bool FIN>
__global__ void __launch_bounds__(kEwThreads)
bnrelu_apply_kernel(const uint4* __restrict__ y, const float* __restrict__ scale,
                    const float* __restrict__ shift, uint4* __restrict__ a, long nvec, int cgroups,
                    const BnFinalizeArgs fin) {
  pdl_trigger();
  pdl_wait();
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  const int cg = static_cast<int>(tid % cgroups);
  float sc[8], sh[8];
  if (FIN) {
    bn_finalize8(fin, cgroups * 8, cg, blockIdx.x == 0 && threadIdx.x < cgroups, sc, sh);
  } else {
    load8(scale + cg * 8, sc);
    load8(shift + cg * 8, sh);
  }
  for (long i = tid; i < nvec; i += nthreads) {
    float v[8];
    unpack8(ldg_nc_v4(y + i), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f);
    a[i] = pack8(v);
  }
}

// ----------------------------------------------------------------------------------------------
// a = relu(y * scale + shift) (optional), pooled = maxpool2x2(a), idx = argmax position (0..3,
// scan order (0,0),(0,1),(1,0),(1,1), first maximum wins like ATen's max_pool2d_with_indices).
// apply == 0: the input is already activated (eval path) and `a` is not written.
// ----------------------------------------------------------------------------------------------
template <bool FIN>
__global__ void __launch_bounds__(kEwThreads)
bnrelu_pool_kernel(const uint4* __restrict__ y, const float* __restrict__ scale,
                   const float* __restrict__ shift, uint4* __restrict__ a,
                   uint4* __restrict__ pooled, uint2* __restrict__ idx, int N, int H, int W,
                   int cgroups, int apply, const BnFinalizeArgs fin) {
  pdl_trigger();
  pdl_wait();
  const int Hp = H / 2, Wp = W / 2;
  const long nvec = static_cast<long>(N) * Hp * Wp * cgroups;
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  const int cg = static_cast<int>(tid % cgroups);
  float sc[8], sh[8];
  if (FIN) {
    bn_finalize8(fin, cgroups * 8, cg, blockIdx.x == 0 && threadIdx.x < cgroups, sc, sh);
  } else if (apply) {
    load8(scale + cg * 8, sc);
    load8(shift + cg * 8, sh);
  }
  for (long i = tid; i < nvec; i += nthreads) {
    long pp = i / cgroups;
    const int px = static_cast<int>(pp % Wp);
    pp /= Wp;
    const int py = static_cast<int>(pp % Hp);
    const int n = static_cast<int>(pp / Hp);
    const long base = ((static_cast<long>(n) * H + 2 * py) * W + 2 * px) * cgroups + cg;
    const long offs[4] = {base, base + cgroups, base + static_cast<long>(W) * cgroups,
                          base + static_cast<long>(W) * cgroups + cgroups};
    float best[8];
    uint32_t bi[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v[8];
      unpack8(ldg_nc_v4(y + offs[k]), v);
      if (apply) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f);
        a[offs[k]] = pack8(v);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (k == 0 || v[j] > best[j]) {
          best[j] = v[j];
          bi[j] = k;
        }
      }
    }
    pooled[i] = pack8(best);
    if (idx != nullptr) {
      uint2 u;
      u.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      u.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      idx[i] = u;
    }
  }
}

// dA[argmax position] (+)= dP (+ extra), other positions (+)= 0 (+ extra).  When `bn_y` is given, dA is final after this
// kernel and the first pass of the producing layer's BatchNorm+ReLU backward is fused in:
// sums[0..C) += sum g, sums[C..2C) += sum g*xhat  (g = dA*[y*scale+shift > 0], xhat = (y-mean)*invstd).
__global__ void __launch_bounds__(kEwThreads)
maxpool_bwd_kernel(const uint4* __restrict__ dp, const uint2* __restrict__ idx, uint4* da, int N,
                   int H, int W, int cgroups, int accumulate, const uint4* __restrict__ extra,
                   const uint4* __restrict__ bn_y,
                   const float* __restrict__ scale, const float* __restrict__ shift,
                   const float* __restrict__ mean, const float* __restrict__ invstd,
                   float* __restrict__ sums) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sacc[];  // [2][C] when bn_y != nullptr
  const int C = cgroups * 8;
  const bool bn = bn_y != nullptr;
  if (bn) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
  }
  const int Hp = H / 2, Wp = W / 2;
  const long nvec = static_cast<long>(N) * Hp * Wp * cgroups;
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  const int cg = static_cast<int>(tid % cgroups);
  float sc[8], sh[8], mu[8], is[8], s1[8], s2[8];
  if (bn) {
    load8(scale + cg * 8, sc);
    load8(shift + cg * 8, sh);
    load8(mean + cg * 8, mu);
    load8(invstd + cg * 8, is);
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  }
  for (long i = tid; i < nvec; i += nthreads) {
    long pp = i / cgroups;
    const int px = static_cast<int>(pp % Wp);
    pp /= Wp;
    const int py = static_cast<int>(pp % Hp);
    const int n = static_cast<int>(pp / Hp);
    const long base = ((static_cast<long>(n) * H + 2 * py) * W + 2 * px) * cgroups + cg;
    const long offs[4] = {base, base + cgroups, base + static_cast<long>(W) * cgroups,
                          base + static_cast<long>(W) * cgroups + cgroups};
    float g[8];
    unpack8(ldg_nc_v4(dp + i), g);
    const uint2 u = idx[i];
    uint32_t bi[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bi[j] = (u.x >> (8 * j)) & 0xff;
      bi[4 + j] = (u.y >> (8 * j)) & 0xff;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v[8];
      if (accumulate) {
        unpack8(da[offs[k]], v);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
      if (extra != nullptr) {       // a second gradient of the same activation (parallel decoder)
        float e[8];
        unpack8(ldg_nc_v4(extra + offs[k]), e);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += e[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += (bi[j] == static_cast<uint32_t>(k)) ? g[j] : 0.f;
      const uint4 packed = pack8(v);
      da[offs[k]] = packed;
      if (bn) {
        float r[8], yv[8];
        unpack8(packed, r);                       // the stored (bf16) gradient
        unpack8(ldg_nc_v4(bn_y + offs[k]), yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gg = (fmaf(yv[j], sc[j], sh[j]) > 0.f) ? r[j] : 0.f;
          s1[j] += gg;
          s2[j] = fmaf(gg, (yv[j] - mu[j]) * is[j], s2[j]);
        }
      }
    }
  }
  if (bn) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sacc[cg * 8 + j], s1[j]);
      atomicAdd(&sacc[C + cg * 8 + j], s2[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(sums + i, sacc[i]);
  }
}

// ----------------------------------------------------------------------------------------------
// BatchNorm + ReLU backward, pass 1: per-channel  s1 = sum g,  s2 = sum g * xhat
//   with g = dA * [y*scale + shift > 0],  xhat = (y - mean) * invstd.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads)
bn_bwd_reduce_kernel(const uint4* __restrict__ da, const uint4* __restrict__ y,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ mean, const float* __restrict__ invstd,
                     float* __restrict__ sums, long nvec, int cgroups) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sacc[];  // [2][C]
  const int C = cgroups * 8;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  const int cg = static_cast<int>(tid % cgroups);
  // The kernel is issue bound, not bandwidth bound (3.9 TB/s with ~9 instructions per element), so the
  // inner loop accumulates sum g and sum g*y only (5 instructions per element) and the affine map to
  // sum g*xhat = invstd * (sum g*y - mean * sum g) is applied once per thread.
  float sc[8], sh[8], s1[8], s2[8];
  load8(scale + cg * 8, sc);
  load8(shift + cg * 8, sh);
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  for (long i = tid; i < nvec; i += nthreads) {
    float g[8], v[8];
    unpack8(ldg_nc_v4(da + i), g);
    unpack8(ldg_nc_v4(y + i), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gg = (fmaf(v[j], sc[j], sh[j]) > 0.f) ? g[j] : 0.f;
      s1[j] += gg;
      s2[j] = fmaf(gg, v[j], s2[j]);
    }
  }
  {
    float mu[8], is[8];
    load8(mean + cg * 8, mu);
    load8(invstd + cg * 8, is);
#pragma unroll
    for (int j = 0; j < 8; ++j) s2[j] = is[j] * (s2[j] - mu[j] * s1[j]);
  }
  // lanes that share a channel group fold through shuffles first (when cgroups is a power of two < 32,
  // lanes l and l ^ o hold the same channels for every o >= cgroups)
  const bool fold = cgroups < 32 && (cgroups & (cgroups - 1)) == 0;
  for (int o = 16; fold && o >= cgroups; o >>= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
      s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
    }
  }
  if (!fold || (threadIdx.x & 31) < cgroups) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sacc[cg * 8 + j], s1[j]);
      atomicAdd(&sacc[C + cg * 8 + j], s2[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(sums + i, sacc[i]);
}

// pass 2: dY = scale * (g - s1/M - xhat * s2/M);  block 0 also writes dgamma = s2, dbeta = s1.
__global__ void __launch_bounds__(kEwThreads)
bn_bwd_apply_kernel(const uint4* __restrict__ da, const uint4* __restrict__ y,
                    const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ sums, const float* __restrict__ param_sums,
                    uint4* __restrict__ dy, float* __restrict__ dgamma, float* __restrict__ dbeta, long nvec,
                    int cgroups, float inv_count) {
  pdl_trigger();
  pdl_wait();
  const int C = cgroups * 8;
  if (blockIdx.x == 0) {      // parameter gradients: this rank's sums (`sums` are the cross-rank means under SyncBN)
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      dbeta[i] = param_sums[i];
      dgamma[i] = param_sums[C + i];
    }
  }
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  const int cg = static_cast<int>(tid % cgroups);
  // dY = scale*(g - m1 - xhat*m2) = scale*g + A*y + B with A = -scale*m2*invstd, B = -scale*m1 - A*mean:
  // 5 instructions per element instead of 9 (the kernel is issue bound like the reduction)
  float sc[8], sh[8], ca[8], cb[8];
  {
    float mu[8], is[8], m1[8], m2[8];
    load8(scale + cg * 8, sc);
    load8(shift + cg * 8, sh);
    load8(mean + cg * 8, mu);
    load8(invstd + cg * 8, is);
    load8(sums + cg * 8, m1);
    load8(sums + C + cg * 8, m2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ca[j] = -sc[j] * (m2[j] * inv_count) * is[j];
      cb[j] = -sc[j] * (m1[j] * inv_count) - ca[j] * mu[j];
    }
  }
  for (long i = tid; i < nvec; i += nthreads) {
    float g[8], v[8];
    unpack8(ldg_nc_v4(da + i), g);
    unpack8(ldg_nc_v4(y + i), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gg = (fmaf(v[j], sc[j], sh[j]) > 0.f) ? g[j] : 0.f;
      v[j] = fmaf(sc[j], gg, fmaf(ca[j], v[j], cb[j]));
    }
    dy[i] = pack8(v);
  }
}

// ----------------------------------------------------------------------------------------------
// layout conversion (for standalone blocks and tests): NCHW fp32 <-> NHWC bf16, 32x32 smem tiles
// ----------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                    int C, long HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long p0 = static_cast<long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r;
    const long p = p0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && p < HW) ? in[(static_cast<long>(n) * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long p = p0 + r;
    const int c = c0 + threadIdx.x;
    if (c < C && p < HW) out[(static_cast<long>(n) * HW + p) * C + c] = __float2bfloat16(tile[threadIdx.x][r]);
  }
}
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out,
                                    int C, long HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long p0 = static_cast<long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long p = p0 + r;
    const int c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && p < HW) ? __bfloat162float(in[(static_cast<long>(n) * HW + p) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r;
    const long p = p0 + threadIdx.x;
    if (c < C && p < HW) out[(static_cast<long>(n) * C + c) * HW + p] = tile[threadIdx.x][r];
  }
}

// ----------------------------------------------------------------------------------------------
// host launchers
// ----------------------------------------------------------------------------------------------
static int check_c(int C, const char* who) {
  UNETK_CHECK_ARG(C > 0 && C % 8 == 0, "%s: channel count must be a multiple of 8 (got %d)", who, C);
  return UNETK_OK;
}

int bn_finalize(const float* stats, const float* gamma, const float* beta, float* running_mean,
                float* running_var, long long* nbt, float* scale, float* shift, float* mean,
                float* invstd, int C, double count, float eps, float momentum, cudaStream_t s) {
  UNETK_CHECK_ARG(stats && gamma && beta && scale && shift && mean && invstd, "bn_finalize: null pointer");
  UNETK_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running_mean and running_var go together");
  UNETK_CHECK_ARG(count >= 1.0, "bn_finalize: empty batch");
  BnFinalizeArgs f{stats, gamma, beta, running_mean, running_var, nbt, scale, shift, mean, invstd, count, 0.0, 0.f, eps, momentum};
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(with_reciprocals(f), C);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int bn_fold(const float* gamma, const float* beta, const float* rm, const float* rv, float* scale,
            float* shift, int C, float eps, cudaStream_t s) {
  UNETK_CHECK_ARG(gamma && beta && rm && rv && scale && shift, "bn_fold: null pointer");
  bn_fold_kernel<<<(C + 127) / 128, 128, 0, s>>>(gamma, beta, rm, rv, scale, shift, C, eps);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int bnrelu_apply(const void* y, const float* scale, const float* shift, void* a, long npix, int C,
                 cudaStream_t s) {
  UNETK_CHECK_ARG(y && scale && shift && a, "bnrelu_apply: null pointer");
  if (int e = check_c(C, "bnrelu_apply")) return e;
  const long nvec = npix * (C / 8);
  UNETK_CUDA(launch_pdl(bnrelu_apply_kernel<false>, dim3(ew_grid(nvec, C / 8)), dim3(kEwThreads), 0, s, static_cast<const uint4*>(y), scale, shift,
                                                                         static_cast<uint4*>(a), nvec, C / 8, BnFinalizeArgs{}));
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

static int check_fin(const BnFinalizeArgs* f, int C, const char* who) {
  UNETK_CHECK_ARG(f && f->stats && f->gamma && f->beta && f->scale && f->shift && f->mean && f->invstd,
                  "%s: incomplete BatchNorm finalize arguments", who);
  UNETK_CHECK_ARG((f->running_mean == nullptr) == (f->running_var == nullptr), "%s: running_mean and running_var go together", who);
  UNETK_CHECK_ARG(C / 8 <= kEwThreads, "%s: at most %d channels", who, 8 * kEwThreads);
  UNETK_CHECK_ARG(f->count >= 1.0, "%s: empty batch", who);
  return UNETK_OK;
}

int bnrelu_apply_fin(const void* y, const BnFinalizeArgs* fin, void* a, long npix, int C, cudaStream_t s) {
  UNETK_CHECK_ARG(y && a, "bnrelu_apply_fin: null pointer");
  if (int e = check_c(C, "bnrelu_apply_fin")) return e;
  if (int e = check_fin(fin, C, "bnrelu_apply_fin")) return e;
  const long nvec = npix * (C / 8);
  UNETK_CUDA(launch_pdl(bnrelu_apply_kernel<true>, dim3(ew_grid(nvec, C / 8)), dim3(kEwThreads), 0, s, static_cast<const uint4*>(y), nullptr, nullptr,
                                                                        static_cast<uint4*>(a), nvec, C / 8, with_reciprocals(*fin)));
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int bnrelu_pool_fin(const void* y, const BnFinalizeArgs* fin, void* a, void* pooled, void* idx, int N, int H, int W,
                    int C, cudaStream_t s) {
  UNETK_CHECK_ARG(y && a && pooled, "bnrelu_pool_fin: null pointer");
  UNETK_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "bnrelu_pool_fin: H and W must be even (got %dx%d)", H, W);
  if (int e = check_c(C, "bnrelu_pool_fin")) return e;
  if (int e = check_fin(fin, C, "bnrelu_pool_fin")) return e;
  const long nvec = static_cast<long>(N) * (H / 2) * (W / 2) * (C / 8);
  UNETK_CUDA(launch_pdl(bnrelu_pool_kernel<true>, dim3(ew_grid(nvec, C / 8)), dim3(kEwThreads), 0, s, static_cast<const uint4*>(y), nullptr, nullptr,
                                                                       static_cast<uint4*>(a), static_cast<uint4*>(pooled),
                                                                       static_cast<uint2*>(idx), N, H, W, C / 8, 1, with_reciprocals(*fin)));
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int bnrelu_pool(const void* y, const float* scale, const float* shift, void* a, void* pooled,
                void* idx, int N, int H, int W, int C, int apply, cudaStream_t s) {
  UNETK_CHECK_ARG(y && pooled, "bnrelu_pool: null pointer");
  UNETK_CHECK_ARG(!apply || (scale && shift && a), "bnrelu_pool: apply needs scale, shift and a");
  // odd sizes pool with floor semantics (nn.MaxPool2d(2)); the last row / column is then outside every
  // window, so the fused BatchNorm+ReLU variant, which writes `a` window by window, needs even sizes
  UNETK_CHECK_ARG(!apply || (H % 2 == 0 && W % 2 == 0), "bnrelu_pool: the fused apply needs even H and W (got %dx%d): apply first, then pool", H, W);
  UNETK_CHECK_ARG(H >= 2 && W >= 2, "bnrelu_pool: input smaller than the 2x2 window");
  if (int e = check_c(C, "bnrelu_pool")) return e;
  const long nvec = static_cast<long>(N) * (H / 2) * (W / 2) * (C / 8);
  UNETK_CUDA(launch_pdl(bnrelu_pool_kernel<false>, dim3(ew_grid(nvec, C / 8)), dim3(kEwThreads), 0, s, static_cast<const uint4*>(y), scale, shift,
                                                                        static_cast<uint4*>(a), static_cast<uint4*>(pooled),
                                                                        static_cast<uint2*>(idx), N, H, W, C / 8, apply, BnFinalizeArgs{}));
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int maxpool_bwd(const void* dp, const void* idx, void* da, int N, int H, int W, int C, int accumulate,
                const void* extra, const BnReduceCtx* bn, cudaStream_t s) {
  UNETK_CHECK_ARG(dp && idx && da, "maxpool_bwd: null pointer");
  // odd sizes: the last row / column lies outside every window and keeps what it holds, which is only
  // right when the kernel accumulates into an existing gradient and adds no second tensor
  UNETK_CHECK_ARG((H % 2 == 0 && W % 2 == 0) || (accumulate && extra == nullptr),
                  "maxpool_bwd: odd H or W (%dx%d) needs accumulate != 0 and no extra tensor", H, W);
  if (int e = check_c(C, "maxpool_bwd")) return e;
  const long nvec = static_cast<long>(N) * (H / 2) * (W / 2) * (C / 8);
  const bool fuse = bn != nullptr && bn->y != nullptr;
  UNETK_CHECK_ARG(!fuse || (bn->scale && bn->shift && bn->mean && bn->invstd && bn->sums), "maxpool_bwd: incomplete BatchNorm context");
  int grid = ew_grid(nvec, C / 8);
  const int cap = num_sms() * 4;
  if (fuse && grid > cap && (static_cast<long>(cap) * kEwThreads) % (C / 8) == 0) grid = cap;   // fewer global atomics
  UNETK_CUDA(launch_pdl(maxpool_bwd_kernel, dim3(grid), dim3(kEwThreads), fuse ? 2 * C * sizeof(float) : 0, s,
      static_cast<const uint4*>(dp), static_cast<const uint2*>(idx), static_cast<uint4*>(da), N, H, W, C / 8, accumulate,
      static_cast<const uint4*>(extra), fuse ? static_cast<const uint4*>(bn->y) : nullptr, fuse ? bn->scale : nullptr, fuse ? bn->shift : nullptr,
      fuse ? bn->mean : nullptr, fuse ? bn->invstd : nullptr, fuse ? bn->sums : nullptr));
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int bn_bwd_reduce(const void* da, const void* y, const float* scale, const float* shift,
                  const float* mean, const float* invstd, float* sums, long npix, int C,
                  cudaStream_t s) {
  UNETK_CHECK_ARG(da && y && scale && shift && mean && invstd && sums, "bn_bwd_reduce: null pointer");
  if (int e = check_c(C, "bn_bwd_reduce")) return e;
  const long nvec = npix * (C / 8);
  // full grid (8 blocks per SM): at 4 per SM the two read streams ran at 3.9 TB/s, latency bound; an
  // unrolled variant with four loads in flight per thread was slower still (3.2 TB/s)
  // ... but every block ends with 2*C global atomics, so the deep layers (few pixels, many channels: 16x16x1024
  // x 16 images is 2 vectors per thread of a full grid against 2.4 M atomics) get a grid with at least 16
  // vectors per thread
  int grid = ew_grid(nvec, C / 8);
  {
    const long want = (nvec + 16L * kEwThreads - 1) / (16L * kEwThreads);
    if (want < grid) grid = ew_grid(want * kEwThreads, C / 8);
  }
  UNETK_CUDA(launch_pdl(bn_bwd_reduce_kernel, dim3(grid), dim3(kEwThreads), 2 * C * sizeof(float), s, static_cast<const uint4*>(da), static_cast<const uint4*>(y),
                                                                      scale, shift, mean, invstd, sums, nvec, C / 8));
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int bn_bwd_apply(const void* da, const void* y, const float* scale, const float* shift,
                 const float* mean, const float* invstd, const float* sums, const float* param_sums, void* dy,
                 float* dgamma, float* dbeta, long npix, int C, cudaStream_t s) {
  UNETK_CHECK_ARG(da && y && scale && shift && mean && invstd && sums && dy && dgamma && dbeta, "bn_bwd_apply: null pointer");
  if (int e = check_c(C, "bn_bwd_apply")) return e;
  const long nvec = npix * (C / 8);
  UNETK_CUDA(launch_pdl(bn_bwd_apply_kernel, dim3(ew_grid(nvec, C / 8)), dim3(kEwThreads), 0, s,
      static_cast<const uint4*>(da), static_cast<const uint4*>(y), scale, shift, mean, invstd, sums,
      param_sums != nullptr ? param_sums : sums, static_cast<uint4*>(dy), dgamma, dbeta, nvec, C / 8,
      1.0f / static_cast<float>(npix)));
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int nchw_to_nhwc(const float* in, void* out, int N, int C, int H, int W, cudaStream_t s) {
  UNETK_CHECK_ARG(in && out, "nchw_to_nhwc: null pointer");
  const long HW = static_cast<long>(H) * W;
  dim3 grid(static_cast<unsigned>((HW + 31) / 32), (C + 31) / 32, N), block(32, 8);
  nchw_to_nhwc_kernel<<<grid, block, 0, s>>>(in, static_cast<__nv_bfloat16*>(out), C, HW);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int nhwc_to_nchw(const void* in, float* out, int N, int C, int H, int W, cudaStream_t s) {
  UNETK_CHECK_ARG(in && out, "nhwc_to_nchw: null pointer");
  const long HW = static_cast<long>(H) * W;
  dim3 grid(static_cast<unsigned>((HW + 31) / 32), (C + 31) / 32, N), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, s>>>(static_cast<const __nv_bfloat16*>(in), out, C, HW);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

}  // namespace unetk
