// Optimizer step and weight (re)packing (sm_100a, bandwidth-bound).
//
//   adam_step      : torch.optim.Adam / AdamW / SGD(momentum) single-tensor semantics over one
//                    FLAT fp32 parameter buffer (all 98 tensors of AnomalyUNet are views of it),
//                    replacing get_optimizer at /root/reference/src/train_utils.py:263-272.
//   permute_batched: one launch that re-lays-out many small tensors, used for
//                    (a) fp32 OIHW weights -> bf16 kernel layouts ([Cout][tap][Cin] etc.) and
//                    (b) fp32 [Cout][tap][Cin] weight gradients -> fp32 OIHW (.grad layout).
#include "common.cuh"
#include "kernels.h"

namespace unetk {

// mode 0: Adam with L2 weight decay folded into the gradient (torch.optim.Adam)
// mode 1: AdamW (decoupled decay)          mode 2: SGD with momentum (torch.optim.SGD, dampening 0)
// one element of the update; returns the new parameter, m / v are updated in place
__device__ __forceinline__ float adam_update1(float p, float g, float& m, float& v, float lr, float beta1, float beta2,
                                              float eps, float wd, float step_size, float bc2_sqrt, float grad_scale,
                                              int mode, int first_step) {
  float gg = __fmul_rn(g, grad_scale);
  if (mode == 0) {
    gg = fmaf(wd, p, gg);
  } else if (mode == 1) {
    p = __fmul_rn(p, __fsub_rn(1.f, __fmul_rn(lr, wd)));
  }
  if (mode == 2) {
    gg = fmaf(wd, p, gg);
    m = first_step ? gg : fmaf(beta1, m, gg);
    return __fmaf_rn(-lr, m, p);
  }
  // explicit roundings (no compiler-chosen contractions): the stand-alone and the fused kernels agree bit for bit
  m = __fmaf_rn(1.f - beta1, __fsub_rn(gg, m), m);                                    // lerp_, as torch does
  v = __fmaf_rn(__fmul_rn(1.f - beta2, gg), gg, __fmul_rn(beta2, v));                 // mul_ + addcmul_
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);
  return __fmaf_rn(-step_size, __fdiv_rn(m, denom), p);                               // addcdiv_
}

// Two float4 groups per thread and iteration: eight 16-byte loads in flight before the first store (one group
// per iteration ran at 4.2 TB/s of the 28 bytes per parameter, latency bound).
__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                 float* __restrict__ v, long n, float lr, float beta1, float beta2, float eps,
                 float wd, float bc1, float bc2_sqrt, float grad_scale, int mode, int first_step) {
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  const float step_size = lr / bc1;
  const long n4 = n / 4;                                // full float4 groups
  constexpr int U = 2;
  for (long i0 = tid; i0 < n4; i0 += nthreads * U) {
    float4 pv[U], gv[U], mv[U], vv[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long i = i0 + u * nthreads;
      ok[u] = i < n4;
      if (ok[u]) {
        pv[u] = reinterpret_cast<const float4*>(p)[i];
        gv[u] = reinterpret_cast<const float4*>(g)[i];
        mv[u] = reinterpret_cast<const float4*>(m)[i];
        vv[u] = (mode != 2) ? reinterpret_cast<const float4*>(v)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      const long i = i0 + u * nthreads;
      pv[u].x = adam_update1(pv[u].x, gv[u].x, mv[u].x, vv[u].x, lr, beta1, beta2, eps, wd, step_size, bc2_sqrt, grad_scale, mode, first_step);
      pv[u].y = adam_update1(pv[u].y, gv[u].y, mv[u].y, vv[u].y, lr, beta1, beta2, eps, wd, step_size, bc2_sqrt, grad_scale, mode, first_step);
      pv[u].z = adam_update1(pv[u].z, gv[u].z, mv[u].z, vv[u].z, lr, beta1, beta2, eps, wd, step_size, bc2_sqrt, grad_scale, mode, first_step);
      pv[u].w = adam_update1(pv[u].w, gv[u].w, mv[u].w, vv[u].w, lr, beta1, beta2, eps, wd, step_size, bc2_sqrt, grad_scale, mode, first_step);
      reinterpret_cast<float4*>(p)[i] = pv[u];
      reinterpret_cast<float4*>(m)[i] = mv[u];
      if (mode != 2) reinterpret_cast<float4*>(v)[i] = vv[u];
    }
  }
  // the last n % 4 elements
  for (long i = n4 * 4 + tid; i < n; i += nthreads) {
    float mm = m[i], vv1 = (mode != 2) ? v[i] : 0.f;
    p[i] = adam_update1(p[i], g[i], mm, vv1, lr, beta1, beta2, eps, wd, step_size, bc2_sqrt, grad_scale, mode, first_step);
    m[i] = mm;
    if (mode != 2) v[i] = vv1;
  }
}

// out (contiguous [d0][d1][d2]) <- in[i0*s0 + f(i1)*s1 + i2*s2], f flips i1 when flip1 != 0.
struct PermuteDesc {
  long long in_off;   // element offset into the fp32 source buffer
  long long out_off;  // element offset into the destination buffer
  long long s0, s1, s2;
  int d0, d1, d2;
  int flip1;
  int chunk_begin;    // first chunk id of this tensor
  int pad;
};
static_assert(sizeof(PermuteDesc) == 64, "PermuteDesc layout is part of the C ABI");

constexpr int kPermChunk = 2048;

template <typename OutT>
__global__ void __launch_bounds__(256)
permute_batched_kernel(const PermuteDesc* __restrict__ descs, int n_desc,
                       const float* __restrict__ in, OutT* __restrict__ out, float scale) {
  // binary search the tensor that owns this chunk
  int lo = 0, hi = n_desc - 1;
  const int chunk = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].chunk_begin <= chunk) lo = mid; else hi = mid - 1;
  }
  const PermuteDesc d = descs[lo];
  const long total = static_cast<long>(d.d0) * d.d1 * d.d2;
  const long begin = static_cast<long>(chunk - d.chunk_begin) * kPermChunk;
  const long end = min(begin + kPermChunk, total);
  for (long i = begin + threadIdx.x; i < end; i += blockDim.x) {
    const int i2 = static_cast<int>(i % d.d2);
    const long t = i / d.d2;
    int i1 = static_cast<int>(t % d.d1);
    const int i0 = static_cast<int>(t / d.d1);
    if (d.flip1) i1 = d.d1 - 1 - i1;
    const float v = in[d.in_off + i0 * d.s0 + i1 * d.s1 + i2 * d.s2] * scale;
    if constexpr (sizeof(OutT) == 2) out[d.out_off + i] = __float2bfloat16(v);
    else out[d.out_off + i] = v;
  }
}

// ----------------------------------------------------------------------------------------------
// Tiled conv3x3 weight packing / gradient un-packing (the generic permute above reads with a
// 36-byte stride; these move >= 64-byte runs on both sides through a shared-memory tile).
// ----------------------------------------------------------------------------------------------
struct ConvPackDesc {
  long long src_off;   // pack: OIHW fp32 weight in the flat parameter buffer | unpack: staged [co][tap][ci] gradient
  long long dst_off;   // pack: Wf [co][tap][ci] bf16                        | unpack: OIHW fp32 gradient
  long long dst2_off;  // pack: Wd [ci][8-tap][co] bf16                      | unpack: unused
  int cout, cin;
  int block_begin;     // first block id of this layer
  int pad[7];
};
static_assert(sizeof(ConvPackDesc) == 64, "ConvPackDesc layout is part of the C ABI");

__device__ __forceinline__ ConvPackDesc find_layer(const ConvPackDesc* descs, int n, int block) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].block_begin <= block) lo = mid; else hi = mid - 1;
  }
  return descs[lo];
}

// one block = 32 output channels x 32 input channels x 9 taps
__global__ void __launch_bounds__(256)
pack_conv_weights_kernel(const ConvPackDesc* __restrict__ descs, int n_desc, const float* __restrict__ flat_p,
                         __nv_bfloat16* __restrict__ wpack) {
  __shared__ float tile[32][289];
  const ConvPackDesc d = find_layer(descs, n_desc, blockIdx.x);
  const int tiles_ci = d.cin / 32;
  const int tix = blockIdx.x - d.block_begin;
  const int co0 = (tix / tiles_ci) * 32, ci0 = (tix % tiles_ci) * 32;
  for (int i = threadIdx.x; i < 32 * 288; i += blockDim.x) {
    const int r = i / 288, j = i - r * 288;
    tile[r][j] = flat_p[d.src_off + (static_cast<long long>(co0 + r) * d.cin + ci0) * 9 + j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 288; i += blockDim.x) {
    const int c = i & 31, rt = i >> 5;          // rt = r * 9 + tap
    const int r = rt / 9, tap = rt - r * 9;
    wpack[d.dst_off + (static_cast<long long>(co0 + r) * 9 + tap) * d.cin + ci0 + c] = __float2bfloat16(tile[r][c * 9 + tap]);
  }
  for (int i = threadIdx.x; i < 32 * 288; i += blockDim.x) {
    const int r = i & 31, ct = i >> 5;          // ct = c * 9 + tap'
    const int c = ct / 9, tapf = ct - c * 9;    // tap' = 8 - tap (180-degree rotation for dgrad)
    wpack[d.dst2_off + (static_cast<long long>(ci0 + c) * 9 + tapf) * d.cout + co0 + r] = __float2bfloat16(tile[r][c * 9 + (8 - tapf)]);
  }
}

// one block = one output channel: [tap][ci] -> [ci][tap], scaled
__global__ void __launch_bounds__(256)
unpack_conv_grads_kernel(const ConvPackDesc* __restrict__ descs, int n_desc, const float* __restrict__ gstage,
                         float* __restrict__ flat_g, float scale) {
  extern __shared__ float row[];               // [9][cin + 1]
  const ConvPackDesc d = find_layer(descs, n_desc, blockIdx.x);
  const int co = blockIdx.x - d.block_begin;
  const int n = 9 * d.cin, pitch = d.cin + 1;
  const float* src = gstage + d.src_off + static_cast<long long>(co) * n;
  float* dst = flat_g + d.dst_off + static_cast<long long>(co) * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int tap = i / d.cin, ci = i - tap * d.cin;
    row[tap * pitch + ci] = src[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ci = i / 9, tap = i - ci * 9;
    dst[i] = row[tap * pitch + ci] * scale;
  }
}

// ----------------------------------------------------------------------------------------------
// Fused parameter update: ONE pass over a group of tensors does what used to be three --
//   (1) staged gradient -> `.grad` in the reference layout, times the data-parallel average,
//   (2) the optimizer step (adam_update1: bit-identical arithmetic to adam_step_kernel),
//   (3) the new fp32 weights -> the bf16 kernel layouts the NEXT forward / backward pass reads,
// and zeroes the staging range it has read (the weight-gradient kernels accumulate into it with
// red.global.add), so the next step needs neither a packing pass nor a memset.  36 bytes per conv
// weight instead of 8 (un-pack) + 28 (Adam) + 8 (pack) + 4 (memset), in one launch per group that
// the engine releases as soon as the group's gradients are final, under the rest of the backward pass.
// The hyper-parameters come from device memory (hyper[]: the launch is part of a CUDA graph whose
// replays see a new step count / learning rate every time).
// ----------------------------------------------------------------------------------------------
struct UpdateDesc {
  long long off;      // element offset of the tensor in the flat p / g / m / v / staging buffers
  long long wf_off;   // bf16 element offset in wpack: conv Wf[co][tap][ci] | convT Wt[tap][co][ci]
  long long wd_off;   //                               conv Wd[ci][8-tap][co] | convT Wtd[ci][tap][co]
  int kind;           // 0 plain (same layout everywhere, no packed copy), 1 conv3x3 OIHW, 2 convT [ci][co][2][2]
  int cout, cin;      // plain: cout = number of elements
  int block_begin;    // first work unit of this tensor
  int pad[6];
};
static_assert(sizeof(UpdateDesc) == 64, "UpdateDesc layout is part of the C ABI");

constexpr int kUpdPlainChunk = 2048;
constexpr int kUpdPitch9 = 290;                  // conv3x3 tile: [32 co][32 ci x 9 taps]; even pitch: 8-byte shared accesses
constexpr int kUpdPitch4 = 130;                  // transposed-conv tile: [32 ci][32 co x 4 taps]
constexpr int kUpdMaxDesc = 128;

struct Hyper {
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, grad_scale;
  int first_step;
};

__device__ __forceinline__ Hyper load_hyper(const float* __restrict__ h) {
  Hyper y;
  y.lr = h[0]; y.beta1 = h[1]; y.beta2 = h[2]; y.eps = h[3]; y.wd = h[4]; y.bc1 = h[5]; y.bc2_sqrt = h[6];
  y.first_step = h[7] != 0.f;
  y.grad_scale = h[8];
  return y;
}

// Staged gradients of one tile (ROWS rows of 32 contiguous floats, row i at gtile + i * row_stride) -> the shared
// tile, transposed into the parameter layout (slot(row, c) = position of element c of row `row`), and zeroes over
// what was read.  Every thread issues all of its 16-byte loads before the first store (the staging rows are 128-byte
// pieces a few KB apart: DRAM latency is what this pass waits for).
template <int ROWS, typename Slot>
__device__ __forceinline__ void stage_tile(float* __restrict__ gtile, int row_stride, float* tile, Slot slot) {
  constexpr int PER = ROWS * 8 / 256;           // float4 groups per thread: 9 / 4
  static_assert(ROWS * 8 % 256 == 0, "tile size");
  float4 val[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int q = threadIdx.x + k * 256;
    val[k] = *reinterpret_cast<const float4*>(gtile + (q >> 3) * row_stride + (q & 7) * 4);
  }
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int q = threadIdx.x + k * 256;
    const int row = q >> 3, c = (q & 7) * 4;
    *reinterpret_cast<float4*>(gtile + row * row_stride + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    tile[slot(row, c)] = val[k].x;
    tile[slot(row, c + 1)] = val[k].y;
    tile[slot(row, c + 2)] = val[k].z;
    tile[slot(row, c + 3)] = val[k].w;
  }
}

// 32 rows x ROWLEN contiguous fp32 parameters (row r at p + r * row_stride); tile[r * (ROWLEN + 2) + j] holds the
// element's staged gradient on entry and its new value on exit (same thread, same slot: updated in place).
template <int ROWLEN>
__device__ __forceinline__ void update_rows(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                            float* __restrict__ v, int row_stride, float* tile, const Hyper& y, int mode) {
  constexpr int Q = ROWLEN / 4;                 // float4 groups per row
  constexpr int PER = 32 * Q / 256;             // groups per thread: 9 / 4
  static_assert(32 * Q % 256 == 0, "tile size");
  const float step_size = y.lr / y.bc1;
  constexpr int U = (PER % 3 == 0) ? 3 : 2;     // groups in flight per thread: 9 / 6 16-byte loads before the first store
#pragma unroll
  for (int k0 = 0; k0 < PER; k0 += U) {
    float4 pv[U], mv[U], vv[U];
    int at[U];                                  // (offsets inside the tile's 32 rows: < 2^31)
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int q = threadIdx.x + (k0 + u) * 256;
      const int r = q / Q, j4 = q - r * Q;
      at[u] = r * row_stride + j4 * 4;
      pv[u] = *reinterpret_cast<const float4*>(p + at[u]);
      mv[u] = *reinterpret_cast<const float4*>(m + at[u]);
      vv[u] = (mode != 2) ? *reinterpret_cast<const float4*>(v + at[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int q = threadIdx.x + (k0 + u) * 256;
      const int r = q / Q, j = (q - r * Q) * 4;
      float2* slot = reinterpret_cast<float2*>(tile + r * (ROWLEN + 2) + j);      // (even offset: 8-byte aligned)
      const float2 g01 = slot[0], g23 = slot[1];
      float4 gg;
      gg.x = g01.x * y.grad_scale;
      gg.y = g01.y * y.grad_scale;
      gg.z = g23.x * y.grad_scale;
      gg.w = g23.y * y.grad_scale;
      pv[u].x = adam_update1(pv[u].x, gg.x, mv[u].x, vv[u].x, y.lr, y.beta1, y.beta2, y.eps, y.wd, step_size, y.bc2_sqrt, 1.f, mode, y.first_step);
      pv[u].y = adam_update1(pv[u].y, gg.y, mv[u].y, vv[u].y, y.lr, y.beta1, y.beta2, y.eps, y.wd, step_size, y.bc2_sqrt, 1.f, mode, y.first_step);
      pv[u].z = adam_update1(pv[u].z, gg.z, mv[u].z, vv[u].z, y.lr, y.beta1, y.beta2, y.eps, y.wd, step_size, y.bc2_sqrt, 1.f, mode, y.first_step);
      pv[u].w = adam_update1(pv[u].w, gg.w, mv[u].w, vv[u].w, y.lr, y.beta1, y.beta2, y.eps, y.wd, step_size, y.bc2_sqrt, 1.f, mode, y.first_step);
      *reinterpret_cast<float4*>(g + at[u]) = gg;
      *reinterpret_cast<float4*>(p + at[u]) = pv[u];
      *reinterpret_cast<float4*>(m + at[u]) = mv[u];
      if (mode != 2) *reinterpret_cast<float4*>(v + at[u]) = vv[u];
      slot[0] = make_float2(pv[u].x, pv[u].y);
      slot[1] = make_float2(pv[u].z, pv[u].w);
    }
  }
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
}

// The lines the NEXT work unit of this block will load (its staged gradient tile and its rows of p / m / v), pulled
// into L2 while the current unit packs its weights: a block's loads then cost an L2 round trip instead of a DRAM one
// (the kernel was latency bound at 42 % of the DRAM bandwidth with four blocks per SM: profiles/ncu_fused_update_r02_summary.txt).
__device__ __forceinline__ void prefetch_unit(const UpdateDesc* __restrict__ descs, const int* begins, int n_desc, int unit,
                                              const float* p, const float* m, const float* v, const float* gstage, int mode) {
  int lo = 0, hi = n_desc - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (begins[mid] <= unit) lo = mid; else hi = mid - 1;
  }
  const long long off = descs[lo].off;
  const int kind = descs[lo].kind, cout = descs[lo].cout, cin = descs[lo].cin;
  const int tix = unit - begins[lo];
  if (kind == 0) return;                       // (small tensors)
  int taps, rows;                              // rows of 128 bytes in the staging tile
  long long gbase, pbase, prow;
  if (kind == 1) {
    const int tiles_ci = cin / 32;
    const int co0 = (tix / tiles_ci) * 32, ci0 = (tix % tiles_ci) * 32;
    taps = 9; rows = 288;
    gbase = off + static_cast<long long>(co0) * 9 * cin + ci0;
    pbase = off + (static_cast<long long>(co0) * cin + ci0) * 9;
    prow = static_cast<long long>(cin) * 9;
  } else {
    const int tiles_co = cout / 32;
    const int ci0 = (tix / tiles_co) * 32, co0 = (tix % tiles_co) * 32;
    taps = 4; rows = 128;
    gbase = off + static_cast<long long>(co0) * 4 * cin + ci0;
    pbase = off + (static_cast<long long>(ci0) * cout + co0) * 4;
    prow = static_cast<long long>(cout) * 4;
  }
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    prefetch_l2(gstage + gbase + static_cast<long long>(i) * cin);
    // the same count of 128-byte lines in each of p / m / v: 32 rows of `taps` lines
    const long long a = pbase + (i / taps) * prow + (i % taps) * 32;
    prefetch_l2(p + a);
    prefetch_l2(m + a);
    if (mode != 2) prefetch_l2(v + a);
  }
}

__global__ void __launch_bounds__(256, 4)
fused_update_kernel(const UpdateDesc* __restrict__ descs, int n_desc, int n_units, float* __restrict__ p,
                    float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float* __restrict__ gstage,
                    __nv_bfloat16* __restrict__ wpack, const float* __restrict__ hyper, int mode) {
  __shared__ __align__(16) float tile[32 * kUpdPitch9];
  __shared__ int begins[kUpdMaxDesc];
  for (int i = threadIdx.x; i < n_desc; i += blockDim.x) begins[i] = descs[i].block_begin;
  const Hyper y = load_hyper(hyper);
  __syncthreads();
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    int lo = 0, hi = n_desc - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (begins[mid] <= unit) lo = mid; else hi = mid - 1;
    }
    const long long off = descs[lo].off;
    const int kind = descs[lo].kind, cout = descs[lo].cout, cin = descs[lo].cin;
    const int tix = unit - begins[lo];
    if (kind == 0) {
      const float step_size = y.lr / y.bc1;
      const int begin = tix * kUpdPlainChunk;
      const int end = min(begin + kUpdPlainChunk, cout);
      if (end - begin == kUpdPlainChunk) {
        // a full chunk (tensor offsets are 256-byte aligned): two 16-byte groups per thread, every load before the first store
        constexpr int G = kUpdPlainChunk / 4 / 256;
        float4 gv[G], pv[G], mv[G], vv[G];
#pragma unroll
        for (int k = 0; k < G; ++k) {
          const long long a = off + begin + (threadIdx.x + k * 256) * 4;
          gv[k] = *reinterpret_cast<const float4*>(gstage + a);
          pv[k] = *reinterpret_cast<const float4*>(p + a);
          mv[k] = *reinterpret_cast<const float4*>(m + a);
          vv[k] = (mode != 2) ? *reinterpret_cast<const float4*>(v + a) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
          const long long a = off + begin + (threadIdx.x + k * 256) * 4;
          gv[k].x *= y.grad_scale; gv[k].y *= y.grad_scale; gv[k].z *= y.grad_scale; gv[k].w *= y.grad_scale;
          pv[k].x = adam_update1(pv[k].x, gv[k].x, mv[k].x, vv[k].x, y.lr, y.beta1, y.beta2, y.eps, y.wd, step_size, y.bc2_sqrt, 1.f, mode, y.first_step);
          pv[k].y = adam_update1(pv[k].y, gv[k].y, mv[k].y, vv[k].y, y.lr, y.beta1, y.beta2, y.eps, y.wd, step_size, y.bc2_sqrt, 1.f, mode, y.first_step);
          pv[k].z = adam_update1(pv[k].z, gv[k].z, mv[k].z, vv[k].z, y.lr, y.beta1, y.beta2, y.eps, y.wd, step_size, y.bc2_sqrt, 1.f, mode, y.first_step);
          pv[k].w = adam_update1(pv[k].w, gv[k].w, mv[k].w, vv[k].w, y.lr, y.beta1, y.beta2, y.eps, y.wd, step_size, y.bc2_sqrt, 1.f, mode, y.first_step);
          *reinterpret_cast<float4*>(gstage + a) = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(g + a) = gv[k];
          *reinterpret_cast<float4*>(p + a) = pv[k];
          *reinterpret_cast<float4*>(m + a) = mv[k];
          if (mode != 2) *reinterpret_cast<float4*>(v + a) = vv[k];
        }
        continue;
      }
      for (int i = begin + threadIdx.x; i < end; i += blockDim.x) {
        const long long a = off + i;
        const float gv = gstage[a] * y.grad_scale;
        gstage[a] = 0.f;
        float mm = m[a], vv = (mode != 2) ? v[a] : 0.f;
        g[a] = gv;
        p[a] = adam_update1(p[a], gv, mm, vv, y.lr, y.beta1, y.beta2, y.eps, y.wd, step_size, y.bc2_sqrt, 1.f, mode, y.first_step);
        m[a] = mm;
        if (mode != 2) v[a] = vv;
      }
      continue;                   // (no shared-memory tile involved)
    }
    const long long wf_off = descs[lo].wf_off, wd_off = descs[lo].wd_off;
    if (kind == 1) {
      // conv3x3: 32 output channels x 32 input channels x 9 taps.  staged dW[co][tap][ci]; parameter OIHW
      const int tiles_ci = cin / 32;
      const int co0 = (tix / tiles_ci) * 32, ci0 = (tix % tiles_ci) * 32;
      // staged row i = r * 9 + tap (rows of one output channel are consecutive taps, cin apart)
      stage_tile<288>(gstage + off + static_cast<long long>(co0) * 9 * cin + ci0, cin, tile,
                      [](int row, int c) { const int r = row / 9; return r * kUpdPitch9 + c * 9 + (row - r * 9); });
      __syncthreads();
      const long long pb = off + (static_cast<long long>(co0) * cin + ci0) * 9;
      update_rows<288>(p + pb, g + pb, m + pb, mode != 2 ? v + pb : v, cin * 9, tile, y, mode);
      __syncthreads();
      if (unit + static_cast<int>(gridDim.x) < n_units) prefetch_unit(descs, begins, n_desc, unit + gridDim.x, p, m, v, gstage, mode);
      // packed copies, eight bf16 (16 bytes) per store: Wf[co][tap][ci0 + 8 q ..], Wd[ci][8 - tap][co0 + 8 q ..]
      for (int i = threadIdx.x; i < 32 * 9 * 4; i += blockDim.x) {
        const int q = i & 3, rt = i >> 2;
        const int r = rt / 9, tap = rt - r * 9;
        const float* src = tile + r * kUpdPitch9 + q * 72 + tap;          // element c = 8 q + e at + 9 e
        uint4 w;
        w.x = pack_bf16x2(src[0], src[9]);  w.y = pack_bf16x2(src[18], src[27]);
        w.z = pack_bf16x2(src[36], src[45]); w.w = pack_bf16x2(src[54], src[63]);
        *reinterpret_cast<uint4*>(wpack + wf_off + (static_cast<long long>(co0 + r) * 9 + tap) * cin + ci0 + q * 8) = w;
      }
      for (int i = threadIdx.x; i < 32 * 9 * 4; i += blockDim.x) {
        const int q = i & 3, ct = i >> 2;
        const int c = ct / 9, tapf = ct - c * 9;
        const float* src = tile + (q * 8) * kUpdPitch9 + c * 9 + (8 - tapf);   // element r = 8 q + e at + pitch e
        uint4 w;
        w.x = pack_bf16x2(src[0], src[kUpdPitch9]);
        w.y = pack_bf16x2(src[2 * kUpdPitch9], src[3 * kUpdPitch9]);
        w.z = pack_bf16x2(src[4 * kUpdPitch9], src[5 * kUpdPitch9]);
        w.w = pack_bf16x2(src[6 * kUpdPitch9], src[7 * kUpdPitch9]);
        *reinterpret_cast<uint4*>(wpack + wd_off + (static_cast<long long>(ci0 + c) * 9 + tapf) * cout + co0 + q * 8) = w;
      }
    } else {
      // transposed conv 2x2: 32 input channels x 32 output channels x 4 taps.  staged dW[co][tap][ci]; parameter [ci][co][tap]
      const int tiles_co = cout / 32;
      const int ci0 = (tix / tiles_co) * 32, co0 = (tix % tiles_co) * 32;
      // staged row i = c * 4 + tap holds 32 input channels r: tile[r][c * 4 + tap]
      stage_tile<128>(gstage + off + static_cast<long long>(co0) * 4 * cin + ci0, cin, tile,
                      [](int row, int c) { return c * kUpdPitch4 + row; });
      __syncthreads();
      const long long pb = off + (static_cast<long long>(ci0) * cout + co0) * 4;
      update_rows<128>(p + pb, g + pb, m + pb, mode != 2 ? v + pb : v, cout * 4, tile, y, mode);
      __syncthreads();
      if (unit + static_cast<int>(gridDim.x) < n_units) prefetch_unit(descs, begins, n_desc, unit + gridDim.x, p, m, v, gstage, mode);
      // Wt[tap][co][ci0 + 8 q ..] and Wtd[ci][tap][co0 + 8 q ..], 16 bytes per store
      for (int i = threadIdx.x; i < 128 * 4; i += blockDim.x) {
        const int q = i & 3, ct = i >> 2;      // ct = c * 4 + tap
        const int c = ct >> 2, tap = ct & 3;
        const float* src = tile + (q * 8) * kUpdPitch4 + ct;                   // element r = 8 q + e at + pitch e
        uint4 w;
        w.x = pack_bf16x2(src[0], src[kUpdPitch4]);
        w.y = pack_bf16x2(src[2 * kUpdPitch4], src[3 * kUpdPitch4]);
        w.z = pack_bf16x2(src[4 * kUpdPitch4], src[5 * kUpdPitch4]);
        w.w = pack_bf16x2(src[6 * kUpdPitch4], src[7 * kUpdPitch4]);
        *reinterpret_cast<uint4*>(wpack + wf_off + (static_cast<long long>(tap) * cout + co0 + c) * cin + ci0 + q * 8) = w;
      }
      for (int i = threadIdx.x; i < 128 * 4; i += blockDim.x) {
        const int q = i & 3, rt = i >> 2;      // rt = r * 4 + tap
        const int r = rt >> 2, tap = rt & 3;
        const float* src = tile + r * kUpdPitch4 + q * 32 + tap;               // element c = 8 q + e at + 4 e
        uint4 w;
        w.x = pack_bf16x2(src[0], src[4]);   w.y = pack_bf16x2(src[8], src[12]);
        w.z = pack_bf16x2(src[16], src[20]); w.w = pack_bf16x2(src[24], src[28]);
        *reinterpret_cast<uint4*>(wpack + wd_off + (static_cast<long long>(ci0 + r) * 4 + tap) * cout + co0 + q * 8) = w;
      }
    }
    __syncthreads();        // the tile is reused by the next unit
  }
}

struct FloatVals { float v[16]; };
__global__ void set_floats_kernel(float* dst, FloatVals vals, int n) {
  if (threadIdx.x < n) dst[threadIdx.x] = vals.v[threadIdx.x];
}

int set_floats(float* dst, const float* host_vals, int n, cudaStream_t s) {
  UNETK_CHECK_ARG(dst && host_vals && n > 0 && n <= 16, "set_floats: 1..16 values");
  FloatVals vals{};
  for (int i = 0; i < n; ++i) vals.v[i] = host_vals[i];
  set_floats_kernel<<<1, 32, 0, s>>>(dst, vals, n);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int fused_update(const void* descs, int n_desc, int n_blocks, int max_blocks, float* p, float* g, float* m, float* v,
                 float* gstage, void* wpack, const float* hyper, int mode, cudaStream_t s) {
  UNETK_CHECK_ARG(descs && p && g && m && gstage && wpack && hyper && n_desc > 0 && n_blocks > 0, "fused_update: bad arguments");
  UNETK_CHECK_ARG(mode >= 0 && mode <= 2, "fused_update: unknown mode %d", mode);
  UNETK_CHECK_ARG(mode == 2 || v != nullptr, "fused_update: second-moment buffer missing");
  UNETK_CHECK_ARG(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                    reinterpret_cast<uintptr_t>(v)) & 15) == 0, "fused_update: buffers must be 16-byte aligned");
  // max_blocks > 0: a bounded grid that walks the work units (a launch that runs beside the backward pass must leave
  // the SMs' block slots to the kernels other streams wait for); 0: one block per unit
  // ... and at most four blocks per SM (what is resident at once): every block then walks several units and can
  // prefetch its next one
  int grid = (max_blocks > 0 && max_blocks < n_blocks) ? max_blocks : n_blocks;
  if (grid > 4 * num_sms()) grid = 4 * num_sms();
  UNETK_CHECK_ARG(n_desc <= kUpdMaxDesc, "fused_update: at most %d tensors per launch", kUpdMaxDesc);
  UNETK_CUDA(cudaFuncSetAttribute(fused_update_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  fused_update_kernel<<<grid, 256, 0, s>>>(static_cast<const UpdateDesc*>(descs), n_desc, n_blocks, p, g, m, v, gstage,
                                          static_cast<__nv_bfloat16*>(wpack), hyper, mode);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int update_plain_chunk_elems() { return kUpdPlainChunk; }


int pack_conv_weights(const void* descs, int n_desc, int n_blocks, const float* flat_p, void* wpack,
                      cudaStream_t s) {
  UNETK_CHECK_ARG(descs && flat_p && wpack && n_desc > 0 && n_blocks > 0, "pack_conv_weights: bad arguments");
  pack_conv_weights_kernel<<<n_blocks, 256, 0, s>>>(static_cast<const ConvPackDesc*>(descs), n_desc, flat_p,
                                                   static_cast<__nv_bfloat16*>(wpack));
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int unpack_conv_grads(const void* descs, int n_desc, int n_blocks, int max_cin, const float* gstage,
                      float* flat_g, float scale, cudaStream_t s) {
  UNETK_CHECK_ARG(descs && gstage && flat_g && n_desc > 0 && n_blocks > 0, "unpack_conv_grads: bad arguments");
  UNETK_CHECK_ARG(max_cin > 0 && max_cin <= 1024, "unpack_conv_grads: cin up to 1024 supported");
  const size_t smem = static_cast<size_t>(9) * (max_cin + 1) * sizeof(float);
  unpack_conv_grads_kernel<<<n_blocks, 256, smem, s>>>(static_cast<const ConvPackDesc*>(descs), n_desc, gstage, flat_g, scale);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int adam_step(float* p, const float* g, float* m, float* v, long n, float lr, float beta1,
              float beta2, float eps, float wd, long step, float grad_scale, int mode,
              cudaStream_t s) {
  UNETK_CHECK_ARG(p && g && m, "adam_step: null pointer");
  UNETK_CHECK_ARG(mode == 2 || v != nullptr, "adam_step: second-moment buffer missing");
  UNETK_CHECK_ARG(mode >= 0 && mode <= 2, "adam_step: unknown mode %d", mode);
  UNETK_CHECK_ARG(step >= 1, "adam_step: step counts from 1");
  UNETK_CHECK_ARG((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(m) & 15) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0,
                  "adam_step: buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
  long blocks = (n / 4 + 255) / 256;
  const long cap = static_cast<long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_step_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(p, g, m, v, n, lr, beta1, beta2, eps, wd, static_cast<float>(bc1),
                                                           static_cast<float>(sqrt(bc2)), grad_scale, mode, step == 1);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int permute_batched(const void* descs, int n_desc, int n_chunks, const float* in, void* out,
                    int out_bf16, float scale, cudaStream_t s) {
  UNETK_CHECK_ARG(descs && in && out, "permute_batched: null pointer");
  UNETK_CHECK_ARG(n_desc > 0 && n_chunks > 0, "permute_batched: empty table");
  if (out_bf16)
    permute_batched_kernel<__nv_bfloat16><<<n_chunks, 256, 0, s>>>(static_cast<const PermuteDesc*>(descs), n_desc, in,
                                                                   static_cast<__nv_bfloat16*>(out), scale);
  else
    permute_batched_kernel<float><<<n_chunks, 256, 0, s>>>(static_cast<const PermuteDesc*>(descs), n_desc, in,
                                                          static_cast<float*>(out), scale);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int permute_chunk_elems() { return kPermChunk; }

}  // namespace unetk
