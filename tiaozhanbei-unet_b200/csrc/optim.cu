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
__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                 float* __restrict__ v, long n, float lr, float beta1, float beta2, float eps,
                 float wd, float bc1, float bc2_sqrt, float grad_scale, int mode, int first_step) {
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  const float step_size = lr / bc1;
  for (long i = tid * 4; i < n; i += nthreads * 4) {
    float pv[4], gv[4], mv[4], vv[4];
    const int cnt = (i + 4 <= n) ? 4 : static_cast<int>(n - i);
    if (cnt == 4) {
      *reinterpret_cast<float4*>(pv) = *reinterpret_cast<const float4*>(p + i);
      *reinterpret_cast<float4*>(gv) = *reinterpret_cast<const float4*>(g + i);
      *reinterpret_cast<float4*>(mv) = *reinterpret_cast<const float4*>(m + i);
      if (mode != 2) *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(v + i);
    } else {
      for (int k = 0; k < cnt; ++k) {
        pv[k] = p[i + k]; gv[k] = g[i + k]; mv[k] = m[i + k];
        if (mode != 2) vv[k] = v[i + k];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k >= cnt) break;
      float gg = gv[k] * grad_scale;
      if (mode == 0) {
        gg = fmaf(wd, pv[k], gg);
      } else if (mode == 1) {
        pv[k] *= (1.f - lr * wd);
      }
      if (mode == 2) {
        gg = fmaf(wd, pv[k], gg);
        mv[k] = first_step ? gg : fmaf(beta1, mv[k], gg);
        pv[k] -= lr * mv[k];
      } else {
        mv[k] = mv[k] + (1.f - beta1) * (gg - mv[k]);               // lerp_, as torch does
        vv[k] = beta2 * vv[k] + (1.f - beta2) * gg * gg;            // mul_ + addcmul_
        const float denom = sqrtf(vv[k]) / bc2_sqrt + eps;
        pv[k] -= step_size * (mv[k] / denom);                       // addcdiv_
      }
    }
    if (cnt == 4) {
      *reinterpret_cast<float4*>(p + i) = *reinterpret_cast<float4*>(pv);
      *reinterpret_cast<float4*>(m + i) = *reinterpret_cast<float4*>(mv);
      if (mode != 2) *reinterpret_cast<float4*>(v + i) = *reinterpret_cast<float4*>(vv);
    } else {
      for (int k = 0; k < cnt; ++k) {
        p[i + k] = pv[k]; m[i + k] = mv[k];
        if (mode != 2) v[i + k] = vv[k];
      }
    }
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
