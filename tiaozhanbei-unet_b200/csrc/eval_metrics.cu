// Evaluation reductions of validate_epoch / calculate_pixel_metrics on the device:
//
//   pixel_confusion   per-threshold confusion counts of (mask_pred > t) against (mask_true > 0.5)
//                     over the anomalous images only            src/train_utils.py:232-246, src/utils.py:97-108
//   select_kth        exact k-th / (k+1)-th order statistics of an fp32 array: the two neighbours
//                     np.percentile(all_scores, 95) interpolates  src/train_utils.py:213
//
// Both are integer / selection work: results are bit-exact against numpy.  HBM bound: the confusion
// counts read each map once (8 B per pixel for up to 8 thresholds), the selection reads the scores five
// times (4 radix passes of 8 bits + one pass for the successor).
#include "common.cuh"
#include "kernels.h"

namespace unetk {

namespace {

constexpr int kMaxThresholds = 8;

struct ConfusionArgs {
  float thr[kMaxThresholds];
  int nt;
  float truth_thr;
};

__device__ __forceinline__ unsigned warp_sum(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// grid (blocks per image, images).  counts[t*4 + {tn, fp, fn, tp}] (sklearn's confusion_matrix().ravel()).
template <bool VEC>
__global__ void __launch_bounds__(256, 4)
pixel_confusion_kernel(const float* __restrict__ pred, const float* __restrict__ truth,
                       const unsigned char* __restrict__ include, ConfusionArgs a,
                       unsigned long long* __restrict__ counts, long epi) {
  const int img = blockIdx.y;
  if (include != nullptr && include[img] == 0) return;
  const float* p = pred + static_cast<long>(img) * epi;
  const float* t = truth + static_cast<long>(img) * epi;
  unsigned pos = 0, cnt = 0, pp[kMaxThresholds], tp[kMaxThresholds];
#pragma unroll
  for (int j = 0; j < kMaxThresholds; ++j) pp[j] = tp[j] = 0;

  auto one = [&](float pv, float tv) {
    const unsigned is_pos = tv > a.truth_thr ? 1u : 0u;
    pos += is_pos;
    cnt += 1;
#pragma unroll
    for (int j = 0; j < kMaxThresholds; ++j) {
      if (j < a.nt) {
        const unsigned hit = pv > a.thr[j] ? 1u : 0u;
        pp[j] += hit;
        tp[j] += hit & is_pos;
      }
    }
  };

  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  if (VEC) {
    const long nv = epi >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(p);
    const float4* t4 = reinterpret_cast<const float4*>(t);
#pragma unroll 2
    for (long i = tid; i < nv; i += stride) {
      const float4 pv = __ldg(p4 + i), tv = __ldg(t4 + i);
      one(pv.x, tv.x);
      one(pv.y, tv.y);
      one(pv.z, tv.z);
      one(pv.w, tv.w);
    }
  } else {
    for (long i = tid; i < epi; i += stride) one(__ldg(p + i), __ldg(t + i));
  }

  __shared__ unsigned s_acc[2 + 2 * kMaxThresholds];
  if (threadIdx.x < 2 + 2 * kMaxThresholds) s_acc[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  pos = warp_sum(pos);
  cnt = warp_sum(cnt);
  if (lane == 0) {
    atomicAdd(&s_acc[0], pos);
    atomicAdd(&s_acc[1], cnt);
  }
#pragma unroll
  for (int j = 0; j < kMaxThresholds; ++j) {
    if (j < a.nt) {
      const unsigned x = warp_sum(pp[j]), y = warp_sum(tp[j]);
      if (lane == 0) {
        atomicAdd(&s_acc[2 + j], x);
        atomicAdd(&s_acc[2 + kMaxThresholds + j], y);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < a.nt) {
    const int j = threadIdx.x;
    const unsigned long long P = s_acc[0], N = s_acc[1], predp = s_acc[2 + j], tpos = s_acc[2 + kMaxThresholds + j];
    const unsigned long long fp = predp - tpos, fn = P - tpos, tn = N - P - fp;
    if (tn) atomicAdd(counts + j * 4 + 0, tn);
    if (fp) atomicAdd(counts + j * 4 + 1, fp);
    if (fn) atomicAdd(counts + j * 4 + 2, fn);
    if (tpos) atomicAdd(counts + j * 4 + 3, tpos);
  }
}

// ---------------------------------------------------------------------------------------------
// SegmentationMetrics.update (src/metrics.py:22-45): torch.argmax over the class axis fused with
// sklearn's confusion_matrix(target, pred, labels=range(C)): cm[t, p] += 1 for the pixels whose label and
// prediction are both in [0, C); pixels with target == ignore_index are dropped first and `kept` counts the
// rest (total_samples).  Integer work, bit-exact.  One pass over the logits (or the predicted indices) and
// the labels; equal (t, p) pairs of a warp are combined before the shared-memory histogram (the background
// class dominates).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxClasses = 8;

template <bool LOGITS>
__global__ void __launch_bounds__(256)
class_confusion_kernel(const float* __restrict__ logits, const long long* __restrict__ preds,
                       const long long* __restrict__ target, int C, long hw, long total, long long ignore_index,
                       int has_ignore, unsigned long long* __restrict__ cm, unsigned long long* __restrict__ kept) {
  __shared__ unsigned h[kMaxClasses * kMaxClasses + 1];
  for (int i = threadIdx.x; i <= kMaxClasses * kMaxClasses; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long rounded = (total + 31) & ~31L;      // whole warps stay in the loop together (warp-wide votes below)
  for (long p = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; p < rounded;
       p += static_cast<long>(gridDim.x) * blockDim.x) {
    int bin = -1;
    bool keep = false;
    if (p < total) {
      const long long t = target[p];
      if (!(has_ignore && t == ignore_index)) {
        keep = true;
        long long pr;
        if (LOGITS) {
          const long n = p / hw, r = p - n * hw;
          const float* b = logits + n * C * hw + r;
          float best = __ldg(b);
          int bi = 0;
          for (int c = 1; c < C; ++c) {
            const float v = __ldg(b + c * hw);
            if (v > best || (v != v && best == best)) {   // first maximum wins, NaN counts as the maximum (ATen)
              best = v;
              bi = c;
            }
          }
          pr = bi;
        } else {
          pr = preds[p];
        }
        if (t >= 0 && t < C && pr >= 0 && pr < C) bin = static_cast<int>(t) * C + static_cast<int>(pr);
      }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0 && lane == __ffs(peers) - 1) atomicAdd(&h[bin], static_cast<unsigned>(__popc(peers)));
    const unsigned kept_lanes = __ballot_sync(0xffffffffu, keep);
    if (lane == 0 && kept_lanes) atomicAdd(&h[kMaxClasses * kMaxClasses], static_cast<unsigned>(__popc(kept_lanes)));
  }
  __syncthreads();
  if (threadIdx.x < C * C && h[threadIdx.x]) atomicAdd(cm + threadIdx.x, static_cast<unsigned long long>(h[threadIdx.x]));
  if (threadIdx.x == kMaxClasses * kMaxClasses && h[threadIdx.x])
    atomicAdd(kept, static_cast<unsigned long long>(h[threadIdx.x]));
}

// ---------------------------------------------------------------------------------------------
// radix selection on the order-preserving 32-bit key of an fp32 value (NaN orders last, like numpy)
// ---------------------------------------------------------------------------------------------
struct SelectState {
  unsigned long long k;          // rank still to find inside the current prefix
  unsigned long long dup;        // after the last pass: elements equal to the answer at ranks > k
  unsigned prefix;               // key bits fixed so far
  unsigned next_min;             // smallest key greater than the answer
  unsigned long long hist[4][256];
};

__device__ __forceinline__ unsigned float_key(float v) {
  if (v != v) return 0xFFFFFFFFu;
  const unsigned u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned key) {
  if (key == 0xFFFFFFFFu) return __uint_as_float(0x7FC00000u);
  return __uint_as_float((key & 0x80000000u) ? (key & 0x7FFFFFFFu) : ~key);
}

__global__ void select_init_kernel(SelectState* st, unsigned long long k) {
  st->k = k;
  st->dup = 0;
  st->prefix = 0;
  st->next_min = 0xFFFFFFFFu;
}

template <bool VEC, typename F>
__device__ __forceinline__ void for_each_value(const float* __restrict__ x, long n, F&& f) {
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  if (VEC) {
    const long nv = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (long i = tid; i < nv; i += stride) {
      const float4 v = __ldg(x4 + i);
      f(v.x);
      f(v.y);
      f(v.z);
      f(v.w);
    }
    for (long i = (nv << 2) + tid; i < n; i += stride) f(__ldg(x + i));
  } else {
    for (long i = tid; i < n; i += stride) f(__ldg(x + i));
  }
}

// histogram of the 8 key bits of pass `pass` (most significant first) over the values whose higher
// bits equal the prefix found so far
template <bool VEC>
__global__ void __launch_bounds__(256)
select_hist_kernel(const float* __restrict__ x, long n, SelectState* __restrict__ st, int pass) {
  __shared__ unsigned h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int shift = 24 - 8 * pass;
  const unsigned mask = pass == 0 ? 0u : (0xFFFFFFFFu << (32 - 8 * pass));
  const unsigned prefix = st->prefix;
  for_each_value<VEC>(x, n, [&](float v) {
    const unsigned key = float_key(v);
    if ((key & mask) == prefix) atomicAdd(&h[(key >> shift) & 0xFFu], 1u);
  });
  __syncthreads();
  const unsigned c = h[threadIdx.x];
  if (c) atomicAdd(&st->hist[pass][threadIdx.x], static_cast<unsigned long long>(c));
}

__global__ void select_scan_kernel(SelectState* st, int pass) {
  if (threadIdx.x != 0) return;
  unsigned long long k = st->k, before = 0;
  int bin = 255;
  for (int b = 0; b < 256; ++b) {
    const unsigned long long c = st->hist[pass][b];
    if (k < before + c) {
      bin = b;
      break;
    }
    before += c;
  }
  st->prefix |= static_cast<unsigned>(bin) << (24 - 8 * pass);
  st->k = k - before;
  if (pass == 3) st->dup = st->hist[3][bin] - (k - before) - 1;
}

template <bool VEC>
__global__ void __launch_bounds__(256)
select_next_kernel(const float* __restrict__ x, long n, SelectState* __restrict__ st) {
  const unsigned answer = st->prefix;
  unsigned best = 0xFFFFFFFFu;   // also the key of NaN: when nothing finite follows, the successor is NaN
  for_each_value<VEC>(x, n, [&](float v) {
    const unsigned key = float_key(v);
    if (key > answer && key < best) best = key;
  });
  best = __reduce_min_sync(0xffffffffu, best);
  if ((threadIdx.x & 31) == 0 && best != 0xFFFFFFFFu) atomicMin(&st->next_min, best);
}

// out[0] = k-th smallest, out[1] = (k+1)-th smallest (= out[0] when k is the last rank)
__global__ void select_finish_kernel(const SelectState* st, float* out, int has_next) {
  const float lo = key_float(st->prefix);
  out[0] = lo;
  out[1] = (st->dup > 0 || !has_next) ? lo : key_float(st->next_min);
}

inline int capped_blocks(long work_items, int per_sm) {
  long blocks = (work_items + 255) / 256;
  const long cap = static_cast<long>(num_sms()) * per_sm;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

}  // namespace

int pixel_confusion(const float* pred, const float* truth, const unsigned char* include, const float* thresholds,
                    int n_thresholds, float truth_threshold, long long* counts, int n, long elems_per_image,
                    cudaStream_t s) {
  UNETK_CHECK_ARG(pred && truth && thresholds && counts, "pixel_confusion: null pointer");
  UNETK_CHECK_ARG(n_thresholds >= 1 && n_thresholds <= kMaxThresholds, "pixel_confusion: 1..%d thresholds, got %d",
                  kMaxThresholds, n_thresholds);
  UNETK_CHECK_ARG(n >= 0 && n <= 65535 && elems_per_image >= 0, "pixel_confusion: bad sizes n=%d elems=%ld", n,
                  elems_per_image);
  UNETK_CHECK_ARG(elems_per_image < (1L << 32), "pixel_confusion: more than 2^32 pixels per image");
  if (n == 0 || elems_per_image == 0) return UNETK_OK;
  ConfusionArgs a{};
  for (int j = 0; j < n_thresholds; ++j) a.thr[j] = thresholds[j];
  a.nt = n_thresholds;
  a.truth_thr = truth_threshold;
  const bool vec = (elems_per_image % 4 == 0) && ((reinterpret_cast<uintptr_t>(pred) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(truth) & 15) == 0);
  const long per_thread = vec ? 16 : 4;
  long bx = (elems_per_image + 256 * per_thread - 1) / (256 * per_thread);
  long cap = static_cast<long>(num_sms()) * 4 / n;   // one wave: 4 blocks of 256 threads resident per SM
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(n));
  auto* c = reinterpret_cast<unsigned long long*>(counts);
  if (vec)
    pixel_confusion_kernel<true><<<grid, 256, 0, s>>>(pred, truth, include, a, c, elems_per_image);
  else
    pixel_confusion_kernel<false><<<grid, 256, 0, s>>>(pred, truth, include, a, c, elems_per_image);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

size_t select_workspace_bytes() { return sizeof(SelectState); }

int select_kth(const float* x, long n, long k, float* out2, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  UNETK_CHECK_ARG(x && out2 && workspace, "select_kth: null pointer");
  UNETK_CHECK_ARG(n >= 1 && k >= 0 && k < n, "select_kth: rank %ld outside [0, %ld)", k, n);
  if (workspace_bytes < sizeof(SelectState)) {
    set_last_error("select_kth: workspace of %zu bytes, %zu needed", workspace_bytes, sizeof(SelectState));
    return UNETK_ERR_WORKSPACE;
  }
  UNETK_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "select_kth: workspace must be 8-byte aligned");
  auto* st = static_cast<SelectState*>(workspace);
  UNETK_CUDA(cudaMemsetAsync(st, 0, sizeof(SelectState), s));
  select_init_kernel<<<1, 1, 0, s>>>(st, static_cast<unsigned long long>(k));
  const bool vec = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const int blocks = capped_blocks((n + 7) / 8, 8);
  for (int pass = 0; pass < 4; ++pass) {
    if (vec)
      select_hist_kernel<true><<<blocks, 256, 0, s>>>(x, n, st, pass);
    else
      select_hist_kernel<false><<<blocks, 256, 0, s>>>(x, n, st, pass);
    select_scan_kernel<<<1, 32, 0, s>>>(st, pass);
  }
  const int has_next = k + 1 < n ? 1 : 0;
  if (has_next) {
    if (vec)
      select_next_kernel<true><<<blocks, 256, 0, s>>>(x, n, st);
    else
      select_next_kernel<false><<<blocks, 256, 0, s>>>(x, n, st);
  }
  select_finish_kernel<<<1, 1, 0, s>>>(st, out2, has_next);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int class_confusion(const float* logits, const long long* preds, const long long* target, int n, int c, long hw,
                    long long ignore_index, int has_ignore, long long* cm, long long* kept, cudaStream_t s) {
  UNETK_CHECK_ARG((logits != nullptr) != (preds != nullptr), "class_confusion: exactly one of logits / preds");
  UNETK_CHECK_ARG(target && cm && kept, "class_confusion: null pointer");
  UNETK_CHECK_ARG(c >= 1 && c <= kMaxClasses, "class_confusion: 1..%d classes, got %d", kMaxClasses, c);
  UNETK_CHECK_ARG(n >= 0 && hw >= 0, "class_confusion: bad sizes");
  const long total = static_cast<long>(n) * hw;
  if (total == 0) return UNETK_OK;
  const int blocks = capped_blocks(total, 8);
  auto* m = reinterpret_cast<unsigned long long*>(cm);
  auto* k = reinterpret_cast<unsigned long long*>(kept);
  if (logits != nullptr)
    class_confusion_kernel<true><<<blocks, 256, 0, s>>>(logits, nullptr, target, c, hw, total, ignore_index, has_ignore, m, k);
  else
    class_confusion_kernel<false><<<blocks, 256, 0, s>>>(nullptr, preds, target, c, hw, total, ignore_index, has_ignore, m, k);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

}  // namespace unetk
