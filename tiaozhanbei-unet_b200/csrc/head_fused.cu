// Fused output stage of one AnomalyUNet decoder for a training step (sm_100a, bandwidth bound):
//
//   y  = raw output of the decoder's last conv3x3                        (bf16 NHWC, read ONCE)
//   a  = relu(BatchNorm_train(y))      finalize folded in, running statistics updated   src/model.py:18-19
//   p  = sigmoid(conv1x1(a) + b)       OutConv + torch.sigmoid                          src/model.py:69-75,201,208
//   L  = weight * mean((p - x)^2)            reconstruction term of CombinedLoss        src/train_utils.py:20,36
//      | weight * mean(alpha (1-pt)^g bce)   focal term                                  src/train_utils.py:23-28,39
//   dA = dL/da (bf16 NHWC), dW, db of the 1x1 conv, and the first pass of the BatchNorm backward of the
//        layer that produced y (sum g, sum g*xhat)                               autograd of the above
//
// i.e. head_fwd_bn + combined_loss + head_bwd of csrc/head_loss.cu in ONE pass over the 64-channel
// full-resolution tensor.  The gradient of a mean-reduced loss w.r.t. one output element does not depend on
// the value of the mean, so forward and backward of the head need no grid-wide step between them; the loss
// VALUE is a deterministic two-stage reduction (per-block partials in double, folded in block order by the
// last block), and the later of the two decoders' kernels combines the two terms into total_loss.
// The unfused sequence took 0.57 ms between the last convolution of the forward pass and the first
// convolution of the backward pass at 16 x 256^2 (five dependent bandwidth-bound launches, three of them
// reading the same 134 MB tensor, plus the boundary between the forward and the backward CUDA graph).
//
// Thread mapping as head_bwd_kernel: 8 lanes cover one pixel (16-byte vectors of 8 channels), a warp covers
// 4 pixels per step.  The 1x1 weights live in shared memory, and so do the loads in flight (cp.async ring).
#include "common.cuh"
#include "kernels.h"

namespace unetk {

constexpr int kFC = 64;          // channels of the decoder's last layer
constexpr int kFMaxBlocks = 1024;

struct HeadLossWork {            // zero-initialised once by the caller; self-cleaning afterwards
  unsigned int ticket[2];        // blocks of decoder `which` that have stored their partial
  unsigned int both;             // decoders whose term is final
  unsigned int pad;
  float term[2];                 // mean loss of each decoder (unweighted)
  double partial[2][kFMaxBlocks];
};

// cp.async (LDGSTS) helpers: 16 / 4 bytes global -> shared, zero-filled when !ok (src-size 0: nothing is read)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool ok) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(ok ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, bool ok) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(ok ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// D: depth of the per-lane prefetch ring.  The kernel sits at its register cap (two blocks of 256 threads per SM:
// 40 accumulators per thread for dW and the BatchNorm sums), so loads in flight cannot live in registers: with four
// 16-byte loads per thread it ran at 2.3 TB/s, latency bound.  Every lane instead keeps D steps of its own
// 16-byte vector (and the warp D steps of its 4 x NOUT targets) in flight through cp.async into a private ring
// in shared memory -- a lane only ever reads back what it issued itself, so no barrier is involved.
template <int NOUT, int LOSS, int D>   // LOSS 0: squared error against `target`, 1: focal BCE
__global__ void __launch_bounds__(256, 2)
head_loss_bwd_kernel(const uint4* __restrict__ y, const BnFinalizeArgs fin, const float* __restrict__ w,
                     const float* __restrict__ b, const float* __restrict__ target, float* __restrict__ out,
                     uint4* __restrict__ da, float* __restrict__ dw, float* __restrict__ db, float* __restrict__ bn_sums,
                     HeadLossWork* __restrict__ work, float* __restrict__ out3, int which, long npix_total, long hw,
                     float gcoef, float w0, float w1, float alpha, float gamma) {
  __shared__ float sw[NOUT * kFC];
  __shared__ float sb[NOUT];
  __shared__ float sbn[4 * kFC];     // scale | shift | mean | invstd
  __shared__ float sdw[NOUT * kFC];
  __shared__ float sdb[NOUT];
  __shared__ float ssum[2 * kFC];
  __shared__ float sloss[8];
  __shared__ bool is_last;
  __shared__ uint4 yring[8 * D * 32];      // [warp][slot][lane]
  __shared__ float tring[8 * D * 32];      // [warp][slot][pixel slot * 8 + j]
  pdl_trigger();
  pdl_wait();
  for (int i = threadIdx.x; i < NOUT * kFC; i += blockDim.x) {
    sw[i] = w[i];
    sdw[i] = 0.f;
  }
  if (threadIdx.x < NOUT) {
    sb[threadIdx.x] = b[threadIdx.x];
    sdb[threadIdx.x] = 0.f;
  }
  if (threadIdx.x < 2 * kFC) ssum[threadIdx.x] = 0.f;
  if (threadIdx.x < kFC) {
    float sc, sh, mu, is;
    bn_finalize1(fin, kFC, threadIdx.x, blockIdx.x == 0, sc, sh, &mu, &is);
    sbn[threadIdx.x] = sc;
    sbn[kFC + threadIdx.x] = sh;
    sbn[2 * kFC + threadIdx.x] = mu;
    sbn[3 * kFC + threadIdx.x] = is;
    if (blockIdx.x == 0 && threadIdx.x == 0 && fin.num_batches_tracked != nullptr) *fin.num_batches_tracked += 1;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int cg = lane & 7;          // channel group: channels 8*cg .. 8*cg+7
  const int pl = lane >> 3;         // pixel slot inside the warp step
  const float* wsc = sbn + cg * 8;
  const float* wsh = sbn + kFC + cg * 8;
  float gw[NOUT][8], gb[NOUT], s1[8], s2[8];
  bool degenerate = false;          // a channel with scale == 0 cannot give xhat back from its activation
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s1[k] = s2[k] = 0.f;
    if (wsc[k] == 0.f) degenerate = true;
  }
#pragma unroll
  for (int j = 0; j < NOUT; ++j) {
    gb[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) gw[j][k] = 0.f;
  }
  float lsum = 0.f;
  const long warp = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  uint4* myy = yring + (threadIdx.x >> 5) * (D * 32) + lane;
  float* myt = tring + (threadIdx.x >> 5) * (D * 32);
  // step i of this warp covers the four pixels (warp + i * nwarps) * 4 + {0..3}.  (image, offset) of the pixel the next
  // issue() fetches are carried from step to step: a 64-bit division per step and lane was a quarter of the
  // kernel's instructions, and the kernel is issue bound
  const long pstep = nwarps * 4;
  const long dq = pstep / hw, dr = pstep - dq * hw;
  long ip = warp * 4 + pl;                   // pixel of the next issue()
  long in = ip / hw, ir = ip - in * hw;
  auto issue = [&](int slot) {
    const bool okp = ip < npix_total;
    cp_async16(myy + slot * 32, y + (okp ? ip : 0) * 8 + cg, okp);
    if (cg < NOUT) cp_async4(myt + slot * 32 + lane, target + (okp ? (in * NOUT + cg) * hw + ir : 0), okp);
    cp_async_commit();
    ip += pstep;
    in += dq;
    ir += dr;
    if (ir >= hw) {
      ir -= hw;
      ++in;
    }
  };
#pragma unroll
  for (int s = 0; s < D; ++s) issue(s);
  int slot = 0;
  for (long i = 0; (warp + i * nwarps) * 4 < npix_total; ++i) {
    cp_async_wait<D - 1>();
    __syncwarp();                      // (the targets of a pixel were fetched by its lanes 0..NOUT-1)
    const long p = (warp + i * nwarps) * 4 + pl;
    const bool oku = p < npix_total;
    const uint4 yq0 = myy[slot * 32];
    float tgt[NOUT];
#pragma unroll
    for (int j = 0; j < NOUT; ++j) tgt[j] = myt[slot * 32 + pl * 8 + j];
    __syncwarp();
    issue(slot);                       // refill the slot just read (step i + D)
    slot = (slot + 1 == D) ? 0 : slot + 1;
    {
      float x[8], g[8];
      {
        const float2 f0 = unpack_bf16x2(yq0.x), f1 = unpack_bf16x2(yq0.y), f2 = unpack_bf16x2(yq0.z), f3 = unpack_bf16x2(yq0.w);
        x[0] = f0.x; x[1] = f0.y; x[2] = f1.x; x[3] = f1.y; x[4] = f2.x; x[5] = f2.y; x[6] = f3.x; x[7] = f3.y;
      }
      // BatchNorm + ReLU, rounded to bf16 exactly as the stand-alone pass would have stored the activation
#pragma unroll
      for (int k = 0; k < 8; ++k)
        x[k] = oku ? __bfloat162float(__float2bfloat16(fmaxf(fmaf(x[k], wsc[k], wsh[k]), 0.f))) : 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = 0.f;
#pragma unroll
      for (int j = 0; j < NOUT; ++j) {
        // 1x1 conv: this lane's 8 channels, then the 8 lanes of the pixel
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(x[k], sw[j * kFC + cg * 8 + k], acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        const float pr = 1.f / (1.f + __expf(-(acc + sb[j])));
        const float t = tgt[j];
        float dldp, term;
        if (LOSS == 0) {
          const float d = pr - t;
          term = d * d;
          dldp = gcoef * d;
        } else {
          // F.binary_cross_entropy (log clamped at -100) inside the focal weighting, and ATen's BCE backward
          const float lp = fmaxf(logf(pr), -100.f), l1p = fmaxf(log1pf(-pr), -100.f);
          const float bce = -(t * lp + (1.f - t) * l1p);
          const float pt = expf(-bce);
          const float om = 1.f - pt;
          const float omg = (gamma == 2.f) ? om * om : powf(om, gamma);
          term = alpha * omg * bce;
          const float dbce = (pr - t) / fmaxf((1.f - pr) * pr, 1e-12f);
          const float omg1 = (gamma == 2.f) ? om : ((om > 0.f) ? powf(om, gamma - 1.f) : 0.f);
          dldp = gcoef * alpha * (omg + gamma * omg1 * pt * bce) * dbce;
        }
        const float dl = oku ? dldp * pr * (1.f - pr) : 0.f;      // through the sigmoid
        if (cg == 0 && oku) {
          lsum += term;
          gb[j] += dl;
        }
        if (out != nullptr && cg == j && oku) {
          const long n = p / hw;
          out[(n * NOUT + j) * hw + (p - n * hw)] = pr;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          g[k] = fmaf(dl, sw[j * kFC + cg * 8 + k], g[k]);
          gw[j][k] = fmaf(dl, x[k], gw[j][k]);
        }
      }
      if (oku) {
        uint4 o4;
        o4.x = pack_bf16x2(g[0], g[1]); o4.y = pack_bf16x2(g[2], g[3]);
        o4.z = pack_bf16x2(g[4], g[5]); o4.w = pack_bf16x2(g[6], g[7]);
        da[p * 8 + cg] = o4;
        const float2 r0 = unpack_bf16x2(o4.x), r1 = unpack_bf16x2(o4.y), r2 = unpack_bf16x2(o4.z), r3 = unpack_bf16x2(o4.w);
        const float rr[8] = {r0.x, r0.y, r1.x, r1.y, r2.x, r2.y, r3.x, r3.y};   // the stored (bf16) gradient
        if (!degenerate) {
          // where a > 0:  xhat = A*a + B  (A = invstd/scale), so sum g and sum g*a suffice; mapped once per thread
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float gg = (x[k] > 0.f) ? rr[k] : 0.f;
            s1[k] += gg;
            s2[k] = fmaf(gg, x[k], s2[k]);
          }
        } else {
          const uint4 yq = ldg_nc_v4(y + p * 8 + cg);
          const float2 y0 = unpack_bf16x2(yq.x), y1 = unpack_bf16x2(yq.y), y2 = unpack_bf16x2(yq.z), y3 = unpack_bf16x2(yq.w);
          const float yy[8] = {y0.x, y0.y, y1.x, y1.y, y2.x, y2.y, y3.x, y3.y};
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = cg * 8 + k;
            const float gg = (fmaf(yy[k], sbn[c], sbn[kFC + c]) > 0.f) ? rr[k] : 0.f;
            s1[k] += gg;
            s2[k] = fmaf(gg, (yy[k] - sbn[2 * kFC + c]) * sbn[3 * kFC + c], s2[k]);
          }
        }
      }
    }
  }
  cp_async_wait<0>();
  // fold the four pixel slots of the warp, then one shared atomic per (j, channel) per warp
#pragma unroll
  for (int j = 0; j < NOUT; ++j) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = gw[j][k];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (pl == 0) atomicAdd(&sdw[j * kFC + cg * 8 + k], v);
    }
    float bsum = gb[j];
    bsum += __shfl_xor_sync(0xffffffffu, bsum, 8);
    bsum += __shfl_xor_sync(0xffffffffu, bsum, 16);
    if (lane == 0) atomicAdd(&sdb[j], bsum);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float v1 = s1[k], v2 = s2[k];
    if (!degenerate) {          // sum g*a -> sum g*xhat
      const int c = cg * 8 + k;
      const float inv = sbn[3 * kFC + c] / sbn[c];
      v2 = inv * (v2 - sbn[kFC + c] * v1) - sbn[2 * kFC + c] * sbn[3 * kFC + c] * v1;
    }
    v1 += __shfl_xor_sync(0xffffffffu, v1, 8);
    v1 += __shfl_xor_sync(0xffffffffu, v1, 16);
    v2 += __shfl_xor_sync(0xffffffffu, v2, 8);
    v2 += __shfl_xor_sync(0xffffffffu, v2, 16);
    if (pl == 0) {
      atomicAdd(&ssum[cg * 8 + k], v1);
      atomicAdd(&ssum[kFC + cg * 8 + k], v2);
    }
  }
  lsum = warp_sum(lsum);
  if (lane == 0) sloss[threadIdx.x >> 5] = lsum;
  __syncthreads();
  for (int i = threadIdx.x; i < NOUT * kFC; i += blockDim.x) atomicAdd(dw + i, sdw[i]);
  if (threadIdx.x < NOUT) atomicAdd(db + threadIdx.x, sdb[threadIdx.x]);
  if (threadIdx.x < 2 * kFC) atomicAdd(bn_sums + threadIdx.x, ssum[threadIdx.x]);
  // loss value: per-block partial, the last block of this decoder folds them in block order
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < 8; ++i) a += sloss[i];
    work->partial[which][blockIdx.x] = a;
    __threadfence();
    const unsigned int t = atomicAdd(&work->ticket[which], 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __shared__ double fold[256];
  __threadfence();
  double a = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x)
    a += reinterpret_cast<volatile double*>(work->partial[which])[i];
  fold[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) fold[threadIdx.x] += fold[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float term = static_cast<float>(fold[0] / (static_cast<double>(npix_total) * NOUT));
    work->term[which] = term;
    out3[1 + which] = term;
    work->ticket[which] = 0;                      // self-cleaning for the next step
    __threadfence();
    if (atomicAdd(&work->both, 1u) == 1u) {       // the other decoder's term is final too
      __threadfence();
      const volatile float* tv = work->term;
      out3[0] = w0 * tv[0] + w1 * tv[1];
      work->both = 0;
    }
  }
}

size_t head_loss_workspace_bytes() { return sizeof(HeadLossWork); }

// which: 0 = reconstruction decoder (LOSS 0, n_out channels), 1 = segmentation decoder (LOSS 1, one channel).
// out3 = {total, recon, seg}: the decoder whose kernel finishes second writes total = w0*recon + w1*seg.
int head_loss_bwd(const void* y, const BnFinalizeArgs* fin, const float* w, const float* b, const float* target,
                  float* out, void* da, float* dw, float* db, float* bn_sums, void* work, size_t work_bytes,
                  float* out3, int which, int N, int H, int W, int C, int n_out, float recon_weight, float seg_weight,
                  float alpha, float gamma, cudaStream_t s) {
  UNETK_CHECK_ARG(y && fin && w && b && target && da && dw && db && bn_sums && work && out3, "head_loss_bwd: null pointer");
  UNETK_CHECK_ARG(C == kFC, "head_loss_bwd: input channels must be 64 (got %d)", C);
  UNETK_CHECK_ARG(which == 0 || which == 1, "head_loss_bwd: which must be 0 (reconstruction) or 1 (segmentation)");
  UNETK_CHECK_ARG((which == 0 && (n_out == 1 || n_out == 3)) || (which == 1 && n_out == 1),
                  "head_loss_bwd: reconstruction heads with 1 or 3 channels, segmentation heads with 1 channel (got %d)", n_out);
  UNETK_CHECK_ARG(work_bytes >= sizeof(HeadLossWork), "head_loss_bwd: workspace too small");
  UNETK_CHECK_ARG(fin->stats && fin->gamma && fin->beta && fin->scale && fin->shift && fin->mean && fin->invstd,
                  "head_loss_bwd: incomplete BatchNorm finalize arguments");
  UNETK_CHECK_ARG((fin->running_mean == nullptr) == (fin->running_var == nullptr), "head_loss_bwd: running_mean and running_var go together");
  UNETK_CHECK_ARG(fin->count >= 1.0, "head_loss_bwd: empty batch");
  const long hw = static_cast<long>(H) * W, np = hw * N;
  long blocks = (np + 127) / 128;
  const long cap = static_cast<long>(num_sms()) * 2;     // two resident blocks per SM, grid-stride
  if (blocks > cap) blocks = cap;
  if (blocks > kFMaxBlocks) blocks = kFMaxBlocks;
  if (blocks < 1) blocks = 1;
  const BnFinalizeArgs f = with_reciprocals(*fin);
  const double n_el = static_cast<double>(np) * n_out;
  // gradient of the weighted mean w.r.t. one element (the expressions of combined_loss_kernel)
  const float gcoef = which == 0 ? recon_weight * 2.f / static_cast<float>(n_el) : seg_weight / static_cast<float>(n_el);
  const uint4* y4 = static_cast<const uint4*>(y);
  uint4* da4 = static_cast<uint4*>(da);
  HeadLossWork* wk = static_cast<HeadLossWork*>(work);
  const dim3 grid(static_cast<unsigned int>(blocks)), block(256);
  if (which == 1)
    UNETK_CUDA(launch_pdl(head_loss_bwd_kernel<1, 1, 8>, grid, block, 0, s, y4, f, w, b, target, out, da4, dw, db, bn_sums, wk, out3,
                          which, np, hw, gcoef, recon_weight, seg_weight, alpha, gamma));
  else if (n_out == 3)
    UNETK_CUDA(launch_pdl(head_loss_bwd_kernel<3, 0, 8>, grid, block, 0, s, y4, f, w, b, target, out, da4, dw, db, bn_sums, wk, out3,
                          which, np, hw, gcoef, recon_weight, seg_weight, alpha, gamma));
  else
    UNETK_CUDA(launch_pdl(head_loss_bwd_kernel<1, 0, 8>, grid, block, 0, s, y4, f, w, b, target, out, da4, dw, db, bn_sums, wk, out3,
                          which, np, hw, gcoef, recon_weight, seg_weight, alpha, gamma));
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

}  // namespace unetk
