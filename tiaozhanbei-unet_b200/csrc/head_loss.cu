// Bandwidth-bound kernels at the network output (sm_100a): 1x1 head conv (+sigmoid) forward and
// backward, CombinedLoss (MSE + focal) forward+backward, SSIM loss forward/backward, anomaly score.
//
// Reference operators being replaced:
//   OutConv = nn.Conv2d(64, n, 1) (+ torch.sigmoid)      /root/reference/src/model.py:69-75,201,208
//   CombinedLoss.forward / focal_loss                     src/train_utils.py:10-44
//   SSIMLoss._ssim                                         src/train_utils.py:67-87
//   compute_anomaly_score                                  src/utils.py:205-215
#include "common.cuh"
#include "kernels.h"

namespace unetk {

// ----------------------------------------------------------------------------------------------
// head forward: out[n, j, y, x] = act( sum_c a[n, y, x, c] * w[j, c] + b[j] ),  C == 64, n_out <= 4
// one thread per pixel: 128 B contiguous bf16 in, coalesced fp32 planes out
// ----------------------------------------------------------------------------------------------
constexpr int kHeadC = 64;
constexpr int kHeadMaxOut = 4;

// BN: `a` is the RAW output y of the last conv; the kernel applies that layer's training-mode BatchNorm (finalize
// folded in, see bn_finalize1) + ReLU on the fly, rounding the activation to bf16 exactly as the stand-alone
// BN+ReLU pass would have stored it -- the activated tensor of the last layer is never written (a 134 MB read +
// 134 MB write pass per decoder at 16 x 256^2 that sat, un-overlapped, at the end of the forward pass).
template <bool BN>
__global__ void __launch_bounds__(256)
head_fwd_kernel(const uint4* __restrict__ a, const float* __restrict__ w, const float* __restrict__ b,
                float* __restrict__ out, long npix_total, long hw, int n_out, int sigmoid, const BnFinalizeArgs fin) {
  __shared__ float sw[kHeadMaxOut * kHeadC];
  __shared__ float sb[kHeadMaxOut];
  __shared__ float ssc[BN ? kHeadC : 1], ssh[BN ? kHeadC : 1];
  pdl_trigger();
  pdl_wait();
  for (int i = threadIdx.x; i < n_out * kHeadC; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < n_out) sb[threadIdx.x] = b[threadIdx.x];
  if (BN && threadIdx.x < kHeadC) {
    float sc, sh;
    bn_finalize1(fin, kHeadC, threadIdx.x, blockIdx.x == 0, sc, sh);
    ssc[threadIdx.x] = sc;
    ssh[threadIdx.x] = sh;
    if (blockIdx.x == 0 && threadIdx.x == 0 && fin.num_batches_tracked != nullptr) *fin.num_batches_tracked += 1;
  }
  __syncthreads();
  for (long p = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; p < npix_total;
       p += static_cast<long>(gridDim.x) * blockDim.x) {
    float acc[kHeadMaxOut] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const uint4 u = ldg_nc_v4(a + p * 8 + v);
      const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      float x[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
      if (BN) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          x[k] = __bfloat162float(__float2bfloat16(fmaxf(fmaf(x[k], ssc[v * 8 + k], ssh[v * 8 + k]), 0.f)));
      }
#pragma unroll
      for (int j = 0; j < kHeadMaxOut; ++j) {
        if (j < n_out) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[j] = fmaf(x[k], sw[j * kHeadC + v * 8 + k], acc[j]);
        }
      }
    }
    const long n = p / hw, r = p - n * hw;
    for (int j = 0; j < n_out; ++j) {
      float z = acc[j] + sb[j];
      if (sigmoid) z = 1.f / (1.f + __expf(-z));
      out[(n * n_out + j) * hw + r] = z;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// head backward.  dlogit = dout * out * (1 - out) if sigmoid else dout.
//   dA[p, c] = sum_j dlogit[p, j] * w[j, c]          (bf16 NHWC)
//   dW[j, c] = sum_p dlogit[p, j] * a[p, c],  db[j] = sum_p dlogit[p, j]   (fp32, atomically added)
// 8 lanes cover one pixel (16-byte vectors of 8 channels), a warp covers 4 pixels per step and every
// thread keeps four steps in flight; weights and the dW partial sums live in registers.
// ----------------------------------------------------------------------------------------------
// Optional fusion (bn_sums != nullptr): first pass of the BatchNorm+ReLU backward of the layer that
// produced `a` (y = its raw conv output): sums[0..64) += sum g, sums[64..128) += sum g*xhat.
// a == nullptr (the forward pass used head_fwd_bn and never stored the activation): a is recomputed from
// bn_y as bf16(relu(y*scale + shift)), bit for bit what head_fwd_bn fed the 1x1 convolution.
template <int NOUT>
__global__ void __launch_bounds__(256, 2)
head_bwd_kernel(const uint4* __restrict__ a, const float* __restrict__ w,
                const float* __restrict__ dout, const float* __restrict__ out,
                uint4* __restrict__ da, float* __restrict__ dw, float* __restrict__ db,
                long npix_total, long hw, int sigmoid, const uint4* __restrict__ bn_y,
                const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                const float* __restrict__ bn_mean, const float* __restrict__ bn_invstd,
                float* __restrict__ bn_sums) {
  __shared__ float sdw[NOUT * kHeadC];
  __shared__ float sdb[NOUT];
  __shared__ float sbn[2 * kHeadC];
  __shared__ float sysc[2 * kHeadC];
  pdl_trigger();
  pdl_wait();
  const bool bn = bn_sums != nullptr;
  const bool from_y = a == nullptr;
  for (int i = threadIdx.x; i < NOUT * kHeadC; i += blockDim.x) sdw[i] = 0.f;
  if (threadIdx.x < NOUT) sdb[threadIdx.x] = 0.f;
  if (threadIdx.x < 2 * kHeadC) {
    sbn[threadIdx.x] = 0.f;
    if (from_y) sysc[threadIdx.x] = threadIdx.x < kHeadC ? bn_scale[threadIdx.x] : bn_shift[threadIdx.x - kHeadC];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int cg = lane & 7;          // channel group: channels 8*cg .. 8*cg+7
  const int pl = lane >> 3;         // pixel slot inside the warp step
  float wr[NOUT][8], gw[NOUT][8], gb[NOUT];
  // BatchNorm fusion: a = relu(y*scale + shift) is already in registers, and where a > 0
  //   xhat = (y - mean)*invstd = (a - shift)/scale*invstd - mean*invstd = A*a + B,
  // so the reduction needs only s1 = sum g and s2 = sum g*a (g = dA*[a > 0]); A, B are applied once per
  // block.  A channel whose scale is exactly 0 (gamma == 0) cannot be inverted: it takes the direct route
  // through y (never in practice; kept for exactness).
  float s1[8], s2[8];
  bool degenerate = false;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s1[k] = s2[k] = 0.f;
    if (bn && bn_scale[cg * 8 + k] == 0.f) degenerate = true;
  }
  const uint4* __restrict__ src = from_y ? bn_y : a;
  // (scale, shift of this thread's 8 channels stay in shared memory: the kernel sits at its register cap)
  const float* ysc = sysc + cg * 8;
  const float* ysh = sysc + kHeadC + cg * 8;
#pragma unroll
  for (int j = 0; j < NOUT; ++j) {
    gb[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      wr[j][k] = w[j * kHeadC + cg * 8 + k];
      gw[j][k] = 0.f;
    }
  }
  const long warp = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  constexpr int U = (NOUT >= 3) ? 2 : 4;   // steps in flight per thread (two blocks per SM)
  for (long base = warp * (4 * U); base < npix_total; base += nwarps * (4 * U)) {
    uint4 av[U];
    float dl[U][NOUT];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long p = base + u * 4 + pl;
      ok[u] = p < npix_total;
      av[u] = ok[u] ? ldg_nc_v4(src + p * 8 + cg) : make_uint4(0, 0, 0, 0);
      const long n = ok[u] ? p / hw : 0, r = ok[u] ? p - n * hw : 0;
#pragma unroll
      for (int j = 0; j < NOUT; ++j) {
        const long o = (n * NOUT + j) * hw + r;
        float d = ok[u] ? __ldg(dout + o) : 0.f;
        if (sigmoid && ok[u]) {
          const float sg = __ldg(out + o);
          d *= sg * (1.f - sg);
        }
        dl[u][j] = d;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float x[8], g[8];
      const float2 f0 = unpack_bf16x2(av[u].x), f1 = unpack_bf16x2(av[u].y), f2 = unpack_bf16x2(av[u].z), f3 = unpack_bf16x2(av[u].w);
      x[0] = f0.x; x[1] = f0.y; x[2] = f1.x; x[3] = f1.y; x[4] = f2.x; x[5] = f2.y; x[6] = f3.x; x[7] = f3.y;
      if (from_y) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          x[k] = ok[u] ? __bfloat162float(__float2bfloat16(fmaxf(fmaf(x[k], ysc[k], ysh[k]), 0.f))) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = 0.f;
#pragma unroll
      for (int j = 0; j < NOUT; ++j) {
        const float d = dl[u][j];
        if (cg == 0) gb[j] += d;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          g[k] = fmaf(d, wr[j][k], g[k]);
          gw[j][k] = fmaf(d, x[k], gw[j][k]);
        }
      }
      if (ok[u]) {
        uint4 o4;
        o4.x = pack_bf16x2(g[0], g[1]); o4.y = pack_bf16x2(g[2], g[3]);
        o4.z = pack_bf16x2(g[4], g[5]); o4.w = pack_bf16x2(g[6], g[7]);
        da[(base + u * 4 + pl) * 8 + cg] = o4;
        if (bn) {
          const float2 r0 = unpack_bf16x2(o4.x), r1 = unpack_bf16x2(o4.y), r2 = unpack_bf16x2(o4.z), r3 = unpack_bf16x2(o4.w);
          const float rr[8] = {r0.x, r0.y, r1.x, r1.y, r2.x, r2.y, r3.x, r3.y};   // the stored (bf16) gradient
          if (!degenerate) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float gg = (x[k] > 0.f) ? rr[k] : 0.f;
              s1[k] += gg;
              s2[k] = fmaf(gg, x[k], s2[k]);
            }
          } else {
            const uint4 yv4 = ldg_nc_v4(bn_y + (base + u * 4 + pl) * 8 + cg);
            const float2 y0 = unpack_bf16x2(yv4.x), y1 = unpack_bf16x2(yv4.y), y2 = unpack_bf16x2(yv4.z), y3 = unpack_bf16x2(yv4.w);
            const float yy[8] = {y0.x, y0.y, y1.x, y1.y, y2.x, y2.y, y3.x, y3.y};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int c = cg * 8 + k;
              const float gg = (fmaf(yy[k], bn_scale[c], bn_shift[c]) > 0.f) ? rr[k] : 0.f;
              s1[k] += gg;
              s2[k] = fmaf(gg, (yy[k] - bn_mean[c]) * bn_invstd[c], s2[k]);
            }
          }
        }
      }
    }
  }
  // fold the four pixel slots of the warp, then one shared atomic per (j, channel) per warp
#pragma unroll
  for (int j = 0; j < NOUT; ++j) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = gw[j][k];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (pl == 0) atomicAdd(&sdw[j * kHeadC + cg * 8 + k], v);
    }
    float bsum = gb[j];
    bsum += __shfl_xor_sync(0xffffffffu, bsum, 8);
    bsum += __shfl_xor_sync(0xffffffffu, bsum, 16);
    if (lane == 0) atomicAdd(&sdb[j], bsum);
  }
  if (bn) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v1 = s1[k], v2 = s2[k];
      if (!degenerate) {          // sum g*a -> sum g*xhat
        const int c = cg * 8 + k;
        const float inv = bn_invstd[c] / bn_scale[c];
        v2 = inv * (v2 - bn_shift[c] * v1) - bn_mean[c] * bn_invstd[c] * v1;
      }
      v1 += __shfl_xor_sync(0xffffffffu, v1, 8);
      v1 += __shfl_xor_sync(0xffffffffu, v1, 16);
      v2 += __shfl_xor_sync(0xffffffffu, v2, 8);
      v2 += __shfl_xor_sync(0xffffffffu, v2, 16);
      if (pl == 0) {
        atomicAdd(&sbn[cg * 8 + k], v1);
        atomicAdd(&sbn[kHeadC + cg * 8 + k], v2);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NOUT * kHeadC; i += blockDim.x) atomicAdd(dw + i, sdw[i]);
  if (threadIdx.x < NOUT) atomicAdd(db + threadIdx.x, sdb[threadIdx.x]);
  if (bn && threadIdx.x < 2 * kHeadC) atomicAdd(bn_sums + threadIdx.x, sbn[threadIdx.x]);
}

// ----------------------------------------------------------------------------------------------
// CombinedLoss forward + backward in one pass.
//   recon_loss = mean((r - x)^2)
//   bce = -(t * max(log p, -100) + (1 - t) * max(log(1 - p), -100))      (F.binary_cross_entropy)
//   seg_loss = mean(alpha * (1 - exp(-bce))^gamma * bce)
//   total = rw * recon_loss + sw * seg_loss
// Gradients (of `total`): d/dr = rw * 2 (r - x) / Nr ;  d/dp = sw * f'(p) / Ns with ATen's BCE
// backward  dbce/dp = (p - t) / max((1 - p) p, 1e-12).
// Deterministic two-stage reduction: block partials, last block (ticket) folds them in order.
// ----------------------------------------------------------------------------------------------
struct LossWork {
  unsigned int ticket;
  unsigned int pad;
  double partial[2 * 1024];  // [blocks][2]
};

__global__ void __launch_bounds__(256)
combined_loss_kernel(const float* __restrict__ recon, const float* __restrict__ amap,
                     const float* __restrict__ image, const float* __restrict__ mask,
                     float* __restrict__ d_recon, float* __restrict__ d_amap, float* __restrict__ out3,
                     LossWork* __restrict__ work, long n_recon, long n_seg, float rw, float sw,
                     float alpha, float gamma) {
  const long tid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long nthreads = static_cast<long>(gridDim.x) * blockDim.x;
  float s_mse = 0.f, s_foc = 0.f;
  const float gr = rw * 2.f / static_cast<float>(n_recon);
  for (long i = tid * 4; i < n_recon; i += nthreads * 4) {
    if (i + 4 <= n_recon) {
      const float4 r = *reinterpret_cast<const float4*>(recon + i);
      const float4 x = *reinterpret_cast<const float4*>(image + i);
      const float4 d = make_float4(r.x - x.x, r.y - x.y, r.z - x.z, r.w - x.w);
      s_mse += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
      if (d_recon) *reinterpret_cast<float4*>(d_recon + i) = make_float4(gr * d.x, gr * d.y, gr * d.z, gr * d.w);
    } else {
      for (long k = i; k < n_recon; ++k) {
        const float d = recon[k] - image[k];
        s_mse += d * d;
        if (d_recon) d_recon[k] = gr * d;
      }
    }
  }
  const float gs = sw / static_cast<float>(n_seg);
  for (long i = tid; i < n_seg; i += nthreads) {
    const float p = amap[i], t = mask[i];
    const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(log1pf(-p), -100.f);
    const float bce = -(t * lp + (1.f - t) * l1p);
    const float pt = expf(-bce);
    const float om = 1.f - pt;
    const float omg = (gamma == 2.f) ? om * om : powf(om, gamma);
    s_foc += alpha * omg * bce;
    if (d_amap) {
      const float dbce = (p - t) / fmaxf((1.f - p) * p, 1e-12f);
      // d/dbce [alpha (1-pt)^g bce] = alpha [(1-pt)^g + g (1-pt)^(g-1) pt bce]
      const float omg1 = (gamma == 2.f) ? om : ((om > 0.f) ? powf(om, gamma - 1.f) : 0.f);
      d_amap[i] = gs * alpha * (omg + gamma * omg1 * pt * bce) * dbce;
    }
  }
  __shared__ float red[2][8];
  s_mse = warp_sum(s_mse);
  s_foc = warp_sum(s_foc);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s_mse;
    red[1][threadIdx.x >> 5] = s_foc;
  }
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) {
      a += red[0][i];
      b += red[1][i];
    }
    work->partial[2 * blockIdx.x] = a;
    work->partial[2 * blockIdx.x + 1] = b;
    __threadfence();
    const unsigned int t = atomicAdd(&work->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  __shared__ double fold[2][256];
  if (is_last) {
    // the last block folds the per-block partials: every thread a strided slice, then a fixed-order tree
    // (deterministic; a single thread walking 1024 partials cost ~40 us on the forward->backward critical path)
    __threadfence();
    double a = 0.0, b = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
      a += reinterpret_cast<volatile double*>(work->partial)[2 * i];
      b += reinterpret_cast<volatile double*>(work->partial)[2 * i + 1];
    }
    fold[0][threadIdx.x] = a;
    fold[1][threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        fold[0][threadIdx.x] += fold[0][threadIdx.x + o];
        fold[1][threadIdx.x] += fold[1][threadIdx.x + o];
      }
      __syncthreads();
    }
  }
  if (is_last && threadIdx.x == 0) {
    const double a = fold[0][0], b = fold[1][0];
    const float rl = static_cast<float>(a / static_cast<double>(n_recon));
    const float sl = static_cast<float>(b / static_cast<double>(n_seg));
    out3[0] = rw * rl + sw * sl;
    out3[1] = rl;
    out3[2] = sl;
    work->ticket = 0;  // self-cleaning for the next launch
  }
}

// y[i] *= s[0]  (chain-rule scale when the upstream gradient is not 1)
__global__ void scale_by_scalar_kernel(float* __restrict__ y, const float* __restrict__ s, long n) {
  const float k = *s;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x)
    y[i] *= k;
}

// ----------------------------------------------------------------------------------------------
// anomaly score: out[n, y, x] = mean_c (r - x)^2   (method 0)  |  mean_c |r - x|  (method 1)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
anomaly_score_kernel(const float* __restrict__ recon, const float* __restrict__ orig,
                     float* __restrict__ out, long n_img, int C, long hw, int method) {
  const long total = n_img * hw;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long n = i / hw, r = i - n * hw;
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      const float d = recon[(n * C + c) * hw + r] - orig[(n * C + c) * hw + r];
      s += method == 0 ? d * d : fabsf(d);
    }
    out[i] = s / static_cast<float>(C);
  }
}

// ----------------------------------------------------------------------------------------------
// SSIM.  11-tap separable gaussian (sigma 1.5), zero padding 5, per (n, c) plane.
// forward: per-image sums of the SSIM map and, when `maps` != null, the three derivative planes
//   M1 = dS/dmu1, M11 = dS/dE[x1^2], M12 = dS/dE[x1 x2]   (and M2, M22 for the second input)
// backward: grad1 = k * ( G*M1 + 2 x1 (G*M11) + x2 (G*M12) ), k = -gout / count.
// ----------------------------------------------------------------------------------------------
constexpr int kSsimT = 32;            // output tile edge
constexpr int kSsimR = 5;             // window radius
constexpr int kSsimI = kSsimT + 2 * kSsimR;  // 42

struct SsimWin {
  float w[11];
};

template <int NMAPS>
__device__ __forceinline__ void blur_rows(const float (*src)[kSsimI][kSsimI + 1],
                                          float (*dst)[kSsimI][kSsimT + 1], const SsimWin& win) {
  // horizontal pass: dst[m][r][x] = sum_k w[k] * src[m][r][x + k]
  for (int i = threadIdx.x; i < kSsimI * kSsimT; i += blockDim.x) {
    const int r = i / kSsimT, x = i - r * kSsimT;
#pragma unroll
    for (int m = 0; m < NMAPS; ++m) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) s = fmaf(win.w[k], src[m][r][x + k], s);
      dst[m][r][x] = s;
    }
  }
}

__global__ void __launch_bounds__(256)
ssim_fwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2,
                float* __restrict__ sums, float* __restrict__ maps, int C, int H, int W,
                SsimWin win, int need2) {
  __shared__ float sin_[2][kSsimI][kSsimI + 1];
  __shared__ float sh[5][kSsimI][kSsimT + 1];
  const int plane = blockIdx.z;
  const int x0 = blockIdx.x * kSsimT, y0 = blockIdx.y * kSsimT;
  const long pbase = static_cast<long>(plane) * H * W;
  for (int i = threadIdx.x; i < kSsimI * kSsimI; i += blockDim.x) {
    const int r = i / kSsimI, c = i - r * kSsimI;
    const int y = y0 + r - kSsimR, x = x0 + c - kSsimR;
    const bool in = (y >= 0 && y < H && x >= 0 && x < W);
    sin_[0][r][c] = in ? img1[pbase + static_cast<long>(y) * W + x] : 0.f;
    sin_[1][r][c] = in ? img2[pbase + static_cast<long>(y) * W + x] : 0.f;
  }
  __syncthreads();
  // horizontal pass of the five product maps
  for (int i = threadIdx.x; i < kSsimI * kSsimT; i += blockDim.x) {
    const int r = i / kSsimT, x = i - r * kSsimT;
    float s1 = 0.f, s2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float a = sin_[0][r][x + k], b = sin_[1][r][x + k], wk = win.w[k];
      s1 = fmaf(wk, a, s1);
      s2 = fmaf(wk, b, s2);
      s11 = fmaf(wk, a * a, s11);
      s22 = fmaf(wk, b * b, s22);
      s12 = fmaf(wk, a * b, s12);
    }
    sh[0][r][x] = s1; sh[1][r][x] = s2; sh[2][r][x] = s11; sh[3][r][x] = s22; sh[4][r][x] = s12;
  }
  __syncthreads();
  float local = 0.f;
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  for (int i = threadIdx.x; i < kSsimT * kSsimT; i += blockDim.x) {
    const int r = i / kSsimT, x = i - r * kSsimT;
    const int gy = y0 + r, gx = x0 + x;
    if (gy >= H || gx >= W) continue;
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float wk = win.w[k];
      mu1 = fmaf(wk, sh[0][r + k][x], mu1);
      mu2 = fmaf(wk, sh[1][r + k][x], mu2);
      e11 = fmaf(wk, sh[2][r + k][x], e11);
      e22 = fmaf(wk, sh[3][r + k][x], e22);
      e12 = fmaf(wk, sh[4][r + k][x], e12);
    }
    const float mu1s = mu1 * mu1, mu2s = mu2 * mu2, mu12 = mu1 * mu2;
    const float sg1 = e11 - mu1s, sg2 = e22 - mu2s, sg12 = e12 - mu12;
    const float A1 = 2.f * mu12 + C1, A2 = 2.f * sg12 + C2;
    const float B1 = mu1s + mu2s + C1, B2 = sg1 + sg2 + C2;
    const float inv = 1.f / (B1 * B2);
    const float S = A1 * A2 * inv;
    local += S;
    if (maps != nullptr) {
      const long planes = static_cast<long>(gridDim.z);
      const long o = pbase + static_cast<long>(gy) * W + gx;
      const long ps = planes * H * W;
      const float dA = 2.f * (A2 - A1) * inv;      // coefficient of mu_other
      const float dB = 2.f * S * (B2 - B1) * inv;  // coefficient of mu_self
      maps[o] = mu2 * dA - mu1 * dB;               // dS/dmu1
      maps[ps + o] = -S / B2;                       // dS/dE11
      maps[2 * ps + o] = 2.f * A1 * inv;            // dS/dE12
      if (need2) {
        maps[3 * ps + o] = mu1 * dA - mu2 * dB;     // dS/dmu2
        maps[4 * ps + o] = -S / B2;                 // dS/dE22
      }
    }
  }
  __shared__ float red[8];
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(sums + plane / C, s);
  }
}

// which == 0: gradient wrt img1 (maps 0,1,2); which == 1: wrt img2 (maps 3,4,2)
__global__ void __launch_bounds__(256)
ssim_bwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2,
                const float* __restrict__ maps, const float* __restrict__ gout,
                float* __restrict__ grad, int C, int H, int W, SsimWin win, int which,
                float inv_count, int per_image) {
  __shared__ float sm[3][kSsimI][kSsimI + 1];
  __shared__ float sh[3][kSsimI][kSsimT + 1];
  const int plane = blockIdx.z;
  const long planes = static_cast<long>(gridDim.z);
  const long ps = planes * H * W;
  const int x0 = blockIdx.x * kSsimT, y0 = blockIdx.y * kSsimT;
  const long pbase = static_cast<long>(plane) * H * W;
  const float* m0 = maps + (which == 0 ? 0 : 3) * ps;
  const float* m1 = maps + (which == 0 ? 1 : 4) * ps;
  const float* m2 = maps + 2 * ps;
  for (int i = threadIdx.x; i < kSsimI * kSsimI; i += blockDim.x) {
    const int r = i / kSsimI, c = i - r * kSsimI;
    const int y = y0 + r - kSsimR, x = x0 + c - kSsimR;
    const bool in = (y >= 0 && y < H && x >= 0 && x < W);
    const long o = pbase + static_cast<long>(y) * W + x;
    sm[0][r][c] = in ? m0[o] : 0.f;
    sm[1][r][c] = in ? m1[o] : 0.f;
    sm[2][r][c] = in ? m2[o] : 0.f;
  }
  __syncthreads();
  blur_rows<3>(sm, sh, win);
  __syncthreads();
  const float k = -gout[per_image ? plane / C : 0] * inv_count;
  for (int i = threadIdx.x; i < kSsimT * kSsimT; i += blockDim.x) {
    const int r = i / kSsimT, x = i - r * kSsimT;
    const int gy = y0 + r, gx = x0 + x;
    if (gy >= H || gx >= W) continue;
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
    for (int t = 0; t < 11; ++t) {
      const float wk = win.w[t];
      b0 = fmaf(wk, sh[0][r + t][x], b0);
      b1 = fmaf(wk, sh[1][r + t][x], b1);
      b2 = fmaf(wk, sh[2][r + t][x], b2);
    }
    const long o = pbase + static_cast<long>(gy) * W + gx;
    const float xs = which == 0 ? img1[o] : img2[o];
    const float xo = which == 0 ? img2[o] : img1[o];
    grad[o] = k * (b0 + 2.f * xs * b1 + xo * b2);
  }
}

// out = 1 - mean(ssim_map): one scalar (size_average) or one value per image
__global__ void ssim_finalize_kernel(const float* __restrict__ sums, float* __restrict__ out, int n,
                                     float per_image_count, int size_average) {
  if (size_average) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      double s = 0.0;
      for (int i = 0; i < n; ++i) s += sums[i];
      out[0] = static_cast<float>(1.0 - s / (static_cast<double>(per_image_count) * n));
    }
  } else {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = 1.f - sums[i] / per_image_count;
  }
}

// ----------------------------------------------------------------------------------------------
// host launchers
// ----------------------------------------------------------------------------------------------
static int ew_blocks(long n, int per_block) {
  long b = (n + per_block - 1) / per_block;
  const long cap = static_cast<long>(num_sms()) * 8;
  if (b > cap) b = cap;
  return static_cast<int>(b < 1 ? 1 : b);
}

int head_fwd(const void* a, const float* w, const float* b, float* out, int N, int H, int W, int C,
             int n_out, int sigmoid, cudaStream_t s) {
  UNETK_CHECK_ARG(a && w && b && out, "head_fwd: null pointer");
  UNETK_CHECK_ARG(C == kHeadC, "head_fwd: input channels must be 64 (got %d)", C);
  UNETK_CHECK_ARG(n_out >= 1 && n_out <= kHeadMaxOut, "head_fwd: 1..4 output channels supported (got %d)", n_out);
  const long hw = static_cast<long>(H) * W, np = hw * N;
  UNETK_CUDA(launch_pdl(head_fwd_kernel<false>, dim3(ew_blocks(np, 256)), dim3(256), 0, s, static_cast<const uint4*>(a), w, b, out, np, hw, n_out, sigmoid,
                        BnFinalizeArgs{}));
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int head_fwd_bn(const void* y, const BnFinalizeArgs* fin, const float* w, const float* b, float* out, int N, int H,
                int W, int C, int n_out, int sigmoid, cudaStream_t s) {
  UNETK_CHECK_ARG(y && fin && w && b && out, "head_fwd_bn: null pointer");
  UNETK_CHECK_ARG(C == kHeadC, "head_fwd_bn: input channels must be 64 (got %d)", C);
  UNETK_CHECK_ARG(n_out >= 1 && n_out <= kHeadMaxOut, "head_fwd_bn: 1..4 output channels supported (got %d)", n_out);
  UNETK_CHECK_ARG(fin->stats && fin->gamma && fin->beta && fin->scale && fin->shift && fin->mean && fin->invstd,
                  "head_fwd_bn: incomplete BatchNorm finalize arguments");
  UNETK_CHECK_ARG((fin->running_mean == nullptr) == (fin->running_var == nullptr), "head_fwd_bn: running_mean and running_var go together");
  UNETK_CHECK_ARG(fin->count >= 1.0, "head_fwd_bn: empty batch");
  const long hw = static_cast<long>(H) * W, np = hw * N;
  UNETK_CUDA(launch_pdl(head_fwd_kernel<true>, dim3(ew_blocks(np, 256)), dim3(256), 0, s, static_cast<const uint4*>(y), w, b, out, np, hw, n_out, sigmoid,
                        with_reciprocals(*fin)));
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int head_bwd(const void* a, const float* w, const float* dout, const float* out, void* da, float* dw,
             float* db, int N, int H, int W, int C, int n_out, int sigmoid, const BnReduceCtx* bn, cudaStream_t s) {
  UNETK_CHECK_ARG(w && dout && da && dw && db, "head_bwd: null pointer");
  UNETK_CHECK_ARG(a != nullptr || (bn != nullptr && bn->y && bn->scale && bn->shift),
                  "head_bwd: without the stored activation the BatchNorm context (y, scale, shift) is required");
  UNETK_CHECK_ARG(!sigmoid || out, "head_bwd: sigmoid backward needs the forward output");
  UNETK_CHECK_ARG(C == kHeadC, "head_bwd: input channels must be 64 (got %d)", C);
  UNETK_CHECK_ARG(n_out >= 1 && n_out <= kHeadMaxOut, "head_bwd: 1..4 output channels supported");
  const long hw = static_cast<long>(H) * W, np = hw * N;
  UNETK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * n_out * kHeadC, s));
  UNETK_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * n_out, s));
  int blocks = ew_blocks((np + 127) / 128, 1);
  const int cap = num_sms() * 2;     // two resident blocks per SM, grid-stride
  if (blocks > cap) blocks = cap;
  const uint4* a4 = static_cast<const uint4*>(a);
  uint4* da4 = static_cast<uint4*>(da);
  const bool fuse = bn != nullptr && bn->y != nullptr && bn->sums != nullptr;     // + BatchNorm-backward sums
  const bool ctx = bn != nullptr && bn->y != nullptr && (fuse || a == nullptr);
  UNETK_CHECK_ARG(!fuse || (bn->scale && bn->shift && bn->mean && bn->invstd), "head_bwd: incomplete BatchNorm context");
  const uint4* by = ctx ? static_cast<const uint4*>(bn->y) : nullptr;
  const float* b1 = ctx ? bn->scale : nullptr; const float* b2 = ctx ? bn->shift : nullptr;
  const float* b3 = ctx ? bn->mean : nullptr;  const float* b4 = ctx ? bn->invstd : nullptr;
  float* b5 = fuse ? bn->sums : nullptr;
  switch (n_out) {
    case 1: UNETK_CUDA(launch_pdl(head_bwd_kernel<1>, dim3(blocks), dim3(256), 0, s, a4, w, dout, out, da4, dw, db, np, hw, sigmoid, by, b1, b2, b3, b4, b5)); break;
    case 2: UNETK_CUDA(launch_pdl(head_bwd_kernel<2>, dim3(blocks), dim3(256), 0, s, a4, w, dout, out, da4, dw, db, np, hw, sigmoid, by, b1, b2, b3, b4, b5)); break;
    case 3: UNETK_CUDA(launch_pdl(head_bwd_kernel<3>, dim3(blocks), dim3(256), 0, s, a4, w, dout, out, da4, dw, db, np, hw, sigmoid, by, b1, b2, b3, b4, b5)); break;
    default: UNETK_CUDA(launch_pdl(head_bwd_kernel<4>, dim3(blocks), dim3(256), 0, s, a4, w, dout, out, da4, dw, db, np, hw, sigmoid, by, b1, b2, b3, b4, b5)); break;
  }
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

size_t combined_loss_workspace_bytes() { return sizeof(LossWork); }

int combined_loss(const float* recon, const float* amap, const float* image, const float* mask,
                  float* d_recon, float* d_amap, float* out3, void* work, size_t work_bytes,
                  long n_recon, long n_seg, float rw, float sw, float alpha, float gamma,
                  cudaStream_t s) {
  UNETK_CHECK_ARG(recon && amap && image && mask && out3 && work, "combined_loss: null pointer");
  if (work_bytes < sizeof(LossWork)) {
    set_last_error("combined_loss: workspace too small (%zu < %zu)", work_bytes, sizeof(LossWork));
    return UNETK_ERR_WORKSPACE;
  }
  UNETK_CHECK_ARG((reinterpret_cast<uintptr_t>(recon) & 15) == 0 && (reinterpret_cast<uintptr_t>(image) & 15) == 0 &&
                      (d_recon == nullptr || (reinterpret_cast<uintptr_t>(d_recon) & 15) == 0),
                  "combined_loss: tensors must be 16-byte aligned");
  int blocks = ew_blocks(n_recon / 4 + n_seg, 256);
  if (blocks > 1024) blocks = 1024;
  combined_loss_kernel<<<blocks, 256, 0, s>>>(recon, amap, image, mask, d_recon, d_amap, out3, static_cast<LossWork*>(work),
                                             n_recon, n_seg, rw, sw, alpha, gamma);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int scale_by_scalar(float* y, const float* s, long n, cudaStream_t st) {
  UNETK_CHECK_ARG(y && s, "scale_by_scalar: null pointer");
  scale_by_scalar_kernel<<<ew_blocks(n, 256), 256, 0, st>>>(y, s, n);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int anomaly_score(const float* recon, const float* orig, float* out, int N, int C, int H, int W,
                  int method, cudaStream_t s) {
  UNETK_CHECK_ARG(recon && orig && out, "anomaly_score: null pointer");
  UNETK_CHECK_ARG(method == 0 || method == 1, "anomaly_score: unknown method %d", method);
  const long hw = static_cast<long>(H) * W;
  anomaly_score_kernel<<<ew_blocks(hw * N, 256), 256, 0, s>>>(recon, orig, out, N, C, hw, method);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

static SsimWin make_window(int window_size, int* err) {
  SsimWin win{};
  *err = 0;
  if (window_size != 11) {
    *err = 1;
    return win;
  }
  // gaussian(11, 1.5) as in SSIMLoss.gaussian (src/train_utils.py:57-59), normalised in fp32
  float g[11], sum = 0.f;
  for (int i = 0; i < 11; ++i) {
    g[i] = static_cast<float>(exp(-static_cast<double>((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5)));
    sum += g[i];
  }
  for (int i = 0; i < 11; ++i) win.w[i] = g[i] / sum;
  return win;
}

int ssim_fwd(const float* img1, const float* img2, float* sums, float* maps, int N, int C, int H,
             int W, int window_size, int need2, cudaStream_t s) {
  UNETK_CHECK_ARG(img1 && img2 && sums, "ssim_fwd: null pointer");
  int err;
  const SsimWin win = make_window(window_size, &err);
  UNETK_CHECK_ARG(!err, "ssim_fwd: only window_size 11 is supported (got %d)", window_size);
  UNETK_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * N, s));
  dim3 grid((W + kSsimT - 1) / kSsimT, (H + kSsimT - 1) / kSsimT, N * C);
  ssim_fwd_kernel<<<grid, 256, 0, s>>>(img1, img2, sums, maps, C, H, W, win, need2);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int ssim_finalize(const float* sums, float* out, int N, int C, int H, int W, int size_average,
                  cudaStream_t s) {
  UNETK_CHECK_ARG(sums && out, "ssim_finalize: null pointer");
  ssim_finalize_kernel<<<(N + 127) / 128, 128, 0, s>>>(sums, out, N, static_cast<float>(C) * H * W, size_average);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

int ssim_bwd(const float* img1, const float* img2, const float* maps, const float* gout, float* grad,
             int N, int C, int H, int W, int window_size, int which, int per_image, cudaStream_t s) {
  UNETK_CHECK_ARG(img1 && img2 && maps && gout && grad, "ssim_bwd: null pointer");
  int err;
  const SsimWin win = make_window(window_size, &err);
  UNETK_CHECK_ARG(!err, "ssim_bwd: only window_size 11 is supported (got %d)", window_size);
  const double count = per_image ? static_cast<double>(C) * H * W : static_cast<double>(N) * C * H * W;
  dim3 grid((W + kSsimT - 1) / kSsimT, (H + kSsimT - 1) / kSsimT, N * C);
  ssim_bwd_kernel<<<grid, 256, 0, s>>>(img1, img2, maps, gout, grad, C, H, W, win, which, static_cast<float>(1.0 / count),
                                      per_image);
  UNETK_CUDA(cudaGetLastError());
  return UNETK_OK;
}

}  // namespace unetk
