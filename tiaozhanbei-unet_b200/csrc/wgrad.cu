// Weight-gradient contractions on tcgen05 / TMEM fed by TMA (sm_100a).
//
//   conv3x3 wgrad   dW[co, tap, ci]  = sum_pix dY[pix, co] * X[pix + tap, ci]
//   convT2x2 wgrad  dW[co, tap, ci]  = sum_pix dUp[2*pix + tap, co] * A[pix, ci]
// (autograd of nn.Conv2d at /root/reference/src/model.py:14,17 and nn.ConvTranspose2d at :51).
//
// The reduction runs over pixels, which is the OUTER dimension of NHWC tensors, so both operands
// reach shared memory pixel-row by pixel-row (one TMA box = 64 pixels x 64 channels = 64 rows of
// 128 B, SWIZZLE_128B) and are consumed through MN-major UMMA descriptors: no transposing copy.
// UMMA M = 128 input channels (or two taps x 64 channels when Cin == 64), N = BN output channels,
// K = 16 pixels per instruction.  One CTA owns one (tap row | 4 taps) x 128 ci x BN co block of dW
// in TMEM and a slice of the pixel range (split-K); partial sums are added to the fp32 gradient
// with red.global.add.f32.
// conv3x3: the K tile is an 8x8 pixel patch; the X operand of the three dx taps of a tap row comes
// from ONE 10x8 halo box per 64-channel chunk (80 rows x 128 B): a dx shift is a descriptor start
// offset of dx*128 B, 8-pixel patch rows are the 8-row K groups with SBO = 10*128 B (the 128B
// swizzle is address based, profiles/umma_probe_r01.txt).  X traffic drops 2.4x.
// MODE 2 (narrow layers, Cin*Cout <= 64*128, where the tap-row kernel is L2-bandwidth bound because
// three CTAs re-read the same pixels): one CTA owns ALL NINE taps of a 64 ci x 64 co block -- M = two
// taps x 64 ci per instruction (five tap pairs, the last half empty), X from ONE 10x10 halo box per
// K tile.  The second tap of a pair is an LBO away (128 B for a dx step, 1024 B from tap (0,2) to (1,0)).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace unetk {

struct WgradParams {
  int n_img, H, W;  // pixel grid that is reduced over
  int tw, th, tn;   // 64-pixel K tile
  int tiles_x, tiles_y, tiles_n;
  int cin, cout;
  int c1_chunks;    // 64-channel chunks of the X operand that live in tensor map 1 (rest: map 2)
  int ci_tiles, co_tiles;
  int ksplit;
  int T;            // taps per weight: 9 or 4
  float* dw;        // [cout][T][cin]
};

constexpr int kWgThreads = 192;
constexpr int kBoxBytes = 64 * 128;   // 64 pixels x 64 bf16
constexpr int kHaloBytes = 80 * 128;  // 10 x 8 pixels x 64 bf16 (conv3x3 X operand)
constexpr int kHaloPitch = 10 * 128;  // bytes between patch rows inside the halo box
constexpr int kHalo9Bytes = 100 * 128;  // 10 x 10 pixels x 64 bf16 (MODE 2)

// CTAS == 2 (conv3x3, BN == 128): cta_group::2 pairs.  The two CTAs own different (tap row, ci tile)
// units (M = 256 per instruction) and each stages only 64 of the 128 dY channels: per CTA an MMA
// fetches A 4 KB + B 2 KB from shared memory instead of 4 + 4 (SS-mode operand fetch is the limiter).
template <int MODE, int BN, bool M64, int CTAS = 1>
struct WgCfg {
  // per stage: A (X operand) and B (dY / dUp operand) regions
  static constexpr int kALoads = (MODE == 2) ? 1 : (MODE == 0) ? (M64 ? 1 : 2) : 2;
  static constexpr int kABoxBytes = (MODE == 2) ? kHalo9Bytes : (MODE == 0) ? kHaloBytes : kBoxBytes;
  // M64 / MODE 2: the second tap of the last pair reads a few rows past the box; keep that inside the stage
  static constexpr int kABytes = (MODE == 2) ? 13 * 1024 : kALoads * kABoxBytes + ((MODE == 0 && M64) ? 1024 : 0);
  static constexpr int kBBoxes = (MODE == 2) ? 1 : (MODE == 0) ? (BN / 64 / CTAS) : 4 * (BN / 64);   // per CTA
  static constexpr int kMmas = (MODE == 2) ? 5 : (MODE == 0) ? (M64 ? 2 : 3) : 4;
  static constexpr int kStageBytes = kABytes + kBBoxes * kBoxBytes;
  static constexpr int kMaxFit = (210 * 1024) / kStageBytes;
  static constexpr int kStages = kMaxFit > 10 ? 10 : (kMaxFit < 2 ? 2 : kMaxFit);
  static constexpr int kAccCols = kMmas * BN;
  static constexpr int kTmemCols = kAccCols <= 64 ? 64 : (kAccCols <= 128 ? 128 : (kAccCols <= 256 ? 256 : 512));
};

template <int MODE, int BN, bool M64, int CTAS>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
             const __grid_constant__ CUtensorMap tmDY, const WgradParams p) {
  using Cfg = WgCfg<MODE, BN, M64, CTAS>;
  static_assert(CTAS == 1 || (MODE == 0 && BN == 128), "CTA pairs: conv3x3 with 128-wide co tiles only");
  static_assert(MODE != 2 || (BN == 64 && CTAS == 1), "nine-tap mode: 64 x 64 blocks, single CTA");
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work decomposition: unit = (tap row r, ci tile); a CTA pair takes two consecutive units
  const uint32_t rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  int bx = blockIdx.x / CTAS;
  const int cot = bx % p.co_tiles;
  bx /= p.co_tiles;
  const int unit = bx * CTAS + static_cast<int>(rank);
  const int n_units = (MODE == 0 ? 3 : 1) * p.ci_tiles;   // MODE 2: ci_tiles counts 64-channel chunks
  const bool live = unit < n_units;          // odd unit count: the last pair has one idle CTA
  const int cit = unit % p.ci_tiles;
  const int r = unit / p.ci_tiles;  // tap row (MODE 0), 0 otherwise
  const int total_k = p.tiles_x * p.tiles_y * p.tiles_n;
  const int k_begin = static_cast<int>(static_cast<long>(total_k) * blockIdx.y / p.ksplit);
  const int k_end = static_cast<int>(static_cast<long>(total_k) * (blockIdx.y + 1) / p.ksplit);
  pdl_trigger();
  if (k_end <= k_begin) return;  // uniform for the whole CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX1);
    tma_prefetch_desc(&tmX2);
    tma_prefetch_desc(&tmDY);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (CTAS == 2) cluster_sync_all();
  if (warp == 1) {
    if (CTAS == 2) tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    else tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();          // set-up above overlapped the previous kernel of the stream; global memory from here on

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = k_begin; kt < k_end; ++kt) {
        int m = kt;
        const int tx = m % p.tiles_x;
        m /= p.tiles_x;
        const int ty = m % p.tiles_y;
        const int tz = m / p.tiles_y;
        const int x0 = tx * p.tw, y0 = ty * p.th, n0 = tz * p.tn;
        const int n0a = live ? n0 : p.n_img;   // idle CTA of a pair: X out of bounds -> zeros
        mbar_wait(&empty[stage], phase ^ 1);
        if (rank == 0) mbar_expect_tx(&full[stage], (Cfg::kALoads * Cfg::kABoxBytes + Cfg::kBBoxes * kBoxBytes) * CTAS);
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        if (MODE == 2) {
          const CUtensorMap* mp = (cit < p.c1_chunks) ? &tmX1 : &tmX2;
          const int cc = ((cit < p.c1_chunks) ? cit : cit - p.c1_chunks) * 64;
          tma_load_5d(sa, mp, &full[stage], cc, x0 - 1, y0 - 1, n0, 0);      // 10x10 halo: all nine taps
          tma_load_5d(sb, &tmDY, &full[stage], cot * BN, x0, y0, n0, 0);
        } else if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < Cfg::kALoads; ++j) {
            const int chunk = M64 ? 0 : cit * 2 + j;
            const CUtensorMap* mp = (chunk < p.c1_chunks) ? &tmX1 : &tmX2;
            const int cc = ((chunk < p.c1_chunks) ? chunk : chunk - p.c1_chunks) * 64;
            // halo box of tap row r: pixels (x0-1 .. x0+8, y0+r-1 .. y0+r+6)
            if (CTAS == 2) tma_load_5d_pair(sa + j * kHaloBytes, mp, &full[stage], cc, x0 - 1, y0 + r - 1, n0a, 0);
            else tma_load_5d(sa + j * kHaloBytes, mp, &full[stage], cc, x0 - 1, y0 + r - 1, n0a, 0);
          }
#pragma unroll
          for (int j = 0; j < Cfg::kBBoxes; ++j) {
            const int co = cot * BN + (static_cast<int>(rank) * Cfg::kBBoxes + j) * 64;   // this CTA's share of dY
            if (CTAS == 2) tma_load_5d_pair(sb + j * kBoxBytes, &tmDY, &full[stage], co, x0, y0, n0, 0);
            else tma_load_5d(sb + j * kBoxBytes, &tmDY, &full[stage], co, x0, y0, n0, 0);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_5d(sa + j * kBoxBytes, &tmX1, &full[stage], (cit * 2 + j) * 64, x0, y0, n0, 0);
#pragma unroll
          for (int j = 0; j < Cfg::kBBoxes; ++j) {
            const int tap = j / (BN / 64);
            tma_load_5d(sb + j * kBoxBytes, &tmDY, &full[stage], cot * BN + (j % (BN / 64)) * 64,
                        tap & 1, x0, tap >> 1, n0 * p.H + y0);
          }
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // warp-uniform loop, one elected lane issues (keeps descriptors in uniform registers)
      constexpr uint32_t idesc = umma_idesc_bf16(128 * CTAS, BN, true, true);
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = k_begin; kt < k_end; ++kt) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t sb = sa + Cfg::kABytes;
        // base descriptors; per-MMA variation is a start-address offset in 16-byte units
        const uint64_t a_desc0 = (MODE == 2) ? umma_smem_desc_sw128(sa, 0, kHaloPitch)
                                 : (MODE == 0) ? umma_smem_desc_sw128(sa, M64 ? 128 : kHaloBytes, kHaloPitch)
                                               : umma_smem_desc_sw128(sa, kBoxBytes, 1024);
        const uint64_t b_desc0 = umma_smem_desc_sw128(sb, kBoxBytes, 1024);
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < Cfg::kMmas; ++j) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // B: K advances by 16 pixel rows = 2048 B; 64-wide MN atoms are one box (8 KB) apart.
              const uint64_t db = b_desc0 + (((MODE == 1) ? j * (BN / 64) : 0) * (kBoxBytes / 16) + k * 128);
              // X halo (MODE 0): tap dx starts dx rows in; 16 pixels = two patch rows of the 10-wide
              // halo.  !M64: M = 128 ci = two halo boxes (LBO).  M64: M = taps (2j, 2j+1) x 64 ci.
              // MODE 2: tap t = 2j sits ((t/3)*10 + t%3) pixel rows into the 10x10 halo, tap t+1 an LBO further
              const int t0 = 2 * j, t1 = 2 * j + 1;                                       // (8, -): the pad half is discarded
              const uint32_t off0 = static_cast<uint32_t>((t0 / 3) * 10 + t0 % 3) * 8;     // 16-byte units
              const uint32_t off1 = (t1 < 9) ? static_cast<uint32_t>((t1 / 3) * 10 + t1 % 3) * 8 : off0 + 8;
              const uint64_t da = (MODE == 2) ? a_desc0 + (off0 + k * 2 * (kHaloPitch / 16)) + (static_cast<uint64_t>(off1 - off0) << 16)
                                : (MODE == 0) ? a_desc0 + ((M64 ? 2 * j : j) * 8 + k * 2 * (kHaloPitch / 16))
                                              : a_desc0 + (k * 128);
              if (CTAS == 2) tc_mma_bf16_pair(tmem_base + j * BN, da, db, idesc, (kt > k_begin) || (k > 0));
              else tc_mma_bf16(tmem_base + j * BN, da, db, idesc, (kt > k_begin) || (k > 0));
            }
          }
          if (CTAS == 2) {
            tc_commit_pair(&empty[stage]);
            if (kt == k_end - 1) tc_commit_pair(tfull);
          } else {
            tc_commit(&empty[stage]);
            if (kt == k_end - 1) tc_commit(tfull);
          }
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(tfull, 0);
    tc_fence_after();
#pragma unroll 1
    for (int j = 0; j < Cfg::kMmas; ++j) {
      int tap, ci;
      bool valid = true;
      if (MODE == 2) {
        tap = 2 * j + (row >> 6);
        ci = cit * 64 + (row & 63);
        valid = tap < 9;
      } else if (MODE == 0) {
        if (M64) {
          const int tj = 2 * j + (row >> 6);
          tap = 3 * r + tj;
          ci = row & 63;
          valid = tj < 3;
        } else {
          tap = 3 * r + j;
          ci = cit * 128 + row;
        }
      } else {
        tap = j;
        ci = cit * 128 + row;
      }
      valid = valid && live && (ci < p.cin);
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + j * BN + ch * 32, v);
        tmem_ld_wait();
        const int co0 = cot * BN + ch * 32;
        if (valid) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const int co = co0 + jj;
            if (co < p.cout) {
              float* dst = p.dw + (static_cast<size_t>(co) * p.T + tap) * p.cin + ci;
              if (p.ksplit == 1) *dst = v[jj]; else atomicAdd(dst, v[jj]);
            }
          }
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// from igemm.cu
void pick_pixel_tile(int n_img, int H, int W, int pixels, int* tw, int* th, int* tn);
int make_act_tmap(CUtensorMap* tm, const void* base, int N, int H, int W, int C, int tw, int th, int tn);
int make_act_tmap_up2(CUtensorMap* tm, const void* base, int N, int H, int W, int C, int tw, int th, int tn);

template <int MODE, int BN, bool M64, int CTAS = 1>
static int launch_wgrad_t(const CUtensorMap& x1, const CUtensorMap& x2, const CUtensorMap& dy,
                          const WgradParams& p, dim3 grid, cudaStream_t stream) {
  using Cfg = WgCfg<MODE, BN, M64, CTAS>;
  const size_t smem = 1024 + Cfg::kStages * Cfg::kStageBytes + (2 * Cfg::kStages + 2) * 8;
  auto kern = wgrad_kernel<MODE, BN, M64, CTAS>;
  UNETK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kWgThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  UNETK_CUDA(cudaLaunchKernelEx(&cfg, kern, x1, x2, dy, p));
  return UNETK_OK;
}

// Split-K factor from a small cost model (units: microseconds), fitted to a sweep of every layer shape of
// the B=16 @256^2 step on a B200 (tools/wgrad_sweep.py -> profiles/wgrad_sweep_r01.txt): a CTA costs a fixed
// ~10 us (launch, prologue, pipeline fill, drain) + its K tiles + its epilogue (red.global.add when
// ksplit > 1, plain stores otherwise); the model picks the measured optimum for all 13 conv shapes.
static int choose_ksplit(int work_tiles, int total_k, double us_per_ktile, double out_elems_per_cta) {
  if (const char* f = getenv("UNETK_WGRAD_KSPLIT")) {       // tuning override (tools/wgrad_sweep.py)
    const int v = atoi(f);
    if (v >= 1) return v < total_k ? v : total_k;
  }
  const int sms = num_sms();
  int best = 1;
  double best_t = 1e30;
  const int max_ks = total_k < 512 ? total_k : 512;
  for (int ks = 1; ks <= max_ks; ++ks) {
    const long ctas = static_cast<long>(work_tiles) * ks;
    const double waves = static_cast<double>((ctas + sms - 1) / sms);
    const double main_us = (static_cast<double>((total_k + ks - 1) / ks)) * us_per_ktile;
    const double epi_us = out_elems_per_cta * ((ks == 1) ? 4e-5 : 1e-4);
    const double t = waves * (main_us + epi_us + 10.0);
    if (t < best_t) {
      best_t = t;
      best = ks;
    }
  }
  return best;
}

int conv3x3_wgrad(const void* dy, const void* src1, int c1, const void* src2, int c2, float* dw,
                  int n, int h, int w, int cout, int h2, int w2, cudaStream_t stream) {
  if (h2 <= 0 || w2 <= 0) { h2 = h; w2 = w; }      // extent of src2 inside the pixel grid (zero beyond: src/model.py:57-61)
  UNETK_CHECK_ARG(h2 <= h && w2 <= w, "conv3x3_wgrad: the second source (%dx%d) must fit inside the pixel grid (%dx%d)", h2, w2, h, w);
  UNETK_CHECK_ARG(dy && src1 && dw, "conv3x3_wgrad: null pointer");
  UNETK_CHECK_ARG(c2 == 0 || src2 != nullptr, "conv3x3_wgrad: second source missing");
  const int cin = c1 + c2;
  UNETK_CHECK_ARG(c1 % 64 == 0 && c2 % 64 == 0 && cout % 64 == 0, "conv3x3_wgrad: channels must be multiples of 64");
  UNETK_CHECK_ARG(cin == 64 || cin % 128 == 0, "conv3x3_wgrad: cin must be 64 or a multiple of 128 (got %d)", cin);
  if (int e = check_device_sm100()) return e;
  WgradParams p{};
  p.n_img = n; p.H = h; p.W = w; p.cin = cin; p.cout = cout; p.c1_chunks = c1 / 64; p.T = 9; p.dw = dw;
  p.tw = 8; p.th = 8; p.tn = 1;   // 8x8 pixel K tiles, X read through 10x8 halo boxes
  p.tiles_x = (w + 7) / 8; p.tiles_y = (h + 7) / 8; p.tiles_n = n;
  static const bool nine_ok = (getenv("UNETK_WGRAD_NO9") == nullptr);
  if (nine_ok && static_cast<long>(cin) * cout <= 64 * 128) {
    // narrow layer: nine-tap units of 64 ci x 64 co (see the header comment)
    p.ci_tiles = cin / 64;
    p.co_tiles = cout / 64;
    const int work = p.ci_tiles * p.co_tiles;
    // one K tile = 5 x 4 MMAs of 128 x 64 x 16, shared-memory bound at ~48 cycles each
    p.ksplit = choose_ksplit(work, p.tiles_x * p.tiles_y * p.tiles_n, 0.65, 5 * 128 * 64.0);
    CUtensorMap x1, x2, tdy;
    int e;
    if ((e = make_act_tmap(&x1, src1, n, h, w, c1, 10, 10, 1))) return e;
    if (c2 > 0) { if ((e = make_act_tmap(&x2, src2, n, h2, w2, c2, 10, 10, 1))) return e; } else x2 = x1;
    if ((e = make_act_tmap(&tdy, dy, n, h, w, cout, 8, 8, 1))) return e;
    return launch_wgrad_t<2, 64, true>(x1, x2, tdy, p, dim3(work, p.ksplit), stream);
  }
  const bool m64 = (cin == 64);
  const int bn = (cout % 128 == 0) ? 128 : 64;
  p.ci_tiles = m64 ? 1 : cin / 128;
  p.co_tiles = cout / bn;
  static const bool pair_ok = (getenv("UNETK_WGRAD_1CTA") == nullptr);
  const int ctas = (pair_ok && bn == 128) ? 2 : 1;
  const int unit_slots = (3 * p.ci_tiles + ctas - 1) / ctas;       // (tap row, ci tile) units per CTA (pair)
  const int work = unit_slots * ctas * p.co_tiles;                  // CTAs in grid.x
  // one K tile = (2 or 3) x 4 MMAs of 128 x bn x 16 (M64: N = bn at half the tile rows): 0.40 us measured
  // for the pair configuration, proportionally less for the narrower ones
  p.ksplit = choose_ksplit(work, p.tiles_x * p.tiles_y * p.tiles_n, 0.40 * (ctas == 2 ? 1.0 : bn / 128.0) * (m64 ? 0.7 : 1.0),
                           (m64 ? 3 * 64 : 3 * 128) * static_cast<double>(bn));
  CUtensorMap x1, x2, tdy;
  int e;
  if ((e = make_act_tmap(&x1, src1, n, h, w, c1, 10, 8, 1))) return e;
  if (c2 > 0) { if ((e = make_act_tmap(&x2, src2, n, h2, w2, c2, 10, 8, 1))) return e; } else x2 = x1;
  if ((e = make_act_tmap(&tdy, dy, n, h, w, cout, 8, 8, 1))) return e;
  dim3 grid(work, p.ksplit);
  if (ctas == 2) return m64 ? launch_wgrad_t<0, 128, true, 2>(x1, x2, tdy, p, grid, stream)
                            : launch_wgrad_t<0, 128, false, 2>(x1, x2, tdy, p, grid, stream);
  if (m64) return bn == 128 ? launch_wgrad_t<0, 128, true>(x1, x2, tdy, p, grid, stream)
                            : launch_wgrad_t<0, 64, true>(x1, x2, tdy, p, grid, stream);
  return bn == 128 ? launch_wgrad_t<0, 128, false>(x1, x2, tdy, p, grid, stream)
                   : launch_wgrad_t<0, 64, false>(x1, x2, tdy, p, grid, stream);
}

int convt2x2_wgrad(const void* a, int cin, const void* dup, int cout, float* dw, int n, int h, int w,
                   cudaStream_t stream) {
  UNETK_CHECK_ARG(a && dup && dw, "convt2x2_wgrad: null pointer");
  UNETK_CHECK_ARG(cin % 128 == 0 && cout % 64 == 0, "convt2x2_wgrad: cin %% 128 and cout %% 64 required");
  if (int e = check_device_sm100()) return e;
  WgradParams p{};
  // no spatial coupling: merge N and H so that any tile shape is legal for the 5-D gather map
  p.n_img = 1; p.H = n * h; p.W = w; p.cin = cin; p.cout = cout; p.c1_chunks = cin / 64; p.T = 4; p.dw = dw;
  pick_pixel_tile(1, p.H, w, 64, &p.tw, &p.th, &p.tn);
  p.tiles_x = (w + p.tw - 1) / p.tw; p.tiles_y = (p.H + p.th - 1) / p.th; p.tiles_n = 1;
  const int bn = (cout % 128 == 0) ? 128 : 64;
  p.ci_tiles = cin / 128;
  p.co_tiles = cout / bn;
  const int work = p.ci_tiles * p.co_tiles;
  p.ksplit = choose_ksplit(work, p.tiles_x * p.tiles_y, 0.55 * bn / 128.0, 4 * 128.0 * bn);
  CUtensorMap x1, tdy;
  int e;
  if ((e = make_act_tmap(&x1, a, 1, p.H, w, cin, p.tw, p.th, p.tn))) return e;
  if ((e = make_act_tmap_up2(&tdy, dup, 1, p.H, w, cout, p.tw, p.th, p.tn))) return e;
  dim3 grid(work, p.ksplit);
  return bn == 128 ? launch_wgrad_t<1, 128, false>(x1, x1, tdy, p, grid, stream)
                   : launch_wgrad_t<1, 64, false>(x1, x1, tdy, p, grid, stream);
}

}  // namespace unetk
