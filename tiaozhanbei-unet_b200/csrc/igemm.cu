// Implicit-GEMM convolution on tcgen05 / TMEM fed by TMA (sm_100a).
//
// One kernel covers every K-major contraction of the U-Net hot path:
//   conv3x3 forward   D[pix, co]  = sum_{tap, ci} A[pix + tap, ci]   * Wf[co, tap, ci]
//   conv3x3 dgrad     D[pix, ci]  = sum_{tap, co} dY[pix + tap, co]  * Wd[ci, tap, co]   (taps pre-flipped)
//   convT2x2 forward  D[pix, (tap, co)] = sum_ci A[pix, ci] * Wt[(tap, co), ci]          (pixel-shuffle store)
//   convT2x2 dgrad    D[pix, ci]  = sum_{tap, co} dUp[2*pix + tap, co] * Wtd[ci, (tap, co)]
// Reference operators being replaced: nn.Conv2d(3x3, pad 1, bias=False) at
// /root/reference/src/model.py:14,17 and nn.ConvTranspose2d(k=2, s=2) at src/model.py:51, plus
// torch.cat([x2, x1], dim=1) at src/model.py:65 (folded into the K loop: two A tensor maps).
//
// Data movement (the kernel is L2-bandwidth bound, not tensor bound, unless operands are reused):
//  * One CTA tile is a 16x16-pixel region = two 8x16 sub-tiles of 128 GEMM rows each.
//  * For a 3x3 conv the A operand of ALL NINE taps comes from ONE TMA load per 64-channel chunk:
//    the 18x18-pixel halo of the region (324 rows x 128 B, SWIZZLE_128B, zero-filled outside the
//    image = the conv padding).  A tap shift is a descriptor START OFFSET of (dy*18 + dx) rows:
//    the 128B swizzle is a function of absolute shared-memory address bits, so a start that is a
//    multiple of 128 B (not 1024 B) addresses the same swizzled data (measured:
//    profiles/umma_probe_r01.txt).  8-pixel tile rows are the 8-row core-matrix groups, the next
//    tile row is SBO = 18*128 B away.  A traffic drops 7x versus one load per tap.
//  * Each weight stage (BN x 64, one tap x one channel chunk) is multiplied into BOTH sub-tiles
//    (two TMEM accumulators), halving weight traffic per FLOP.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (one elected lane), warps 2..5 epilogue
// (TMEM -> registers -> bf16 global, optional per-channel sum / sum-of-squares for BatchNorm,
// optional per-channel affine + ReLU for eval-mode folded BatchNorm).  Persistent over tiles with
// double-buffered TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace unetk {

struct IgemmParams {
  int n_img, H, W;      // GEMM pixel grid (rows of D)
  int H2, W2;           // extent of the second source / destination inside that grid (origin-aligned; the
                        // reference zero-pads the up-sampled tensor to the skip's size, src/model.py:57-61)
  int regions_x, regions_y;
  int n_blocks;         // Ntot / BN
  int Ntot;             // columns of D
  int num_taps;         // 9 (3x3), 1 (1x1 view), 4 (convT dgrad)
  int a_mode;           // 0: A box at (c, x+dx, y+dy, n, 0); 1: A box at (c, kx, x, ky, n*H+y)
  int halo;             // 1 for the 3x3 conv (one halo load serves 9 taps), else 0
  int c1_chunks, c2_chunks;  // 64-channel chunks taken from tmA1 / tmA2
  int epi_mode;         // 0: NHWC store, 1: convT pixel-shuffle store (+bias)
  int split;            // columns [0,split) -> out1, [split,Ntot) -> out2
  __nv_bfloat16* out1;
  __nv_bfloat16* out2;
  int ldc1, ldc2;
  int acc1, acc2;       // 1: read-modify-write (gradient accumulation)
  int tma1, tma2;       // 1: this destination is written by TMA stores from a swizzled staging tile (tmO1 / tmO2)
  float* stats;         // nullable: per-column sums for columns [stats_from, Ntot), stats[col - stats_from]
  float* stats_sq;      // nullable: per-column sums of squares, same indexing
  int stats_from;
  const float* scale;   // nullable per-column affine (eval-mode BatchNorm folded in)
  const float* shift;
  int relu;
  const float* bias;    // epi_mode 1: [cout_t]
  int cout_t;           // epi_mode 1: channels of the up-sampled tensor
  // Fused first pass of the BatchNorm+ReLU backward of the layer that PRODUCED this gradient's
  // activation (epi_mode 0, single destination): with y that layer's raw conv output,
  //   g = d * [y*scale + shift > 0],  stats += sum g,  stats_sq += sum g * (y - mean) * invstd
  const __nv_bfloat16* bn_y;
  const float* bn_scale;
  const float* bn_shift;
  const float* bn_mean;
  const float* bn_invstd;
};

// warps: 0 TMA, 1 MMA, 2.. epilogue (EPW of them).  One epilogue warp per TMEM lane quarter by default: two
// per quarter (8 warps, 320 threads, 168-register cap) measured SLOWER on the K=576 conv layers (fwd
// 0.107 -> 0.113 ms, dgrad 0.086 -> 0.099 ms at 64 ch x 256^2 x 16), whose limit is the operand fetch.  The
// transposed-conv forward is different: K = Cin is short (8..64 MMAs per tile) and its pixel-shuffle
// epilogue moves 256 columns per tile, so it is epilogue bound (ncu: 9 % tensor pipe, 21 % DRAM at
// Cin = 128) and takes EPW = 8.
constexpr int igemm_threads(int epw) { return 64 + 32 * epw; }
constexpr int kRegionH = 16;                  // CTA tile: (8*SUBS) x 16 output pixels
constexpr int kSubW = 8;                      // sub-tile: 8 wide x 16 tall = 128 GEMM rows

// BN   : columns of D per tile (per CTA pair when CTAS == 2)
// SUBS : 8x16 sub-tiles per CTA (1 or 2), each with its own TMEM accumulator
// CTAS : 1, or 2 = cta_group::2: two CTAs of a cluster (one region each, M = 256 per instruction) share
//        every weight stage: each CTA stages only BN/2 weight rows, the leader CTA issues the MMAs.
//        SS-mode UMMA fetches its operands from shared memory at ~64 B/clk/SM (measured: N=128 -> 50 %,
//        N=64 -> 33 % tensor-pipe activity with everything else waiting on the MMA), so halving the B
//        bytes per CTA is what lifts the tensor pipe: 256x256 pair tiles read A 4 KB + B 4 KB per
//        128-cycle instruction.
template <int BN, int SUBS, int CTAS, bool BNRED = false, int EPW_ = 4, bool OSTG = false>
struct IgemmCfg {
  static constexpr int kBNLocal = BN / CTAS;
  static constexpr int kBBytes = kBNLocal * 128;
  static constexpr int kABufBytes = (SUBS == 2) ? 41 * 1024 : 23 * 1024;   // >= (8*SUBS+2)*18*128, 1 KB multiple
  static constexpr int kBudget = 196 * 1024;
  // halo buffers in flight (a third buffer for the narrow BN = 64 tiles was measured: no change)
  static constexpr int kNumABufs = 2;
  // BNRED: the tile of the producing layer's raw output y that belongs to the OUTPUT tile (region x BN columns),
  // fetched by TMA while the tile's MMAs run: one 64-channel chunk = (8*SUBS x 16 pixels) x 128 B
  static constexpr int kYChunkBytes = SUBS * 16 * 1024;
  static constexpr int kYBytes = (BN / 64) * kYChunkBytes;
  static constexpr int kNumYBufs = BNRED ? (BN == 64 ? 2 : 1) : 0;     // narrow tiles are epilogue bound: two in flight
  // Output staging for the TMA stores: per epilogue warp two buffers of SUBS x (32 pixel rows x 64 channels)
  static constexpr int kOWarpBytes = 2 * SUBS * 4096;
  static constexpr int kOBytes = OSTG ? EPW_ * kOWarpBytes : 0;
  static constexpr int kFit = (kBudget - kNumABufs * kABufBytes - kNumYBufs * kYBytes - kOBytes) / kBBytes;
  static constexpr int kBStages = kFit > 16 ? 16 : kFit;
  static constexpr int kTmemCols = 2 * SUBS * BN;    // 2 accumulator stages x SUBS sub-tiles
  static constexpr int kSmemMain = kNumABufs * kABufBytes + kBStages * kBBytes + kNumYBufs * kYBytes + kOBytes;
  static constexpr int kNumBars = 2 * kNumABufs + 2 * kBStages + 4 + 4;
  static_assert(kTmemCols <= 512 && kBStages >= 3, "bad igemm configuration");
};

// BNRED: the epilogue also runs the first pass of the BatchNorm+ReLU backward of the layer that produced
//        the activation whose gradient this kernel completes (IgemmParams::bn_*), see unetk_bn_ctx.
// OSTG: the kernel has output staging tiles and writes the destinations flagged in IgemmParams::tma1/2 by TMA
template <int BN, int SUBS, int CTAS, bool BNRED, int EPW, bool OSTG>
__global__ void __launch_bounds__(igemm_threads(EPW), 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY,
             const __grid_constant__ CUtensorMap tmO1, const __grid_constant__ CUtensorMap tmO2, const IgemmParams p) {
  using Cfg = IgemmCfg<BN, SUBS, CTAS, BNRED, EPW, OSTG>;
  static_assert(!(BNRED && OSTG), "the BatchNorm-reduction variants keep per-thread stores (no room for staging)");
  constexpr int kNumYBufs = Cfg::kNumYBufs;
  constexpr int kBStages = Cfg::kBStages;
  constexpr int kABufBytes = Cfg::kABufBytes;
  constexpr int kNumABufs = Cfg::kNumABufs;
  constexpr int kRegionW = kSubW * SUBS;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kNumABufs * kABufBytes;
  uint8_t* sY = sB + kBStages * Cfg::kBBytes;            // BNRED only (1 KB aligned: every region above is)
  uint8_t* sO = sY + kNumYBufs * Cfg::kYBytes;           // output staging of the epilogue warps
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kSmemMain);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kNumABufs;
  uint64_t* b_full = a_empty + kNumABufs;
  uint64_t* b_empty = b_full + kBStages;
  uint64_t* tfull = b_empty + kBStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* yfull = tempty + 2;
  uint64_t* yempty = yfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(yempty + 2);
  float* colbuf = reinterpret_cast<float*>(yempty + 4);  // affine: [2][Ntot]; stats: [4][2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  pdl_trigger();       // the next kernel of the stream may set itself up while this one runs
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kNumABufs; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], EPW * CTAS);  // one arrive per epilogue warp (of both CTAs of a pair)
      mbar_init(&yfull[i], 1);
      mbar_init(&yempty[i], EPW);         // this CTA's epilogue warps
    }
    if (BNRED) tma_prefetch_desc(&tmY);
    if (p.tma1) tma_prefetch_desc(&tmO1);
    if (p.tma2) tma_prefetch_desc(&tmO2);
    fence_barrier_init();
  }
  if (CTAS == 2) cluster_sync_all();     // pair TMEM allocation needs both CTAs resident
  if (warp == 1) {
    if (CTAS == 2) tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    else tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  // everything above (barriers, tensor-map prefetch, TMEM allocation) overlapped the previous kernel of the
  // stream; from here on global memory is touched
  pdl_wait();
  if (p.scale != nullptr) {          // [2][Ntot]: scale, shift
    for (int i = threadIdx.x; i < 2 * p.Ntot; i += blockDim.x)
      colbuf[i] = (i < p.Ntot) ? p.scale[i] : p.shift[i - p.Ntot];
  } else if (p.stats != nullptr) {   // [4 lane quarters][2][BN]: sum, sum of squares of the current column block
    for (int i = threadIdx.x; i < 8 * BN; i += blockDim.x) colbuf[i] = 0.f;
    if (BNRED) {                     // + per-column BatchNorm constants {scale, shift, mean, invstd}
      float4* bnc = reinterpret_cast<float4*>(colbuf + 8 * BN);
      for (int i = threadIdx.x; i < p.Ntot; i += blockDim.x)
        bnc[i] = make_float4(p.bn_scale[i], p.bn_shift[i], p.bn_mean[i], p.bn_invstd[i]);
    }
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();   // barrier inits visible to the peer's TMA / arrives
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int regions = p.regions_x * p.regions_y * p.n_img;
  const int region_slots = (regions + CTAS - 1) / CTAS;      // one slot = one region per CTA of the pair
  const int total_tiles = region_slots * p.n_blocks;
  const int tile0 = blockIdx.x / CTAS, tile_step = gridDim.x / CTAS;
  const int chunks = p.c1_chunks + p.c2_chunks;
  const int pitch = kRegionW + 2 * p.halo;              // halo tile row pitch in pixels
  const uint32_t a_bytes = static_cast<uint32_t>(pitch * (kRegionH + 2 * p.halo) * 128);
  // A loads per tile: one per channel chunk (halo mode: serves all taps) or one per (tap, chunk)
  const int taps_per_load = p.halo ? p.num_taps : 1;
  const int a_loads = p.halo ? chunks : p.num_taps * chunks;

  if (warp == 0) {
    // ===================== TMA producer (each CTA feeds its own shared memory) =====================
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int pit = 0;
      for (int t = tile0; t < total_tiles; t += tile_step, ++pit) {
        // region-fastest order: CTAs running together share the weight tile (and the statistics slot)
        const int nb = t / region_slots;
        int m = (t - nb * region_slots) * CTAS + static_cast<int>(rank);
        const bool live = m < regions;                  // odd region count: the last pair has one idle CTA
        const int rx = m % p.regions_x;
        m /= p.regions_x;
        const int ry = m % p.regions_y;
        const int n0 = live ? m / p.regions_y : p.n_img;   // idle CTA: everything out of bounds -> zeros
        const int x0 = rx * kRegionW, y0 = ry * kRegionH;
        const int col0 = nb * BN + static_cast<int>(rank) * Cfg::kBNLocal;
        for (int al = 0; al < a_loads; ++al) {
          const int c = p.halo ? al : al % chunks;
          const int tap0 = p.halo ? 0 : al / chunks;
          mbar_wait(&a_empty[as], aph ^ 1);
          if (rank == 0) mbar_expect_tx(&a_full[as], a_bytes * CTAS);
          const CUtensorMap* mp = (c < p.c1_chunks) ? &tmA1 : &tmA2;
          const int cc = ((c < p.c1_chunks) ? c : c - p.c1_chunks) * 64;
          uint8_t* dstA = sA + as * kABufBytes;
          if (p.a_mode == 0) {
            if (CTAS == 2) tma_load_5d_pair(dstA, mp, &a_full[as], cc, x0 - p.halo, y0 - p.halo, n0, 0);
            else tma_load_5d(dstA, mp, &a_full[as], cc, x0 - p.halo, y0 - p.halo, n0, 0);
          } else {
            if (CTAS == 2) tma_load_5d_pair(dstA, mp, &a_full[as], cc, tap0 & 1, x0, tap0 >> 1, n0 * p.H + y0);
            else tma_load_5d(dstA, mp, &a_full[as], cc, tap0 & 1, x0, tap0 >> 1, n0 * p.H + y0);
          }
          if (++as == kNumABufs) {
            as = 0;
            aph ^= 1;
          }
          for (int tt = 0; tt < taps_per_load; ++tt) {
            const int tap = tap0 + tt;
            mbar_wait(&b_empty[bs], bph ^ 1);
            if (rank == 0) mbar_expect_tx(&b_full[bs], Cfg::kBBytes * CTAS);
            if (CTAS == 2) tma_load_2d_pair(sB + bs * Cfg::kBBytes, &tmB, &b_full[bs], (tap * chunks + c) * 64, col0);
            else tma_load_2d(sB + bs * Cfg::kBBytes, &tmB, &b_full[bs], (tap * chunks + c) * 64, col0);
            if (++bs == kBStages) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
        if constexpr (BNRED) {
          // the y tile of this OUTPUT tile, for its epilogue: issued behind the tile's operand loads, i.e. while
          // its MMAs run; the buffer is free once the epilogue that last read it is done
          const int yb = (kNumYBufs == 2) ? (pit & 1) : 0;
          const uint32_t yph = (kNumYBufs == 2) ? ((pit >> 1) & 1) : (pit & 1);
          mbar_wait(&yempty[yb], yph ^ 1);
          mbar_expect_tx(&yfull[yb], Cfg::kYBytes);
#pragma unroll
          for (int cc = 0; cc < BN / 64; ++cc)
            tma_load_5d(sY + yb * Cfg::kYBytes + cc * Cfg::kYChunkBytes, &tmY, &yfull[yb], nb * BN + cc * 64, x0, y0, n0, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only when CTAS == 2) =====================
    // The whole warp walks the loops (warp-uniform control flow keeps the descriptors in uniform
    // registers); one elected lane issues the tcgen05 instructions.
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128 * CTAS, BN, false, false);
      const uint32_t sbo = static_cast<uint32_t>(pitch * 128);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      for (int t = tile0; t < total_tiles; t += tile_step, ++it) {
        const int acc = it & 1;
        const uint32_t accph = (it >> 1) & 1;
        mbar_wait(&tempty[acc], accph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (SUBS * BN);
        uint32_t first = 1;
        for (int al = 0; al < a_loads; ++al) {
          mbar_wait(&a_full[as], aph);
          // descriptor of sub-tile 0, tap (0,0), k = 0; everything else is a start-address offset
          const uint64_t a_desc0 = umma_smem_desc_sw128(smem_u32(sA + as * kABufBytes), 0, sbo);
          for (int tt = 0; tt < taps_per_load; ++tt) {
            // tap (dy, dx) in {-1,0,1}^2 = rows (dy+1)*pitch + (dx+1) into the halo tile
            const uint32_t tap_off16 = p.halo ? static_cast<uint32_t>(((tt / 3) * pitch + (tt % 3)) * 8) : 0u;
            mbar_wait(&b_full[bs], bph);
            tc_fence_after();
            const uint64_t b_desc0 = umma_smem_desc_sw128(smem_u32(sB + bs * Cfg::kBBytes), 0, 1024);
            if (elect_one()) {
#pragma unroll
              for (int sub = 0; sub < SUBS; ++sub) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  // start-address field is (addr >> 4): offsets in 16-byte units
                  const uint64_t da = a_desc0 + (tap_off16 + sub * (kSubW * 8) + k * 2);
                  const uint64_t db = b_desc0 + (k * 2);
                  if (CTAS == 2) tc_mma_bf16_pair(d_tmem + sub * BN, da, db, idesc, (first && k == 0) ? 0u : 1u);
                  else tc_mma_bf16(d_tmem + sub * BN, da, db, idesc, (first && k == 0) ? 0u : 1u);
                }
              }
              const bool last_tap = tt == taps_per_load - 1;
              if (CTAS == 2) {
                tc_commit_pair(&b_empty[bs]);
                if (last_tap) tc_commit_pair(&a_empty[as]);
                if (last_tap && al == a_loads - 1) tc_commit_pair(&tfull[acc]);
              } else {
                tc_commit(&b_empty[bs]);
                if (last_tap) tc_commit(&a_empty[as]);
                if (last_tap && al == a_loads - 1) tc_commit(&tfull[acc]);
              }
            }
            __syncwarp();
            first = 0;
            if (++bs == kBStages) {
              bs = 0;
              bph ^= 1;
            }
          }
          if (++as == kNumABufs) {
            as = 0;
            aph ^= 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..) =====================
    // A warp may only touch the TMEM lane quarter (warp % 4); the warps of a quarter interleave the
    // 32-column chunks of the tile.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    float* wbuf = colbuf + q * 2 * BN;       // this lane quarter's [2][BN] statistics slot (chunks are disjoint per warp)
    int cur_nb = -1;
    // add this warp's partial column sums of column block `nbk` to the global statistics
    auto flush_stats = [&](int nbk) {
      for (int c = 2 * half; c < BN / 32; c += ((c & 1) ? 2 * (EPW / 4) - 1 : 1)) {      // the 32-column chunks of this warp's 64-column blocks
        const int i = c * 32 + lane;
        const int col = nbk * BN + i;
        if (col >= p.stats_from) {
          const float v = wbuf[i];
          if (v != 0.f) atomicAdd(p.stats + (col - p.stats_from), v);
          if (p.stats_sq != nullptr) {
            const float v2 = wbuf[BN + i];
            if (v2 != 0.f) atomicAdd(p.stats_sq + (col - p.stats_from), v2);
          }
        }
        wbuf[i] = 0.f;
        wbuf[BN + i] = 0.f;
      }
      __syncwarp();
    };
    const int sx = row & 7;   // pixel inside the 8 x 16 sub-tile
    const int sy = row >> 3;
    int obuf = 0;             // staging buffer of the next TMA store
    int it = 0;
    for (int t = tile0; t < total_tiles; t += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t accph = (it >> 1) & 1;
      const int nb = t / region_slots;
      int m = (t - nb * region_slots) * CTAS + static_cast<int>(rank);
      const bool live = m < regions;
      const int rx = m % p.regions_x;
      m /= p.regions_x;
      const int ry = m % p.regions_y;
      const int n = m / p.regions_y;
      const int y = ry * kRegionH + sy;
      if (p.stats != nullptr && nb != cur_nb) {
        if (cur_nb >= 0) flush_stats(cur_nb);
        cur_nb = nb;
      }
      const int col0 = nb * BN;
      int xs[SUBS];
      bool valids[SUBS], valids2[SUBS];      // inside the pixel grid / inside the second destination
#pragma unroll
      for (int sub = 0; sub < SUBS; ++sub) {
        xs[sub] = rx * kRegionW + sub * kSubW + sx;
        valids[sub] = live && (xs[sub] < p.W) && (y < p.H);
        valids2[sub] = valids[sub] && (xs[sub] < p.W2) && (y < p.H2);
      }
      mbar_wait(&tfull[acc], accph);
      tc_fence_after();
      const int yb = (kNumYBufs == 2) ? (it & 1) : 0;
      if constexpr (BNRED) mbar_wait(&yfull[yb], (kNumYBufs == 2) ? ((it >> 1) & 1) : (it & 1));
      if (live) {
#pragma unroll 1
        for (int c64 = half; c64 < BN / 64; c64 += EPW / 4) {
        // One 64-column block = one destination = (with TMA stores) one 128-byte-row staging tile per sub-tile:
        // the warp's 32 pixel rows x 64 channels, SWIZZLE_128B, written by conflict-free 16-byte shared stores and
        // sent by ONE bulk tensor store.  Per-thread global stores touch 32 different lines per instruction
        // (33 L1 data-pipe wavefronts each, the pipe the tensor core fetches its operands through: ncu,
        // profiles/ncu_igemm64_r02_summary.txt); accumulating destinations keep them (read-modify-write).
        // Optional (OSTG kernels, UNETK_TMA_STORE): see launch_igemm for why it is not the default.
        const bool second_dst = (BNRED || p.epi_mode != 0) ? false : (col0 + c64 * 64 >= p.split);
        const bool use_tma = OSTG && (second_dst ? (p.tma2 != 0) : (p.tma1 != 0));
        uint8_t* stg = sO + (warp - 2) * Cfg::kOWarpBytes + obuf * (SUBS * 4096);
        if (use_tma) {
          if (lane == 0) tma_store_wait_read<1>();       // the store that last read this buffer (two groups ago)
          __syncwarp();
        }
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          const int ch = c64 * 2 + hf;
          const int col = col0 + ch * 32;
          // BNRED: the producing layer's raw conv output, from the TMA-fetched tile (SWIZZLE_128B rows of 64 channels,
          // row = pixel (sy, sub*8 + sx) of the region, 16-byte column XOR (row & 7))
          uint4 yv[BNRED ? SUBS : 1][4];
          if constexpr (BNRED) {
#pragma unroll
            for (int sub = 0; sub < SUBS; ++sub) {
              const int ry_ = sy * kRegionW + sub * kSubW + sx;
              const uint8_t* src = sY + yb * Cfg::kYBytes + (ch >> 1) * Cfg::kYChunkBytes + ry_ * 128;
#pragma unroll
              for (int g = 0; g < 4; ++g)
                yv[sub][g] = *reinterpret_cast<const uint4*>(src + ((((ch & 1) * 4 + g) ^ (ry_ & 7)) << 4));
            }
          }
          float v2[SUBS][32];
#pragma unroll
          for (int sub = 0; sub < SUBS; ++sub)
            tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * (SUBS * BN) + sub * BN + ch * 32, v2[sub]);
          tmem_ld_wait();
          if (!BNRED && p.stats != nullptr && col >= p.stats_from) {
            // the sub-tiles share the columns: add them first, then ONE transposing butterfly each
            // for the sum and the sum of squares; per-quarter smem slots (no shared-memory atomics)
            float s[32];
            const bool second = col >= p.split;      // columns of the second destination: its own extent
            const bool ok0 = second ? valids2[0] : valids[0];
            const bool ok1 = second ? valids2[SUBS - 1] : valids[SUBS - 1];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float a = ok0 ? v2[0][j] : 0.f;
              if (SUBS == 2) a += ok1 ? v2[SUBS - 1][j] : 0.f;
              s[j] = a;
            }
            const float cs = warp_colsum32(s);
            wbuf[ch * 32 + lane] += cs;
            if (p.stats_sq != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float a = ok0 ? v2[0][j] : 0.f;
                float sq = a * a;
                if (SUBS == 2) {
                  const float b = ok1 ? v2[SUBS - 1][j] : 0.f;
                  sq = fmaf(b, b, sq);
                }
                s[j] = sq;
              }
              const float cs2 = warp_colsum32(s);
              wbuf[BN + ch * 32 + lane] += cs2;
            }
          }
          float gx[32];                  // BNRED: sum over the sub-tiles of g * xhat (dead otherwise)
          if (BNRED) {
#pragma unroll
            for (int j = 0; j < 32; ++j) gx[j] = 0.f;
          }
#pragma unroll
          for (int sub = 0; sub < SUBS; ++sub) {
            float (&v)[32] = v2[sub];
            if (!BNRED && p.scale != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float a = fmaf(v[j], colbuf[col + j], colbuf[p.Ntot + col + j]);
                v[j] = p.relu ? fmaxf(a, 0.f) : a;
              }
            }
            if (!valids[sub]) {
              if (BNRED) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
              }
              continue;
            }
            const int x = xs[sub];
            __nv_bfloat16* dst;
            bool accum;
            if (BNRED || p.epi_mode == 0) {
              const size_t pix = (static_cast<size_t>(n) * p.H + y) * p.W + x;
              if (BNRED || col < p.split) {
                dst = p.out1 + pix * p.ldc1 + col;
                accum = p.acc1 != 0;
              } else {
                if (!valids2[sub]) continue;
                const size_t pix2 = (static_cast<size_t>(n) * p.H2 + y) * p.W2 + x;
                dst = p.out2 + pix2 * p.ldc2 + (col - p.split);
                accum = p.acc2 != 0;
              }
            } else {
              const int tap = col / p.cout_t;
              const int co = col - tap * p.cout_t;
              const size_t opix = (static_cast<size_t>(n) * (2 * p.H) + (2 * y + (tap >> 1))) * (2 * p.W) +
                                  (2 * x + (tap & 1));
              dst = p.out1 + opix * p.cout_t + co;
              accum = false;
              if (p.bias != nullptr) {
                const float4* b4 = reinterpret_cast<const float4*>(p.bias + co);   // co is a multiple of 32
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 b = __ldg(b4 + j);
                  v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                }
              }
            }
            uint4* d4 = reinterpret_cast<uint4*>(dst);
            uint8_t* srow = stg + sub * 4096 + lane * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = v[g * 8 + j];
              if (accum && !use_tma) {
                const uint4 o = d4[g];
                const float2 o0 = unpack_bf16x2(o.x), o1 = unpack_bf16x2(o.y);
                const float2 o2 = unpack_bf16x2(o.z), o3 = unpack_bf16x2(o.w);
                f[0] += o0.x; f[1] += o0.y; f[2] += o1.x; f[3] += o1.y;
                f[4] += o2.x; f[5] += o2.y; f[6] += o3.x; f[7] += o3.y;
              }
              uint4 w;
              w.x = pack_bf16x2(f[0], f[1]);
              w.y = pack_bf16x2(f[2], f[3]);
              w.z = pack_bf16x2(f[4], f[5]);
              w.w = pack_bf16x2(f[6], f[7]);
              if (use_tma) *reinterpret_cast<uint4*>(srow + (((hf * 4 + g) ^ (lane & 7)) << 4)) = w;
              else d4[g] = w;
              if constexpr (BNRED) {
                // g = the STORED (accumulated, bf16-rounded) gradient masked by the ReLU of the producing
                // layer; keep g in v[] and g * xhat in gx[] for the column reduction below
                const uint4 yq = yv[sub][g];
                const float2 y0 = unpack_bf16x2(yq.x), y1 = unpack_bf16x2(yq.y), y2 = unpack_bf16x2(yq.z), y3 = unpack_bf16x2(yq.w);
                const float2 r0 = unpack_bf16x2(w.x), r1 = unpack_bf16x2(w.y), r2 = unpack_bf16x2(w.z), r3 = unpack_bf16x2(w.w);
                const float yy[8] = {y0.x, y0.y, y1.x, y1.y, y2.x, y2.y, y3.x, y3.y};
                const float rr[8] = {r0.x, r0.y, r1.x, r1.y, r2.x, r2.y, r3.x, r3.y};
                const float4* bnc = reinterpret_cast<const float4*>(colbuf + 8 * BN) + col + g * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 c = bnc[j];                  // {scale, shift, mean, invstd}
                  const float gg = (fmaf(yy[j], c.x, c.y) > 0.f) ? rr[j] : 0.f;
                  v[g * 8 + j] = gg;
                  gx[g * 8 + j] = fmaf(gg, (yy[j] - c.z) * c.w, gx[g * 8 + j]);
                }
              }
            }
          }
          if constexpr (BNRED) {
            if (SUBS == 2) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v2[0][j] += v2[SUBS - 1][j];
            }
            const float cs = warp_colsum32(v2[0]);
            wbuf[ch * 32 + lane] += cs;
            const float cs2 = warp_colsum32(gx);
            wbuf[BN + ch * 32 + lane] += cs2;
          }
        }
        if (use_tma) {
          fence_proxy_async_smem();          // the staging tile, written through the generic proxy, is read by the TMA
          __syncwarp();
          if (lane == 0) {
            const int colb = col0 + c64 * 64;
            const int yw = ry * kRegionH + q * 4;           // first of this warp's four pixel rows
#pragma unroll
            for (int sub = 0; sub < SUBS; ++sub) {
              const int xw = rx * kRegionW + sub * kSubW;
              if (p.epi_mode == 0) {
                if (second_dst) tma_store_5d(&tmO2, stg + sub * 4096, colb - p.split, xw, yw, n, 0);
                else tma_store_5d(&tmO1, stg + sub * 4096, colb, xw, yw, n, 0);
              } else {                                      // transposed conv: pixel shuffle (co, kx, x, ky, n*H + y)
                const int tap = colb / p.cout_t;
                tma_store_5d(&tmO1, stg + sub * 4096, colb - tap * p.cout_t, tap & 1, xw, tap >> 1, yw);
              }
            }
            tma_store_commit();
          }
          obuf ^= 1;
        }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTAS == 2) mbar_arrive_cluster(&tempty[acc], 0);   // the leader's MMA warp waits on its own barrier
        else mbar_arrive(&tempty[acc]);
        if (BNRED) mbar_arrive(&yempty[yb]);                   // this CTA's producer may refill the y tile
      }
    }
    if (p.stats != nullptr && cur_nb >= 0) flush_stats(cur_nb);
    if (OSTG && lane == 0) tma_store_wait<0>();      // shared memory must outlive the bulk stores that read it
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

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static int floor_pow2(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}

// Pick a (tw, th, tn) patch with tw*th*tn == `pixels` that minimises the padded pixel count
// (used by the weight-gradient kernel for its K tiles).
void pick_pixel_tile(int n_img, int H, int W, int pixels, int* tw, int* th, int* tn) {
  long best = -1;
  for (int w = 1; w <= pixels && w <= 16; w *= 2) {
    if (w > floor_pow2(W) && w > 1) break;
    int h = pixels / w;
    if (h > floor_pow2(H)) h = floor_pow2(H);
    const int n = pixels / (w * h);
    const long tiles = static_cast<long>((W + w - 1) / w) * ((H + h - 1) / h) * ((n_img + n - 1) / n);
    // prefer fewer tiles, then wider rows (longer contiguous runs in NHWC)
    if (best < 0 || tiles <= best) {
      best = tiles;
      *tw = w;
      *th = h;
      *tn = n;
    }
  }
}

// Activation tensor map, mode 0: NHWC viewed as (C, W, H, N, 1), box (64, tw, th, tn, 1).
int make_act_tmap(CUtensorMap* tm, const void* base, int N, int H, int W, int C, int tw, int th,
                  int tn) {
  const uint64_t dims[5] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W),
                            static_cast<uint64_t>(H), static_cast<uint64_t>(N), 1};
  const uint64_t strides[4] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(W) * C * 2,
                               static_cast<uint64_t>(H) * W * C * 2,
                               static_cast<uint64_t>(N) * H * W * C * 2};
  const uint32_t box[5] = {64, static_cast<uint32_t>(tw), static_cast<uint32_t>(th),
                           static_cast<uint32_t>(tn), 1};
  return encode_tmap_bf16(tm, base, 5, dims, strides, box, true);
}

// Activation tensor map, mode 1 (gather for the 2x2/s2 transposed conv): the fine tensor
// [N, 2H, 2W, C] viewed as (C, kx(2), X(W), ky(2), N*H), box (64, 1, tw, 1, th*tn).
int make_act_tmap_up2(CUtensorMap* tm, const void* base, int N, int H, int W, int C, int tw, int th,
                      int tn) {
  const uint64_t dims[5] = {static_cast<uint64_t>(C), 2, static_cast<uint64_t>(W), 2,
                            static_cast<uint64_t>(N) * H};
  const uint64_t strides[4] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(C) * 4,
                               static_cast<uint64_t>(W) * C * 4, static_cast<uint64_t>(W) * C * 8};
  const uint32_t box[5] = {64, 1, static_cast<uint32_t>(tw), 1, static_cast<uint32_t>(th * tn)};
  return encode_tmap_bf16(tm, base, 5, dims, strides, box, true);
}

static int make_b_tmap(CUtensorMap* tm, const void* base, int rows, int k, int box_rows) {
  const uint64_t dims[2] = {static_cast<uint64_t>(k), static_cast<uint64_t>(rows)};
  const uint64_t strides[1] = {static_cast<uint64_t>(k) * 2};
  const uint32_t box[2] = {64, static_cast<uint32_t>(box_rows)};
  return encode_tmap_bf16(tm, base, 2, dims, strides, box, true);
}

struct IgemmLaunch {
  const void* a1;
  int c1;
  const void* a2;
  int c2;
  const void* b;   // [Ntot][num_taps * (c1 + c2)] bf16
  IgemmParams p;
};

template <int BN, int SUBS, int CTAS, bool BNRED, int EPW = 4, bool OSTG = false>
static int launch_cfg(const CUtensorMap* tm, const IgemmParams& p, int grid, size_t colbytes, cudaStream_t stream) {
  using Cfg = IgemmCfg<BN, SUBS, CTAS, BNRED, EPW, OSTG>;
  const CUtensorMap &tmA1 = tm[0], &tmA2 = tm[1], &tmB = tm[2], &tmY = tm[3], &tmO1 = tm[4], &tmO2 = tm[5];
  const size_t smem = 1024 + Cfg::kSmemMain + (Cfg::kNumBars + 2) * 8 + colbytes;
  auto kern = igemm_kernel<BN, SUBS, CTAS, BNRED, EPW, OSTG>;
  UNETK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(igemm_threads(EPW));
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
  UNETK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA1, tmA2, tmB, tmY, tmO1, tmO2, p));
  return UNETK_OK;
}

static int launch_igemm(IgemmLaunch& L, cudaStream_t stream) {
  IgemmParams& p = L.p;
  UNETK_CHECK_ARG(L.c1 % 64 == 0 && L.c2 % 64 == 0 && L.c1 > 0, "igemm: channel counts must be multiples of 64 (got %d, %d)", L.c1, L.c2);
  UNETK_CHECK_ARG(p.Ntot % 64 == 0, "igemm: output columns must be a multiple of 64 (got %d)", p.Ntot);
  UNETK_CHECK_ARG(p.split % 32 == 0, "igemm: column split must be a multiple of 32");
  UNETK_CHECK_ARG((reinterpret_cast<uintptr_t>(L.a1) & 15) == 0 && (reinterpret_cast<uintptr_t>(L.b) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(p.out1) & 15) == 0, "igemm: pointers must be 16-byte aligned");
  if (int e = check_device_sm100()) return e;
  if (p.H2 <= 0 || p.W2 <= 0) { p.H2 = p.H; p.W2 = p.W; }
  UNETK_CHECK_ARG(p.H2 <= p.H && p.W2 <= p.W, "igemm: the second tensor (%dx%d) must fit inside the pixel grid (%dx%d)", p.H2, p.W2, p.H, p.W);
  p.halo = (p.num_taps == 9) ? 1 : 0;
  p.c1_chunks = L.c1 / 64;
  p.c2_chunks = L.c2 / 64;
  // configuration: CTA pairs with 256-wide column blocks where the layer is wide enough
  int bn, subs;
  const int ctas = 2;
  if (p.Ntot % 256 == 0) { bn = 256; subs = 1; }
  else { bn = (p.Ntot % 128 == 0) ? 128 : 64; subs = 2; }
  const int region_w = kSubW * subs;
  p.regions_x = (p.W + region_w - 1) / region_w;
  p.regions_y = (p.H + kRegionH - 1) / kRegionH;
  p.n_blocks = p.Ntot / bn;
  const bool use_colbuf = p.stats != nullptr || p.scale != nullptr;
  UNETK_CHECK_ARG(!(p.stats != nullptr && p.scale != nullptr), "igemm: stats and affine epilogues are exclusive");
  UNETK_CHECK_ARG(!use_colbuf || p.Ntot <= 1024, "igemm: stats/affine epilogue supports at most 1024 columns");

  const int box_w = region_w + 2 * p.halo, box_h = kRegionH + 2 * p.halo;
  CUtensorMap tmA1, tmA2, tmB;
  int e;
  if (p.a_mode == 0) {
    if ((e = make_act_tmap(&tmA1, L.a1, p.n_img, p.H, p.W, L.c1, box_w, box_h, 1))) return e;
    if (L.c2 > 0) {
      if ((e = make_act_tmap(&tmA2, L.a2, p.n_img, p.H2, p.W2, L.c2, box_w, box_h, 1))) return e;
    } else {
      tmA2 = tmA1;
    }
  } else {
    if ((e = make_act_tmap_up2(&tmA1, L.a1, p.n_img, p.H, p.W, L.c1, region_w, kRegionH, 1))) return e;
    tmA2 = tmA1;
  }
  const int K = p.num_taps * (L.c1 + L.c2);
  if ((e = make_b_tmap(&tmB, L.b, p.Ntot, K, bn / ctas))) return e;

  const long regions = static_cast<long>(p.regions_x) * p.regions_y * p.n_img;
  const long total_tiles = ((regions + ctas - 1) / ctas) * p.n_blocks;     // per CTA (pair)
  const int max_units = num_sms() / ctas;
  const int grid = static_cast<int>(total_tiles < max_units ? total_tiles : max_units) * ctas;
  const size_t colbytes = (p.stats != nullptr ? static_cast<size_t>(8) * bn + (p.bn_y != nullptr ? 4 * static_cast<size_t>(p.Ntot) : 0)
                                              : (use_colbuf ? static_cast<size_t>(2) * p.Ntot : 0)) * sizeof(float);
  const bool red = p.bn_y != nullptr;
  CUtensorMap tmY = tmA1;
  if (red) {      // y of the layer whose BatchNorm backward is fused: same pixel grid and columns as the output
    UNETK_CHECK_ARG((reinterpret_cast<uintptr_t>(p.bn_y) & 15) == 0, "igemm: bn_ctx.y must be 16-byte aligned");
    if ((e = make_act_tmap(&tmY, p.bn_y, p.n_img, p.H, p.W, p.Ntot, region_w, kRegionH, 1))) return e;
  }
  // TMA stores (see the epilogue): every non-accumulating destination made of whole 64-channel blocks; box =
  // one epilogue warp's 8 x 4 pixels x 64 channels
  // UNETK_TMA_STORE: 0 = never (default), 1 = every eligible destination, 2 = the transposed conv's pixel-shuffle
  // store only.  Measured on a B200 inside the training step (B=16 @256^2, two runs of 40 steps each): 10.15 / 10.20 ms
  // without, 10.17 / 10.17 ms with (1), 10.19 / 10.22 ms with (2) -- no gain: the staging tiles cost B-operand
  // stages (9 -> 7 for 256-wide tiles, 14 -> 6 for 128-wide ones) and the stores share the TMA unit with the operand
  // loads; under the board's power cap the saved L1 wavefronts do not show.  (Read per launch: tests toggle it.)
  const char* tma_env = getenv("UNETK_TMA_STORE");
  const int tma_mode = tma_env ? atoi(tma_env) : 0;
  CUtensorMap tmO1 = tmA1, tmO2 = tmA1;
  p.tma1 = p.tma2 = 0;
  if (tma_mode == 1 && !red && p.epi_mode == 0) {
    const bool two = p.split < p.Ntot;
    if (!p.acc1 && p.ldc1 % 64 == 0 && p.split % 64 == 0) {
      if ((e = make_act_tmap(&tmO1, p.out1, p.n_img, p.H, p.W, p.ldc1, kSubW, 4, 1))) return e;
      p.tma1 = 1;
    }
    if (two && !p.acc2 && p.ldc2 % 64 == 0 && p.split % 64 == 0 && (reinterpret_cast<uintptr_t>(p.out2) & 15) == 0) {
      if ((e = make_act_tmap(&tmO2, p.out2, p.n_img, p.H2, p.W2, p.ldc2, kSubW, 4, 1))) return e;
      p.tma2 = 1;
    }
  } else if (tma_mode >= 1 && !red && p.epi_mode == 1 && p.cout_t % 64 == 0) {
    if ((e = make_act_tmap_up2(&tmO1, p.out1, 1, p.H, p.W, p.cout_t, kSubW, 4, 1))) return e;
    p.tma1 = 1;
  }
  const CUtensorMap tm[6] = {tmA1, tmA2, tmB, tmY, tmO1, tmO2};
  const bool ostg = p.tma1 || p.tma2;
#define UNETK_IGEMM_LAUNCH(BN_, SUBS_, CTAS_)                                                             \
  (red ? launch_cfg<BN_, SUBS_, CTAS_, true>(tm, p, grid, colbytes, stream)                              \
       : ostg ? launch_cfg<BN_, SUBS_, CTAS_, false, 4, true>(tm, p, grid, colbytes, stream)             \
              : launch_cfg<BN_, SUBS_, CTAS_, false>(tm, p, grid, colbytes, stream))
  if (bn == 256 && p.epi_mode == 1 && !red)      // transposed-conv forward: epilogue bound, eight epilogue warps
    return ostg ? launch_cfg<256, 1, 2, false, 8, true>(tm, p, grid, colbytes, stream)
                : launch_cfg<256, 1, 2, false, 8>(tm, p, grid, colbytes, stream);
  if (bn == 256) return UNETK_IGEMM_LAUNCH(256, 1, 2);
  return bn == 128 ? UNETK_IGEMM_LAUNCH(128, 2, 2) : UNETK_IGEMM_LAUNCH(64, 2, 2);
#undef UNETK_IGEMM_LAUNCH
}

int conv3x3_fwd(const void* src1, int c1, const void* src2, int c2, const void* wpk, void* y,
                float* stats, const float* scale, const float* shift, int relu, int n, int h, int w,
                int cout, int h2, int w2, cudaStream_t stream) {
  UNETK_CHECK_ARG(src1 && wpk && y, "conv3x3_fwd: null pointer");
  UNETK_CHECK_ARG(c2 == 0 || src2 != nullptr, "conv3x3_fwd: second source missing");
  IgemmLaunch L{};
  L.a1 = src1; L.c1 = c1; L.a2 = src2; L.c2 = c2; L.b = wpk;
  IgemmParams& p = L.p;
  p.n_img = n; p.H = h; p.W = w; p.H2 = h2; p.W2 = w2; p.Ntot = cout; p.num_taps = 9; p.a_mode = 0; p.epi_mode = 0;
  p.split = cout; p.out1 = static_cast<__nv_bfloat16*>(y); p.out2 = nullptr; p.ldc1 = cout; p.ldc2 = 0;
  p.stats = stats; p.stats_sq = stats ? stats + cout : nullptr; p.stats_from = 0;
  p.scale = scale; p.shift = shift; p.relu = relu;
  return launch_igemm(L, stream);
}

static int apply_bn_ctx(IgemmParams& p, const BnReduceCtx* bn, const char* who) {
  if (bn == nullptr || bn->y == nullptr) return UNETK_OK;
  UNETK_CHECK_ARG(bn->scale && bn->shift && bn->mean && bn->invstd && bn->sums, "%s: incomplete BatchNorm context", who);
  UNETK_CHECK_ARG(p.stats == nullptr && p.split == p.Ntot, "%s: BatchNorm reduction needs a single destination and no bias sums", who);
  p.bn_y = static_cast<const __nv_bfloat16*>(bn->y);
  p.bn_scale = bn->scale; p.bn_shift = bn->shift; p.bn_mean = bn->mean; p.bn_invstd = bn->invstd;
  p.stats = bn->sums; p.stats_sq = bn->sums + p.Ntot; p.stats_from = 0;
  return UNETK_OK;
}

int conv3x3_dgrad(const void* dy, int cout, const void* wdpk, void* dx1, int c1, int acc1, void* dx2,
                  int c2, int acc2, float* stats, const BnReduceCtx* bn, int n, int h, int w, int h2, int w2,
                  cudaStream_t stream) {
  UNETK_CHECK_ARG(dy && wdpk && dx1, "conv3x3_dgrad: null pointer");
  UNETK_CHECK_ARG(c2 == 0 || dx2 != nullptr, "conv3x3_dgrad: second destination missing");
  UNETK_CHECK_ARG(stats == nullptr || c2 > 0, "conv3x3_dgrad: stats requested without a second destination");
  IgemmLaunch L{};
  L.a1 = dy; L.c1 = cout; L.a2 = nullptr; L.c2 = 0; L.b = wdpk;
  IgemmParams& p = L.p;
  p.n_img = n; p.H = h; p.W = w; p.H2 = h2; p.W2 = w2; p.Ntot = c1 + c2; p.num_taps = 9; p.a_mode = 0; p.epi_mode = 0;
  p.split = c1; p.out1 = static_cast<__nv_bfloat16*>(dx1); p.out2 = static_cast<__nv_bfloat16*>(dx2);
  p.ldc1 = c1; p.ldc2 = c2; p.acc1 = acc1; p.acc2 = acc2;
  // bias gradient of the ConvTranspose2d that produced the second concat source: sum over pixels of dx2
  p.stats = stats; p.stats_sq = nullptr; p.stats_from = c1;
  if (int e = apply_bn_ctx(p, bn, "conv3x3_dgrad")) return e;
  return launch_igemm(L, stream);
}

int convt2x2_fwd(const void* a, int cin, const void* wpk, const float* bias, void* up, int n, int h,
                 int w, int cout, cudaStream_t stream) {
  UNETK_CHECK_ARG(a && wpk && up, "convt2x2_fwd: null pointer");
  UNETK_CHECK_ARG(cout % 32 == 0, "convt2x2_fwd: cout must be a multiple of 32");
  IgemmLaunch L{};
  L.a1 = a; L.c1 = cin; L.b = wpk;
  IgemmParams& p = L.p;
  // a 2x2/s2 transposed conv has no spatial coupling: merge N and H into one row dimension
  p.n_img = 1; p.H = n * h; p.W = w; p.Ntot = 4 * cout; p.num_taps = 1; p.a_mode = 0; p.epi_mode = 1;
  p.split = p.Ntot; p.out1 = static_cast<__nv_bfloat16*>(up); p.bias = bias; p.cout_t = cout;
  return launch_igemm(L, stream);
}

int convt2x2_dgrad(const void* dup, int cout, const void* wdpk, void* da, int cin, int acc,
                   const BnReduceCtx* bn, int n, int h, int w, cudaStream_t stream) {
  UNETK_CHECK_ARG(dup && wdpk && da, "convt2x2_dgrad: null pointer");
  IgemmLaunch L{};
  L.a1 = dup; L.c1 = cout; L.b = wdpk;
  IgemmParams& p = L.p;
  p.n_img = 1; p.H = n * h; p.W = w; p.Ntot = cin; p.num_taps = 4; p.a_mode = 1; p.epi_mode = 0;
  p.split = cin; p.out1 = static_cast<__nv_bfloat16*>(da); p.ldc1 = cin; p.acc1 = acc;
  if (int e = apply_bn_ctx(p, bn, "convt2x2_dgrad")) return e;
  return launch_igemm(L, stream);
}

}  // namespace unetk
