// Micro-experiment (SURVEY.md §7.2 hard part 1): may a SWIZZLE_128B UMMA shared-memory descriptor
// start at an address that is a multiple of 128 B but NOT of 1024 B (a pixel-row shift inside a
// TMA-written halo tile)?  If the hardware derives the swizzle XOR from absolute smem address bits,
// a conv tap shift is just a descriptor start offset and one halo load can feed all 9 taps.
//   build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I tiaozhanbei-unet_b200/csrc \
//           tools/umma_probe.cu tiaozhanbei-unet_b200/csrc/runtime.cu -o tools/umma_probe
//   run on a B200; prints one line per (mode, row shift, base_offset variant).
#include <stdlib.h>
#include <vector>

#include "common.cuh"

using namespace unetk;

constexpr int ROWS = 160;

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  return umma_smem_desc_sw128(saddr, lbo, sbo) | (static_cast<uint64_t>(base_off & 7) << 49);
}

// mode 0: K-major A (rows = M index), B K-major [64 n][64 k].  D[m][n] = sum_k A[r+m][k] B[n][k]
// mode 1: MN-major A: X [pix][128 ch] as two 64-ch boxes, B = Y [pix][64 ch] MN-major.
//         D[m][n] = sum_{p<64} X[p+r][m] * Y[p][n]
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int mode,
             int r, int use_bo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                       // mode 0: 160 x 128 B; mode 1: 2 boxes of 160 x 128 B
  uint8_t* sB = smem + 2 * ROWS * 128;      // 64 (or 160) rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + ROWS * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    if (mode == 0) {
      mbar_expect_tx(&bars[0], ROWS * 128 + 64 * 128);
      tma_load_2d(sA, &tmA, &bars[0], 0, 0);
      tma_load_2d(sB, &tmB, &bars[0], 0, 0);
    } else {
      mbar_expect_tx(&bars[0], 2 * ROWS * 128 + ROWS * 128);
      tma_load_2d(sA, &tmA, &bars[0], 0, 0);
      tma_load_2d(sA + ROWS * 128, &tmA, &bars[0], 64, 0);
      tma_load_2d(sB, &tmB, &bars[0], 0, 0);
    }
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA) + r * 128, b0 = smem_u32(sB);
    const uint32_t bo = use_bo ? (r & 7) : 0;
    if (mode == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 64, false, false);
      for (int k = 0; k < 4; ++k)
        tc_mma_bf16(tmem, desc_bo(a0 + k * 32, 0, 1024, bo), desc_bo(b0 + k * 32, 0, 1024, 0), idesc, k > 0);
    } else {
      const uint32_t idesc = umma_idesc_bf16(128, 64, true, true);
      for (int k = 0; k < 4; ++k)
        tc_mma_bf16(tmem, desc_bo(a0 + k * 2048, ROWS * 128, 1024, bo), desc_bo(b0 + k * 2048, 0, 1024, 0), idesc, k > 0);
    }
    tc_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int ch = 0; ch < 2; ++ch) {
    float v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + ch * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[row * 64 + ch * 32 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static float bf(uint16_t v) {
  uint32_t u = static_cast<uint32_t>(v) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static uint16_t tobf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return static_cast<uint16_t>(u >> 16);
}

int main() {
  std::vector<uint16_t> A(ROWS * 128), B(ROWS * 64);
  srand(1);
  for (auto& v : A) v = tobf(static_cast<float>(rand() % 7 - 3));
  for (auto& v : B) v = tobf(static_cast<float>(rand() % 5 - 2));
  uint16_t *dA, *dB;
  float* dOut;
  cudaMalloc(&dA, A.size() * 2);
  cudaMalloc(&dB, B.size() * 2);
  cudaMalloc(&dOut, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = 1024 + 3 * ROWS * 128 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  int all_ok[2][2] = {{1, 1}, {1, 1}};
  for (int mode = 0; mode < 2; ++mode) {
    CUtensorMap tmA, tmB;
    if (mode == 0) {
      // A viewed as [160 rows][64 k] (first 64 columns of the 128-wide array), B [64 n][64 k]
      uint64_t da[2] = {64, ROWS}, sa[1] = {128 * 2};
      uint32_t ba[2] = {64, ROWS};
      encode_tmap_bf16(&tmA, dA, 2, da, sa, ba, true);
      uint64_t db[2] = {64, 64}, sb[1] = {64 * 2};
      uint32_t bb[2] = {64, 64};
      encode_tmap_bf16(&tmB, dB, 2, db, sb, bb, true);
    } else {
      uint64_t da[2] = {128, ROWS}, sa[1] = {128 * 2};
      uint32_t ba[2] = {64, ROWS};
      encode_tmap_bf16(&tmA, dA, 2, da, sa, ba, true);
      uint64_t db[2] = {64, ROWS}, sb[1] = {64 * 2};
      uint32_t bb[2] = {64, ROWS};
      encode_tmap_bf16(&tmB, dB, 2, db, sb, bb, true);
    }
    for (int use_bo = 0; use_bo < 2; ++use_bo) {
      for (int r = 0; r <= 18; ++r) {
        cudaMemset(dOut, 0, 128 * 64 * 4);
        probe_kernel<<<1, 128, smem>>>(tmA, tmB, dOut, mode, r, use_bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mode %d r %d bo %d: CUDA error %s\n", mode, r, use_bo, cudaGetErrorString(e));
          return 1;
        }
        std::vector<float> out(128 * 64);
        cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            float ref = 0.f;
            if (mode == 0) {
              for (int k = 0; k < 64; ++k) ref += bf(A[(r + m) * 128 + k]) * bf(B[n * 64 + k]);
            } else {
              for (int p = 0; p < 64; ++p) ref += bf(A[(p + r) * 128 + m]) * bf(B[p * 64 + n]);
            }
            if (ref != out[m * 64 + n]) ++bad;
          }
        printf("mode %d (%s) row-shift %2d base_offset_field %d: %s (%d mismatches)\n", mode,
               mode == 0 ? "K-major A" : "MN-major A", r, use_bo ? (r & 7) : 0, bad ? "MISMATCH" : "exact", bad);
        if (bad) all_ok[mode][use_bo] = 0;
      }
    }
  }
  printf("SUMMARY K-major: bo=0 %s, bo=r&7 %s | MN-major: bo=0 %s, bo=r&7 %s\n", all_ok[0][0] ? "ALL EXACT" : "fails",
         all_ok[0][1] ? "ALL EXACT" : "fails", all_ok[1][0] ? "ALL EXACT" : "fails", all_ok[1][1] ? "ALL EXACT" : "fails");
  return 0;
}
