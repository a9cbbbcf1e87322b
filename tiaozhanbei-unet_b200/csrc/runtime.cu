// Host runtime shared by all launchers: thread-local error message, device check (sm_100 only,
// no fallback), SM count cache and TMA tensor-map encoding through the driver entry point.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace unetk {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

static int g_dev_checked[64];  // 0 unknown, 1 ok, -1 not sm_100
static int g_dev_sms[64];

int check_device_sm100() {
  int dev = 0;
  UNETK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) {
    set_last_error("device ordinal %d out of range", dev);
    return UNETK_ERR_ARG;
  }
  if (g_dev_checked[dev] == 0) {
    int major = 0, sms = 0;
    UNETK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    UNETK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    g_dev_sms[dev] = sms;
    g_dev_checked[dev] = (major == 10) ? 1 : -1;
  }
  if (g_dev_checked[dev] != 1) {
    set_last_error("this library is built for sm_100a only; device %d is not compute capability 10.x (no fallback path)", dev);
    return UNETK_ERR_ARCH;
  }
  return UNETK_OK;
}

bool pdl_enabled() {
  static const bool on = [] {
    // Off by default: measured on a B200 (AnomalyUNet step, B=16 @256^2) it gains 1 % on a single stream
    // (11.37 -> 11.25 ms) and nothing once the engine's side streams are on (10.6 ms either way) -- the
    // front end already hides launch latency; the per-launch fixed cost is pipeline fill and drain.
    const char* e = getenv("UNETK_PDL");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

int num_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (g_dev_checked[dev] == 0) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) g_dev_sms[dev] = sms;
  }
  return g_dev_sms[dev] > 0 ? g_dev_sms[dev] : 148;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int resolve_encode() {
  if (g_encode != nullptr) return UNETK_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  UNETK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    set_last_error("cuTensorMapEncodeTiled is not available from the driver");
    return UNETK_ERR_CUDA;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return UNETK_OK;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  if (int e = resolve_encode()) return e;
  cuuint64_t gdims[5], gstrides[4];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstrides[i] = strides_bytes[i];
  const CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                              gdims, gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu %llu %llu, box %u %u %u %u %u)",
                   static_cast<int>(r), rank, (unsigned long long)gdims[0], (unsigned long long)(rank > 1 ? gdims[1] : 0),
                   (unsigned long long)(rank > 2 ? gdims[2] : 0), (unsigned long long)(rank > 3 ? gdims[3] : 0),
                   (unsigned long long)(rank > 4 ? gdims[4] : 0), gbox[0], rank > 1 ? gbox[1] : 0, rank > 2 ? gbox[2] : 0,
                   rank > 3 ? gbox[3] : 0, rank > 4 ? gbox[4] : 0);
    return UNETK_ERR_CUDA;
  }
  return UNETK_OK;
}

}  // namespace unetk
