// Host-side plumbing shared by all kernels: error text, TMA descriptor encoding, device info.
#include "common.cuh"
#include "../../include/ggpt_b200.h"

#include <stdarg.h>
#include <string.h>

namespace ggpt {

static thread_local char g_err[1024] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
    set_error("cuTensorMapEncodeTiled not available from the driver: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -3;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d rows=%llu cols=%llu ld=%llu box=%ux%u base=%p) failed: CUresult %d",
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols, base,
              (int)r);
    return -3;
  }
  return 0;
}

int make_tmap_3d_bf16(CUtensorMap* map, const void* base, uint64_t d2, uint64_t rows, uint64_t cols, uint64_t ld2,
                      uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -3;
  cuuint64_t gdim[3] = {cols, rows, d2};
  cuuint64_t gstr[2] = {ld * 2, ld2 * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d %llu x %llu x %llu) failed: CUresult %d", (unsigned long long)d2,
              (unsigned long long)rows, (unsigned long long)cols, (int)r);
    return -3;
  }
  return 0;
}

int make_tmap_3d_f32(CUtensorMap* map, const void* base, uint64_t d2, uint64_t rows, uint64_t cols, uint64_t ld2,
                     uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  const CUtensorMapSwizzle sw = (box_cols * 4 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B
                              : (box_cols * 4 == 64) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
  EncodeTiledFn enc = get_encode();
  if (!enc) return -3;
  cuuint64_t gdim[3] = {cols, rows, d2};
  cuuint64_t gstr[2] = {ld * 4, ld2 * 4};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d f32 %llu x %llu x %llu) failed: CUresult %d", (unsigned long long)d2,
              (unsigned long long)rows, (unsigned long long)cols, (int)r);
    return -3;
  }
  return 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace ggpt

extern "C" {

const char* ggpt_last_error(void) { return ggpt::g_err; }

int ggpt_abi_version(void) { return GGPT_ABI_VERSION; }

int ggpt_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    ggpt::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return -2;
  }
  cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
  return 0;
}

}  // extern "C"
