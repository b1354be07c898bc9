// Host-side helpers of the C ABI: version / device queries and the TMA tensor-map encoder, which is
// resolved through cudaGetDriverEntryPoint so that the shared library has no link-time dependency on
// libcuda.so (it must load on a CPU-only box for the symbol-export test).
#include "common.cuh"
#include "../../include/omni_avsr.h"

omni_cuTensorMapEncodeTiled_t omni_get_tmap_encoder() {
  static omni_cuTensorMapEncodeTiled_t fn = nullptr;  // resolved once; the value is immutable afterwards
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess) return nullptr;
  if (qres != cudaDriverEntryPointSuccess || !p) return nullptr;
  fn = reinterpret_cast<omni_cuTensorMapEncodeTiled_t>(p);
  return fn;
}

int omni_make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                           uint32_t box_rows, uint32_t box_cols, int swizzle128) {
  omni_cuTensorMapEncodeTiled_t enc = omni_get_tmap_encoder();
  if (!enc) return OMNI_ERR_NO_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld % 8) != 0) return OMNI_ERR_BAD_ARG;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * sizeof(bf16)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? OMNI_OK : OMNI_ERR_CUDA;
}

extern "C" int omni_abi_version(void) { return OMNI_ABI_VERSION; }

extern "C" int omni_device_cc(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return OMNI_ERR_CUDA;
  int major = 0, minor = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return OMNI_ERR_CUDA;
  if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return OMNI_ERR_CUDA;
  return major * 10 + minor;
}
