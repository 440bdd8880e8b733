// common.cu — host plumbing shared by every solver in libtau_b200.so.
#include "common.cuh"
#include "../../include/tau_b200.h"

#include <stdarg.h>

static thread_local char g_err[1024] = "";

void tau_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *tau_last_error(void) { return g_err; }
extern "C" int tau_abi_version(void) { return TAU_B200_ABI_VERSION; }

extern "C" int tau_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();  // clear the sticky "no device" error
    return 0;
  }
  return n;
}

// cuTensorMapEncodeTiled is a driver-API symbol.  It is resolved through the runtime so that the
// shared library does not link libcuda (it must dlopen fine on the GPU-less build box).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int tau_make_tensor_map(CUtensorMap *out, const void *base, int elem_bytes, int rank,
                        const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box) {
  static PFN_encodeTiled encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      tau_set_error("cuTensorMapEncodeTiled not available from the driver (%s)",
                    cudaGetErrorString(e));
      return TAU_ERR_CUDA;
    }
    encode = (PFN_encodeTiled)fn;
  }
  TAU_REQUIRE(rank >= 1 && rank <= 5, "tensor map rank %d unsupported", rank);
  TAU_REQUIRE(elem_bytes == 4 || elem_bytes == 8 || elem_bytes == 1,
              "tensor map element size %d unsupported", elem_bytes);
  cuuint64_t gdim[5], gstride[5];
  cuuint32_t bdim[5], estride[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estride[i] = 1;
    if (i + 1 < rank) gstride[i] = strides_bytes[i];
  }
  CUtensorMapDataType dt = elem_bytes == 8   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64
                           : elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                             : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUresult r = encode(out, dt, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstride, bdim,
                      estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tau_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, box "
                  "%u x %u)",
                  (int)r, rank, (unsigned long long)dims[0],
                  (unsigned long long)(rank > 1 ? dims[1] : 1), box[0], rank > 1 ? box[1] : 1);
    return TAU_ERR_CUDA;
  }
  return TAU_OK;
}
