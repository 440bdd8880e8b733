// common.cuh — shared device/host plumbing for the sm_100a update-path kernels.
//
// Everything here is Blackwell-only (compiled with -gencode arch=compute_100a,code=sm_100a):
//   * mbarrier + TMA (cp.async.bulk.tensor) wrappers used to stage grid tiles in shared memory,
//   * 128-bit streaming global load/store helpers,
//   * warp-shuffle reductions,
//   * the host-side tensor-map encoder (driver entry point fetched at run time, so the library has
//     no link-time dependency on libcuda and still loads on a box without a GPU),
//   * the error convention of the C-ABI (int return code + thread-local message).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

// ----------------------------------------------------------------------------------------------
// Error convention: every extern "C" entry point returns 0 or a negative code and leaves a message
// retrievable through tau_last_error().  The CLI binaries wrap calls in an or-die macro that prints
// the message and exits, which is the reference's CK()/gpuAssert() policy
// (tau_hypersonic_cuda.cu:69-75, tau_gray_scott.cu:29-41).
// ----------------------------------------------------------------------------------------------
#define TAU_OK 0
#define TAU_ERR_INVALID -22   // -EINVAL
#define TAU_ERR_NOMEM -12     // -ENOMEM
#define TAU_ERR_CUDA -5       // -EIO
#define TAU_ERR_NODEV -19     // -ENODEV

void tau_set_error(const char *fmt, ...);

#define TAU_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      tau_set_error("CUDA error: %s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                    __LINE__);                                                             \
      return (_e == cudaErrorMemoryAllocation) ? TAU_ERR_NOMEM : TAU_ERR_CUDA;             \
    }                                                                                      \
  } while (0)

#define TAU_REQUIRE(cond, ...)    \
  do {                            \
    if (!(cond)) {                \
      tau_set_error(__VA_ARGS__); \
      return TAU_ERR_INVALID;     \
    }                             \
  } while (0)

// Host: encode a 2-D/3-D tiled tensor map over a row-major plane (innermost dimension first).
// elem_bytes is 4 or 8.  Out-of-bounds box elements are zero-filled by the hardware.
int tau_make_tensor_map(CUtensorMap *out, const void *base, int elem_bytes, int rank,
                        const uint64_t *dims, const uint64_t *strides_bytes /* rank-1 */,
                        const uint32_t *box);

#ifdef __CUDACC__

namespace tau {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  // make the init visible to the async (TMA) proxy
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- TMA tile loads (global -> shared, completion on an mbarrier) ------------------------------
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int x, int y,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *tmap, int x, int y,
                                            int z, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// order generic-proxy smem writes before later async-proxy (TMA) accesses of the same bytes
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- 128-bit streaming global access -----------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream_f4(const float *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_f4(float *p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_stream_d2(double *p, double a, double b) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b)
               : "memory");
}

// ---- warp-shuffle reductions -------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T w = __shfl_xor_sync(0xffffffffu, v, o);
    v = (w > v) ? w : v;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Non-negative floating-point values order like their bit patterns, so a max over them can use the
// integer atomicMax (exactly associative => the reduced value is independent of arrival order).
__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long *>(addr),
            static_cast<unsigned long long>(__double_as_longlong(v)));
}
__device__ __forceinline__ void atomic_max_nonneg(float *addr, float v) {
  atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

}  // namespace tau

#endif  // __CUDACC__
