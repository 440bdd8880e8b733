/* Fake <cuda_runtime.h> for HOST EXECUTION of CUDA kernels — TEST INFRASTRUCTURE ONLY.
 * Lets g++ compile a reference .cu translation unit (cut before its main(), whose <<< >>> launches
 * are not C++) so that oracle/ref_drivers/ref_*_host.cpp can run the reference's OWN kernel bodies on
 * the CPU, one emulated thread at a time: the driver sets blockIdx / threadIdx and calls the kernel
 * as a plain function.  Works for kernels without __syncthreads() (every thread runs to completion
 * before the next starts); kernels that synchronise are not driven this way.
 * "Device memory" is host memory.  Contains no reference code. */
#ifndef TAU_HOSTCUDA_RUNTIME_H
#define TAU_HOSTCUDA_RUNTIME_H
#include <stdlib.h>
#include <string.h>
#define __global__
#define __device__
#define __host__
#define __shared__ /* `extern __shared__ T name[]` then needs a definition of `name` in the driver */
#define __forceinline__ inline
struct tau_hc_uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static tau_hc_uint3 blockIdx, threadIdx;
static dim3 blockDim, gridDim;
static inline void __syncthreads(void) { abort(); /* see header comment */ }
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline const char *cudaGetErrorString(cudaError_t) { return "hostcuda"; }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)calloc(1, n ? n : 1); return *p ? 0 : 2; }
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaPeekAtLastError(void) { return 0; }
static inline cudaError_t cudaGetLastError(void) { return 0; }
static inline cudaError_t cudaDeviceSynchronize(void) { return 0; }
/* launch emulation: TAU_HC_LAUNCH(gs, bs, kernel(args...)) */
#define TAU_HC_LAUNCH(gs, bs, call)                                                   \
  do {                                                                                \
    gridDim = (gs); blockDim = (bs);                                                  \
    for (blockIdx.z = 0; blockIdx.z < gridDim.z; ++blockIdx.z)                        \
      for (blockIdx.y = 0; blockIdx.y < gridDim.y; ++blockIdx.y)                      \
        for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)                    \
          for (threadIdx.z = 0; threadIdx.z < blockDim.z; ++threadIdx.z)              \
            for (threadIdx.y = 0; threadIdx.y < blockDim.y; ++threadIdx.y)            \
              for (threadIdx.x = 0; threadIdx.x < blockDim.x; ++threadIdx.x) call;    \
  } while (0)
#endif
