// hostemu.h — TEST INFRASTRUCTURE ONLY: runs the PRODUCT's simple CUDA translation units on the CPU so
// that their logic (tile/halo indexing, buffer rotation, clock slots, barrier placement) is checked on
// a GPU-less box.  tests/hostemu/build.py rewrites `k<<<g, b, s, st>>>(args);` into
// tau_hc::launch(g, b, [&]{ k(args); }) and `#include "common.cuh"` into this header, then g++ builds
// build/hostemu/lib<name>_hostemu.so, which ONLY tests/test_hostemu_cpu.py loads.  Nothing in the
// package, bench.py or __graft_entry__ knows this exists: it is a checker for code, not a fallback.
//
// Execution model: blocks run one after another; the threads of a block are ucontext fibers on one
// OS thread.  __syncthreads() and the warp shuffles yield to a round-robin scheduler, which aborts if
// the threads of a block do not all reach the same number of synchronisation points (barrier
// divergence).  Arithmetic is the host's (libm, -ffp-contract=off), so a kernel whose expression
// trees match the CPU oracle's must reproduce it BIT FOR BIT.  Only what burgers.cu / shallow_water.cu
// use is implemented (no TMA, no mbarrier, no inline PTX).
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __grid_constant__

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static uint3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;

namespace tau_hc {
constexpr size_t STACK = 256 << 10;
struct Fiber {
  ucontext_t ctx;
  uint3 tid;
  bool done;
  int nshfl;
};
static std::vector<Fiber> fibers;
static char *stacks = nullptr;
static ucontext_t sched;
static int cur = -1;
static const std::function<void()> *body = nullptr;
static uint64_t shfl_buf[2][1024];
static long long launches = 0;

static void entry() {
  (*body)();
  fibers[cur].done = true;
  swapcontext(&fibers[cur].ctx, &sched);
}
static inline void yield() {  // every synchronisation point of a block
  swapcontext(&fibers[cur].ctx, &sched);
}
static inline void launch(dim3 g, dim3 b, const std::function<void()> &fn) {
  const unsigned nt = b.x * b.y * b.z;
  if (nt == 0 || nt > 1024) { fprintf(stderr, "hostemu: bad block size %u\n", nt); abort(); }
  if (!stacks) stacks = (char *)malloc(STACK * 1024);
  fibers.resize(nt);
  body = &fn;
  gridDim = g;
  blockDim = b;
  launches++;
  for (unsigned bz = 0; bz < g.z; ++bz)
    for (unsigned by = 0; by < g.y; ++by)
      for (unsigned bx = 0; bx < g.x; ++bx) {
        blockIdx = {bx, by, bz};
        for (unsigned t = 0; t < nt; ++t) {
          Fiber &f = fibers[t];
          f.tid = {t % b.x, (t / b.x) % b.y, t / (b.x * b.y)};
          f.done = false;
          f.nshfl = 0;
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = stacks + STACK * t;
          f.ctx.uc_stack.ss_size = STACK;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, entry, 0);
        }
        for (;;) {  // one round = every live thread runs to its next synchronisation point
          unsigned ndone = 0;
          for (unsigned t = 0; t < nt; ++t) {
            cur = (int)t;
            threadIdx = fibers[t].tid;
            swapcontext(&sched, &fibers[t].ctx);
            ndone += fibers[t].done;
          }
          if (ndone == nt) break;
          if (ndone != 0) {
            fprintf(stderr, "hostemu: barrier divergence in block (%u,%u,%u): %u of %u threads exited\n", bx, by,
                    bz, ndone, nt);
            abort();
          }
        }
      }
  cur = -1;
}
template <class T> static inline T shfl(T v, unsigned src_lane_of_me) {
  Fiber &f = fibers[cur];
  const int par = f.nshfl++ & 1;
  const unsigned me = (unsigned)cur;
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  shfl_buf[par][me] = bits;
  yield();
  threadIdx = fibers[cur].tid;
  const unsigned src = (me & ~31u) | (src_lane_of_me & 31u);
  T r;
  memcpy(&r, &shfl_buf[par][src], sizeof(T));
  return r;
}
}  // namespace tau_hc

static inline void __syncthreads() { tau_hc::yield(); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
  return tau_hc::shfl(v, ((unsigned)tau_hc::cur & 31u) ^ (unsigned)m);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
  const unsigned l = (unsigned)tau_hc::cur & 31u;
  return tau_hc::shfl(v, l + d < 32 ? l + d : l);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
  const unsigned l = (unsigned)tau_hc::cur & 31u;
  return tau_hc::shfl(v, l >= d ? l - d : l);
}
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
template <class T> static inline T atomicMax(T *a, T v) { T o = *a; if (v > o) *a = v; return o; }
template <class T> static inline T atomicAdd(T *a, T v) { T o = *a; *a = o + v; return o; }

// ---- the slice of the CUDA runtime API the host functions use; "device memory" is host memory ----
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef struct tau_hc_stream *cudaStream_t;
typedef struct tau_hc_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1 };
static inline const char *cudaGetErrorString(cudaError_t) { return "hostemu"; }
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) {
  *p = (T *)malloc(n ? n : 1);
  if (*p) memset(*p, 0xCB, n);   // garbage, like fresh device memory: reads before writes show up
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) {
  memmove(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

// ---- the slice of common.cuh / common.cu these translation units use ----------------------------
#define TAU_OK 0
#define TAU_ERR_INVALID -22
#define TAU_ERR_NOMEM -12
#define TAU_ERR_CUDA -5
#define TAU_ERR_NODEV -19
static char tau_hc_err[512];
static inline void tau_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tau_hc_err, sizeof(tau_hc_err), fmt, ap);
  va_end(ap);
}
extern "C" const char *tau_hostemu_last_error(void) { return tau_hc_err; }
extern "C" long long tau_hostemu_launches(void) { return tau_hc::launches; }
extern "C" int tau_device_count(void) { return 1; }
#define TAU_CUDA(expr)                                                   \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) {                                             \
      tau_set_error("CUDA error: %s", #expr);                            \
      return (_e == cudaErrorMemoryAllocation) ? TAU_ERR_NOMEM : TAU_ERR_CUDA; \
    }                                                                    \
  } while (0)
#define TAU_REQUIRE(cond, ...)    \
  do {                            \
    if (!(cond)) {                \
      tau_set_error(__VA_ARGS__); \
      return TAU_ERR_INVALID;     \
    }                             \
  } while (0)
namespace tau {
template <typename T> static inline T warp_max(T v) {   // common.cuh:135-143
  for (int o = 16; o > 0; o >>= 1) {
    T w = __shfl_xor_sync(0xffffffffu, v, o);
    v = (w > v) ? w : v;
  }
  return v;
}
template <typename T> static inline T warp_sum(T v) {   // common.cuh:144-149
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
static inline void atomic_max_nonneg(double *addr, double v) {   // common.cuh:153-156
  atomicMax(reinterpret_cast<unsigned long long *>(addr), static_cast<unsigned long long>(__double_as_longlong(v)));
}
static inline void atomic_max_nonneg(float *addr, float v) {     // common.cuh:157-159
  atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
}  // namespace tau
