// hostemu.h — TEST INFRASTRUCTURE ONLY: runs the PRODUCT's simple CUDA translation units on the CPU so
// that their logic (tile/halo indexing, buffer rotation, clock slots, barrier placement) is checked on
// a GPU-less box.  tests/hostemu/hostemu_build.py rewrites `k<<<g, b, s, st>>>(args);` into
// tau_hc::launch(g, b, [&]{ k(args); }) and `#include "common.cuh"` into this header, then g++ builds
// build/hostemu/lib<name>_hostemu.so, which ONLY tests/test_hostemu_cpu.py loads.  Nothing in the
// package, bench.py or __graft_entry__ knows this exists: it is a checker for code, not a fallback.
//
// Execution model: blocks run one after another; the threads of a block are fibers on one OS thread,
// resumed round-robin.  __syncthreads() blocks a fiber until every live thread of the block arrived,
// warp collectives (__shfl_*_sync, __ballot_sync, __syncwarp; full mask only) until every live lane of
// the warp arrived, polling loops (__nanosleep, mbarrier waits) just yield.  A block whose live threads
// are all blocked aborts ("deadlock").  Fresh "device" memory is filled with garbage.  Arithmetic is the
// host's (libm, -ffp-contract=off), so a kernel whose expression trees match the CPU oracle's must
// reproduce it BIT FOR BIT.  TMA tile loads are executed synchronously with the hardware's zero fill of
// out-of-bounds elements; mbarriers keep the phase/arrival/tx-count protocol.  Inline PTX is rewritten
// statement by statement by hostemu_build.py (a statement it does not know is a build error).
// NOT modelled: memory ordering, inter-block concurrency, partial-mask collectives, -use_fast_math, speed.
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <new>
#include <string>
#include <type_traits>
#include <vector>
// (every std header first: the attribute-like macros below must not reach libstdc++)

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
// `__shared__` variables are function-local statics collected in one ELF section, so that the launcher can
// poison ALL of them (0xFF bytes: NaN floats, huge integers) before every block: shared memory is
// uninitialised on a GPU and does not survive from one block to the next.
#define __shared__ static __attribute__((section("tau_hc_smem")))
extern "C" char __start_tau_hc_smem[] __attribute__((weak)), __stop_tau_hc_smem[] __attribute__((weak));
#define __launch_bounds__(...)
#define __grid_constant__
#define __constant__   /* a plain global; cudaMemcpyToSymbol is a memcpy */
#define __align__(n) __attribute__((aligned(n)))
#ifndef TAU_HC_SMEM_BYTES
#define TAU_HC_SMEM_BYTES 232448   // `extern __shared__ T name[]` becomes a static array of this size
#endif

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
// CUDA's alignments: a misaligned vector access faults on the GPU; built with -fsanitize=alignment
// (hostemu_build.py sanitize=True) the emulated run reports it
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(4) uchar4 { unsigned char x, y, z, w; };
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return {x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return {x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }
using std::max;
using std::min;
static uint3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;

// ---- fibers: a 6-register x86-64 context switch (swapcontext costs a sigprocmask syscall per switch) ----
#if !defined(__x86_64__)
#error "tests/hostemu needs x86-64"
#endif
extern "C" void tau_hc_switch(void **save_sp, void *load_sp);
asm(".text\n.weak tau_hc_switch\n.type tau_hc_switch,@function\ntau_hc_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size tau_hc_switch,.-tau_hc_switch\n");

namespace tau_hc {
constexpr size_t STACK = 512 << 10;
enum Wait { RUN = 0, GEN, SPIN };
struct Fiber {
  void *sp;
  uint3 tid;
  bool done;
  Wait wait;
  const volatile unsigned *gen_ptr;  // GEN: runnable once *gen_ptr != gen_val
  unsigned gen_val;
};
static std::vector<Fiber> fibers;
static char *stacks = nullptr;
static void *sched_sp = nullptr;
static int cur = -1;
static unsigned nthreads = 0, nlive = 0;
static const std::function<void()> *body = nullptr;
static long long launches = 0;
// block-level barrier
static unsigned blk_arrived, blk_gen;
static int blk_or, blk_or_result[2];
// warp-level collectives (full-mask only)
static unsigned warp_arrived[32], warp_gen[32], warp_live[32], warp_par[32];
static uint64_t warp_buf[2][32][32];
static unsigned long long fake_timer = 0;

static void fiber_exit();
static void entry() {
  (*body)();
  fiber_exit();
}
static inline void to_sched() {
  tau_hc_switch(&fibers[cur].sp, sched_sp);
  threadIdx = fibers[cur].tid;
}
static void fiber_exit() {
  Fiber &f = fibers[cur];
  f.done = true;
  nlive--;
  const unsigned w = (unsigned)cur >> 5;
  warp_live[w]--;
  // threads that exit stop counting for barriers that others are already waiting at
  if (warp_live[w] && warp_arrived[w] == warp_live[w]) { warp_arrived[w] = 0; warp_gen[w]++; }
  if (nlive && blk_arrived == nlive) { blk_or_result[blk_gen & 1] = blk_or; blk_or = 0; blk_arrived = 0; blk_gen++; }
  tau_hc_switch(&f.sp, sched_sp);
  abort();
}
static inline void wait_gen(const volatile unsigned *g, unsigned v) {
  Fiber &f = fibers[cur];
  f.wait = GEN;
  f.gen_ptr = g;
  f.gen_val = v;
  to_sched();
}
static inline void spin_yield() {  // a polling loop's body: let everybody else run once
  fibers[cur].wait = SPIN;
  to_sched();
}
static inline void warp_barrier() {
  const unsigned w = (unsigned)cur >> 5, g = warp_gen[w];
  if (++warp_arrived[w] == warp_live[w]) { warp_arrived[w] = 0; warp_gen[w]++; }
  else wait_gen(&warp_gen[w], g);
}
static inline int block_barrier(int pred) {
  const unsigned g = blk_gen;
  blk_or |= (pred != 0);
  if (++blk_arrived == nlive) {
    blk_or_result[g & 1] = blk_or;   // read by the waiters of generation g; next written at g + 2
    blk_or = 0;
    blk_arrived = 0;
    blk_gen++;
  } else {
    wait_gen(&blk_gen, g);
  }
  return blk_or_result[g & 1];
}
template <class T> static inline T shfl(T v, unsigned src_lane) {
  static_assert(sizeof(T) <= 8, "shfl: type too wide");
  const unsigned w = (unsigned)cur >> 5, l = (unsigned)cur & 31u;
  const unsigned par = warp_par[w];   // flipped by the lane that completes the barrier
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  warp_buf[par][w][l] = bits;
  const unsigned g = warp_gen[w];
  if (++warp_arrived[w] == warp_live[w]) { warp_arrived[w] = 0; warp_par[w] ^= 1; warp_gen[w]++; }
  else wait_gen(&warp_gen[w], g);
  T r;
  memcpy(&r, &warp_buf[par][w][src_lane & 31u], sizeof(T));
  return r;
}
static inline void launch(dim3 g, dim3 b, const std::function<void()> &fn) {
  const unsigned nt = b.x * b.y * b.z;
  if (nt == 0 || nt > 1024) { fprintf(stderr, "hostemu: bad block size %u\n", nt); abort(); }
  if (!stacks) stacks = (char *)malloc(STACK * 1024);
  fibers.resize(nt);
  body = &fn;
  gridDim = g;
  blockDim = b;
  nthreads = nt;
  launches++;
  // block order: ascending by default; TAU_HC_BLOCK_ORDER=reverse | random (seeded per launch) — a GPU
  // promises no order, so results must not depend on it
  const size_t nblocks = (size_t)g.x * g.y * g.z;
  std::vector<size_t> order(nblocks);
  for (size_t i = 0; i < nblocks; ++i) order[i] = i;
  if (const char *bo = getenv("TAU_HC_BLOCK_ORDER")) {
    if (!strcmp(bo, "reverse")) std::reverse(order.begin(), order.end());
    else if (!strncmp(bo, "random", 6)) {
      uint64_t st = 0x9E3779B97F4A7C15ull * (uint64_t)(launches + 1);
      for (size_t i = nblocks; i > 1; --i) {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        std::swap(order[i - 1], order[st % i]);
      }
    }
  }
  for (size_t oi = 0; oi < nblocks; ++oi) {
      {
        const unsigned bx = (unsigned)(order[oi] % g.x), by = (unsigned)((order[oi] / g.x) % g.y),
                       bz = (unsigned)(order[oi] / ((size_t)g.x * g.y));
        blockIdx = {bx, by, bz};
        if (__start_tau_hc_smem) memset(__start_tau_hc_smem, 0xFF, (size_t)(__stop_tau_hc_smem - __start_tau_hc_smem));
        nlive = nt;
        blk_arrived = 0;
        blk_or = 0;
        for (unsigned w = 0; w < 32; ++w) {
          warp_arrived[w] = 0;
          warp_par[w] = 0;
          warp_live[w] = nt > w * 32 ? std::min(32u, nt - w * 32) : 0;
        }
        for (unsigned t = 0; t < nt; ++t) {
          Fiber &f = fibers[t];
          f.tid = {t % b.x, (t / b.x) % b.y, t / (b.x * b.y)};
          f.done = false;
          f.wait = RUN;
          uintptr_t top = ((uintptr_t)(stacks + STACK * (t + 1))) & ~(uintptr_t)15;
          void **sp = (void **)top;
          *--sp = nullptr;          // the "return address" entry() would return to (it never returns)
          *--sp = (void *)entry;    // popped by tau_hc_switch's ret
          for (int k = 0; k < 6; ++k) *--sp = nullptr;
          f.sp = sp;
        }
        unsigned long long idle_spins = 0;
        while (nlive) {
          unsigned ran = 0, ran_nonspin = 0;
          for (unsigned t = 0; t < nt; ++t) {
            Fiber &f = fibers[t];
            if (f.done) continue;
            if (f.wait == GEN && *f.gen_ptr == f.gen_val) continue;
            ran_nonspin += f.wait != SPIN;
            f.wait = RUN;
            cur = (int)t;
            threadIdx = f.tid;
            tau_hc_switch(&sched_sp, f.sp);
            ran++;
          }
          if (ran == 0) {
            fprintf(stderr, "hostemu: deadlock in block (%u,%u,%u): %u live threads all blocked at a barrier\n", bx,
                    by, bz, nlive);
            abort();
          }
          idle_spins = ran_nonspin ? 0 : idle_spins + 1;
          if (idle_spins > 2000000ull) {
            fprintf(stderr, "hostemu: livelock in block (%u,%u,%u): only polling loops are running\n", bx, by, bz);
            abort();
          }
        }
      }
  }
  cur = -1;
}
}  // namespace tau_hc

static inline void __syncthreads() { tau_hc::block_barrier(0); }
static inline int __syncthreads_or(int p) { return tau_hc::block_barrier(p); }
static inline void __syncwarp(unsigned = 0xffffffffu) { tau_hc::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) { tau_hc::spin_yield(); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return tau_hc::shfl(v, (unsigned)src); }
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
static inline unsigned __ballot_sync(unsigned, int p) {
  const unsigned me = (unsigned)tau_hc::cur & 31u, w = (unsigned)tau_hc::cur >> 5;
  const unsigned par = tau_hc::warp_par[w];
  tau_hc::shfl(p ? (1u << me) : 0u, me);   // every lane publishes its bit; returns after the warp arrived
  unsigned r = 0;
  for (unsigned l = 0; l < 32; ++l) r |= (unsigned)tau_hc::warp_buf[par][w][l];
  return r;
}
// __activemask() inside divergent code cannot be emulated (fibers have no notion of convergence): it
// returns 0, and a vote over a mask that is not the full warp answers "not unanimous" without
// synchronising.  That is exact for the one way the product uses it — an early-out that the general
// path it skips would decide identically (hypersonic3d.cu hllc: "whole warp supersonic").
static inline unsigned __activemask() { return 0u; }
static inline int __all_sync(unsigned m, int p) { return m == 0xffffffffu ? __ballot_sync(m, !p) == 0u : 0; }
static inline int __any_sync(unsigned m, int p) { return m == 0xffffffffu ? __ballot_sync(m, p) != 0u : (p != 0); }
static inline unsigned __match_any_sync(unsigned, unsigned long long v) {
  const unsigned me = (unsigned)tau_hc::cur & 31u, w = (unsigned)tau_hc::cur >> 5;
  const unsigned par = tau_hc::warp_par[w];
  tau_hc::shfl(v, me);
  unsigned r = 0;
  for (unsigned l = 0; l < tau_hc::warp_live[w] && l < 32; ++l) r |= (tau_hc::warp_buf[par][w][l] == v) << l;
  return r;
}
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
#ifdef TAU_HC_ROUGH_FASTMATH
// A rough stand-in for -use_fast_math (ex2.approx(x*log2e), lg2.approx(x)*ln2: the argument scaling is
// rounded to fp32, so the error grows with |x| like the GPU's) — ONLY to size test tolerances for code that
// has not run on hardware yet; nothing is compared bit for bit in this mode.
static inline float tau_hc_fast_expf(float x) { return exp2f(x * 1.4426950408889634f); }
static inline float tau_hc_fast_logf(float x) { return log2f(x) * 0.6931471805599453f; }
#define expf(x) tau_hc_fast_expf(x)
#define logf(x) tau_hc_fast_logf(x)
#endif
#define __expf(x) expf(x)   /* glibc declares __expf / __logf itself: macros, not functions */
#define __logf(x) logf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(long long u) { double d; memcpy(&d, &u, 8); return d; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return {a.x + b.x, a.y + b.y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return {a.x * b.x, a.y * b.y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return {fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
template <class T> static inline T atomicMax(T *a, T v) { T o = *a; if (v > o) *a = v; return o; }
template <class T> static inline T atomicMin(T *a, T v) { T o = *a; if (v < o) *a = v; return o; }
template <class T, class U> static inline T atomicAdd(T *a, U v) { T o = *a; *a = o + (T)v; return o; }
template <class T> static inline T atomicExch(T *a, T v) { T o = *a; *a = v; return o; }
template <class T, class U> static inline T atomicOr(T *a, U v) { T o = *a; *a = o | (T)v; return o; }
template <class T> static inline T atomicCAS(T *a, T c, T v) { T o = *a; if (o == c) *a = v; return o; }

// ---- the slice of the CUDA runtime API the host functions use; "device memory" is host memory ----
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorPeerAccessAlreadyEnabled = 704 };
typedef struct tau_hc_stream *cudaStream_t;
typedef struct tau_hc_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1 };
static inline const char *cudaGetErrorString(cudaError_t) { return "hostemu"; }
// "device" allocations: garbage-filled (reads before writes show up), 256-byte aligned like cudaMalloc's,
// with a guard zone on either side that cudaFree verifies (out-of-bounds WRITES of a kernel abort the run)
constexpr size_t TAU_HC_GUARD = 4096;
struct tau_hc_alloc { char *raw; size_t n; };
static std::vector<std::pair<void *, tau_hc_alloc>> tau_hc_allocs;
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) {
  char *raw = (char *)aligned_alloc(256, ((n + 2 * TAU_HC_GUARD + 255) / 256) * 256);
  if (!raw) { *p = nullptr; return cudaErrorMemoryAllocation; }
  memset(raw, 0xA5, TAU_HC_GUARD);
  memset(raw + TAU_HC_GUARD, 0xCB, n);
  memset(raw + TAU_HC_GUARD + n, 0xA5, TAU_HC_GUARD);
  *p = (T *)(raw + TAU_HC_GUARD);
  tau_hc_allocs.push_back({(void *)*p, {raw, n}});
  return cudaSuccess;
}
static inline cudaError_t cudaFree(void *p) {
  if (!p) return cudaSuccess;
  for (size_t i = 0; i < tau_hc_allocs.size(); ++i)
    if (tau_hc_allocs[i].first == p) {
      const tau_hc_alloc a = tau_hc_allocs[i].second;
      for (size_t k = 0; k < TAU_HC_GUARD; ++k)
        if ((unsigned char)a.raw[k] != 0xA5 || (unsigned char)a.raw[TAU_HC_GUARD + a.n + k] != 0xA5) {
          fprintf(stderr, "hostemu: out-of-bounds write next to a %zu-byte device allocation (%s it, offset %zu)\n",
                  a.n, (unsigned char)a.raw[k] != 0xA5 ? "before" : "after", k);
          abort();
        }
      free(a.raw);
      tau_hc_allocs.erase(tau_hc_allocs.begin() + i);
      return cudaSuccess;
    }
  fprintf(stderr, "hostemu: cudaFree of a pointer cudaMalloc did not return\n");
  abort();
}
enum { cudaHostAllocDefault = 0 };
template <class T> static inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned) { *p = (T *)malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
extern "C" __attribute__((weak)) long long tau_hostemu_live_allocations(void) { return (long long)tau_hc_allocs.size(); }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) {
  memmove(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <class T> static inline cudaError_t cudaMemcpyToSymbol(T &sym, const void *src, size_t n) {
  memcpy(&sym, src, n);
  return cudaSuccess;
}
struct cudaDeviceProp {
  char name[256];
  int major, minor, multiProcessorCount, maxThreadsPerBlock, warpSize;
  size_t sharedMemPerBlock, sharedMemPerBlockOptin, totalGlobalMem;
};
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline int tau_hc_env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e && *e ? atoi(e) : dflt;
}
// pretend devices of one process (tau_hyp2d_group): all share the host's memory, "peer access" is a no-op
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = tau_hc_env_int("TAU_HC_DEVICES", 1); return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaMemcpyPeerAsync(void *d, int, const void *s, int, size_t n, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
// a small pretend device keeps persistent grids small: TAU_HC_SMS "SMs" x TAU_HC_CTAS_PER_SM resident CTAs
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = tau_hc_env_int("TAU_HC_SMS", 3); return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
  memset(p, 0, sizeof(*p));
  strcpy(p->name, "hostemu");
  p->major = 10; p->minor = 0; p->multiProcessorCount = tau_hc_env_int("TAU_HC_SMS", 3);
  p->maxThreadsPerBlock = 1024; p->warpSize = 32;
  p->sharedMemPerBlock = 48 << 10; p->sharedMemPerBlockOptin = TAU_HC_SMEM_BYTES; p->totalGlobalMem = (size_t)1 << 34;
  return cudaSuccess;
}
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int v) {
  return v <= TAU_HC_SMEM_BYTES ? cudaSuccess : 1;
}
template <class K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t smem) {
  *n = smem <= TAU_HC_SMEM_BYTES ? tau_hc_env_int("TAU_HC_CTAS_PER_SM", 2) : 0;
  return cudaSuccess;
}
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 4 };
struct cudaLaunchAttribute {
  cudaLaunchAttributeID id;
  struct { int programmaticStreamSerializationAllowed; } val;
};
struct cudaLaunchConfig_t {
  dim3 gridDim, blockDim;
  size_t dynamicSmemBytes;
  cudaStream_t stream;
  cudaLaunchAttribute *attrs;
  unsigned numAttrs;
};
template <class... KA, class... A>
static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t *lc, void (*k)(KA...), A &&...a) {
  if (lc->dynamicSmemBytes > TAU_HC_SMEM_BYTES) return 1;
  tau_hc::launch(lc->gridDim, lc->blockDim, [&] { k(a...); });
  return cudaSuccess;
}
// CUDA IPC: every "device" lives in this process, so a handle is the pointer itself.  Several handles
// stepped in turn by one host thread emulate one-process-per-GPU peers (tests/hostemu/hyp2d_emu.py).
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
static long long tau_hc_ipc_open = 0;
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
  memset(h, 0, sizeof(*h));
  memcpy(h->reserved, &p, sizeof(p));
  return cudaSuccess;
}
static inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
  memcpy(p, h.reserved, sizeof(*p));
  tau_hc_ipc_open++;
  return *p ? cudaSuccess : 1;
}
static inline cudaError_t cudaIpcCloseMemHandle(void *) { tau_hc_ipc_open--; return cudaSuccess; }
extern "C" __attribute__((weak)) long long tau_hostemu_ipc_open_mappings(void) { return tau_hc_ipc_open; }

// ---- tensor maps, TMA tile loads, mbarriers (common.cuh:50-113, common.cu) ------------------------------
struct CUtensorMap {
  const char *base;
  int elem, rank;
  uint64_t dims[5], strides[5];   // strides[d] = bytes between consecutive indices of dimension d
  uint32_t box[5];
};
static inline int tau_make_tensor_map(CUtensorMap *out, const void *base, int elem_bytes, int rank, const uint64_t *dims,
                                      const uint64_t *strides_bytes, const uint32_t *box) {
  memset(out, 0, sizeof(*out));
  out->base = (const char *)base;
  out->elem = elem_bytes;
  out->rank = rank;
  for (int d = 0; d < rank; ++d) {
    out->dims[d] = dims[d];
    out->box[d] = box[d];
    out->strides[d] = d == 0 ? (uint64_t)elem_bytes : strides_bytes[d - 1];
    // the hardware's rules: 16-byte aligned base and strides, box <= 256 per dimension, inner box bytes % 16 == 0
    if (box[d] == 0 || box[d] > 256 || (d > 0 && out->strides[d] % 16)) return -22;
  }
  if (((uintptr_t)base & 15) || ((uint64_t)box[0] * elem_bytes) % 16) return -22;
  return 0;
}
namespace tau {
struct MBar { uint32_t phase : 1, count : 15, pending : 16; int32_t tx; };
static_assert(sizeof(MBar) == 8, "mbarrier emulation state must fit the 8-byte barrier word");
static inline void mbar_check(MBar *b) {
  if (b->pending == 0 && b->tx == 0) { b->phase ^= 1; b->pending = b->count; }
}
static inline void mbar_init(uint64_t *bar, uint32_t count) {
  MBar *b = (MBar *)bar;
  b->phase = 0; b->count = count; b->pending = count; b->tx = 0;
}
static inline void mbar_fence_init() {}
static inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {   // arrive.expect_tx
  MBar *b = (MBar *)bar;
  if (b->pending == 0) { fprintf(stderr, "hostemu: mbarrier over-arrival\n"); abort(); }
  b->tx += (int32_t)bytes;
  b->pending--;
  mbar_check(b);
}
static inline void mbar_wait(uint64_t *bar, uint32_t parity) {       // try_wait.parity loop
  while (((volatile MBar *)bar)->phase == (parity & 1u)) tau_hc::spin_yield();
}
static inline void tma_copy(void *dst, const CUtensorMap *m, const int *c, uint64_t *bar) {
  if ((uintptr_t)dst & 127) { fprintf(stderr, "hostemu: TMA destination not 128-byte aligned\n"); abort(); }
  if (((int64_t)c[0] * m->elem) % 16) { fprintf(stderr, "hostemu: TMA box origin not 16-byte aligned (x = %d)\n", c[0]); abort(); }
  uint32_t bx[5] = {1, 1, 1, 1, 1};
  for (int d = 0; d < m->rank; ++d) bx[d] = m->box[d];
  char *o = (char *)dst;
  size_t bytes = 0;
  for (uint32_t k = 0; k < bx[2]; ++k)
    for (uint32_t j = 0; j < bx[1]; ++j)
      for (uint32_t i = 0; i < bx[0]; ++i, o += m->elem, bytes += m->elem) {
        const int64_t g[3] = {(int64_t)c[0] + i, (int64_t)c[1] + j, (int64_t)c[2] + k};
        bool in = true;
        for (int d = 0; d < m->rank; ++d) in = in && g[d] >= 0 && (uint64_t)g[d] < m->dims[d];
        if (!in) { memset(o, 0, m->elem); continue; }   // out-of-bounds box elements are zero-filled
        const char *src = m->base;
        for (int d = 0; d < m->rank; ++d) src += (uint64_t)g[d] * m->strides[d];
        memcpy(o, src, m->elem);
      }
  MBar *b = (MBar *)bar;
  b->tx -= (int32_t)bytes;
  mbar_check(b);
}
static inline void tma_load_2d(void *dst, const CUtensorMap *m, int x, int y, uint64_t *bar) {
  const int c[3] = {x, y, 0};
  tma_copy(dst, m, c, bar);
}
static inline void tma_load_3d(void *dst, const CUtensorMap *m, int x, int y, int z, uint64_t *bar) {
  const int c[3] = {x, y, z};
  tma_copy(dst, m, c, bar);
}
static inline void tma_prefetch_desc(const CUtensorMap *) {}
static inline void fence_proxy_async() {}
static inline float4 ldg_stream_f4(const float *p) { return {p[0], p[1], p[2], p[3]}; }
static inline void stg_stream_f4(float *p, float4 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; }
static inline void stg_stream_d2(double *p, double a, double b) { p[0] = a; p[1] = b; }
}  // namespace tau

// ---- the slice of common.cuh / common.cu these translation units use ----------------------------
#define TAU_OK 0
#define TAU_ERR_INVALID -22
#define TAU_ERR_NOMEM -12
#define TAU_ERR_CUDA -5
#define TAU_ERR_NODEV -19
// (weak: several emulated translation units can be linked into one library — hostemu_build.build_all)
#define TAU_HC_WEAK __attribute__((weak))
TAU_HC_WEAK char tau_hc_err[512];
TAU_HC_WEAK void tau_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tau_hc_err, sizeof(tau_hc_err), fmt, ap);
  va_end(ap);
}
extern "C" TAU_HC_WEAK const char *tau_hostemu_last_error(void) { return tau_hc_err; }
extern "C" TAU_HC_WEAK const char *tau_last_error(void) { return tau_hc_err; }
extern "C" TAU_HC_WEAK int tau_abi_version(void) { return 1; }
extern "C" TAU_HC_WEAK long long tau_hostemu_launches(void) { return tau_hc::launches; }
extern "C" TAU_HC_WEAK int tau_device_count(void) { return tau_hc_env_int("TAU_HC_DEVICES", 1); }
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
