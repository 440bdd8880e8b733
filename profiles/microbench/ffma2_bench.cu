// Microbenchmark: is the packed fp32x2 FMA (FFMA2, sm_100 `fma.rn.f32x2`) issued at the same rate as
// the scalar FFMA?  If so, pairing two independent fp32 streams halves their issue slots.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a ffma2_bench.cu -o ffma2_bench && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096, ILP = 8;

__global__ void k_scalar(float *out, float a, float b) {
  float x[2 * ILP];
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float *out, float a, float b) {
  float2 x[ILP];
  const float2 A = make_float2(a, a), B = make_float2(b, b);
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + 2 * i, threadIdx.x * 1e-3f + 2 * i + 1);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = __ffma2_rn(x[i], A, B);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: FP work + equal amount of integer work, to see whether freed issue slots are usable
__global__ void k_scalar_mixed(float *out, float a, float b, int m) {
  float x[2 * ILP];
  int y[2 * ILP];
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = threadIdx.x + i; }
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) { x[i] = fmaf(x[i], a, b); y[i] = (y[i] ^ m) + it; }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) s += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed_mixed(float *out, float a, float b, int m) {
  float2 x[ILP];
  int y[2 * ILP];
  const float2 A = make_float2(a, a), B = make_float2(b, b);
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + 2 * i, threadIdx.x * 1e-3f + 2 * i + 1);
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) y[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = __ffma2_rn(x[i], A, B);
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) y[i] = (y[i] ^ m) + it;
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) s += y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float timeit(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / 5;
}

int main() {
  const int blocks = 148 * 8, threads = 256;
  float *out; cudaMalloc(&out, blocks * threads * sizeof(float));
  const double fmas = (double)blocks * threads * ITER * 2 * ILP;
  float t1 = timeit([&] { k_scalar<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
  float t2 = timeit([&] { k_packed<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
  float t3 = timeit([&] { k_scalar_mixed<<<blocks, threads>>>(out, 1.0001f, 0.5f, 5); });
  float t4 = timeit([&] { k_packed_mixed<<<blocks, threads>>>(out, 1.0001f, 0.5f, 5); });
  printf("scalar FFMA : %.3f ms  %.1f TFMA/s\n", t1, fmas / t1 / 1e9);
  printf("packed FFMA2: %.3f ms  %.1f TFMA/s\n", t2, fmas / t2 / 1e9);
  printf("scalar FFMA + int mix : %.3f ms\n", t3);
  printf("packed FFMA2 + int mix: %.3f ms\n", t4);
  return 0;
}
