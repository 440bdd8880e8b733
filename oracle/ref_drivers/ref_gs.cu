// ref_gs.cu — TEST INFRASTRUCTURE ONLY (never linked into the product).
// Compiles the UNMODIFIED reference translation unit tau_gray_scott.cu (from /root/reference, via
// the fake curses header in oracle/shims) and exposes its step_kernel / init_pattern behind a
// C entry point so the GPU parity tests can run the reference's own kernel side by side with the
// product on the same device.  The host loop restates tau_gray_scott.cu:315-329.
#define main ref_gs_main
#include "tau_gray_scott.cu"
#undef main

extern "C" void ref_gs_init_pattern(float *u, float *v, int nx, int ny, unsigned seed) {
  std::vector<float> hu((size_t)nx * ny), hv((size_t)nx * ny);
  init_pattern(hu, hv, nx, ny, seed);
  memcpy(u, hu.data(), hu.size() * sizeof(float));
  memcpy(v, hv.data(), hv.size() * sizeof(float));
}

// u, v: host planes in/out.  Returns 0 or a cudaError_t.
extern "C" int ref_gs_run(float *u, float *v, int nx, int ny, float Du, float Dv, float dt,
                          float dx, float feed, float kill, int steps) {
  size_t N = (size_t)nx * ny;
  float *d_u0, *d_u1, *d_v0, *d_v1;
  cudaError_t e;
  if ((e = cudaMalloc(&d_u0, N * 4))) return e;
  if ((e = cudaMalloc(&d_u1, N * 4))) return e;
  if ((e = cudaMalloc(&d_v0, N * 4))) return e;
  if ((e = cudaMalloc(&d_v1, N * 4))) return e;
  cudaMemcpy(d_u0, u, N * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_v0, v, N * 4, cudaMemcpyHostToDevice);
  dim3 block(16, 16);
  dim3 grid((nx + block.x - 1) / block.x, (ny + block.y - 1) / block.y);
  for (int s = 0; s < steps; ++s) {
    step_kernel<<<grid, block>>>(d_u1, d_v1, d_u0, d_v0, nx, ny, Du, Dv, dt, dx, feed, kill);
    std::swap(d_u0, d_u1);
    std::swap(d_v0, d_v1);
  }
  e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaMemcpy(u, d_u0, N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(v, d_v0, N * 4, cudaMemcpyDeviceToHost);
  cudaFree(d_u0); cudaFree(d_u1); cudaFree(d_v0); cudaFree(d_v1);
  return (int)e;
}

// device time of `steps` reference steps on resident planes (ms), for side-by-side reporting
extern "C" int ref_gs_time(int nx, int ny, int steps, float *ms_out) {
  size_t N = (size_t)nx * ny;
  std::vector<float> hu(N), hv(N);
  init_pattern(hu, hv, nx, ny, 1337u);
  float *d_u0, *d_u1, *d_v0, *d_v1;
  cudaMalloc(&d_u0, N * 4); cudaMalloc(&d_u1, N * 4); cudaMalloc(&d_v0, N * 4); cudaMalloc(&d_v1, N * 4);
  cudaMemcpy(d_u0, hu.data(), N * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_v0, hv.data(), N * 4, cudaMemcpyHostToDevice);
  dim3 block(16, 16);
  dim3 grid((nx + block.x - 1) / block.x, (ny + block.y - 1) / block.y);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int s = 0; s < 3; ++s) {
    step_kernel<<<grid, block>>>(d_u1, d_v1, d_u0, d_v0, nx, ny, 0.2f, 0.1f, 1.f, 1.f, 0.03f, 0.06f);
    std::swap(d_u0, d_u1); std::swap(d_v0, d_v1);
  }
  cudaEventRecord(a);
  for (int s = 0; s < steps; ++s) {
    step_kernel<<<grid, block>>>(d_u1, d_v1, d_u0, d_v0, nx, ny, 0.2f, 0.1f, 1.f, 1.f, 0.03f, 0.06f);
    std::swap(d_u0, d_u1); std::swap(d_v0, d_v1);
  }
  cudaEventRecord(b);
  cudaError_t e = cudaEventSynchronize(b);
  cudaEventElapsedTime(ms_out, a, b);
  cudaFree(d_u0); cudaFree(d_u1); cudaFree(d_v0); cudaFree(d_v1);
  return (int)e;
}
