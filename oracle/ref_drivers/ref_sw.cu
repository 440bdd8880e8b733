// ref_sw.cu — TEST INFRASTRUCTURE ONLY (never linked into the product).
// Compiles the UNMODIFIED reference translation unit tau_shallow_water.cu (from /root/reference, via
// the fake curses header in oracle/shims) and drives its own kernels with main()'s launch shapes; the
// host loop restates `do_step` + the clock update (tau_shallow_water.cu:669-705, :767-768).
#define main ref_sw_main
#include "tau_shallow_water.cu"
#undef main

// pf: nx ny dx dy g f0 nu H0 bumpAmp bumpSigma CFL offx offy asym swirl swirlRc tau0 t0 dtau
static Params params_from(const float *pf) {
  Params P;
  P.nx = (int)pf[0]; P.ny = (int)pf[1]; P.dx = pf[2]; P.dy = pf[3]; P.g = pf[4]; P.f0 = pf[5]; P.nu = pf[6];
  P.H0 = pf[7]; P.bumpAmp = pf[8]; P.bumpSigma = pf[9]; P.CFL = pf[10]; P.offx = pf[11]; P.offy = pf[12];
  P.asym = pf[13]; P.swirl = pf[14]; P.swirlRc = pf[15]; P.tau0 = pf[16]; P.t0 = pf[17]; P.dtau = pf[18];
  return P;
}

extern "C" void ref_sw_init(const float *pf, float *sigma, float *u, float *v) {
  Params P = params_from(pf);
  HostState H;
  initialize_host(P, H);
  memcpy(sigma, H.h_sigma.data(), H.h_sigma.size() * sizeof(float));
  memcpy(u, H.h_u.data(), H.h_u.size() * sizeof(float));
  memcpy(v, H.h_v.data(), H.h_v.size() * sizeof(float));
}

// sigma, u, v in/out (host); clock = {t, tau} in/out; dts (optional) receives every dt_eff.
// skip_visc != 0 leaves viscosity_uv out (its in-place update is a data race; the rest is deterministic).
extern "C" int ref_sw_run(const float *pf, float *sigma, float *u, float *v, int steps, float *clock, float *dts,
                          int skip_visc, float *ms) {
  Params P = params_from(pf);
  int nx = P.nx, ny = P.ny, N = nx * ny;
  DeviceState D;
  device_alloc(D, N);
  CUDA_CHECK(cudaMemcpy(D.d_sigma, sigma, N * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(D.d_u, u, N * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(D.d_v, v, N * sizeof(float), cudaMemcpyHostToDevice));
  dim3 bs(16, 16), gs((nx + bs.x - 1) / bs.x, (ny + bs.y - 1) / bs.y);
  CUDA_CHECK(cudaMalloc(&D.d_block_cmax, gs.x * gs.y * sizeof(float)));
  float t = clock[0], tau = clock[1], dtau = P.dtau;
  std::vector<float> h_blk(gs.x * gs.y);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int step = 0; step < steps; ++step) {
    size_t shmem = bs.x * bs.y * sizeof(float);
    wavespeed_block_max<<<gs, bs, shmem>>>(D.d_sigma, D.d_u, D.d_v, P.g, nx, ny, D.d_block_cmax);
    CUDA_CHECK(cudaMemcpy(h_blk.data(), D.d_block_cmax, h_blk.size() * sizeof(float), cudaMemcpyDeviceToHost));
    float cmax = 0.0f;
    for (float c : h_blk) cmax = std::max(cmax, c);
    if (cmax < 1e-12f) cmax = 1e-12f;
    float dt_cfl = P.CFL * fminf(P.dx, P.dy) / cmax;
    float dt_eff = fminf(t * dtau, dt_cfl);
    flux_x_kernel<<<gs, bs>>>(D.d_sigma, D.d_u, D.d_v, D.d_Fh_x, D.d_Fmx_x, D.d_Fmy_x, nx, ny, P.g);
    flux_y_kernel<<<gs, bs>>>(D.d_sigma, D.d_u, D.d_v, D.d_Gh_y, D.d_Gmx_y, D.d_Gmy_y, nx, ny, P.g);
    update_kernel<<<gs, bs>>>(D.d_sigma, D.d_u, D.d_v, D.d_Fh_x, D.d_Fmx_x, D.d_Fmy_x, D.d_Gh_y, D.d_Gmx_y,
                              D.d_Gmy_y, nx, ny, P.dx, P.dy, dt_eff, P.g);
    if (P.nu > 0.0f && !skip_visc) viscosity_uv<<<gs, bs>>>(D.d_u, D.d_v, nx, ny, P.dx, P.dy, P.nu, dt_eff);
    if (dts) dts[step] = dt_eff;
    tau += dtau;
    t *= expf(dtau);
  }
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (ms) cudaEventElapsedTime(ms, e0, e1);
  CUDA_CHECK(cudaMemcpy(sigma, D.d_sigma, N * sizeof(float), cudaMemcpyDeviceToHost));
  CUDA_CHECK(cudaMemcpy(u, D.d_u, N * sizeof(float), cudaMemcpyDeviceToHost));
  CUDA_CHECK(cudaMemcpy(v, D.d_v, N * sizeof(float), cudaMemcpyDeviceToHost));
  clock[0] = t; clock[1] = tau;
  device_free(D);
  return (int)e;
}
