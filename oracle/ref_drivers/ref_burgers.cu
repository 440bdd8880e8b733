// ref_burgers.cu — TEST INFRASTRUCTURE ONLY (never linked into the product).
// Compiles the UNMODIFIED reference translation unit tau_burgers.cu (from /root/reference, via the
// fake curses header in oracle/shims) and drives its own kernels with main()'s launch shapes; the
// host loop restates `do_step` + the clock update (tau_burgers.cu:677-718, :768-769).
#define main ref_burgers_main
#include "tau_burgers.cu"
#undef main

// pf: nx ny dx dy nu u0 amp bsig swirl rc offx offy asym CFL tau0 t0 dtau muscl visc_substeps colehopf ck ca
static Params params_from(const float *pf) {
  Params P;
  P.nx = (int)pf[0]; P.ny = (int)pf[1]; P.dx = pf[2]; P.dy = pf[3]; P.nu = pf[4]; P.u0 = pf[5];
  P.amp = pf[6]; P.bsig = pf[7]; P.swirl = pf[8]; P.rc = pf[9]; P.offx = pf[10]; P.offy = pf[11];
  P.asym = pf[12]; P.CFL = pf[13]; P.tau0 = pf[14]; P.t0 = pf[15]; P.dtau = pf[16];
  P.muscl = pf[17] != 0.f; P.visc_substeps = (int)pf[18]; P.colehopf = pf[19] != 0.f;
  P.ck = (int)pf[20]; P.ca = pf[21];
  if (P.colehopf) P.ny = 1;  // :649-650
  return P;
}

extern "C" void ref_burgers_init(const float *pf, float *phi_u, float *phi_v) {
  Params P = params_from(pf);
  HostState H;
  initialize_host(P, H);
  memcpy(phi_u, H.h_phi_u.data(), H.h_phi_u.size() * sizeof(float));
  memcpy(phi_v, H.h_phi_v.data(), H.h_phi_v.size() * sizeof(float));
}

// phi_u, phi_v in/out (host); clock = {t, tau} in/out; dts (optional) receives every dt_eff.
// skip_visc != 0 leaves viscosity_step out (its in-place update is a data race; the convective part
// alone is deterministic).
extern "C" int ref_burgers_run(const float *pf, float *phi_u, float *phi_v, int steps, float *clock,
                               float *dts, int skip_visc, float *ms) {
  Params P = params_from(pf);
  int nx = P.nx, ny = P.ny, N = nx * ny;
  DeviceState D;
  device_alloc(D, N);
  CUDA_CHECK(cudaMemcpy(D.d_phi_u, phi_u, N * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(D.d_phi_v, phi_v, N * sizeof(float), cudaMemcpyHostToDevice));
  dim3 bs(16, 16), gs((nx + bs.x - 1) / bs.x, (ny + bs.y - 1) / bs.y);
  CUDA_CHECK(cudaMalloc(&D.d_block_max, gs.x * gs.y * sizeof(float)));
  float t = clock[0], tau = clock[1], dtau = P.dtau;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int step = 0; step < steps; ++step) {
    size_t shmem = bs.x * bs.y * sizeof(float);
    wavespeed_block_max<<<gs, bs, shmem>>>(D.d_phi_u, D.d_phi_v, P.u0, nx, ny, 1.0f / P.dx,
                                           (ny > 1 ? 1.0f / P.dy : 0.0f), D.d_block_max);
    std::vector<float> h_blk(gs.x * gs.y);
    CUDA_CHECK(cudaMemcpy(h_blk.data(), D.d_block_max, h_blk.size() * sizeof(float), cudaMemcpyDeviceToHost));
    float smax = 1e-12f;
    for (float v : h_blk) smax = fmaxf(smax, v);
    float dt_cfl = P.CFL / smax;
    float dt_eff = fminf(t * dtau, dt_cfl);
    flux_x_kernel<<<gs, bs>>>(D.d_phi_u, D.d_phi_v, D.d_Fu_x, D.d_Fv_x, nx, ny, P.u0, P.muscl ? 1 : 0);
    if (!P.colehopf)
      flux_y_kernel<<<gs, bs>>>(D.d_phi_u, D.d_phi_v, D.d_Gu_y, D.d_Gv_y, nx, ny, P.u0, P.muscl ? 1 : 0);
    update_convective<<<gs, bs>>>(D.d_phi_u, D.d_phi_v, D.d_Fu_x, D.d_Fv_x, D.d_Gu_y, D.d_Gv_y, nx, ny, P.dx,
                                  P.dy, dt_eff, P.u0, P.colehopf ? 1 : 0);
    int K = (P.visc_substeps > 0 ? P.visc_substeps : 1);
    float sub = dt_eff / K;
    if (!skip_visc)
      for (int k = 0; k < K; ++k)
        viscosity_step<<<gs, bs>>>(D.d_phi_u, D.d_phi_v, nx, ny, P.dx, P.dy, P.nu, sub, P.u0, P.colehopf ? 1 : 0);
    if (dts) dts[step] = dt_eff;
    tau += dtau;
    t *= expf(dtau);
  }
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (ms) cudaEventElapsedTime(ms, e0, e1);
  CUDA_CHECK(cudaMemcpy(phi_u, D.d_phi_u, N * sizeof(float), cudaMemcpyDeviceToHost));
  CUDA_CHECK(cudaMemcpy(phi_v, D.d_phi_v, N * sizeof(float), cudaMemcpyDeviceToHost));
  clock[0] = t; clock[1] = tau;
  device_free(D);
  return (int)e;
}
