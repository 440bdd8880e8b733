// ref_sw_host.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
// Runs the reference's OWN shallow-water kernel bodies ON THE CPU: g++ compiles the part of
// tau_shallow_water.cu that precedes main() (REF_SRC, produced by oracle/Makefile with sed into a
// scratch file that is deleted after compilation) against oracle/shims/hostcuda/cuda_runtime.h, and
// this driver emulates the launches of `do_step` (:669-705) thread by thread.  That pins
// oracle/sw_oracle.c against the reference's code without a GPU: same expressions, same libm, same
// compiler flags (-O2 -ffp-contract=off) => bit-identical results are demanded by the tests.
// Not emulated: wavespeed_block_max (:394-421, a __syncthreads() tree reduction); the max over cells
// of its per-thread value is taken directly — fmaxf is exact, so the reduction order cannot matter.
// viscosity_uv's in-place update runs in the emulation's sequential thread order (on a GPU the order
// is unspecified).
#include REF_SRC

float sdata[1];  // definition for wavespeed_block_max's `extern __shared__` (the kernel is never called)

static Params params_from(const float *pf) {
  Params P;
  P.nx = (int)pf[0]; P.ny = (int)pf[1]; P.dx = pf[2]; P.dy = pf[3]; P.g = pf[4]; P.f0 = pf[5]; P.nu = pf[6];
  P.H0 = pf[7]; P.bumpAmp = pf[8]; P.bumpSigma = pf[9]; P.CFL = pf[10]; P.offx = pf[11]; P.offy = pf[12];
  P.asym = pf[13]; P.swirl = pf[14]; P.swirlRc = pf[15]; P.tau0 = pf[16]; P.t0 = pf[17]; P.dtau = pf[18];
  return P;
}

extern "C" void ref_sw_host_init(const float *pf, float *sigma, float *u, float *v) {
  Params P = params_from(pf);
  HostState H;
  initialize_host(P, H);
  memcpy(sigma, H.h_sigma.data(), H.h_sigma.size() * sizeof(float));
  memcpy(u, H.h_u.data(), H.h_u.size() * sizeof(float));
  memcpy(v, H.h_v.data(), H.h_v.size() * sizeof(float));
}

extern "C" int ref_sw_host_run(const float *pf, float *sigma, float *u, float *v, int steps, float *clock,
                               float *dts, int skip_visc) {
  Params P = params_from(pf);
  int nx = P.nx, ny = P.ny, N = nx * ny;
  DeviceState D;
  device_alloc(D, N);
  memcpy(D.d_sigma, sigma, N * sizeof(float));
  memcpy(D.d_u, u, N * sizeof(float));
  memcpy(D.d_v, v, N * sizeof(float));
  dim3 bs(16, 16), gs((nx + bs.x - 1) / bs.x, (ny + bs.y - 1) / bs.y);
  float t = clock[0], tau = clock[1], dtau = P.dtau;
  for (int step = 0; step < steps; ++step) {
    float cmax = 0.0f;
    for (int k = 0; k < N; ++k) {  // per-thread value of wavespeed_block_max :404-409
      float h = expf(D.d_sigma[k]);
      float c = sqrtf(P.g * h);
      float uu = fabsf(D.d_u[k]), vv = fabsf(D.d_v[k]);
      cmax = std::max(cmax, fmaxf(uu + c, vv + c));
    }
    if (cmax < 1e-12f) cmax = 1e-12f;
    float dt_cfl = P.CFL * fminf(P.dx, P.dy) / cmax;
    float dt_eff = fminf(t * dtau, dt_cfl);
    TAU_HC_LAUNCH(gs, bs, flux_x_kernel(D.d_sigma, D.d_u, D.d_v, D.d_Fh_x, D.d_Fmx_x, D.d_Fmy_x, nx, ny, P.g));
    TAU_HC_LAUNCH(gs, bs, flux_y_kernel(D.d_sigma, D.d_u, D.d_v, D.d_Gh_y, D.d_Gmx_y, D.d_Gmy_y, nx, ny, P.g));
    TAU_HC_LAUNCH(gs, bs, update_kernel(D.d_sigma, D.d_u, D.d_v, D.d_Fh_x, D.d_Fmx_x, D.d_Fmy_x, D.d_Gh_y,
                                        D.d_Gmx_y, D.d_Gmy_y, nx, ny, P.dx, P.dy, dt_eff, P.g));
    if (P.nu > 0.0f && !skip_visc)
      TAU_HC_LAUNCH(gs, bs, viscosity_uv(D.d_u, D.d_v, nx, ny, P.dx, P.dy, P.nu, dt_eff));
    if (dts) dts[step] = dt_eff;
    tau += dtau;
    t *= expf(dtau);
  }
  memcpy(sigma, D.d_sigma, N * sizeof(float));
  memcpy(u, D.d_u, N * sizeof(float));
  memcpy(v, D.d_v, N * sizeof(float));
  clock[0] = t; clock[1] = tau;
  device_free(D);
  return 0;
}
