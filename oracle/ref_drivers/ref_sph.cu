// ref_sph.cu — TEST INFRASTRUCTURE ONLY (never linked into the product).
// Compiles the UNMODIFIED reference translation unit tau_sph.cu (through the fake curses header,
// main() renamed) and drives its kernels with the host step control of its main loop (:663-722)
// restated below (the reference's headless mode never terminates and prints no state).
#define main ref_sph_main
#include "tau_sph.cu"
#undef main

static Params params_from(const float *a) {
  Params P;
  P.N = (int)a[0]; P.boxX = a[1]; P.boxY = a[2]; P.dTau = a[3]; P.t0 = a[4]; P.CFL = a[5];
  P.rho0 = a[6]; P.c0 = a[7]; P.gammaEOS = a[8]; P.hMul = a[9]; P.viscAlpha = a[10]; P.gravity = a[11];
  P.rain = a[12] != 0; P.useVisc = a[13] != 0; P.useGrav = a[14] != 0; P.viscSub = (int)a[15];
  P.useXSPH = a[16] != 0; P.xsphEps = a[17]; P.seed = (int)a[18];
  return P;
}

extern "C" void ref_sph_reset_particles(const float *pf, float *pos, float *vel) {
  Params P = params_from(pf);
  std::vector<float2> hp(P.N), hv(P.N);
  reset_particles(P, hp, hv);
  memcpy(pos, hp.data(), P.N * sizeof(float2));
  memcpy(vel, hv.data(), P.N * sizeof(float2));
}

// pos/vel (N x 2 floats) in/out; s, press, acc outputs; clock = {t, tau, rain_carry, step} in/out.
extern "C" int ref_sph_run(const float *pf, float *pos, float *vel, float *acc, float *s, float *press,
                           int nframes, float *clock, float *ms) {
  Params P = params_from(pf);
  DevState d{};
  memset(&d, 0, sizeof(d));
  CUDA_CHECK(cudaMalloc(&d.pos, P.N * sizeof(float2)));
  CUDA_CHECK(cudaMalloc(&d.vel, P.N * sizeof(float2)));
  CUDA_CHECK(cudaMalloc(&d.acc, P.N * sizeof(float2)));
  CUDA_CHECK(cudaMalloc(&d.s, P.N * sizeof(float)));
  CUDA_CHECK(cudaMalloc(&d.press, P.N * sizeof(float)));
  CUDA_CHECK(cudaMemcpy(d.pos, pos, P.N * sizeof(float2), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(d.vel, vel, P.N * sizeof(float2), cudaMemcpyHostToDevice));
  const float area = P.boxX * P.boxY;
  const float mass = (P.rho0 * area) / P.N;
  const float spacing = sqrtf(area / P.N);
  float h = P.hMul * spacing;
  ensure_cell_buffers(d, P.N, P.boxX, P.boxY, h);
  float t = clock[0], tau = clock[1], rain_carry = clock[2];
  long long step = (long long)clock[3];
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int f = 0; f < nframes; ++f) {
    float dTau_accum = 0.f;
    int K = (P.viscSub > 0 ? P.viscSub : 1);
    float dt_try = t * P.dTau;
    float dt_cfl = P.CFL * h / (P.c0 * (1.0f + 2.0f * P.viscAlpha));
    float dt_eff = fminf(dt_try, dt_cfl);
    float dt_sub = dt_eff / K;
    int BS = 256, GS = (P.N + BS - 1) / BS, M = d.Gx * d.Gy, GSm = (M + BS - 1) / BS;
    for (int k = 0; k < K; ++k) {
      k_clear_heads<<<GSm, BS>>>(d.cellHead, M);
      k_build_cells<<<GS, BS>>>(d.pos, P.N, d.cellHead, d.next, d.Gx, d.Gy, d.cell);
      k_density_pressure_cell<<<GS, BS>>>(d.pos, d.s, d.press, d.cellHead, d.next, P.N, mass, h, P.rho0,
                                          P.c0, P.gammaEOS, d.Gx, d.Gy, d.cell);
      k_forces_cell<<<GS, BS>>>(d.pos, d.vel, d.s, d.press, d.acc, d.cellHead, d.next, P.N, mass, h,
                                P.viscAlpha, P.c0, 0.f, -(P.useGrav ? P.gravity : 0.f), P.useVisc,
                                P.useGrav, d.Gx, d.Gy, d.cell);
      k_integrate<<<GS, BS>>>(d.pos, d.vel, d.acc, P.N, dt_sub, P.boxX, P.boxY);
      if (P.useXSPH && P.xsphEps > 0.f) {
        k_xsph_cell<<<GS, BS>>>(d.pos, d.vel, d.s, d.acc, d.cellHead, d.next, P.N, mass, h, P.xsphEps,
                                d.Gx, d.Gy, d.cell);
        k_apply_xsph<<<GS, BS>>>(d.vel, d.acc, P.N);
      }
      if (P.rain) {
        rain_carry += 0.02f * P.N * dt_sub;
        int nspawn = (int)rain_carry;
        rain_carry -= nspawn;
        if (nspawn > 0) {
          int BSr = 128, GSr = (nspawn + BSr - 1) / BSr;
          k_rain<<<GSr, BSr>>>(d.pos, d.vel, P.N, nspawn, P.boxX, P.boxY, P.c0, (unsigned)(P.seed + step));
        }
      }
      float dTau_actual = dt_sub / fmaxf(t, 1e-9f);
      dTau_accum += dTau_actual;
      t = P.t0 * expf(tau + dTau_accum);
    }
    tau += dTau_accum;
    step++;
  }
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (ms) cudaEventElapsedTime(ms, e0, e1);
  CUDA_CHECK(cudaMemcpy(pos, d.pos, P.N * sizeof(float2), cudaMemcpyDeviceToHost));
  CUDA_CHECK(cudaMemcpy(vel, d.vel, P.N * sizeof(float2), cudaMemcpyDeviceToHost));
  if (acc) CUDA_CHECK(cudaMemcpy(acc, d.acc, P.N * sizeof(float2), cudaMemcpyDeviceToHost));
  if (s) CUDA_CHECK(cudaMemcpy(s, d.s, P.N * sizeof(float), cudaMemcpyDeviceToHost));
  if (press) CUDA_CHECK(cudaMemcpy(press, d.press, P.N * sizeof(float), cudaMemcpyDeviceToHost));
  clock[0] = t; clock[1] = tau; clock[2] = rain_carry; clock[3] = (float)step;
  cudaFree(d.pos); cudaFree(d.vel); cudaFree(d.acc); cudaFree(d.s); cudaFree(d.press);
  cudaFree(d.cellHead); cudaFree(d.next);
  return (int)e;
}

// render pass :747-755: k_clear_grid + k_rasterize on caller-provided positions
extern "C" int ref_sph_rasterize(const float *pos, int N, int W, int H, float boxX, float boxY, int *grid2) {
  float2 *dpos; int *dgrid;
  const size_t cells = (size_t)W * 2 * H;
  CUDA_CHECK(cudaMalloc(&dpos, (size_t)N * sizeof(float2)));
  CUDA_CHECK(cudaMalloc(&dgrid, cells * sizeof(int)));
  CUDA_CHECK(cudaMemcpy(dpos, pos, (size_t)N * sizeof(float2), cudaMemcpyHostToDevice));
  int BSg = 256, GSg = ((int)cells + BSg - 1) / BSg;
  k_clear_grid<<<GSg, BSg>>>(dgrid, (int)cells);
  int BSp = 256, GSp = (N + BSp - 1) / BSp;
  k_rasterize<<<GSp, BSp>>>(dpos, N, dgrid, W, H, boxX, boxY);
  cudaError_t e = cudaDeviceSynchronize();
  CUDA_CHECK(cudaMemcpy(grid2, dgrid, cells * sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(dpos); cudaFree(dgrid);
  return (int)e;
}
