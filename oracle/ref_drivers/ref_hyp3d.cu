// ref_hyp3d.cu — TEST INFRASTRUCTURE ONLY (never linked into the product).
// Compiles the UNMODIFIED reference translation unit tau_hypersonic_3d_cuda.cu (through the fake
// raylib/rlgl headers in oracle/shims, main() renamed) and drives its k_build_solid_mask / k_init /
// k_step with the host clock/controller of its frame loop (:1678-1712) restated below.
#define main ref_hyp3d_main
#include "tau_hypersonic_3d_cuda.cu"
#undef main

static Params params_from(const float *a, int nx, int ny, int nz) {
  Params hp{};
  hp.nx = nx; hp.ny = ny; hp.nz = nz;
  hp.dx = a[0]; hp.dy = a[1]; hp.dz = a[2]; hp.cfl = a[3]; hp.u_ref = a[4]; hp.R = a[5];
  hp.gamma_floor = a[6]; hp.Twall = a[7]; hp.tau_vib = a[8]; hp.theta_v = a[9];
  hp.sdf_cx = a[10]; hp.sdf_cy = a[11]; hp.sdf_cz = a[12]; hp.sdf_r = a[13];
  hp.inflow_r = a[14]; hp.inflow_p = a[15]; hp.inflow_u = a[16]; hp.inflow_v = a[17]; hp.inflow_w = a[18];
  hp.sponge_n = (int)a[19]; hp.sponge_strength = a[20]; hp.sponge_out_n = (int)a[21];
  hp.sponge_out_strength = a[22];
  return hp;
}

// pf: 23 floats in the order of params_from.  planes: 6 host arrays of N floats (xi, phix, phiy,
// phiz, lam, zet), outputs (and inputs when !do_init).  clock: {t, d_tau} in/out.
extern "C" int ref_hyp3d_run(const float *pf, int nx, int ny, int nz, int steps, int do_init,
                             float *const *planes, uint8_t *solid_out, float *clock, float *dt_hist,
                             float *maxs_hist, float *ms) {
  Params hp = params_from(pf, nx, ny, nz);
  cudaError_t e;
  if ((e = cudaMemcpyToSymbol(P, &hp, sizeof(Params)))) return e;
  size_t N = (size_t)nx * ny * nz, bytes = N * sizeof(float);
  float *d[6], *d2[6], *d_maxs;
  uint8_t *d_solid;
  for (int f = 0; f < 6; ++f) { ck(cudaMalloc(&d[f], bytes), "malloc"); ck(cudaMalloc(&d2[f], bytes), "malloc"); }
  ck(cudaMalloc(&d_maxs, sizeof(float)), "malloc maxs");
  ck(cudaMalloc(&d_solid, N), "malloc solid");
  dim3 block(8, 8, 4);
  dim3 grid((nx + block.x - 1) / block.x, (ny + block.y - 1) / block.y, (nz + block.z - 1) / block.z);
  size_t smem = (size_t)(block.x + 2 * WENO_HALO) * (block.y + 2 * WENO_HALO) * (block.z + 2 * WENO_HALO) *
                (6 * sizeof(float) + sizeof(uint8_t));
  ck(cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "attr");
  k_build_solid_mask<<<grid, block>>>(d_solid);
  if (do_init) {
    k_init<<<grid, block>>>(d[0], d[1], d[2], d[3], d[4], d[5], d_solid);
  } else {
    for (int f = 0; f < 6; ++f) ck(cudaMemcpy(d[f], planes[f], bytes, cudaMemcpyHostToDevice), "h2d");
  }
  ck(cudaDeviceSynchronize(), "init sync");
  float t = clock[0], d_tau = clock[1], maxs = 0.f;
  cudaEvent_t ev0, ev1;
  cudaEventCreate(&ev0); cudaEventCreate(&ev1);
  cudaEventRecord(ev0);
  for (int s = 0; s < steps; s++) {
    t *= expf(d_tau);
    float dt = t * d_tau;
    float ramp = t / 0.02f;
    float inflow_gain = fminf(fmaxf(ramp, 0.f), 1.f);
    float zero = 0.f;
    ck(cudaMemcpy(d_maxs, &zero, sizeof(float), cudaMemcpyHostToDevice), "set maxs");
    k_step<<<grid, block, smem>>>(d[0], d[1], d[2], d[3], d[4], d[5], d2[0], d2[1], d2[2], d2[3], d2[4],
                                  d2[5], dt, inflow_gain, d_maxs, d_solid);
    ck(cudaGetLastError(), "k_step launch");
    ck(cudaMemcpy(&maxs, d_maxs, sizeof(float), cudaMemcpyDeviceToHost), "get maxs");
    float dt_cfl = hp.cfl / fmaxf(maxs, 1e-9f);
    if (dt > 1.10f * dt_cfl) d_tau *= 0.80f;
    else if (dt < 0.85f * dt_cfl) d_tau *= 1.10f;
    d_tau = fminf(fmaxf(d_tau, 1e-7f), 5e-2f);
    if (dt_hist) dt_hist[s] = dt;
    if (maxs_hist) maxs_hist[s] = maxs;
    for (int f = 0; f < 6; ++f) std::swap(d[f], d2[f]);
  }
  cudaEventRecord(ev1);
  e = cudaDeviceSynchronize();
  if (ms) cudaEventElapsedTime(ms, ev0, ev1);
  for (int f = 0; f < 6; ++f) ck(cudaMemcpy(planes[f], d[f], bytes, cudaMemcpyDeviceToHost), "d2h");
  if (solid_out) ck(cudaMemcpy(solid_out, d_solid, N, cudaMemcpyDeviceToHost), "d2h solid");
  clock[0] = t; clock[1] = d_tau;
  for (int f = 0; f < 6; ++f) { cudaFree(d[f]); cudaFree(d2[f]); }
  cudaFree(d_maxs); cudaFree(d_solid);
  return (int)e;
}

// k_vis :800-905 on caller-provided planes, with main()'s launch shape (:1715)
extern "C" int ref_hyp3d_vis(const float *pf, int nx, int ny, int nz, const float *const *planes, int mode,
                             float *out) {
  Params hp = params_from(pf, nx, ny, nz);
  cudaError_t e;
  if ((e = cudaMemcpyToSymbol(P, &hp, sizeof(Params)))) return e;
  size_t N = (size_t)nx * ny * nz, bytes = N * sizeof(float);
  float *d[6], *d_out;
  uint8_t *d_solid;
  for (int f = 0; f < 6; ++f) {
    ck(cudaMalloc(&d[f], bytes), "malloc");
    ck(cudaMemcpy(d[f], planes[f], bytes, cudaMemcpyHostToDevice), "h2d");
  }
  ck(cudaMalloc(&d_out, bytes), "malloc out");
  ck(cudaMalloc(&d_solid, N), "malloc solid");
  dim3 block(8, 8, 4);
  dim3 grid((nx + block.x - 1) / block.x, (ny + block.y - 1) / block.y, (nz + block.z - 1) / block.z);
  k_build_solid_mask<<<grid, block>>>(d_solid);
  k_vis<<<grid, block>>>(d[0], d[1], d[2], d[3], d[4], d[5], d_solid, d_out, mode);
  e = cudaDeviceSynchronize();
  ck(cudaMemcpy(out, d_out, bytes, cudaMemcpyDeviceToHost), "d2h");
  for (int f = 0; f < 6; ++f) cudaFree(d[f]);
  cudaFree(d_out); cudaFree(d_solid);
  return (int)e;
}
