/* sw_oracle.c — CPU restatement (fp32) of the reference 2-D shallow-water step.
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and the bench's baseline legs may
 * call this; the product never links or imports it.
 *
 * Follows tau_shallow_water.cu kernel by kernel (line numbers cited): wavespeed_block_max :394-421
 * + the host dt rule :679-684, hll_x :322-352, hll_y :355-385, flux_x_kernel :424-446,
 * flux_y_kernel :449-470, update_kernel :473-513, viscosity_uv :516-551, clock :767-768,
 * initialize_host :238-277.  One deliberate difference: viscosity_uv updates u, v in place while
 * neighbouring threads read them (a data race); here — as in the product — every cell reads the
 * pre-update state (Jacobi).
 *
 * Pinning: the reference commits no golden values for this solver.  The oracle is pinned against the
 * reference's OWN code: tau_shallow_water.cu's initialize_host and the bodies of flux_x_kernel,
 * flux_y_kernel, update_kernel and viscosity_uv, compiled for the host by g++ against a fake
 * cuda_runtime.h (oracle/shims/hostcuda) and emulated thread by thread by
 * oracle/ref_drivers/ref_sw_host.cpp — same libm, same -ffp-contract=off, so the comparison is
 * BIT-EXACT (fields, every dt, clock): committed as tests/golden/sw_ref_host.npz (generator
 * tests/golden/make_golden_host.py) and repeated live where oracle/_ref is built
 * (tests/test_oracle_cpu.py).  Not yet pinned against the reference's kernels RUN ON A GPU with the
 * -use_fast_math intrinsics (oracle/_ref/libref_sw.so builds; tests/golden/make_golden_gpu.py::sw
 * writes that fixture on the first GPU run): this file was written after the round's GPU time was spent.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int nx, ny;
  float dx, dy;
  float g, f0, nu, H0;
  float bumpAmp, bumpSigma, CFL;
  float offx, offy, asym, swirl, swirlRc;
  float tau0, t0, dtau;
} oracle_sw_params;

static int wraps(int i, int n) { i %= n; if (i < 0) i += n; return i; }              /* :91-96 */

void oracle_sw_default_params(oracle_sw_params *p) {                                   /* :52-89 */
  memset(p, 0, sizeof(*p));
  p->nx = 512; p->ny = 512; p->dx = 1.0f; p->dy = 1.0f;
  p->g = 9.81f; p->f0 = 1.0f; p->nu = 0.001f; p->H0 = 1000.0f;
  p->bumpAmp = 1.0f; p->bumpSigma = 1.0f; p->CFL = 0.5f;
  p->offx = 100.0f; p->offy = 100.0f; p->asym = 10.0f; p->swirl = 1.0f; p->swirlRc = 100.0f;
  p->tau0 = 0.0f; p->t0 = 1.0f; p->dtau = 1.0f;
}

void oracle_sw_init(const oracle_sw_params *P, float *sigma, float *u, float *v) {     /* :238-277 */
  const int nx = P->nx, ny = P->ny;
  float cx = 0.5f * nx + P->offx, cy = 0.5f * ny + P->offy;
  float sig2 = P->bumpSigma * P->bumpSigma;
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      float dx = i - cx, dy = j - cy;
      float r2 = (dx * dx + dy * dy) / sig2;
      float theta = atan2f(dy, dx);
      float mod = 1.0f + P->asym * cosf(theta);
      float bump = P->bumpAmp * mod;
      float h = P->H0 + bump * expf(-0.5f * r2);
      int id = j * nx + i;
      sigma[id] = logf(fmaxf(h, 1e-6f));
      float rx = dx * P->dx, ry = dy * P->dy;
      float r = sqrtf(rx * rx + ry * ry);
      float rc = P->swirlRc * fminf(P->dx, P->dy);
      float u_theta = 0.0f;
      if (r > 0.0f && P->swirl != 0.0f) u_theta = P->swirl * r * expf(-0.5f * (r / rc) * (r / rc));
      u[id] = (r > 0.0f) ? (-u_theta * (ry / r)) : 0.0f;
      v[id] = (r > 0.0f) ? (u_theta * (rx / r)) : 0.0f;
    }
}

/* hll_x :322-352 (d = 0) and hll_y :355-385 (d = 1); L/R are (h, u, v); F = fluxes of (h, h u, h v) */
static void hll(int d, const float L[3], const float R[3], float g, float F[3]) {
  const float hL = L[0], hR = R[0];
  const float qL = d ? L[2] : L[1], qR = d ? R[2] : R[1]; /* normal velocity */
  const float cL = sqrtf(g * hL), cR = sqrtf(g * hR);
  const float sL = fminf(qL - cL, qR - cR), sR = fmaxf(qL + cL, qR + cR);
  const float mL = hL * L[1], mR = hR * R[1], nL = hL * L[2], nR = hR * R[2];
  float FL[3], FR[3];
  if (!d) {
    FL[0] = mL; FL[1] = mL * L[1] + 0.5f * g * hL * hL; FL[2] = mL * L[2];
    FR[0] = mR; FR[1] = mR * R[1] + 0.5f * g * hR * hR; FR[2] = mR * R[2];
  } else {
    FL[0] = nL; FL[1] = mL * L[2]; FL[2] = nL * L[2] + 0.5f * g * hL * hL;
    FR[0] = nR; FR[1] = mR * R[2]; FR[2] = nR * R[2] + 0.5f * g * hR * hR;
  }
  if (sL >= 0.0f) { F[0] = FL[0]; F[1] = FL[1]; F[2] = FL[2]; return; }
  if (sR <= 0.0f) { F[0] = FR[0]; F[1] = FR[1]; F[2] = FR[2]; return; }
  const float inv = 1.0f / (sR - sL);
  F[0] = (sR * FL[0] - sL * FR[0] + sR * sL * (hR - hL)) * inv;
  F[1] = (sR * FL[1] - sL * FR[1] + sR * sL * (mR - mL)) * inv;
  F[2] = (sR * FL[2] - sL * FR[2] + sR * sL * (nR - nL)) * inv;
}

static float sw_cmax(const oracle_sw_params *P, const float *sigma, const float *u, const float *v) {
  float cmax = 0.0f;                                                                    /* :394-421, :676-678 */
  for (int k = 0; k < P->nx * P->ny; ++k) {
    float c = sqrtf(P->g * expf(sigma[k]));
    cmax = fmaxf(cmax, fmaxf(fabsf(u[k]) + c, fabsf(v[k]) + c));
  }
  return cmax;
}

/* one do_step :669-705; returns dt_eff.  W: 6 flux planes + 2 scratch planes of nx*ny floats */
static float sw_step(const oracle_sw_params *P, float *sigma, float *u, float *v, float *W, float t) {
  const int nx = P->nx, ny = P->ny, n = nx * ny;
  float *Fh = W, *Fmx = W + n, *Fmy = W + 2 * n, *Gh = W + 3 * n, *Gmx = W + 4 * n, *Gmy = W + 5 * n;
  float *tu = W + 6 * n, *tv = W + 7 * n;
  float cmax = sw_cmax(P, sigma, u, v);
  if (cmax < 1e-12f) cmax = 1e-12f;
  const float dt_cfl = P->CFL * fminf(P->dx, P->dy) / cmax;
  const float dt = fminf(t * P->dtau, dt_cfl);
#define ID(i, j) (wraps(j, ny) * nx + wraps(i, nx))
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      const int a = ID(i, j), bx = ID(i + 1, j), by = ID(i, j + 1);
      const float A[3] = {expf(sigma[a]), u[a], v[a]};
      const float BX[3] = {expf(sigma[bx]), u[bx], v[bx]};
      const float BY[3] = {expf(sigma[by]), u[by], v[by]};
      float F[3];
      hll(0, A, BX, P->g, F);
      Fh[a] = F[0]; Fmx[a] = F[1]; Fmy[a] = F[2];
      hll(1, A, BY, P->g, F);
      Gh[a] = F[0]; Gmx[a] = F[1]; Gmy[a] = F[2];
    }
  const float invdx = 1.0f / P->dx, invdy = 1.0f / P->dy;
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) { /* update_kernel: each cell touches only its own state */
      const int id = ID(i, j), im = ID(i - 1, j), jm = ID(i, j - 1);
      float h = expf(sigma[id]);
      float mx = h * u[id], my = h * v[id];
      const float dFx_h = Fh[id] - Fh[im], dFx_mx = Fmx[id] - Fmx[im], dFx_my = Fmy[id] - Fmy[im];
      const float dGy_h = Gh[id] - Gh[jm], dGy_mx = Gmx[id] - Gmx[jm], dGy_my = Gmy[id] - Gmy[jm];
      h -= dt * (dFx_h * invdx + dGy_h * invdy);
      mx -= dt * (dFx_mx * invdx + dGy_mx * invdy);
      my -= dt * (dFx_my * invdx + dGy_my * invdy);
      h = fmaxf(h, 1e-6f);
      sigma[id] = logf(h);
      u[id] = mx / h;
      v[id] = my / h;
    }
  if (P->nu > 0.0f) { /* :700-703 */
    const float invdx2 = 1.0f / (P->dx * P->dx), invdy2 = 1.0f / (P->dy * P->dy);
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const int id = ID(i, j), xp = ID(i + 1, j), xm = ID(i - 1, j), yp = ID(i, j + 1), ym = ID(i, j - 1);
        const float du = (u[xp] - 2.0f * u[id] + u[xm]) * invdx2 + (u[yp] - 2.0f * u[id] + u[ym]) * invdy2;
        const float dv = (v[xp] - 2.0f * v[id] + v[xm]) * invdx2 + (v[yp] - 2.0f * v[id] + v[ym]) * invdy2;
        tu[id] = u[id] + P->nu * dt * du;
        tv[id] = v[id] + P->nu * dt * dv;
      }
    memcpy(u, tu, sizeof(float) * (size_t)n);
    memcpy(v, tv, sizeof(float) * (size_t)n);
  }
#undef ID
  return dt;
}

/* steps x { do_step; tau += dtau; t *= expf(dtau) }.  clock2 = {t, tau} in/out; dts[steps] out (or NULL) */
void oracle_sw_run(const oracle_sw_params *P, float *sigma, float *u, float *v, int steps, float *clock2,
                   float *dts) {
  const size_t n = (size_t)P->nx * P->ny;
  float *W = (float *)malloc(sizeof(float) * 8 * n);
  float t = clock2[0], tau = clock2[1];
  for (int s = 0; s < steps; ++s) {
    const float dt = sw_step(P, sigma, u, v, W, t);
    if (dts) dts[s] = dt;
    tau += P->dtau;
    t *= expf(P->dtau);
  }
  clock2[0] = t;
  clock2[1] = tau;
  free(W);
}

float oracle_sw_cmax(const oracle_sw_params *P, const float *sigma, const float *u, const float *v) {
  return sw_cmax(P, sigma, u, v);
}
