/* sph_oracle.c — CPU restatement of the reference SPH sub-step.  TEST INFRASTRUCTURE ONLY: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call
 * this; the product never links or imports it.
 *
 * Follows tau_sph.cu kernel by kernel: cell index (:141-176), density/pressure (:178-213), forces
 * (:215-272), integration (:324-355), XSPH (:274-322), rain (:377-392) and the host step control
 * (:663-722).  Two places where the reference is not a function of its inputs are fixed here (and
 * identically in the product): the per-cell linked list's order — an atomicExch race (:175) — is
 * replaced by ascending particle index, and k_rain's colliding writes (:389-391) by "the highest
 * spawn index wins".  The reference is built with -use_fast_math; this file uses libm, so the
 * agreement with the GPU reference is at fp32 round-off, not bit level.
 *
 * Pinning: the reference has no tests or golden vectors for this solver;
 * tests/golden/sph_ref_*.npz hold outputs of the reference's own kernels run on a B200 through
 * oracle/_ref/libref_sph.so; tests/test_oracle_cpu.py compares this file against them.  The
 * integer part (cell keys and their stable sort) is pinned bit-exactly. */
#define _GNU_SOURCE
#include <math.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int N;
  float boxX, boxY, dTau, t0, CFL, rho0, c0, gammaEOS, hMul, viscAlpha, gravity;
  int rain, useVisc, useGrav, viscSub, useXSPH;
  float xsphEps;
  int seed;
} oracle_sph_params;

typedef struct { /* host-side state of the reference main loop */
  float t, tau, rain_carry;
  long long step;
} oracle_sph_clock;

static int grid_c(float x, float cell, int G) { /* :141-157 */
  int g = (int)floorf(x / cell);
  if (g < 0) g = 0;
  if (g >= G) g = G - 1;
  return g;
}
static float W_cubic(float r, float h) { /* :105-116 */
  float q = r / h;
  const float alpha = (float)(10.0f / (7.0f * M_PI * h * h));
  if (q < 1.0f) { float q2 = q * q, q3 = q2 * q; return alpha * (1.f - 1.5f * q2 + 0.75f * q3); }
  else if (q < 2.0f) { float t = 2.f - q; return alpha * 0.25f * t * t * t; }
  return 0.f;
}
static void gradW_cubic(float rx, float ry, float r, float h, float *gx, float *gy) { /* :118-133 */
  if (r <= 1e-8f || r >= 2.0f * h) { *gx = 0.f; *gy = 0.f; return; }
  float q = r / h;
  const float alpha = (float)(10.0f / (7.0f * M_PI * h * h));
  float dWdq;
  if (q < 1.0f) dWdq = alpha * (-3.0f * q + 2.25f * q * q);
  else { float t = 2.0f - q; dWdq = alpha * (-0.75f * t * t); }
  float invr = 1.0f / r, dWdr = dWdq / h;
  *gx = dWdr * rx * invr;
  *gy = dWdr * ry * invr;
}

void oracle_sph_derived(const oracle_sph_params *P, float *mass, float *h, float *cell, int *Gx, int *Gy) {
  const float area = P->boxX * P->boxY; /* :573-576 */
  *mass = (P->rho0 * area) / P->N;
  *h = P->hMul * sqrtf(area / P->N);
  *cell = 2.0f * *h; /* ensure_cell_buffers :512-540 */
  *Gx = (int)ceilf(P->boxX / *cell);
  *Gy = (int)ceilf(P->boxY / *cell);
  if (*Gx < 1) *Gx = 1;
  if (*Gy < 1) *Gy = 1;
}

/* cell key per particle and the stable sort of (key, index) — counting sort by key keeps ascending
 * particle index inside a cell, i.e. exactly what a stable radix sort produces */
void oracle_sph_cell_sort(const oracle_sph_params *P, const float *pos, uint32_t *keys_sorted,
                          uint32_t *vals_sorted, int *cellStart /* M+1 */) {
  float mass, h, cell; int Gx, Gy;
  oracle_sph_derived(P, &mass, &h, &cell, &Gx, &Gy);
  const int M = Gx * Gy, N = P->N;
  uint32_t *key = (uint32_t *)malloc(sizeof(uint32_t) * N);
  memset(cellStart, 0, sizeof(int) * (M + 1));
  for (int i = 0; i < N; ++i) {
    key[i] = (uint32_t)(grid_c(pos[2 * i + 1], cell, Gy) * Gx + grid_c(pos[2 * i], cell, Gx));
    cellStart[key[i] + 1]++;
  }
  for (int c = 0; c < M; ++c) cellStart[c + 1] += cellStart[c];
  int *fill = (int *)calloc(M, sizeof(int));
  for (int i = 0; i < N; ++i) {
    int d = cellStart[key[i]] + fill[key[i]]++;
    keys_sorted[d] = key[i];
    vals_sorted[d] = (uint32_t)i;
  }
  free(fill);
  free(key);
}

/* one sub-step (:676-716) in place on pos/vel (N x 2); s, press, acc are outputs */
void oracle_sph_substep(const oracle_sph_params *P, float *pos, float *vel, float *acc, float *s,
                        float *press, float dt, oracle_sph_clock *clk) {
  float mass, h, cell; int Gx, Gy;
  oracle_sph_derived(P, &mass, &h, &cell, &Gx, &Gy);
  const int M = Gx * Gy, N = P->N;
  uint32_t *ks = (uint32_t *)malloc(sizeof(uint32_t) * N), *vs = (uint32_t *)malloc(sizeof(uint32_t) * N);
  int *cs = (int *)malloc(sizeof(int) * (M + 1));
  oracle_sph_cell_sort(P, pos, ks, vs, cs);
  const float twoh = 2.f * h, twoh2 = twoh * twoh;
  float *rho_rt = (float *)malloc(sizeof(float) * N);
  /* density + pressure :178-213 */
  for (int i = 0; i < N; ++i) {
    const float xi = pos[2 * i], yi = pos[2 * i + 1];
    const int gx = grid_c(xi, cell, Gx), gy = grid_c(yi, cell, Gy);
    float rho = 0.f;
    for (int oy = -1; oy <= 1; ++oy)
      for (int ox = -1; ox <= 1; ++ox) {
        int cx = gx + ox, cy = gy + oy;
        if ((unsigned)cx >= (unsigned)Gx || (unsigned)cy >= (unsigned)Gy) continue;
        int c = cy * Gx + cx;
        for (int q = cs[c]; q < cs[c + 1]; ++q) {
          int j = (int)vs[q];
          float rx = xi - pos[2 * j], ry = yi - pos[2 * j + 1];
          float r2 = rx * rx + ry * ry;
          if (r2 >= twoh2) continue;
          rho += mass * W_cubic(sqrtf(r2), h);
        }
      }
    float si = logf(fmaxf(rho, 1e-6f));
    s[i] = si;
    rho = expf(si);
    rho_rt[i] = rho;
    float ratio = rho / P->rho0;
    float p = (P->c0 * P->c0) * P->rho0 * (powf(ratio, P->gammaEOS) - 1.0f) / P->gammaEOS;
    press[i] = fmaxf(p, 0.0f);
  }
  /* forces :215-272 */
  const float gyv = -(P->useGrav ? P->gravity : 0.f);
  for (int i = 0; i < N; ++i) {
    const float xi = pos[2 * i], yi = pos[2 * i + 1], vxi = vel[2 * i], vyi = vel[2 * i + 1];
    const float rhoi = rho_rt[i], pi = press[i];
    const int gx = grid_c(xi, cell, Gx), gy = grid_c(yi, cell, Gy);
    float ax = 0.f, ay = 0.f;
    for (int oy = -1; oy <= 1; ++oy)
      for (int ox = -1; ox <= 1; ++ox) {
        int cx = gx + ox, cy = gy + oy;
        if ((unsigned)cx >= (unsigned)Gx || (unsigned)cy >= (unsigned)Gy) continue;
        int c = cy * Gx + cx;
        for (int q = cs[c]; q < cs[c + 1]; ++q) {
          int j = (int)vs[q];
          if (j == i) continue;
          float rx = xi - pos[2 * j], ry = yi - pos[2 * j + 1];
          float r2 = rx * rx + ry * ry;
          if (r2 >= twoh2 || r2 <= 1e-16f) continue;
          float r = sqrtf(r2), gwx, gwy;
          gradW_cubic(rx, ry, r, h, &gwx, &gwy);
          float rhoj = rho_rt[j], pj = press[j];
          float common = -mass * (pi / (rhoi * rhoi) + pj / (rhoj * rhoj));
          ax += common * gwx;
          ay += common * gwy;
          if (P->useVisc) {
            float vx = vxi - vel[2 * j], vy = vyi - vel[2 * j + 1];
            float dot = vx * rx + vy * ry;
            if (dot < 0.f) {
              float mu = (h * dot) / (r2 + 0.01f * h * h);
              float rhoBar = 0.5f * (rhoi + rhoj);
              float Pi_ij = (-P->viscAlpha * P->c0 * mu) / rhoBar;
              ax += -mass * Pi_ij * gwx;
              ay += -mass * Pi_ij * gwy;
            }
          }
        }
      }
    if (P->useGrav) { ax += 0.f; ay += gyv; }
    acc[2 * i] = ax;
    acc[2 * i + 1] = ay;
  }
  /* integrate :324-355 */
  for (int i = 0; i < N; ++i) {
    float vx = vel[2 * i], vy = vel[2 * i + 1], x = pos[2 * i], y = pos[2 * i + 1];
    vx += acc[2 * i] * dt; vy += acc[2 * i + 1] * dt;
    x += vx * dt; y += vy * dt;
    const float e = 0.2f;
    if (x < 0.f) { x = 0.f; vx = -e * vx; }
    if (x > P->boxX) { x = P->boxX; vx = -e * vx; }
    if (y < 0.f) { y = 0.f; vy = -e * vy; }
    if (y > P->boxY) { y = P->boxY; vy = -e * vy; }
    pos[2 * i] = x; pos[2 * i + 1] = y; vel[2 * i] = vx; vel[2 * i + 1] = vy;
  }
  /* XSPH :274-322 — updated positions/velocities, cell structure from before the integration */
  if (P->useXSPH && P->xsphEps > 0.f) {
    for (int i = 0; i < N; ++i) {
      const float xi = pos[2 * i], yi = pos[2 * i + 1], vxi = vel[2 * i], vyi = vel[2 * i + 1];
      const int gx = grid_c(xi, cell, Gx), gy = grid_c(yi, cell, Gy);
      float dx = 0.f, dy = 0.f;
      for (int oy = -1; oy <= 1; ++oy)
        for (int ox = -1; ox <= 1; ++ox) {
          int cx = gx + ox, cy = gy + oy;
          if ((unsigned)cx >= (unsigned)Gx || (unsigned)cy >= (unsigned)Gy) continue;
          int c = cy * Gx + cx;
          for (int q = cs[c]; q < cs[c + 1]; ++q) {
            int j = (int)vs[q];
            if (j == i) continue;
            float rx = xi - pos[2 * j], ry = yi - pos[2 * j + 1];
            float r2 = rx * rx + ry * ry;
            if (r2 >= twoh2) continue;
            float w = W_cubic(sqrtf(r2), h);
            float rhoBar = 0.5f * (rho_rt[i] + rho_rt[j]);
            dx += (mass / rhoBar) * (vel[2 * j] - vxi) * w;
            dy += (mass / rhoBar) * (vel[2 * j + 1] - vyi) * w;
          }
        }
      acc[2 * i] = P->xsphEps * dx;
      acc[2 * i + 1] = P->xsphEps * dy;
    }
    for (int i = 0; i < N; ++i) { vel[2 * i] += acc[2 * i]; vel[2 * i + 1] += acc[2 * i + 1]; }
  }
  /* rain :377-392, :706-716 — sequential order: the highest spawn index wins a collision */
  if (P->rain) {
    clk->rain_carry += 0.02f * P->N * dt;
    int nspawn = (int)clk->rain_carry;
    clk->rain_carry -= nspawn;
    const unsigned seed = (unsigned)(P->seed + clk->step);
    for (int k = 0; k < nspawn; ++k) {
      unsigned sd = seed ^ ((unsigned)k * 1664525u + 1013904223u);
      sd = sd * 1664525u + 1013904223u;
      float rx = (sd & 0x00FFFFFF) / 16777216.f;
      sd = sd * 1664525u + 1013904223u;
      float x = rx * (P->boxX * 0.8f) + 0.1f * P->boxX;
      float ry = (sd & 0x00FFFFFF) / 16777216.f;
      float y = P->boxY * (0.9f + 0.08f * ry);
      int i = (int)(sd % (unsigned)N);
      pos[2 * i] = x; pos[2 * i + 1] = y; vel[2 * i] = 0.f; vel[2 * i + 1] = -0.5f * P->c0;
    }
  }
  free(rho_rt); free(cs); free(vs); free(ks);
}

/* nframes x the doStep block :663-722 */
void oracle_sph_run(const oracle_sph_params *P, float *pos, float *vel, float *acc, float *s, float *press,
                    int nframes, oracle_sph_clock *clk) {
  float mass, h, cell; int Gx, Gy;
  oracle_sph_derived(P, &mass, &h, &cell, &Gx, &Gy);
  for (int f = 0; f < nframes; ++f) {
    float dTau_accum = 0.f;
    const int K = P->viscSub > 0 ? P->viscSub : 1;
    const float dt_try = clk->t * P->dTau;
    const float dt_cfl = P->CFL * h / (P->c0 * (1.0f + 2.0f * P->viscAlpha));
    const float dt_sub = fminf(dt_try, dt_cfl) / K;
    for (int k = 0; k < K; ++k) {
      oracle_sph_substep(P, pos, vel, acc, s, press, dt_sub, clk);
      dTau_accum += dt_sub / fmaxf(clk->t, 1e-9f);
      clk->t = P->t0 * expf(clk->tau + dTau_accum);
    }
    clk->tau += dTau_accum;
    clk->step++;
  }
}

/* k_clear_grid + k_rasterize :357-374 (plain C division; the reference is built with
 * -use_fast_math, so a particle whose scaled coordinate is within an ulp of an integer may land in
 * the neighbouring raster cell there) */
void oracle_sph_rasterize(const float *pos, int N, int W, int H, float boxX, float boxY, int *grid2) {
  for (int i = 0; i < W * 2 * H; ++i) grid2[i] = 0;
  for (int i = 0; i < N; ++i) {
    float px = pos[2 * i], py = pos[2 * i + 1];
    int cx = (int)(px / boxX * (W - 1));
    int sy = (int)((boxY - py) / boxY * (2 * H - 1));
    if ((unsigned)cx < (unsigned)W && (unsigned)sy < (unsigned)(2 * H)) grid2[sy * W + cx] += 1;
  }
}
