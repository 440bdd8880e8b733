// burgers.cu — 2-D viscous Burgers (Rusanov / optional MUSCL-minmod convection, explicit Laplacian
// viscosity, log-time clock) update path for sm_100a.  SURVEY.md 8(f) rank 3.  Replaces the
// per-step host sequence `do_step` of the reference `tau_burgers` (tau_burgers.cu:677-718):
//     wavespeed_block_max -> D2H of all block maxima + host max + dt_eff -> flux_x_kernel ->
//     flux_y_kernel -> update_convective -> K x viscosity_step;   tau += dtau; t *= expf(dtau)
//
// What is restructured:
//   * state is stored as phi = asinh(u/u0) (:11) and the reference re-evaluates u0*sinhf(phi) at
//     every use: ~14 sinhf per cell-step without MUSCL.  Here a CTA stages its 32x16 tile plus halo
//     in shared memory and decodes every tile cell ONCE per kernel.
//   * the four flux planes (Fu_x, Fv_x, Gu_y, Gv_y: 16 B/cell written and re-read twice) never
//     exist: each face flux of the tile is computed once into shared memory and the update reads its
//     four faces from there.  One kernel instead of three.
//   * dt never visits the host: the last kernel of a step reduces max(|u|/dx + |v|/dy) of the state
//     it writes (warp shuffle + one integer atomicMax per warp; max is exactly associative, so the
//     value equals the reference's two-level reduction) and the next step's first kernel evaluates
//     dt_eff = min(t * dtau, CFL / max) (:687-691) itself; (t, tau) live in device memory.
//   * viscosity_step (:490-527) updates phi IN PLACE while neighbouring threads still read it — a
//     data race whose result depends on block scheduling.  Here it is the Jacobi update the kernel's
//     own comment describes ("explicit Laplacian on (u,v)"): read the old buffer, write the other.
//     Parity against the reference kernels is therefore exact-to-round-off for nu = 0 and for the
//     convective part, and tolerance-level for nu > 0; the Cole-Hopf exact solution (:720-737) is the
//     independent check.
// Arithmetic keeps the reference's expression trees and its transcendental calls by name; this TU
// is compiled with the reference's -use_fast_math (reference Makefile:75-76).
#include "common.cuh"
#include "../../include/tau_b200.h"

#include <math.h>
#include <new>
#include <vector>

namespace {

constexpr int B_TX = 32, B_TY = 16;         // cells per tile
constexpr int B_H = 2;                      // halo (MUSCL needs i-1 .. i+2 for face i+1/2)
constexpr int B_SX = B_TX + 2 * B_H, B_SY = B_TY + 2 * B_H;
constexpr int B_THREADS = 256;
constexpr int B_NFX = (B_TX + 1) * B_TY;    // x-faces of a tile (incl. its left edge)
constexpr int B_NFY = B_TX * (B_TY + 1);    // y-faces (incl. its bottom edge)

struct BPar {
  int nx, ny;
  float dx, dy, nu, u0, CFL, dtau;
  int muscl, oneD;  // oneD = Cole-Hopf harness: no y-flux, no y-Laplacian (:706, :712-716)
};

struct BClock {       // device-resident step control
  float t, tau;       // log-time clock (:675, :768-769)
  float dt_last;      // dt_eff of the most recent step
  float smax[2];      // max wavespeed slots: step s reads [s&1], the step's last kernel fills [(s+1)&1]
};

__device__ __forceinline__ int wrap(int i, int n) {  // :92-97, callers are at most one period out
  if (i < 0) i += n;
  else if (i >= n) i -= n;
  if ((unsigned)i >= (unsigned)n) {
    i %= n;
    if (i < 0) i += n;
  }
  return i;
}
__device__ __forceinline__ float minmod(float a, float b) {  // :331-333
  return (a * b <= 0.0f) ? 0.0f : copysignf(fminf(fabsf(a), fabsf(b)), a);
}

// dt_eff of :687-691 from the device-resident wavespeed max
__device__ __forceinline__ float step_dt(const BPar &P, const BClock *clk, int slot) {
  const float smax = fmaxf(1e-12f, clk->smax[slot]);
  const float dt_cfl = P.CFL / smax;
  return fminf(clk->t * P.dtau, dt_cfl);
}

// Stage phi_u, phi_v of the tile + halo (periodic wrap) and their decoded velocities.
__device__ __forceinline__ void stage_tile(const BPar &P, const float *__restrict__ phi_u,
                                           const float *__restrict__ phi_v, int bx0, int by0, int halo,
                                           float *s_pu, float *s_pv, float *s_u, float *s_v) {
  const int sx = B_TX + 2 * halo, sy = B_TY + 2 * halo;
  for (int t = threadIdx.x; t < sx * sy; t += B_THREADS) {
    const int ly = t / sx, lx = t - ly * sx;
    const int gi = wrap(bx0 + lx - halo, P.nx), gj = wrap(by0 + ly - halo, P.ny);
    const size_t id = (size_t)gj * P.nx + gi;
    const float pu = phi_u[id], pv = phi_v[id];
    if (s_pu) { s_pu[t] = pu; s_pv[t] = pv; }
    s_u[t] = P.u0 * sinhf(pu);
    s_v[t] = P.u0 * sinhf(pv);
  }
}

// ---- convection: flux_x_kernel + flux_y_kernel + update_convective (:364-487) in one kernel ----------
__global__ void __launch_bounds__(B_THREADS)
burgers_convect(const BPar P, const float *__restrict__ phi_u, const float *__restrict__ phi_v,
                float *__restrict__ out_u, float *__restrict__ out_v, BClock *__restrict__ clk, int slot) {
  __shared__ float s_pu[B_SX * B_SY], s_pv[B_SX * B_SY], s_u[B_SX * B_SY], s_v[B_SX * B_SY];
  __shared__ float s_Fu[B_NFX], s_Fv[B_NFX], s_Gu[B_NFY], s_Gv[B_NFY];
  const int bx0 = blockIdx.x * B_TX, by0 = blockIdx.y * B_TY;
  const float dt = step_dt(P, clk, slot);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    clk->dt_last = dt;
    clk->smax[slot ^ 1] = 0.f;  // the slot this step's last kernel reduces into
  }
  stage_tile(P, phi_u, phi_v, bx0, by0, B_H, s_pu, s_pv, s_u, s_v);
  __syncthreads();

  // x-faces: face f sits on the RIGHT of tile column fx-1 (fx = 0 is the tile's left edge)
  for (int f = threadIdx.x; f < B_NFX; f += B_THREADS) {
    const int fy = f / (B_TX + 1), fx = f - fy * (B_TX + 1);
    const int cL = (fy + B_H) * B_SX + (fx - 1 + B_H), cR = cL + 1;
    float uL = s_u[cL], vL = s_v[cL], uR = s_u[cR], vR = s_v[cR];
    if (P.muscl) {  // :378-395 — slopes on phi, face states decoded afterwards
      float pUL = s_pu[cL], pUR = s_pu[cR], pVL = s_pv[cL], pVR = s_pv[cR];
      const float sUL = 0.5f * minmod(pUL - s_pu[cL - 1], pUR - pUL);
      const float sUR = 0.5f * minmod(s_pu[cR + 1] - pUR, pUR - pUL);
      const float sVL = 0.5f * minmod(pVL - s_pv[cL - 1], pVR - pVL);
      const float sVR = 0.5f * minmod(s_pv[cR + 1] - pVR, pVR - pVL);
      pUL = pUL + sUL; pUR = pUR - sUR; pVL = pVL + sVL; pVR = pVR - sVR;
      uL = P.u0 * sinhf(pUL); vL = P.u0 * sinhf(pVL);
      uR = P.u0 * sinhf(pUR); vR = P.u0 * sinhf(pVR);
    }
    const float FL_u = 0.5f * uL * uL, FL_v = uL * vL, FR_u = 0.5f * uR * uR, FR_v = uR * vR;
    const float a = fmaxf(fabsf(uL), fabsf(uR));
    s_Fu[f] = 0.5f * (FL_u + FR_u) - 0.5f * a * (uR - uL);
    s_Fv[f] = 0.5f * (FL_v + FR_v) - 0.5f * a * (vR - vL);
  }
  if (!P.oneD) {  // y-faces: face f sits on TOP of tile row fy-1 (:411-455)
    for (int f = threadIdx.x; f < B_NFY; f += B_THREADS) {
      const int fy = f / B_TX, fx = f - fy * B_TX;
      const int cB = (fy - 1 + B_H) * B_SX + (fx + B_H), cT = cB + B_SX;
      float uB = s_u[cB], vB = s_v[cB], uT = s_u[cT], vT = s_v[cT];
      if (P.muscl) {
        float pUB = s_pu[cB], pUT = s_pu[cT], pVB = s_pv[cB], pVT = s_pv[cT];
        const float sUB = 0.5f * minmod(pUB - s_pu[cB - B_SX], pUT - pUB);
        const float sUT = 0.5f * minmod(s_pu[cT + B_SX] - pUT, pUT - pUB);
        const float sVB = 0.5f * minmod(pVB - s_pv[cB - B_SX], pVT - pVB);
        const float sVT = 0.5f * minmod(s_pv[cT + B_SX] - pVT, pVT - pVB);
        pUB = pUB + sUB; pUT = pUT - sUT; pVB = pVB + sVB; pVT = pVT - sVT;
        uB = P.u0 * sinhf(pUB); vB = P.u0 * sinhf(pVB);
        uT = P.u0 * sinhf(pUT); vT = P.u0 * sinhf(pVT);
      }
      const float GL_u = uB * vB, GL_v = 0.5f * vB * vB, GR_u = uT * vT, GR_v = 0.5f * vT * vT;
      const float a = fmaxf(fabsf(vB), fabsf(vT));
      s_Gu[f] = 0.5f * (GL_u + GR_u) - 0.5f * a * (uT - uB);
      s_Gv[f] = 0.5f * (GL_v + GR_v) - 0.5f * a * (vT - vB);
    }
  }
  __syncthreads();

  // update_convective :458-487
  const float invdx = 1.0f / P.dx, invdy = P.oneD ? 0.0f : (1.0f / P.dy);
  for (int t = threadIdx.x; t < B_TX * B_TY; t += B_THREADS) {
    const int ly = t / B_TX, lx = t - ly * B_TX;
    const int i = bx0 + lx, j = by0 + ly;
    if (i >= P.nx || j >= P.ny) continue;
    const int c = (ly + B_H) * B_SX + lx + B_H;
    float u = s_u[c], v = s_v[c];
    const float dFx_u = s_Fu[ly * (B_TX + 1) + lx + 1] - s_Fu[ly * (B_TX + 1) + lx];
    const float dFx_v = s_Fv[ly * (B_TX + 1) + lx + 1] - s_Fv[ly * (B_TX + 1) + lx];
    const float dGy_u = P.oneD ? 0.0f : (s_Gu[(ly + 1) * B_TX + lx] - s_Gu[ly * B_TX + lx]);
    const float dGy_v = P.oneD ? 0.0f : (s_Gv[(ly + 1) * B_TX + lx] - s_Gv[ly * B_TX + lx]);
    u -= dt * (dFx_u * invdx + dGy_u * invdy);
    v -= dt * (dFx_v * invdx + dGy_v * invdy);
    const size_t id = (size_t)j * P.nx + i;
    out_u[id] = asinhf(u / P.u0);
    out_v[id] = asinhf(v / P.u0);
  }
}

// ---- viscosity_step :490-527 as a Jacobi update (see header); the step's last sub-step also reduces
// the wavespeed max of the state it writes (wavespeed_block_max :337-361) and advances the clock ----------
__global__ void __launch_bounds__(B_THREADS)
burgers_viscosity(const BPar P, const float *__restrict__ phi_u, const float *__restrict__ phi_v,
                  float *__restrict__ out_u, float *__restrict__ out_v, BClock *__restrict__ clk, int slot,
                  int nsub, int last) {
  constexpr int SX = B_TX + 2, SY = B_TY + 2;
  __shared__ float s_u[SX * SY], s_v[SX * SY];
  const int bx0 = blockIdx.x * B_TX, by0 = blockIdx.y * B_TY;
  const float sub = clk->dt_last / (float)nsub;  // :711
  stage_tile(P, phi_u, phi_v, bx0, by0, 1, nullptr, nullptr, s_u, s_v);
  __syncthreads();
  const float invdx2 = 1.0f / (P.dx * P.dx), invdy2 = P.oneD ? 0.0f : (1.0f / (P.dy * P.dy));
  const float invdx = 1.0f / P.dx, invdy = (P.ny > 1 ? 1.0f / P.dy : 0.0f);  // :681-682
  float smax = 0.f;
  for (int t = threadIdx.x; t < B_TX * B_TY; t += B_THREADS) {
    const int ly = t / B_TX, lx = t - ly * B_TX;
    const int i = bx0 + lx, j = by0 + ly;
    if (i >= P.nx || j >= P.ny) continue;
    const int c = (ly + 1) * SX + lx + 1;
    const float cu = s_u[c], cv = s_v[c];
    const float lap_u = (s_u[c + 1] - 2.0f * cu + s_u[c - 1]) * invdx2 + (s_u[c + SX] - 2.0f * cu + s_u[c - SX]) * invdy2;
    const float lap_v = (s_v[c + 1] - 2.0f * cv + s_v[c - 1]) * invdx2 + (s_v[c + SX] - 2.0f * cv + s_v[c - SX]) * invdy2;
    const float u = cu + P.nu * sub * lap_u;
    const float v = cv + P.nu * sub * lap_v;
    const size_t id = (size_t)j * P.nx + i;
    const float pu = asinhf(u / P.u0), pv = asinhf(v / P.u0);
    out_u[id] = pu;
    out_v[id] = pv;
    if (last) {  // the next step's CFL scan reads the ENCODED state (:346-347)
      const float un = P.u0 * sinhf(pu), vn = P.u0 * sinhf(pv);
      smax = fmaxf(smax, fabsf(un) * invdx + fabsf(vn) * invdy);
    }
  }
  if (last) {
    smax = tau::warp_max(smax);
    if ((threadIdx.x & 31) == 0 && smax > 0.f) tau::atomic_max_nonneg(&clk->smax[slot ^ 1], smax);
    // tau += dtau; t *= expf(dtau) (:768-769): nothing else in this kernel reads the clock
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
      clk->tau += P.dtau;
      clk->t *= expf(P.dtau);
    }
  }
}

// first CFL scan after init / upload (wavespeed_block_max :337-361)
__global__ void burgers_wavespeed(const BPar P, const float *__restrict__ phi_u, const float *__restrict__ phi_v,
                                  BClock *clk, int slot) {
  const size_t n = (size_t)P.nx * P.ny;
  const float invdx = 1.0f / P.dx, invdy = (P.ny > 1 ? 1.0f / P.dy : 0.0f);
  float smax = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float u = P.u0 * sinhf(phi_u[i]), v = P.u0 * sinhf(phi_v[i]);
    smax = fmaxf(smax, fabsf(u) * invdx + fabsf(v) * invdy);
  }
  smax = tau::warp_max(smax);
  if ((threadIdx.x & 31) == 0 && smax > 0.f) tau::atomic_max_nonneg(&clk->smax[slot], smax);
}

}  // namespace

struct tau_burgers {
  tau_burgers_params p;
  int device;
  cudaStream_t stream;
  bool own_stream;
  float *phi[2][2];  // [buffer][u, v]
  BClock *clk;
  int cur;
  long long steps, launches;
  bool have_state;
  cudaEvent_t ev0, ev1;
  bool timed;
};

namespace {
BPar make_par(const tau_burgers *h) {
  BPar P;
  const tau_burgers_params &p = h->p;
  P.nx = p.nx; P.ny = p.colehopf ? 1 : p.ny;
  P.dx = p.dx; P.dy = p.dy; P.nu = p.nu; P.u0 = p.u0; P.CFL = p.CFL; P.dtau = p.dtau;
  P.muscl = p.muscl ? 1 : 0;
  P.oneD = p.colehopf ? 1 : 0;
  return P;
}
int state_changed(tau_burgers *h) {
  const BPar P = make_par(h);
  const int slot = (int)(h->steps & 1);
  TAU_CUDA(cudaMemsetAsync(h->clk->smax, 0, 2 * sizeof(float), h->stream));
  burgers_wavespeed<<<148 * 4, 256, 0, h->stream>>>(P, h->phi[h->cur][0], h->phi[h->cur][1], h->clk, slot);
  h->launches++;
  TAU_CUDA(cudaGetLastError());
  h->have_state = true;
  return TAU_OK;
}
}  // namespace

extern "C" {

// struct Params :53-90 (simulation fields)
void tau_burgers_default_params(tau_burgers_params *p) {
  memset(p, 0, sizeof(*p));
  p->nx = 512; p->ny = 512; p->dx = 1.0f; p->dy = 1.0f;
  p->nu = 0.1f; p->u0 = 1.0f;
  p->amp = 1.0f; p->bsig = 16.0f; p->swirl = 10.0f; p->rc = 40.0f; p->offx = 0.0f; p->offy = 0.0f; p->asym = 0.0f;
  p->CFL = 0.45f; p->tau0 = 0.0f; p->t0 = 1.0f; p->dtau = 1.0f;
  p->muscl = 0; p->visc_substeps = 1;
  p->colehopf = 0; p->ck = 4; p->ca = 0.5f;
}

// initialize_host :250-304 (host code in the reference too: same libm, same values)
void tau_burgers_init_host(const tau_burgers_params *P, float *phi_u, float *phi_v) {
  const int nx = P->nx, ny = P->colehopf ? 1 : P->ny;
  for (size_t k = 0; k < (size_t)nx * ny; ++k) phi_u[k] = phi_v[k] = 0.0f;
  if (P->colehopf) {
    const float Lx = P->dx * nx;
    const float k = 2.0f * (float)M_PI * P->ck / Lx;
    for (int i = 0; i < nx; ++i) {
      const float x = (i + 0.5f) * P->dx;
      const float denom = 1.0f + P->ca * cosf(k * x);
      const float u = (denom != 0.0f) ? (2.0f * P->nu * P->ca * k * sinf(k * x) / denom) : 0.0f;
      const float phi = asinhf(u / P->u0);
      for (int j = 0; j < ny; ++j) {
        phi_u[(size_t)j * nx + i] = phi;
        phi_v[(size_t)j * nx + i] = 0.0f;
      }
    }
    return;
  }
  const float cx = 0.5f * nx + P->offx, cy = 0.5f * ny + P->offy;
  const float sig2 = P->bsig * P->bsig;
  const float rc = P->rc * fminf(P->dx, P->dy);
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      const float dx = i - cx, dy = j - cy;
      const float r2 = (dx * dx + dy * dy) / fmaxf(sig2, 1e-6f);
      const float theta = atan2f(dy, dx);
      const float mod = 1.0f + P->asym * cosf(theta);
      const float rx = dx * P->dx, ry = dy * P->dy;
      const float r = sqrtf(rx * rx + ry * ry);
      const float u_theta = (r > 0.0f) ? (P->swirl * r * expf(-0.5f * (r / rc) * (r / rc))) : 0.0f;
      float u = (r > 0.0f) ? (-u_theta * (ry / r)) : 0.0f;
      float v = (r > 0.0f) ? (u_theta * (rx / r)) : 0.0f;
      const float g = P->amp * mod * expf(-0.5f * r2);
      u += 0.5f * g;
      v += -0.5f * g;
      phi_u[(size_t)j * nx + i] = asinhf(u / P->u0);
      phi_v[(size_t)j * nx + i] = asinhf(v / P->u0);
    }
}

int tau_burgers_create(const tau_burgers_params *p, int device, void *stream, tau_burgers **out) {
  TAU_REQUIRE(p && out, "tau_burgers_create: null argument");
  TAU_REQUIRE(p->nx >= 1 && p->ny >= 1, "tau_burgers_create: bad grid %d x %d", p->nx, p->ny);
  TAU_REQUIRE(p->dx > 0.f && p->dy > 0.f && p->u0 > 0.f, "tau_burgers_create: dx, dy, u0 must be > 0");
  if (tau_device_count() <= 0) {
    tau_set_error("tau_burgers_create: no CUDA device (this library has no CPU fallback)");
    return TAU_ERR_NODEV;
  }
  TAU_CUDA(cudaSetDevice(device));
  tau_burgers *h = new (std::nothrow) tau_burgers();
  if (!h) return TAU_ERR_NOMEM;
  memset(h, 0, sizeof(*h));
  h->p = *p;
  if (h->p.colehopf) h->p.ny = 1;  // :649-650
  h->device = device;
  if (stream) {
    h->stream = (cudaStream_t)stream;
    h->own_stream = false;
  } else {
    TAU_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  const size_t n = (size_t)h->p.nx * h->p.ny;
  for (int b = 0; b < 2; ++b)
    for (int f = 0; f < 2; ++f) TAU_CUDA(cudaMalloc(&h->phi[b][f], n * sizeof(float)));
  TAU_CUDA(cudaMalloc(&h->clk, sizeof(BClock)));
  TAU_CUDA(cudaEventCreate(&h->ev0));
  TAU_CUDA(cudaEventCreate(&h->ev1));
  *out = h;
  return TAU_OK;
}

// phi_u, phi_v: ny x nx host planes (index j*nx+i, :98-100); clock2 = {t, tau} or NULL for (t0, tau0)
int tau_burgers_upload(tau_burgers *h, const float *phi_u, const float *phi_v, const float *clock2) {
  TAU_REQUIRE(h && phi_u && phi_v, "tau_burgers_upload: null argument");
  TAU_CUDA(cudaSetDevice(h->device));
  const size_t n = (size_t)h->p.nx * h->p.ny;
  TAU_CUDA(cudaMemcpyAsync(h->phi[h->cur][0], phi_u, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  TAU_CUDA(cudaMemcpyAsync(h->phi[h->cur][1], phi_v, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  BClock c;
  memset(&c, 0, sizeof(c));
  c.t = clock2 ? clock2[0] : h->p.t0;
  c.tau = clock2 ? clock2[1] : h->p.tau0;
  TAU_CUDA(cudaMemcpyAsync(h->clk, &c, sizeof(c), cudaMemcpyHostToDevice, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));  // `c` is on the stack
  int rc = state_changed(h);
  if (rc) return rc;
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

// initialize_host + H2D :652-661
int tau_burgers_init(tau_burgers *h) {
  TAU_REQUIRE(h, "tau_burgers_init: null handle");
  const size_t n = (size_t)h->p.nx * h->p.ny;
  std::vector<float> u(n), v(n);
  tau_burgers_init_host(&h->p, u.data(), v.data());
  h->steps = 0;
  return tau_burgers_upload(h, u.data(), v.data(), nullptr);
}

// THE hot path: nsteps x { do_step :677-718; tau += dtau; t *= expf(dtau) :768-769 }, no host sync
int tau_burgers_step(tau_burgers *h, int nsteps) {
  TAU_REQUIRE(h && nsteps >= 0, "tau_burgers_step: bad argument");
  TAU_REQUIRE(h->have_state, "tau_burgers_step: no state (call tau_burgers_init or tau_burgers_upload)");
  TAU_CUDA(cudaSetDevice(h->device));
  const BPar P = make_par(h);
  const dim3 grid((P.nx + B_TX - 1) / B_TX, (P.ny + B_TY - 1) / B_TY);
  const int K = h->p.visc_substeps > 0 ? h->p.visc_substeps : 1;
  TAU_CUDA(cudaEventRecord(h->ev0, h->stream));
  for (int s = 0; s < nsteps; ++s) {
    const int slot = (int)(h->steps & 1);
    int a = h->cur;
    burgers_convect<<<grid, B_THREADS, 0, h->stream>>>(P, h->phi[a][0], h->phi[a][1], h->phi[a ^ 1][0],
                                                       h->phi[a ^ 1][1], h->clk, slot);
    a ^= 1;
    for (int k = 0; k < K; ++k) {
      burgers_viscosity<<<grid, B_THREADS, 0, h->stream>>>(P, h->phi[a][0], h->phi[a][1], h->phi[a ^ 1][0],
                                                           h->phi[a ^ 1][1], h->clk, slot, K, k == K - 1);
      a ^= 1;
    }
    h->launches += 1 + K;
    h->cur = a;
    h->steps++;
  }
  TAU_CUDA(cudaGetLastError());
  TAU_CUDA(cudaEventRecord(h->ev1, h->stream));
  h->timed = true;
  return TAU_OK;
}

int tau_burgers_clock(tau_burgers *h, float *t, float *tau, float *dt_last) {
  TAU_REQUIRE(h, "tau_burgers_clock: null handle");
  BClock c;
  TAU_CUDA(cudaMemcpyAsync(&c, h->clk, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  if (t) *t = c.t;
  if (tau) *tau = c.tau;
  if (dt_last) *dt_last = c.dt_last;
  return TAU_OK;
}

int tau_burgers_download(tau_burgers *h, float *phi_u, float *phi_v) {
  TAU_REQUIRE(h, "tau_burgers_download: null handle");
  TAU_CUDA(cudaSetDevice(h->device));
  const size_t n = (size_t)h->p.nx * h->p.ny;
  if (phi_u) TAU_CUDA(cudaMemcpyAsync(phi_u, h->phi[h->cur][0], n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  if (phi_v) TAU_CUDA(cudaMemcpyAsync(phi_v, h->phi[h->cur][1], n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

// colehopf_relL2 :720-737: relative L2 error of row 0 against the exact 1-D solution at time t
int tau_burgers_colehopf_error(tau_burgers *h, double *rel_l2) {
  TAU_REQUIRE(h && rel_l2, "tau_burgers_colehopf_error: null argument");
  TAU_REQUIRE(h->p.colehopf, "tau_burgers_colehopf_error: the handle was not created with colehopf = 1");
  const tau_burgers_params &P = h->p;
  std::vector<float> pu((size_t)P.nx * P.ny);
  int rc = tau_burgers_download(h, pu.data(), nullptr);
  if (rc) return rc;
  float t_now;
  rc = tau_burgers_clock(h, &t_now, nullptr, nullptr);
  if (rc) return rc;
  const float Lx = P.dx * P.nx;
  const float k = 2.0f * (float)M_PI * P.ck / Lx;
  const float decay = expf(-P.nu * k * k * t_now);
  double num = 0.0, den = 0.0;
  for (int i = 0; i < P.nx; ++i) {
    const float x = (i + 0.5f) * P.dx;
    const float u_ex = (2.0f * P.nu * P.ca * k * decay * sinf(k * x)) / (1.0f + P.ca * decay * cosf(k * x));
    const double u_num = P.u0 * sinh((double)pu[i]);
    const double diff = u_num - u_ex;
    num += diff * diff;
    den += (double)u_ex * u_ex;
  }
  *rel_l2 = (den > 0.0) ? sqrt(num / den) : sqrt(num);
  return TAU_OK;
}

int tau_burgers_sync(tau_burgers *h) {
  TAU_REQUIRE(h, "tau_burgers_sync: null handle");
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}
long long tau_burgers_steps_done(tau_burgers *h) { return h ? h->steps : -1; }
long long tau_burgers_launch_count(tau_burgers *h) { return h ? h->launches : -1; }
int tau_burgers_last_step_ms(tau_burgers *h, float *ms) {
  TAU_REQUIRE(h && ms, "tau_burgers_last_step_ms: null argument");
  TAU_REQUIRE(h->timed, "tau_burgers_last_step_ms: no step has been timed yet");
  TAU_CUDA(cudaEventSynchronize(h->ev1));
  TAU_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return TAU_OK;
}
int tau_burgers_destroy(tau_burgers *h) {
  if (!h) return TAU_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaFree(h->clk);
  for (int b = 1; b >= 0; --b)
    for (int f = 1; f >= 0; --f) cudaFree(h->phi[b][f]);
  cudaEventDestroy(h->ev1);
  cudaEventDestroy(h->ev0);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return TAU_OK;
}

}  // extern "C"
