// shallow_water.cu — 2-D shallow water (HLL fluxes, log-depth state, log-time clock, explicit viscosity
// on u,v) update path for sm_100a.  SURVEY.md 8(f) rank 3, second half.  Replaces the per-step host
// sequence `do_step` of the reference `tau_sw` (tau_shallow_water.cu:669-705):
//     wavespeed_block_max -> D2H of all block maxima + host max + dt_eff -> flux_x_kernel ->
//     flux_y_kernel -> update_kernel -> [viscosity_uv];   tau += dtau; t *= expf(dtau)
//
// STATUS: written after the round-1 GPU budget was spent.  It compiles, its CPU oracle is pinned by
// invariants (tests/test_oracle_cpu.py), the reference-kernel driver builds — but it HAS NOT RUN ON
// HARDWARE yet; its GPU tests (tests/test_sw_gpu.py) are opt-in (TAU_TEST_SW=1) until it has.
//
// Same restructuring as burgers.cu: a CTA stages its 32x16 tile + one-cell halo in shared memory and
// evaluates h = expf(sigma) and the wave speed ONCE per tile cell (the reference: ~7 expf + 5 sqrtf
// per cell-step); every HLL face flux of the tile is computed once into shared memory (the six flux
// planes, 24 B/cell written and re-read, never exist); dt = min(t dtau, CFL min(dx,dy) / cmax) is
// evaluated on the device from a max reduced by the previous step's last kernel (warp shuffle +
// integer atomicMax); the log-time clock lives in device memory.  viscosity_uv (:516-551) updates u,v
// in place while neighbouring threads read them (a data race); here it is the Jacobi update.
// Arithmetic keeps the reference's expression trees; compiled with the reference's -use_fast_math
// (reference Makefile:90-91).
#include "common.cuh"
#include "../../include/tau_b200.h"

#include <math.h>
#include <new>
#include <vector>

namespace {

constexpr int S_TX = 32, S_TY = 16;
constexpr int S_SX = S_TX + 2, S_SY = S_TY + 2;
constexpr int S_THREADS = 256;
constexpr int S_NFX = (S_TX + 1) * S_TY;
constexpr int S_NFY = S_TX * (S_TY + 1);

struct SPar {
  int nx, ny;
  float dx, dy, g, nu, CFL, dtau;
  float edtau;  // expf(dtau) evaluated on the host like the reference's loop (:768), so the clock is bit-equal
};
// Everything a step READS here stays untouched for the whole step (blocks of one kernel run in any
// order): step s reads t[s&1], tau[s&1], cmax[s%3]; its last kernel writes t[(s+1)&1], tau[(s+1)&1] and
// max-reduces into cmax[(s+1)%3]; its first kernel clears cmax[(s+2)%3] (last read by step s-1).
struct SClock {
  float t[2], tau[2], dt_last;
  float cmax[3];
};

__device__ __forceinline__ int wrap(int i, int n) {  // :91-96, callers are at most one period out
  if (i < 0) i += n;
  else if (i >= n) i -= n;
  if ((unsigned)i >= (unsigned)n) {
    i %= n;
    if (i < 0) i += n;
  }
  return i;
}

// HLL flux through a face with normal velocity q and tangential velocity r (x faces: q = u, r = v; y faces:
// q = v, r = u): fluxes of (h, h q, h r).  hll_x :322-352 and hll_y :355-385 are this one function up to the
// association of the triple product in the tangential momentum flux — (h u) v in both, i.e. (h q) r on x
// faces and (h r) q on y faces — which TANGENTIAL_FIRST keeps, so that both round like the reference.
template <bool TANGENTIAL_FIRST>
__device__ __forceinline__ void hll_flux(float hL, float qL, float rL, float hR, float qR, float rR, float g,
                                         float &Fh, float &Fq, float &Fr) {
  const float cL = sqrtf(g * hL), cR = sqrtf(g * hR);
  const float sL = fminf(qL - cL, qR - cR), sR = fmaxf(qL + cL, qR + cR);
  const float hqL = hL * qL, hqR = hR * qR, hrL = hL * rL, hrR = hR * rR;   // conserved momenta
  const float FqL = hqL * qL + 0.5f * g * hL * hL, FqR = hqR * qR + 0.5f * g * hR * hR;
  const float FrL = TANGENTIAL_FIRST ? hrL * qL : hqL * rL, FrR = TANGENTIAL_FIRST ? hrR * qR : hqR * rR;
  if (sL >= 0.0f) { Fh = hqL; Fq = FqL; Fr = FrL; return; }
  if (sR <= 0.0f) { Fh = hqR; Fq = FqR; Fr = FrR; return; }
  const float inv = 1.0f / (sR - sL), sRL = sR * sL;
  Fh = (sR * hqL - sL * hqR + sRL * (hR - hL)) * inv;
  Fq = (sR * FqL - sL * FqR + sRL * (hqR - hqL)) * inv;
  Fr = (sR * FrL - sL * FrR + sRL * (hrR - hrL)) * inv;
}

__device__ __forceinline__ float step_dt(const SPar &P, const SClock *clk, int step3, int step2) {  // :679-684
  float cmax = clk->cmax[step3];
  if (cmax < 1e-12f) cmax = 1e-12f;
  const float dt_cfl = P.CFL * fminf(P.dx, P.dy) / cmax;
  return fminf(clk->t[step2] * P.dtau, dt_cfl);
}
__device__ __forceinline__ void advance_clock(const SPar &P, SClock *clk, int step2) {  // :767-768
  clk->tau[step2 ^ 1] = clk->tau[step2] + P.dtau;
  clk->t[step2 ^ 1] = clk->t[step2] * P.edtau;
}
__device__ __forceinline__ int next3(int s) { return s == 2 ? 0 : s + 1; }

// ---- flux_x_kernel + flux_y_kernel + update_kernel (:424-513) in one kernel -----------------------------
// FINAL != 0 (nu == 0: no viscosity kernel follows): also reduce the next step's cmax and advance the clock.
__global__ void __launch_bounds__(S_THREADS)
sw_update(const SPar P, const float *__restrict__ sig, const float *__restrict__ u, const float *__restrict__ v,
          float *__restrict__ osig, float *__restrict__ ou, float *__restrict__ ov, SClock *__restrict__ clk,
          int step3, int step2, int final_kernel) {
  __shared__ float s_h[S_SX * S_SY], s_u[S_SX * S_SY], s_v[S_SX * S_SY];
  __shared__ float s_Fh[S_NFX], s_Fmx[S_NFX], s_Fmy[S_NFX], s_Gh[S_NFY], s_Gmx[S_NFY], s_Gmy[S_NFY];
  const int bx0 = blockIdx.x * S_TX, by0 = blockIdx.y * S_TY;
  const float dt = step_dt(P, clk, step3, step2);
  const int fill3 = next3(step3);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    clk->dt_last = dt;
    clk->cmax[next3(fill3)] = 0.f;
  }
  for (int t = threadIdx.x; t < S_SX * S_SY; t += S_THREADS) {
    const int ly = t / S_SX, lx = t - ly * S_SX;
    const size_t id = (size_t)wrap(by0 + ly - 1, P.ny) * P.nx + wrap(bx0 + lx - 1, P.nx);
    s_h[t] = expf(sig[id]);
    s_u[t] = u[id];
    s_v[t] = v[id];
  }
  __syncthreads();
  for (int f = threadIdx.x; f < S_NFX; f += S_THREADS) {  // face on the RIGHT of tile column fx-1
    const int fy = f / (S_TX + 1), fx = f - fy * (S_TX + 1);
    const int cL = (fy + 1) * S_SX + fx, cR = cL + 1;
    hll_flux<false>(s_h[cL], s_u[cL], s_v[cL], s_h[cR], s_u[cR], s_v[cR], P.g, s_Fh[f], s_Fmx[f], s_Fmy[f]);
  }
  for (int f = threadIdx.x; f < S_NFY; f += S_THREADS) {  // face on TOP of tile row fy-1
    const int fy = f / S_TX, fx = f - fy * S_TX;
    const int cB = fy * S_SX + fx + 1, cT = cB + S_SX;
    hll_flux<true>(s_h[cB], s_v[cB], s_u[cB], s_h[cT], s_v[cT], s_u[cT], P.g, s_Gh[f], s_Gmy[f], s_Gmx[f]);
  }
  __syncthreads();
  const float invdx = 1.0f / P.dx, invdy = 1.0f / P.dy;
  float cmax = 0.f;
  for (int t = threadIdx.x; t < S_TX * S_TY; t += S_THREADS) {
    const int ly = t / S_TX, lx = t - ly * S_TX;
    const int i = bx0 + lx, j = by0 + ly;
    if (i >= P.nx || j >= P.ny) continue;
    const int c = (ly + 1) * S_SX + lx + 1;
    float h = s_h[c];
    float mx = h * s_u[c], my = h * s_v[c];
    const int fxp = ly * (S_TX + 1) + lx + 1, fxm = fxp - 1, fyp = (ly + 1) * S_TX + lx, fym = fyp - S_TX;
    const float dFx_h = s_Fh[fxp] - s_Fh[fxm], dFx_mx = s_Fmx[fxp] - s_Fmx[fxm], dFx_my = s_Fmy[fxp] - s_Fmy[fxm];
    const float dGy_h = s_Gh[fyp] - s_Gh[fym], dGy_mx = s_Gmx[fyp] - s_Gmx[fym], dGy_my = s_Gmy[fyp] - s_Gmy[fym];
    h -= dt * (dFx_h * invdx + dGy_h * invdy);
    mx -= dt * (dFx_mx * invdx + dGy_mx * invdy);
    my -= dt * (dFx_my * invdx + dGy_my * invdy);
    const float eps = 1e-6f;
    h = fmaxf(h, eps);
    const size_t id = (size_t)j * P.nx + i;
    const float so = logf(h), uo = mx / h, vo = my / h;
    osig[id] = so;
    ou[id] = uo;
    ov[id] = vo;
    if (final_kernel) {  // wavespeed_block_max :394-421 reads the stored state
      const float cc = sqrtf(P.g * expf(so));
      cmax = fmaxf(cmax, fmaxf(fabsf(uo) + cc, fabsf(vo) + cc));
    }
  }
  if (final_kernel) {
    cmax = tau::warp_max(cmax);
    if ((threadIdx.x & 31) == 0 && cmax > 0.f) tau::atomic_max_nonneg(&clk->cmax[fill3], cmax);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) advance_clock(P, clk, step2);
  }
}

// ---- viscosity_uv :516-551 as a Jacobi update; always the step's last kernel ---------------------------------
__global__ void __launch_bounds__(S_THREADS)
sw_viscosity(const SPar P, const float *__restrict__ sig, const float *__restrict__ u, const float *__restrict__ v,
             float *__restrict__ ou, float *__restrict__ ov, SClock *__restrict__ clk, int step3, int step2) {
  __shared__ float s_u[S_SX * S_SY], s_v[S_SX * S_SY];
  const int bx0 = blockIdx.x * S_TX, by0 = blockIdx.y * S_TY;
  const float dt = step_dt(P, clk, step3, step2);  // the value sw_update used: nothing it reads has moved
  for (int t = threadIdx.x; t < S_SX * S_SY; t += S_THREADS) {
    const int ly = t / S_SX, lx = t - ly * S_SX;
    const size_t id = (size_t)wrap(by0 + ly - 1, P.ny) * P.nx + wrap(bx0 + lx - 1, P.nx);
    s_u[t] = u[id];
    s_v[t] = v[id];
  }
  __syncthreads();
  const float invdx2 = 1.0f / (P.dx * P.dx), invdy2 = 1.0f / (P.dy * P.dy);
  float cmax = 0.f;
  for (int t = threadIdx.x; t < S_TX * S_TY; t += S_THREADS) {
    const int ly = t / S_TX, lx = t - ly * S_TX;
    const int i = bx0 + lx, j = by0 + ly;
    if (i >= P.nx || j >= P.ny) continue;
    const int c = (ly + 1) * S_SX + lx + 1;
    const float u_c = s_u[c], v_c = s_v[c];
    const float du = (s_u[c + 1] - 2.0f * u_c + s_u[c - 1]) * invdx2 + (s_u[c + S_SX] - 2.0f * u_c + s_u[c - S_SX]) * invdy2;
    const float dv = (s_v[c + 1] - 2.0f * v_c + s_v[c - 1]) * invdx2 + (s_v[c + S_SX] - 2.0f * v_c + s_v[c - S_SX]) * invdy2;
    const float un = u_c + P.nu * dt * du, vn = v_c + P.nu * dt * dv;
    const size_t id = (size_t)j * P.nx + i;
    ou[id] = un;
    ov[id] = vn;
    const float cc = sqrtf(P.g * expf(sig[id]));
    cmax = fmaxf(cmax, fmaxf(fabsf(un) + cc, fabsf(vn) + cc));
  }
  cmax = tau::warp_max(cmax);
  if ((threadIdx.x & 31) == 0 && cmax > 0.f) tau::atomic_max_nonneg(&clk->cmax[next3(step3)], cmax);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) advance_clock(P, clk, step2);
}

__global__ void sw_wavespeed(const SPar P, const float *__restrict__ sig, const float *__restrict__ u,
                             const float *__restrict__ v, SClock *clk, int step3) {
  const size_t n = (size_t)P.nx * P.ny;
  float cmax = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float c = sqrtf(P.g * expf(sig[i]));
    cmax = fmaxf(cmax, fmaxf(fabsf(u[i]) + c, fabsf(v[i]) + c));
  }
  cmax = tau::warp_max(cmax);
  if ((threadIdx.x & 31) == 0 && cmax > 0.f) tau::atomic_max_nonneg(&clk->cmax[step3], cmax);
}

}  // namespace

struct tau_sw {
  tau_sw_params p;
  int device;
  cudaStream_t stream;
  bool own_stream;
  float *st[2][3];  // [buffer][sigma, u, v]
  SClock *clk;
  int cur;
  long long steps, launches;
  bool have_state, timed;
  cudaEvent_t ev0, ev1;
};

namespace {
SPar make_par(const tau_sw *h) {
  SPar P;
  P.nx = h->p.nx; P.ny = h->p.ny; P.dx = h->p.dx; P.dy = h->p.dy; P.g = h->p.g; P.nu = h->p.nu;
  P.CFL = h->p.CFL; P.dtau = h->p.dtau;
  P.edtau = expf(h->p.dtau);
  return P;
}
}  // namespace

extern "C" {

// struct Params :52-89 (simulation fields, the struct's defaults; the usage text shows other numbers)
void tau_sw_default_params(tau_sw_params *p) {
  memset(p, 0, sizeof(*p));
  p->nx = 512; p->ny = 512; p->dx = 1.0f; p->dy = 1.0f;
  p->g = 9.81f; p->f0 = 1.0f; p->nu = 0.001f; p->H0 = 1000.0f;
  p->bumpAmp = 1.0f; p->bumpSigma = 1.0f; p->CFL = 0.5f;
  p->offx = 100.0f; p->offy = 100.0f; p->asym = 10.0f; p->swirl = 1.0f; p->swirlRc = 100.0f;
  p->tau0 = 0.0f; p->t0 = 1.0f; p->dtau = 1.0f;
}

// initialize_host :238-277 (host code in the reference too)
void tau_sw_init_host(const tau_sw_params *P, float *sigma, float *u, float *v) {
  const int nx = P->nx, ny = P->ny;
  const float cx = 0.5f * nx + P->offx, cy = 0.5f * ny + P->offy;
  const float sig2 = P->bumpSigma * P->bumpSigma;
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      const float dx = i - cx, dy = j - cy;
      const float r2 = (dx * dx + dy * dy) / sig2;
      const float theta = atan2f(dy, dx);
      const float mod = 1.0f + P->asym * cosf(theta);
      const float h = P->H0 + (P->bumpAmp * mod) * expf(-0.5f * r2);
      const size_t id = (size_t)j * nx + i;
      sigma[id] = logf(fmaxf(h, 1e-6f));
      const float rx = dx * P->dx, ry = dy * P->dy;
      const float r = sqrtf(rx * rx + ry * ry);
      const float rc = P->swirlRc * fminf(P->dx, P->dy);
      const float u_theta = (r > 0.0f && P->swirl != 0.0f) ? (P->swirl * r * expf(-0.5f * (r / rc) * (r / rc))) : 0.0f;
      u[id] = (r > 0.0f) ? (-u_theta * (ry / r)) : 0.0f;
      v[id] = (r > 0.0f) ? (u_theta * (rx / r)) : 0.0f;
    }
}

int tau_sw_create(const tau_sw_params *p, int device, void *stream, tau_sw **out) {
  TAU_REQUIRE(p && out, "tau_sw_create: null argument");
  TAU_REQUIRE(p->nx >= 1 && p->ny >= 1, "tau_sw_create: bad grid %d x %d", p->nx, p->ny);
  TAU_REQUIRE(p->dx > 0.f && p->dy > 0.f && p->g > 0.f, "tau_sw_create: dx, dy, g must be > 0");
  if (tau_device_count() <= 0) {
    tau_set_error("tau_sw_create: no CUDA device (this library has no CPU fallback)");
    return TAU_ERR_NODEV;
  }
  TAU_CUDA(cudaSetDevice(device));
  tau_sw *h = new (std::nothrow) tau_sw();
  if (!h) return TAU_ERR_NOMEM;
  memset(h, 0, sizeof(*h));
  h->p = *p;
  h->device = device;
  if (stream) {
    h->stream = (cudaStream_t)stream;
    h->own_stream = false;
  } else {
    TAU_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  const size_t n = (size_t)p->nx * p->ny;
  for (int b = 0; b < 2; ++b)
    for (int f = 0; f < 3; ++f) TAU_CUDA(cudaMalloc(&h->st[b][f], n * sizeof(float)));
  TAU_CUDA(cudaMalloc(&h->clk, sizeof(SClock)));
  TAU_CUDA(cudaEventCreate(&h->ev0));
  TAU_CUDA(cudaEventCreate(&h->ev1));
  *out = h;
  return TAU_OK;
}

int tau_sw_upload(tau_sw *h, const float *sigma, const float *u, const float *v, const float *clock2) {
  TAU_REQUIRE(h && sigma && u && v, "tau_sw_upload: null argument");
  TAU_CUDA(cudaSetDevice(h->device));
  const size_t n = (size_t)h->p.nx * h->p.ny;
  const float *src[3] = {sigma, u, v};
  for (int f = 0; f < 3; ++f)
    TAU_CUDA(cudaMemcpyAsync(h->st[h->cur][f], src[f], n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  SClock c;
  memset(&c, 0, sizeof(c));
  c.t[h->steps & 1] = clock2 ? clock2[0] : h->p.t0;
  c.tau[h->steps & 1] = clock2 ? clock2[1] : h->p.tau0;
  TAU_CUDA(cudaMemcpyAsync(h->clk, &c, sizeof(c), cudaMemcpyHostToDevice, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  sw_wavespeed<<<148 * 4, 256, 0, h->stream>>>(make_par(h), h->st[h->cur][0], h->st[h->cur][1], h->st[h->cur][2],
                                               h->clk, (int)(h->steps % 3));
  h->launches++;
  TAU_CUDA(cudaGetLastError());
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  h->have_state = true;
  return TAU_OK;
}

int tau_sw_init(tau_sw *h) {
  TAU_REQUIRE(h, "tau_sw_init: null handle");
  const size_t n = (size_t)h->p.nx * h->p.ny;
  std::vector<float> s(n), u(n), v(n);
  tau_sw_init_host(&h->p, s.data(), u.data(), v.data());
  h->steps = 0;
  return tau_sw_upload(h, s.data(), u.data(), v.data(), nullptr);
}

// THE hot path: nsteps x { do_step :669-705; tau += dtau; t *= expf(dtau) }, no host sync
int tau_sw_step(tau_sw *h, int nsteps) {
  TAU_REQUIRE(h && nsteps >= 0, "tau_sw_step: bad argument");
  TAU_REQUIRE(h->have_state, "tau_sw_step: no state (call tau_sw_init or tau_sw_upload)");
  TAU_CUDA(cudaSetDevice(h->device));
  const SPar P = make_par(h);
  const dim3 grid((P.nx + S_TX - 1) / S_TX, (P.ny + S_TY - 1) / S_TY);
  const bool visc = h->p.nu > 0.0f;  // :700
  TAU_CUDA(cudaEventRecord(h->ev0, h->stream));
  for (int s = 0; s < nsteps; ++s) {
    const int step3 = (int)(h->steps % 3), step2 = (int)(h->steps & 1), a = h->cur, b = a ^ 1;
    sw_update<<<grid, S_THREADS, 0, h->stream>>>(P, h->st[a][0], h->st[a][1], h->st[a][2], h->st[b][0], h->st[b][1],
                                                 h->st[b][2], h->clk, step3, step2, visc ? 0 : 1);
    h->launches++;
    if (visc) {
      // sigma stays in buffer b; u, v go b -> a; then buffer a needs b's sigma: swap the sigma pointers
      sw_viscosity<<<grid, S_THREADS, 0, h->stream>>>(P, h->st[b][0], h->st[b][1], h->st[b][2], h->st[a][1],
                                                      h->st[a][2], h->clk, step3, step2);
      h->launches++;
      float *tmp = h->st[a][0];
      h->st[a][0] = h->st[b][0];
      h->st[b][0] = tmp;
      // current state: sigma (now st[a][0]), u, v in buffer a
    } else {
      h->cur = b;
    }
    h->steps++;
  }
  TAU_CUDA(cudaGetLastError());
  TAU_CUDA(cudaEventRecord(h->ev1, h->stream));
  h->timed = true;
  return TAU_OK;
}

int tau_sw_clock(tau_sw *h, float *t, float *tau, float *dt_last) {
  TAU_REQUIRE(h, "tau_sw_clock: null handle");
  SClock c;
  TAU_CUDA(cudaMemcpyAsync(&c, h->clk, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  if (t) *t = c.t[h->steps & 1];
  if (tau) *tau = c.tau[h->steps & 1];
  if (dt_last) *dt_last = c.dt_last;
  return TAU_OK;
}

int tau_sw_download(tau_sw *h, float *sigma, float *u, float *v) {
  TAU_REQUIRE(h, "tau_sw_download: null handle");
  TAU_CUDA(cudaSetDevice(h->device));
  const size_t n = (size_t)h->p.nx * h->p.ny;
  float *dst[3] = {sigma, u, v};
  for (int f = 0; f < 3; ++f)
    if (dst[f]) TAU_CUDA(cudaMemcpyAsync(dst[f], h->st[h->cur][f], n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

int tau_sw_sync(tau_sw *h) {
  TAU_REQUIRE(h, "tau_sw_sync: null handle");
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}
long long tau_sw_steps_done(tau_sw *h) { return h ? h->steps : -1; }
long long tau_sw_launch_count(tau_sw *h) { return h ? h->launches : -1; }
int tau_sw_last_step_ms(tau_sw *h, float *ms) {
  TAU_REQUIRE(h && ms, "tau_sw_last_step_ms: null argument");
  TAU_REQUIRE(h->timed, "tau_sw_last_step_ms: no step has been timed yet");
  TAU_CUDA(cudaEventSynchronize(h->ev1));
  TAU_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return TAU_OK;
}
int tau_sw_destroy(tau_sw *h) {
  if (!h) return TAU_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaFree(h->clk);
  for (int b = 1; b >= 0; --b)
    for (int f = 2; f >= 0; --f) cudaFree(h->st[b][f]);
  cudaEventDestroy(h->ev1);
  cudaEventDestroy(h->ev0);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return TAU_OK;
}

}  // extern "C"
