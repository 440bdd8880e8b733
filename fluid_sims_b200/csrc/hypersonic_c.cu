// hypersonic_c.cu — the update path of the reference's CPU solver `tau_hypersonic` (tau_hypersonic.c, BASELINE
// config 1: 256 x 256, "speed mode") on the device, in fp64, behind tau_hypc_* (include/tau_b200.h).
//
// tau_hypersonic.c is a different scheme from tau_hypersonic_cuda.cu (gamma 1.4, CFL 0.3, Mach 15, SLIP wall at a
// circular body, no diffusion, unguarded HLLC, EPS 1e-10), so it gets its own kernels rather than a mode of
// hypersonic2d.cu.  Its per-step host entry point is `static void step_physics(void)` (:500-674) on file-static
// AoS state; this file replaces it with tau_hypc_step(h, n): TWO kernels per step, dt and sim_t on the device.
//
// Bit-exactness is the design constraint here, not speed (the grid is 256 x 256): the reference is compiled by
// `gcc -O3` for x86-64 without FMA, every operation is an IEEE fp64 add / mul / div / sqrt / min / max, and the
// device has all of those correctly rounded.  This translation unit is therefore built with --fmad=false and
// keeps the reference's operation order; results equal the reference's to 0 ulp (tests/test_hypc_gpu.py).
//   * step_physics sweeps x faces then y faces, scattering +-dt*F into Unew.  A cell therefore receives, in this
//     order: + left-face flux, - right-face flux, + bottom-face flux, - top-face flux.  hypc_update gathers the
//     four in exactly that order; each face flux is a pure function of the same inputs on either side, so the two
//     cells that share a face compute identical values (no atomics, no ordering dependence).
//   * the predicted face states of a cell (reconstruct + Hancock half step, once per axis) are computed once per
//     cell by hypc_predict and parked in HBM (16 doubles per cell), instead of 2-4 times per face as the
//     reference's loops do (:540-574).
//   * compute_dt (:477-498) scans the state BEFORE the column-0 inflow overwrite (:509-515): the scan is fused
//     into the previous step's update kernel (max is exactly associative: integer atomicMax on the bit pattern),
//     and the overwrite is applied at read time, so memory keeps what step_physics leaves in U.
// Layout: 4 SoA planes (rho, mx, my, E) of H x W doubles (index y*W+x, :46), ping-pong; mask H x W bytes.
#include "common.cuh"
#include "../../include/tau_b200.h"

#include <math.h>
#include <new>
#include <vector>

namespace {

constexpr double HC_GAMMA = 1.4;   // :15
constexpr double HC_CFL = 0.3;     // :16
constexpr double HC_EPS = 1e-10;   // EPS_RHO == EPS_P :20-21
constexpr double HC_MACH = 15.0;   // :247

struct Cons { double rho, mx, my, E; };
struct Prim { double rho, u, v, p; };

struct HCtrl {
  unsigned long long maxs[3];  // rotating max-wavespeed slots (bit patterns of non-negative doubles)
  double sim_t, dt_last;
};

struct HPar {
  int W, H;
  size_t N;
  double inflow_u;  // 15 * sqrt(1.4), evaluated on the host like the reference does (:250-251)
};

__host__ __device__ __forceinline__ Prim c2p(Cons c) {  // :65-81
  Prim q;
  const double rho = fmax(c.rho, HC_EPS), inv = 1.0 / rho;
  const double u = c.mx * inv, v = c.my * inv;
  const double kin = 0.5 * rho * (u * u + v * v);
  q.rho = rho; q.u = u; q.v = v;
  q.p = (HC_GAMMA - 1.0) * fmax(c.E - kin, HC_EPS);
  return q;
}
__host__ __device__ __forceinline__ Cons p2c(Prim q) {  // :83-93
  Cons c;
  const double rho = fmax(q.rho, HC_EPS), pr = fmax(q.p, HC_EPS);
  c.rho = rho; c.mx = rho * q.u; c.my = rho * q.v;
  c.E = pr / (HC_GAMMA - 1.0) + 0.5 * rho * (q.u * q.u + q.v * q.v);
  return c;
}
__device__ __forceinline__ double sound(Prim q) {  // :95-97
  return sqrt(HC_GAMMA * fmax(q.p, HC_EPS) / fmax(q.rho, HC_EPS));
}

template <int AX> __device__ __forceinline__ Cons flux(Cons c) {  // flux_x :99-107, flux_y :109-117
  const Prim q = c2p(c);
  const double w = AX ? q.v : q.u;
  Cons f;
  f.rho = AX ? c.my : c.mx;
  f.mx = AX ? c.mx * w : c.mx * w + q.p;
  f.my = AX ? c.my * w + q.p : c.my * w;
  f.E = (c.E + q.p) * w;
  return f;
}

template <int AX> __device__ Cons hllc(Cons UL, Cons UR) {  // hllc_x :119-180, hllc_y :182-243
  const Prim L = c2p(UL), R = c2p(UR);
  const double aL = sound(L), aR = sound(R);
  const double nL = AX ? L.v : L.u, nR = AX ? R.v : R.u;
  const double tL = AX ? L.u : L.v, tR = AX ? R.u : R.v;
  const double SL = fmin(nL - aL, nR - aR), SR = fmax(nL + aL, nR + aR);
  const Cons FL = flux<AX>(UL), FR = flux<AX>(UR);
  if (SL >= 0.0) return FL;
  if (SR <= 0.0) return FR;
  const double num = R.p - L.p + L.rho * nL * (SL - nL) - R.rho * nR * (SR - nR);
  const double den = L.rho * (SL - nL) - R.rho * (SR - nR);
  const double SM = num / den;
  double pStar = L.p + L.rho * (SL - nL) * (SM - nL);
  pStar = fmax(pStar, HC_EPS);
  Cons F;
  if (SM >= 0.0) {
    const double rs = L.rho * (SL - nL) / (SL - SM);
    const double sn = rs * SM, st = rs * tL;
    const double Es = ((SL - nL) * UL.E - L.p * nL + pStar * SM) / (SL - SM);
    F.rho = FL.rho + SL * (rs - UL.rho);
    F.mx = FL.mx + SL * ((AX ? st : sn) - UL.mx);
    F.my = FL.my + SL * ((AX ? sn : st) - UL.my);
    F.E = FL.E + SL * (Es - UL.E);
  } else {
    const double rs = R.rho * (SR - nR) / (SR - SM);
    const double sn = rs * SM, st = rs * tR;
    const double Es = ((SR - nR) * UR.E - R.p * nR + pStar * SM) / (SR - SM);
    F.rho = FR.rho + SR * (rs - UR.rho);
    F.mx = FR.mx + SR * ((AX ? st : sn) - UR.mx);
    F.my = FR.my + SR * ((AX ? sn : st) - UR.my);
    F.E = FR.E + SR * (Es - UR.E);
  }
  return F;
}

__host__ __device__ __forceinline__ Prim inflow_prim(double inflow_u) {  // :245-254
  Prim s;
  s.rho = 1.0; s.u = inflow_u; s.v = 0.0; s.p = 1.0;
  return s;
}

// reflect_slip :279-294 — (nx, ny) enter the arithmetic as in the reference (signed zeros included)
__device__ __forceinline__ Cons reflect(Cons inside, double nx, double ny) {
  const Prim q = c2p(inside);
  double vn = q.u * nx + q.v * ny;
  const double ut = -q.u * ny + q.v * nx;
  vn = -vn;
  Prim g;
  g.rho = q.rho; g.p = q.p;
  g.u = vn * nx - ut * ny;
  g.v = vn * ny + ut * nx;
  return p2c(g);
}

struct Grid {
  const double *rho, *mx, *my, *E;
  const uint8_t *mask;
  int W, H;
  Cons inflowC;
  // the state as step_physics' sweeps see it: column 0 of unmasked rows already holds the inflow state (:509-515)
  __device__ __forceinline__ Cons at(int x, int y) const {
    const size_t i = (size_t)y * W + x;
    if (x == 0 && !mask[i]) return inflowC;
    return Cons{rho[i], mx[i], my[i], E[i]};
  }
};

// neighbor_or_wall :295-315
__device__ __forceinline__ Cons neighbor_or_wall(const Grid &g, int x, int y, int dxc, int dyc, double nx, double ny) {
  const int xn = x + dxc;
  int yn = y + dyc;
  if (xn < 0) return g.inflowC;
  if (xn >= g.W) return g.at(g.W - 1, y);
  if (yn < 0) yn = 0;
  if (yn >= g.H) yn = g.H - 1;
  if (g.mask[(size_t)yn * g.W + xn]) return reflect(g.at(x, y), nx, ny);
  return g.at(xn, yn);
}

__device__ __forceinline__ double minmod(double a, double b) {  // :50-54
  if (a * b <= 0.0) return 0.0;
  return fabs(a) < fabs(b) ? a : b;
}
__device__ __forceinline__ double slope(double m, double c, double p) {  // mc_limiter :56-63 on (dl, dc, dr)
  const double dl = c - m, dr = p - c, dc = 0.5 * (p - m);
  const double m1 = minmod(dl, dr), m2 = minmod(dc, 2.0 * dl), m3 = minmod(dc, 2.0 * dr);
  return minmod(m1, minmod(m2, m3));
}

__device__ __forceinline__ void positive_faces(Prim &qm, const Prim &qc, Prim &qp) {  // :320-346
  for (int it = 0; it < 8; ++it) {
    const bool bad = qm.rho <= HC_EPS || qp.rho <= HC_EPS || qm.p <= HC_EPS || qp.p <= HC_EPS;
    if (!bad) return;
    qm.rho = 0.5 * (qm.rho + qc.rho); qm.u = 0.5 * (qm.u + qc.u); qm.v = 0.5 * (qm.v + qc.v); qm.p = 0.5 * (qm.p + qc.p);
    qp.rho = 0.5 * (qp.rho + qc.rho); qp.u = 0.5 * (qp.u + qc.u); qp.v = 0.5 * (qp.v + qc.v); qp.p = 0.5 * (qp.p + qc.p);
  }
  qm.rho = fmax(qm.rho, HC_EPS); qp.rho = fmax(qp.rho, HC_EPS);
  qm.p = fmax(qm.p, HC_EPS);     qp.p = fmax(qp.p, HC_EPS);
}

// reconstruct_x/y :348-418 + half_step_predict_x/y :420-448 as step_physics applies them (:551-574, :617-632)
template <int AX>
__device__ void predict_cell(const Grid &g, int x, int y, double half_dt, Prim &lo, Prim &hi) {
  const double nx = AX ? 0.0 : 1.0, ny = AX ? 1.0 : 0.0;
  const Prim qc = c2p(g.at(x, y));
  const Prim qm = c2p(neighbor_or_wall(g, x, y, AX ? 0 : -1, AX ? -1 : 0, nx, ny));
  const Prim qp = c2p(neighbor_or_wall(g, x, y, AX ? 0 : +1, AX ? +1 : 0, nx, ny));
  const double s_rho = slope(qm.rho, qc.rho, qp.rho), s_u = slope(qm.u, qc.u, qp.u);
  const double s_v = slope(qm.v, qc.v, qp.v), s_p = slope(qm.p, qc.p, qp.p);
  Prim qL{qc.rho - 0.5 * s_rho, qc.u - 0.5 * s_u, qc.v - 0.5 * s_v, qc.p - 0.5 * s_p};
  Prim qR{qc.rho + 0.5 * s_rho, qc.u + 0.5 * s_u, qc.v + 0.5 * s_v, qc.p + 0.5 * s_p};
  positive_faces(qL, qc, qR);
  const Cons Ff = flux<AX>(p2c(qR)), Fb = flux<AX>(p2c(qL));
  const double d_rho = Ff.rho - Fb.rho, d_mx = Ff.mx - Fb.mx, d_my = Ff.my - Fb.my, d_E = Ff.E - Fb.E;
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    Cons c = p2c(side ? qR : qL);
    c.rho -= half_dt * d_rho; c.mx -= half_dt * d_mx; c.my -= half_dt * d_my; c.E -= half_dt * d_E;
    Prim o = c2p(c);
    o.rho = fmax(o.rho, HC_EPS); o.p = fmax(o.p, HC_EPS);
    if (side) hi = o; else lo = o;
  }
}

__device__ __forceinline__ double dt_of(const HCtrl *ctrl, int slot) {  // compute_dt :477-498
  double maxs = __longlong_as_double((long long)ctrl->maxs[slot]);
  if (!(maxs > 1e-12)) maxs = 1e-12;
  return HC_CFL * fmin(1.0, 1.0) / maxs;
}

// pred: 16 planes of N doubles — [axis][side][rho,u,v,p]
__device__ __forceinline__ Prim load_pred(const double *pred, size_t N, int ax, int side, size_t i) {
  const double *b = pred + (size_t)(ax * 8 + side * 4) * N + i;
  return Prim{b[0], b[N], b[2 * N], b[3 * N]};
}
__device__ __forceinline__ void store_pred(double *pred, size_t N, int ax, int side, size_t i, const Prim &q) {
  double *b = pred + (size_t)(ax * 8 + side * 4) * N + i;
  b[0] = q.rho; b[N] = q.u; b[2 * N] = q.v; b[3 * N] = q.p;
}

__global__ void __launch_bounds__(128) hypc_predict(const HPar P, Grid g, double *__restrict__ pred,
                                                    const HCtrl *__restrict__ ctrl, int slot) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= P.W) return;
  const size_t i = (size_t)y * P.W + x;
  if (g.mask[i]) return;
  const double half_dt = 0.5 * (dt_of(ctrl, slot) / 1.0);
  Prim lo, hi;
  predict_cell<0>(g, x, y, half_dt, lo, hi);
  store_pred(pred, P.N, 0, 0, i, lo);
  store_pred(pred, P.N, 0, 1, i, hi);
  predict_cell<1>(g, x, y, half_dt, lo, hi);
  store_pred(pred, P.N, 1, 0, i, lo);
  store_pred(pred, P.N, 1, 1, i, hi);
}

// one face of the sweeps (:524-586 x, :588-650 y): (xl, yl) / (xh, yh) = the cells on its low / high side
template <int AX>
__device__ Cons face_flux(const HPar &P, const Grid &g, const double *pred, int xl, int yl, int xh, int yh) {
  const double nx = AX ? 0.0 : 1.0, ny = AX ? 1.0 : 0.0;
  const size_t il = (size_t)yl * P.W + xl, ih = (size_t)yh * P.W + xh;
  Prim ql = g.mask[il] ? c2p(reflect(g.at(xh, yh), nx, ny)) : load_pred(pred, P.N, AX, 1, il);
  Prim qh = g.mask[ih] ? c2p(reflect(g.at(xl, yl), nx, ny)) : load_pred(pred, P.N, AX, 0, ih);
  ql.rho = fmax(ql.rho, HC_EPS); ql.p = fmax(ql.p, HC_EPS);
  qh.rho = fmax(qh.rho, HC_EPS); qh.p = fmax(qh.p, HC_EPS);
  return hllc<AX>(p2c(ql), p2c(qh));
}

__global__ void __launch_bounds__(128) hypc_update(const HPar P, Grid g, const double *__restrict__ pred,
                                                   double *__restrict__ o_rho, double *__restrict__ o_mx,
                                                   double *__restrict__ o_my, double *__restrict__ o_E,
                                                   HCtrl *__restrict__ ctrl, int slot) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  double ws = 0.0;
  if (x < P.W) {
    const size_t i = (size_t)y * P.W + x;
    if (g.mask[i]) {  // body cells pass through (:653-654)
      o_rho[i] = g.rho[i]; o_mx[i] = g.mx[i]; o_my[i] = g.my[i]; o_E[i] = g.E[i];
    } else {
      const double dt = dt_of(ctrl, slot);
      Cons c = g.at(x, y);  // Unew = U after the inflow overwrite (:517)
      if (x >= 1) {  // + flux through the left face (this cell is iR of face x)
        const Cons F = face_flux<0>(P, g, pred, x - 1, y, x, y);
        c.rho += dt * F.rho; c.mx += dt * F.mx; c.my += dt * F.my; c.E += dt * F.E;
      }
      if (x + 1 < P.W) {  // - flux through the right face (iL of face x+1)
        const Cons F = face_flux<0>(P, g, pred, x, y, x + 1, y);
        c.rho -= dt * F.rho; c.mx -= dt * F.mx; c.my -= dt * F.my; c.E -= dt * F.E;
      }
      if (y >= 1) {
        const Cons F = face_flux<1>(P, g, pred, x, y - 1, x, y);
        c.rho += dt * F.rho; c.mx += dt * F.mx; c.my += dt * F.my; c.E += dt * F.E;
      }
      if (y + 1 < P.H) {
        const Cons F = face_flux<1>(P, g, pred, x, y, x, y + 1);
        c.rho -= dt * F.rho; c.mx -= dt * F.mx; c.my -= dt * F.my; c.E -= dt * F.E;
      }
      c.rho = fmax(c.rho, HC_EPS);  // :656-666
      Prim q = c2p(c);
      if (q.p <= HC_EPS) {
        q.p = HC_EPS;
        c = p2c(q);
      }
      o_rho[i] = c.rho; o_mx[i] = c.mx; o_my[i] = c.my; o_E[i] = c.E;
      // the next step's compute_dt on this cell (:484-493); q is cons_to_prim of what was just stored unless
      // the repair ran — then recompute, as the scan would
      if (q.p <= HC_EPS) q = c2p(c);
      const double a = sound(q);
      ws = fmax(fabs(q.u) + a, fabs(q.v) + a);
    }
  }
  ws = tau::warp_max(ws);
  if ((threadIdx.x & 31) == 0 && ws > 0.0)
    atomicMax(&ctrl->maxs[(slot + 1) % 3], (unsigned long long)__double_as_longlong(ws));
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    const double dt = dt_of(ctrl, slot);
    ctrl->sim_t += dt;  // :673
    ctrl->dt_last = dt;
    ctrl->maxs[(slot + 2) % 3] = 0ull;
  }
}

// standalone compute_dt scan for a state that did not come out of hypc_update (init / upload)
__global__ void hypc_wavespeed(const HPar P, const double *rho, const double *mx, const double *my, const double *E,
                               const uint8_t *mask, HCtrl *ctrl, int slot) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  double ws = 0.0;
  if (i < P.N && !mask[i]) {
    const Prim q = c2p(Cons{rho[i], mx[i], my[i], E[i]});
    const double a = sound(q);
    ws = fmax(fabs(q.u) + a, fabs(q.v) + a);
  }
  ws = tau::warp_max(ws);
  if ((threadIdx.x & 31) == 0 && ws > 0.0)
    atomicMax(&ctrl->maxs[slot], (unsigned long long)__double_as_longlong(ws));
}

// ---- render pass: main()'s loops :713-786 ------------------------------------------------------------------
__device__ __forceinline__ double rho_bc(const HPar &P, const double *rho, int x, int y) {  // get_cell_with_bc :256-277
  if (y < 0) y = 0;
  if (y >= P.H) y = P.H - 1;
  if (x < 0) return c2p(p2c(inflow_prim(P.inflow_u))).rho;
  if (x >= P.W) x = P.W - 1;
  return fmax(rho[(size_t)y * P.W + x], HC_EPS);
}
__device__ double view_value(const HPar &P, const double *rho, const double *mx, const double *my, const double *E,
                             int x, int y, int mode) {
  const size_t i = (size_t)y * P.W + x;
  const Prim q = c2p(Cons{rho[i], mx[i], my[i], E[i]});
  if (mode == 0) return log(q.rho);
  if (mode == 1) return log(q.p);
  if (mode == 2) return sqrt(q.u * q.u + q.v * q.v);
  const double gx = 0.5 * (rho_bc(P, rho, x + 1, y) - rho_bc(P, rho, x - 1, y));
  const double gy = 0.5 * (rho_bc(P, rho, x, y + 1) - rho_bc(P, rho, x, y - 1));
  return log(1e-12 + sqrt(gx * gx + gy * gy));
}
// order-preserving key of a double: min / max by integer atomics are exact and order independent
__device__ __forceinline__ unsigned long long okey(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double okey_inv(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double d;
  memcpy(&d, &b, sizeof(d));
  return d;
}
__global__ void hypc_render_minmax(const HPar P, const double *rho, const double *mx, const double *my, const double *E,
                                   const uint8_t *mask, int mode, unsigned long long *keys /* [min, max] */) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  unsigned long long kmin = ~0ull, kmax = 0ull;
  if (x < P.W && !mask[(size_t)y * P.W + x]) kmin = kmax = okey(view_value(P, rho, mx, my, E, x, y, mode));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o), b = __shfl_xor_sync(0xffffffffu, kmax, o);
    kmin = a < kmin ? a : kmin;
    kmax = b > kmax ? b : kmax;
  }
  if ((threadIdx.x & 31) == 0 && kmax != 0ull) {
    atomicMin(&keys[0], kmin);
    atomicMax(&keys[1], kmax);
  }
}
__global__ void hypc_render_pixels(const HPar P, const double *rho, const double *mx, const double *my, const double *E,
                                   const uint8_t *mask, int mode, double minv, double maxv, uint32_t *rgba) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= P.W) return;
  const size_t i = (size_t)y * P.W + x;
  if (mask[i]) {
    rgba[i] = 110u | (110u << 8) | (110u << 16) | (255u << 24);  // :756-761
    return;
  }
  const double inv = 1.0 / fmax(maxv - minv, 1e-30);
  double t = (view_value(P, rho, mx, my, E, x, y, mode) - minv) * inv;
  if (t < 0) t = 0;  // get_color :676-686
  if (t > 1) t = 1;
  const unsigned r = (unsigned char)(255 * fmin(1.0, fmax(0.0, 3 * t - 1)));
  const unsigned g = (unsigned char)(255 * fmin(1.0, fmax(0.0, 2 - 4 * fabs(t - 0.5))));
  const unsigned b = (unsigned char)(255 * fmin(1.0, fmax(0.0, 2 - 3 * t)));
  rgba[i] = r | (g << 8) | (b << 16) | (255u << 24);
}

}  // namespace

struct tau_hypc {
  int W, H, device;
  size_t N;
  cudaStream_t stream;
  bool own_stream, have_state, timed;
  double *U[2];   // 4 planes each
  double *pred;   // 16 planes
  uint8_t *mask;
  HCtrl *ctrl;
  unsigned long long *keys;
  uint32_t *pixels;
  int cur;
  long long steps, launches;
  double inflow_u;
  cudaEvent_t ev0, ev1;
};

namespace {
HPar make_par(const tau_hypc *h) { return HPar{h->W, h->H, h->N, h->inflow_u}; }
Grid make_grid(const tau_hypc *h, int b) {
  const double *u = h->U[b];
  return Grid{u, u + h->N, u + 2 * h->N, u + 3 * h->N, h->mask, h->W, h->H, p2c(inflow_prim(h->inflow_u))};
}
int state_changed(tau_hypc *h) {
  TAU_CUDA(cudaMemsetAsync(h->ctrl->maxs, 0, sizeof(h->ctrl->maxs), h->stream));
  const double *u = h->U[h->cur];
  hypc_wavespeed<<<(unsigned)((h->N + 255) / 256), 256, 0, h->stream>>>(make_par(h), u, u + h->N, u + 2 * h->N,
                                                                         u + 3 * h->N, h->mask, h->ctrl,
                                                                         (int)(h->steps % 3));
  h->launches++;
  TAU_CUDA(cudaGetLastError());
  h->have_state = true;
  return TAU_OK;
}
}  // namespace

extern "C" {

int tau_hypc_create(int W, int H, int device, void *stream, tau_hypc **out) {
  TAU_REQUIRE(out, "tau_hypc_create: null argument");
  TAU_REQUIRE(W >= 2 && H >= 2 && (long long)W * H < (1ll << 31), "tau_hypc_create: bad grid %d x %d", W, H);
  if (tau_device_count() <= 0) {
    tau_set_error("tau_hypc_create: no CUDA device (this library has no CPU fallback)");
    return TAU_ERR_NODEV;
  }
  TAU_CUDA(cudaSetDevice(device));
  tau_hypc *h = new (std::nothrow) tau_hypc();
  if (!h) return TAU_ERR_NOMEM;
  memset(h, 0, sizeof(*h));
  h->W = W; h->H = H; h->N = (size_t)W * H; h->device = device;
  h->inflow_u = HC_MACH * sqrt(HC_GAMMA * 1.0 / 1.0);  // :248-251
  if (stream) {
    h->stream = (cudaStream_t)stream;
  } else {
    TAU_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  for (int b = 0; b < 2; ++b) TAU_CUDA(cudaMalloc(&h->U[b], 4 * h->N * sizeof(double)));
  TAU_CUDA(cudaMalloc(&h->pred, 16 * h->N * sizeof(double)));
  TAU_CUDA(cudaMalloc(&h->mask, h->N));
  TAU_CUDA(cudaMalloc(&h->ctrl, sizeof(HCtrl)));
  TAU_CUDA(cudaMalloc(&h->keys, 2 * sizeof(unsigned long long)));
  TAU_CUDA(cudaMemsetAsync(h->ctrl, 0, sizeof(HCtrl), h->stream));
  TAU_CUDA(cudaEventCreate(&h->ev0));
  TAU_CUDA(cudaEventCreate(&h->ev1));
  *out = h;
  return TAU_OK;
}

// planes: rho, mx, my, E (H x W doubles, index y*W+x — the SoA view of the reference's `Cons U[W*H]`, :38); mask H x W bytes
int tau_hypc_upload(tau_hypc *h, const double *const planes[4], const uint8_t *mask, double sim_t) {
  TAU_REQUIRE(h && planes && mask, "tau_hypc_upload: null argument");
  TAU_CUDA(cudaSetDevice(h->device));
  for (int f = 0; f < 4; ++f) {
    TAU_REQUIRE(planes[f], "tau_hypc_upload: null plane %d", f);
    TAU_CUDA(cudaMemcpyAsync(h->U[h->cur] + f * h->N, planes[f], h->N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  TAU_CUDA(cudaMemcpyAsync(h->mask, mask, h->N, cudaMemcpyHostToDevice, h->stream));
  HCtrl c;
  memset(&c, 0, sizeof(c));
  c.sim_t = sim_t;
  TAU_CUDA(cudaMemcpyAsync(h->ctrl, &c, sizeof(c), cudaMemcpyHostToDevice, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  h->steps = 0;
  const int rc = state_changed(h);
  if (rc) return rc;
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

// init_sim :450-475 (host, like the reference; body = disc of radius H/6 at (W/3, H/2) in integer arithmetic)
void tau_hypc_init_host(int W, int H, double *rho, double *mx, double *my, double *E, uint8_t *mask) {
  const int cx = W / 3, cy = H / 2, r = H / 6;
  const Prim in = inflow_prim(HC_MACH * sqrt(HC_GAMMA * 1.0 / 1.0));
  const Cons cin = p2c(in), crest = p2c(Prim{in.rho, 0.0, 0.0, in.p});
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const size_t i = (size_t)y * W + x;
      const int dx = x - cx, dy = y - cy;
      mask[i] = (dx * dx + dy * dy < r * r) ? 1 : 0;
      const Cons c = mask[i] ? crest : cin;
      rho[i] = c.rho; mx[i] = c.mx; my[i] = c.my; E[i] = c.E;
    }
}

int tau_hypc_init(tau_hypc *h) {
  TAU_REQUIRE(h, "tau_hypc_init: null handle");
  std::vector<double> u(4 * h->N);
  std::vector<uint8_t> m(h->N);
  tau_hypc_init_host(h->W, h->H, u.data(), u.data() + h->N, u.data() + 2 * h->N, u.data() + 3 * h->N, m.data());
  const double *pl[4] = {u.data(), u.data() + h->N, u.data() + 2 * h->N, u.data() + 3 * h->N};
  return tau_hypc_upload(h, pl, m.data(), 0.0);
}

// THE hot path: nsteps x step_physics (:500-674), two kernels per step, no host synchronisation
int tau_hypc_step(tau_hypc *h, int nsteps) {
  TAU_REQUIRE(h && nsteps >= 0, "tau_hypc_step: bad argument");
  TAU_REQUIRE(h->have_state, "tau_hypc_step: no state (call tau_hypc_init or tau_hypc_upload)");
  TAU_CUDA(cudaSetDevice(h->device));
  const HPar P = make_par(h);
  const dim3 grid((P.W + 127) / 128, P.H);
  TAU_CUDA(cudaEventRecord(h->ev0, h->stream));
  for (int s = 0; s < nsteps; ++s) {
    const int slot = (int)(h->steps % 3), a = h->cur;
    double *o = h->U[a ^ 1];
    hypc_predict<<<grid, 128, 0, h->stream>>>(P, make_grid(h, a), h->pred, h->ctrl, slot);
    hypc_update<<<grid, 128, 0, h->stream>>>(P, make_grid(h, a), h->pred, o, o + h->N, o + 2 * h->N, o + 3 * h->N,
                                             h->ctrl, slot);
    h->launches += 2;
    h->cur = a ^ 1;
    h->steps++;
  }
  TAU_CUDA(cudaGetLastError());
  TAU_CUDA(cudaEventRecord(h->ev1, h->stream));
  h->timed = true;
  return TAU_OK;
}

int tau_hypc_clock(tau_hypc *h, double *sim_t, double *dt_last) {
  TAU_REQUIRE(h, "tau_hypc_clock: null handle");
  HCtrl c;
  TAU_CUDA(cudaMemcpyAsync(&c, h->ctrl, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  if (sim_t) *sim_t = c.sim_t;
  if (dt_last) *dt_last = c.dt_last;
  return TAU_OK;
}

int tau_hypc_download(tau_hypc *h, double *const planes[4], uint8_t *mask) {
  TAU_REQUIRE(h && planes, "tau_hypc_download: null argument");
  TAU_CUDA(cudaSetDevice(h->device));
  for (int f = 0; f < 4; ++f)
    if (planes[f])
      TAU_CUDA(cudaMemcpyAsync(planes[f], h->U[h->cur] + f * h->N, h->N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (mask) TAU_CUDA(cudaMemcpyAsync(mask, h->mask, h->N, cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

// main()'s render pass (:713-786): view_mode 0 log rho, 1 log p, 2 speed ("speed mode"), 3 schlieren; rgba = W*H pixels
int tau_hypc_render(tau_hypc *h, int view_mode, uint32_t *rgba, double minmax_out[2]) {
  TAU_REQUIRE(h && rgba, "tau_hypc_render: null argument");
  TAU_REQUIRE(view_mode >= 0 && view_mode <= 3, "tau_hypc_render: view_mode must be 0..3");
  TAU_REQUIRE(h->have_state, "tau_hypc_render: no state");
  TAU_CUDA(cudaSetDevice(h->device));
  if (!h->pixels) TAU_CUDA(cudaMalloc(&h->pixels, h->N * sizeof(uint32_t)));
  const HPar P = make_par(h);
  const dim3 grid((P.W + 127) / 128, P.H);
  const double *u = h->U[h->cur];
  const unsigned long long init[2] = {~0ull, 0ull};
  TAU_CUDA(cudaMemcpyAsync(h->keys, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
  hypc_render_minmax<<<grid, 128, 0, h->stream>>>(P, u, u + h->N, u + 2 * h->N, u + 3 * h->N, h->mask, view_mode, h->keys);
  unsigned long long k[2];
  TAU_CUDA(cudaMemcpyAsync(k, h->keys, sizeof(k), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  double minv = 1e300, maxv = -1e300;  // :713 (no fluid cell at all)
  if (k[1] != 0ull) {
    minv = okey_inv(k[0]);
    maxv = okey_inv(k[1]);
  }
  hypc_render_pixels<<<grid, 128, 0, h->stream>>>(P, u, u + h->N, u + 2 * h->N, u + 3 * h->N, h->mask, view_mode, minv,
                                                  maxv, h->pixels);
  h->launches += 2;
  TAU_CUDA(cudaGetLastError());
  TAU_CUDA(cudaMemcpyAsync(rgba, h->pixels, h->N * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  if (minmax_out) {
    minmax_out[0] = minv;
    minmax_out[1] = maxv;
  }
  return TAU_OK;
}

int tau_hypc_sync(tau_hypc *h) {
  TAU_REQUIRE(h, "tau_hypc_sync: null handle");
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}
long long tau_hypc_steps_done(tau_hypc *h) { return h ? h->steps : -1; }
long long tau_hypc_launch_count(tau_hypc *h) { return h ? h->launches : -1; }
int tau_hypc_last_step_ms(tau_hypc *h, float *ms) {
  TAU_REQUIRE(h && ms, "tau_hypc_last_step_ms: null argument");
  TAU_REQUIRE(h->timed, "tau_hypc_last_step_ms: no step has been timed yet");
  TAU_CUDA(cudaEventSynchronize(h->ev1));
  TAU_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return TAU_OK;
}
int tau_hypc_destroy(tau_hypc *h) {
  if (!h) return TAU_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->pixels) cudaFree(h->pixels);
  cudaFree(h->keys);
  cudaFree(h->ctrl);
  cudaFree(h->mask);
  cudaFree(h->pred);
  cudaFree(h->U[1]);
  cudaFree(h->U[0]);
  cudaEventDestroy(h->ev1);
  cudaEventDestroy(h->ev0);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return TAU_OK;
}

}  // extern "C"
