// hypersonic3d.cu — 3-D compressible flow with vibrational relaxation (WENO5 + HLLC/HLL blend)
// update path for sm_100a.  Replaces the per-step host sequence of the reference `tau3d`
// (tau_hypersonic_3d_cuda.cu:1678-1712: host clock -> 4-byte H2D -> k_step -> 4-byte D2H -> host
// d_tau controller -> 6 pointer swaps) and its k_step (:987-1359).
//
// What is restructured relative to the reference kernel:
//   * every face flux is computed ONCE.  The reference evaluates 6 faces per cell (each face twice,
//     once from either side; 72 WENO5 evaluations + 6 Riemann solves per cell).  Here a CTA stages
//     its 14x14x10 halo tile as primitives in shared memory (same tile as the reference), then its
//     threads sweep the tile's 896 faces — ordered so that every warp works on one axis — and park
//     the fluxes in shared memory; the cell update gathers its six.  3.5 faces per cell instead of 6.
//   * the tile is LOADED, not decoded: the update tail writes decode(new state) next to the state
//     (two float4 per cell incl. the solid flag) and the next step's tile builds read that instead of
//     running __expf x3 / sinhf x3 on each of their 7.7x-amplified halo cells (see hyp3d_step).
//   * the WENO5 nonlinear weights use one reciprocal per reconstruction instead of six IEEE
//     divisions (algebraically identical: w_k = c_k prod_{j!=k} (eps+b_j)^2 / sum), the rest of the
//     arithmetic keeps the reference's expression trees and its fast intrinsics (__expf/__logf).
//   * the log-time clock, inflow ramp and d_tau controller (:1680-1704) live on the device: a
//     one-thread controller kernel follows each step; no host synchronisation per step.
//   * planes carry 3 ghost planes in z on either side so that z-slabs (multi-GPU ring) need no
//     special casing; a single-GPU handle wraps the plane index instead.
//
// This solver is compute-bound by more than an order of magnitude (885 MUFU and
// ~2e4 instructions per cell in the reference); the HBM roofline fraction is reported as measured.
#include "common.cuh"
#include "../../include/tau_b200.h"

#include <math.h>
#include <stdlib.h>
#include <new>
#include <vector>

// The packed WENO5 pair (weno5_pair below) is the default since its first hardware run (round 2: all 16 parity
// tests of tests/test_hyp3d_gpu.py green, 3.29 -> 3.01 ms/step at 256^3); -DT3_SCALAR_WENO keeps the scalar form.
#if !defined(T3_SCALAR_WENO) && !defined(T3_PACKED_WENO)
#define T3_PACKED_WENO
#endif

namespace {

#ifndef T3_TILE_Z
#define T3_TILE_Z 4
#endif
constexpr int T3_TX = 8, T3_TY = 8, T3_TZ = T3_TILE_Z, T3_H = 3;
constexpr int T3_SX = T3_TX + 2 * T3_H, T3_SY = T3_TY + 2 * T3_H, T3_SZ = T3_TZ + 2 * T3_H;
constexpr int T3_SXY = T3_SX * T3_SY, T3_SVOL = T3_SXY * T3_SZ;
constexpr int T3_THREADS = T3_TX * T3_TY * T3_TZ;
constexpr int T3_STAGE_BATCH = T3_SZ / 2;  // tile planes whose loads one thread keeps in flight while staging
static_assert(T3_SZ % T3_STAGE_BATCH == 0, "whole batches");
constexpr int T3_NFX = (T3_TX + 1) * T3_TY * T3_TZ;  // 288 = 9 warps
constexpr int T3_NFY = T3_TX * (T3_TY + 1) * T3_TZ;  // 288
constexpr int T3_NFZ = T3_TX * T3_TY * (T3_TZ + 1);  // 320
constexpr int T3_NF = T3_NFX + T3_NFY + T3_NFZ;
static_assert(T3_NFX % 32 == 0 && T3_NFY % 32 == 0 && T3_NFZ % 32 == 0, "one axis per warp");

constexpr float RHO_P_FLOOR = 1e-30f;  // tau_hypersonic_3d_cuda.cu:52-58
constexpr float THERMAL_ENERGY_FLOOR = 1e-12f;
constexpr float DENOM_EPS = 1e-12f;
constexpr float NEWTON_TEMP_FLOOR = 1e-6f;
constexpr float WENO_EPS = 1e-6f;
constexpr float TAU_VIB_MIN = 1e-9f;

struct Par {  // struct Params :21-42 + slab geometry
  int nx, ny, nz;        // GLOBAL grid
  float dx, dy, dz, cfl, u_ref, R, gamma_floor, Twall, tau_vib, theta_v;
  float sdf_cx, sdf_cy, sdf_cz, sdf_r;
  float inflow_r, inflow_p, inflow_u, inflow_v, inflow_w;
  int sponge_n;
  float sponge_strength;
  int sponge_out_n;
  float sponge_out_strength;
  int z_begin, nz_local, slab;
  size_t plane;  // elements per plane = nx*ny*(nz_local+6)
};

struct Clock {      // device-resident step control (:1635-1636, :1680-1704)
  float t[2], d_tau[2];
  float dt_last, maxs_last;
  float maxs;       // accumulated by the running step (non-negative: integer atomicMax)
  int pad;
};

struct Q { float r, u, v, w, p, ev; };
struct C6 { float r, mx, my, mz, Et, Ev; };

__device__ __forceinline__ float clampf(float x, float a, float b) { return fminf(fmaxf(x, a), b); }
__device__ __forceinline__ float signed_denom_guard(float x) {  // :147
  return copysignf(fmaxf(fabsf(x), DENOM_EPS), x);
}
__device__ __forceinline__ int wrapi(int i, int n) {  // :156
  // callers are at most one period out of range; the integer modulo (~25 instructions, 7 % of the
  // kernel's stall samples) is only kept for grids narrower than the halo
  if (i < 0) i += n;
  else if (i >= n) i -= n;
  if ((unsigned)i >= (unsigned)n) {
    i %= n;
    if (i < 0) i += n;
  }
  return i;
}
// Quotient without the FCHK slow path.  nvcc expands `a / b` into MUFU.RCP + an FFMA refinement
// guarded by FCHK, and FCHK sends a ZERO NUMERATOR to a ~30-instruction subroutine.  Numerators
// are exactly zero all the time here (v = w = 0 in the free stream, flux differences of uniform
// regions, zero jumps in the shock sensor): measured 1.7 slow-path calls per cell, 19 % of all
// executed instructions (profiles/hyp3d_r1_experiments.md).  rcp.approx + one exact-residual
// correction gives the correctly rounded quotient for operands in the normal range (every divisor
// here is floored away from zero) and exactly 0 for a zero numerator.
__device__ __forceinline__ float fdiv(float a, float b) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  const float q = a * r;
  return fmaf(fmaf(-b, q, a), r, q);
}
__device__ __forceinline__ float asinhf_dev(float x) {  // :121-125
  float ax = fabsf(x);
  return copysignf(logf(ax + sqrtf(ax * ax + 1.0f)), x);
}
__device__ __forceinline__ float evib_eq(const Par &P, float T) {  // :206-211
  float a = fdiv(P.theta_v, fmaxf(T, NEWTON_TEMP_FLOOR));
  float ea = __expf(a);
  // (cold gas: __expf overflows to +inf, and (R theta_v) / inf = 0 — fdiv needs a finite divisor)
  float denom = fminf(fmaxf(ea - 1.f, NEWTON_TEMP_FLOOR), 3.0e38f);
  return fdiv(P.R * P.theta_v, denom);
}
__device__ __forceinline__ Q decode(const Par &P, float xi, float phx, float phy, float phz,
                                    float lam, float zet) {  // log_to_prim_fast :213-225
  Q q;
  q.r = __expf(xi);
  q.u = P.u_ref * sinhf(phx);
  q.v = P.u_ref * sinhf(phy);
  q.w = P.u_ref * sinhf(phz);
  q.p = __expf(lam);
  q.ev = __expf(zet);
  return q;
}
__device__ __forceinline__ C6 prim_to_cons(const Par &P, const Q &q) {  // :234-245
  C6 U;
  U.r = q.r;
  U.mx = q.r * q.u;
  U.my = q.r * q.v;
  U.mz = q.r * q.w;
  float ke = 0.5f * (q.u * q.u + q.v * q.v + q.w * q.w);
  float e_th = fdiv(q.p, fmaxf((P.gamma_floor - 1.f) * q.r, RHO_P_FLOOR));
  U.Ev = q.r * q.ev;
  U.Et = q.r * (ke + e_th + q.ev);
  return U;
}
__device__ __forceinline__ float soundspeed(const Par &P, const Q &q) {  // :264
  return sqrtf(fmaxf(fdiv(P.gamma_floor * q.p, q.r), DENOM_EPS));
}
__device__ __forceinline__ C6 axis_flux(const Par &P, const Q &q, int axis) {  // :268-308
  C6 F;
  float un = axis == 0 ? q.u : (axis == 1 ? q.v : q.w);
  float H = fdiv(q.p, q.r) + (0.5f * (q.u * q.u + q.v * q.v + q.w * q.w) + q.ev) +
            fdiv(q.p, fmaxf((P.gamma_floor - 1.f) * q.r, RHO_P_FLOOR));
  F.r = q.r * un;
  F.mx = q.r * q.u * un + (axis == 0 ? q.p : 0.f);
  F.my = q.r * q.v * un + (axis == 1 ? q.p : 0.f);
  F.mz = q.r * q.w * un + (axis == 2 ? q.p : 0.f);
  F.Et = q.r * H * un;
  F.Ev = q.r * q.ev * un;
  return F;
}
__device__ __forceinline__ float entropy_fix_speed(float s, float a_ref) {  // :366-374
  float d = 0.1f * a_ref, as = fabsf(s);
  float sgn = (s >= 0.f) ? 1.f : -1.f;
  float sm = 0.5f * (fdiv(as * as, fmaxf(d, DENOM_EPS)) + d);
  return (as >= d) ? s : sgn * sm;
}

// hllc_flux_axis :383-460 (HLLC blended towards HLL by shock sensor x alignment)
__device__ __forceinline__ C6 hllc_flux_axis(const Par &P, const Q &L, const Q &R, int axis) {
  float aL = soundspeed(P, L), aR = soundspeed(P, R);
  float unL = axis == 0 ? L.u : (axis == 1 ? L.v : L.w);
  float unR = axis == 0 ? R.u : (axis == 1 ? R.v : R.w);
  float sL = fminf(unL - aL, unR - aR), sR = fmaxf(unL + aL, unR + aR);
  float aRef = fmaxf(aL, aR);
  sL = entropy_fix_speed(sL, aRef);
  sR = entropy_fix_speed(sR, aRef);
  C6 FL = axis_flux(P, L, axis);
  if (__all_sync(__activemask(), sL >= 0.f)) return FL;  // supersonic to the right, whole warp
  C6 UL = prim_to_cons(P, L), UR = prim_to_cons(P, R);
  C6 FR = axis_flux(P, R, axis);
  float rL = L.r, rR = R.r, pL = L.p, pR = R.p;
  float denom = signed_denom_guard(rL * (sL - unL) - rR * (sR - unR));
  float sM = fdiv(pR - pL + rL * unL * (sL - unL) - rR * unR * (sR - unR), denom);
  float pStarL = pL + rL * (sL - unL) * (sM - unL);
  float pStarR = pR + rR * (sR - unR) * (sM - unR);
  float pStar = 0.5f * (pStarL + pStarR);
  float vCarb;  // axis_crossflow_speed :318-325
  if (axis == 0) vCarb = (fabsf(L.v) + fabsf(R.v) + fabsf(L.w) + fabsf(R.w)) * 0.5f;
  else if (axis == 1) vCarb = (fabsf(L.u) + fabsf(R.u) + fabsf(L.w) + fabsf(R.w)) * 0.5f;
  else vCarb = (fabsf(L.u) + fabsf(R.u) + fabsf(L.v) + fabsf(R.v)) * 0.5f;
  float align = clampf(1.f - fdiv(vCarb, fmaxf(aRef, DENOM_EPS)), 0.f, 1.f);
  float dp = fdiv(fabsf(R.p - L.p), fmaxf(R.p + L.p, DENOM_EPS));  // shock_sensor :376-381
  float dr = fdiv(fabsf(R.r - L.r), fmaxf(R.r + L.r, DENOM_EPS));
  float alpha = clampf(5.f * (0.5f * (dp + dr)), 0.f, 1.f) * align;
  float inv_hll = fdiv(1.f, signed_denom_guard(sR - sL)), ss = sL * sR;
  C6 FHLL;
  FHLL.r = ((FL.r * sR - FR.r * sL) + (UR.r - UL.r) * ss) * inv_hll;
  FHLL.mx = ((FL.mx * sR - FR.mx * sL) + (UR.mx - UL.mx) * ss) * inv_hll;
  FHLL.my = ((FL.my * sR - FR.my * sL) + (UR.my - UL.my) * ss) * inv_hll;
  FHLL.mz = ((FL.mz * sR - FR.mz * sL) + (UR.mz - UL.mz) * ss) * inv_hll;
  FHLL.Et = ((FL.Et * sR - FR.Et * sL) + (UR.Et - UL.Et) * ss) * inv_hll;
  FHLL.Ev = ((FL.Ev * sR - FR.Ev * sL) + (UR.Ev - UL.Ev) * ss) * inv_hll;
  const bool left = sM >= 0.f;
  const Q &K = left ? L : R;
  const C6 &UK = left ? UL : UR;
  const C6 &FK = left ? FL : FR;
  float sK = left ? sL : sR, unK = left ? unL : unR;
  float starDenom = signed_denom_guard(sK - sM);
  float rStar = fdiv(K.r * (sK - unK), starDenom);
  float EStar = fdiv((sK - unK) * UK.Et - K.p * unK + pStar * sM, starDenom);
  float EvStar = fdiv(UK.Ev * (sK - unK), starDenom);
  C6 US;  // fill_star_momentum :335-350
  US.r = rStar;
  US.mx = rStar * (axis == 0 ? sM : K.u);
  US.my = rStar * (axis == 1 ? sM : K.v);
  US.mz = rStar * (axis == 2 ? sM : K.w);
  US.Et = EStar;
  US.Ev = EvStar;
  const float om = 1.f - alpha;
  C6 F;
  F.r = (FK.r + (US.r - UK.r) * sK) * om + FHLL.r * alpha;
  F.mx = (FK.mx + (US.mx - UK.mx) * sK) * om + FHLL.mx * alpha;
  F.my = (FK.my + (US.my - UK.my) * sK) * om + FHLL.my * alpha;
  F.mz = (FK.mz + (US.mz - UK.mz) * sK) * om + FHLL.mz * alpha;
  F.Et = (FK.Et + (US.Et - UK.Et) * sK) * om + FHLL.Et * alpha;
  F.Ev = (FK.Ev + (US.Ev - UK.Ev) * sK) * om + FHLL.Ev * alpha;
  if (sL >= 0.f) F = FL;        // :400-403
  else if (sR <= 0.f) F = FR;
  return F;
}

// weno5_left :534-558 with the three weight divisions and the three normalisations folded into
// one reciprocal: w_k = a_k / sum a_j with a_k = c_k / e_k^2  ==  c_k prod_{j != k} e_j^2 / (...).
__device__ __forceinline__ float weno5_left(float v0, float v1, float v2, float v3, float v4) {
  float p0 = (2.f * v0 - 7.f * v1 + 11.f * v2) * (1.f / 6.f);
  float p1 = (-1.f * v1 + 5.f * v2 + 2.f * v3) * (1.f / 6.f);
  float p2 = (2.f * v2 + 5.f * v3 - 1.f * v4) * (1.f / 6.f);
  float t0 = v0 - 2.f * v1 + v2, u0 = v0 - 4.f * v1 + 3.f * v2;
  float t1 = v1 - 2.f * v2 + v3, u1 = v1 - v3;
  float t2 = v2 - 2.f * v3 + v4, u2 = 3.f * v2 - 4.f * v3 + v4;
  float b0 = (13.f / 12.f) * t0 * t0 + 0.25f * u0 * u0;
  float b1 = (13.f / 12.f) * t1 * t1 + 0.25f * u1 * u1;
  float b2 = (13.f / 12.f) * t2 * t2 + 0.25f * u2 * u2;
  float e0 = (WENO_EPS + b0) * (WENO_EPS + b0);
  float e1 = (WENO_EPS + b1) * (WENO_EPS + b1);
  float e2 = (WENO_EPS + b2) * (WENO_EPS + b2);
  float n0 = 0.1f * (e1 * e2), n1 = 0.6f * (e0 * e2), n2 = 0.3f * (e0 * e1);
  return fdiv(n0 * p0 + n1 * p1 + n2 * p2, n0 + n1 + n2);
}
#ifdef T3_PACKED_WENO
// Both one-sided reconstructions of a face use the same six values: L = weno5_left(v0..v4) and
// R = weno5_left(v5..v1).  Evaluated side by side as the two halves of a float2 they run on sm_100's
// packed FADD2/FMUL2/FFMA2 (the body is pure add/mul/fma apart from one reciprocal): same expression
// trees with the contractions spelled out.  Static effect: profiles/hyp3d_r1_experiments.md; measured:
// profiles/r2_first_hw_run.md.
__device__ __forceinline__ float2 weno5_pair(float v0, float v1, float v2, float v3, float v4, float v5) {
  const float2 a0 = make_float2(v0, v5), a1 = make_float2(v1, v4), a2 = make_float2(v2, v3),
               a3 = make_float2(v3, v2), a4 = make_float2(v4, v1);
  auto C = [](float c) { return make_float2(c, c); };
  auto fma2 = [](float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); };
  auto mul2 = [](float2 a, float2 b) { return __fmul2_rn(a, b); };
  auto add2 = [](float2 a, float2 b) { return __fadd2_rn(a, b); };
  const float2 p0 = mul2(fma2(C(11.f), a2, fma2(C(-7.f), a1, mul2(C(2.f), a0))), C(1.f / 6.f));
  const float2 p1 = mul2(fma2(C(2.f), a3, fma2(C(5.f), a2, mul2(C(-1.f), a1))), C(1.f / 6.f));
  const float2 p2 = mul2(fma2(C(-1.f), a4, fma2(C(5.f), a3, mul2(C(2.f), a2))), C(1.f / 6.f));
  const float2 t0 = add2(fma2(C(-2.f), a1, a0), a2), u0 = fma2(C(3.f), a2, fma2(C(-4.f), a1, a0));
  const float2 t1 = add2(fma2(C(-2.f), a2, a1), a3), u1 = fma2(C(-1.f), a3, a1);
  const float2 t2 = add2(fma2(C(-2.f), a3, a2), a4), u2 = add2(fma2(C(-4.f), a3, mul2(C(3.f), a2)), a4);
  const float2 b0 = fma2(mul2(C(0.25f), u0), u0, mul2(mul2(C(13.f / 12.f), t0), t0));
  const float2 b1 = fma2(mul2(C(0.25f), u1), u1, mul2(mul2(C(13.f / 12.f), t1), t1));
  const float2 b2 = fma2(mul2(C(0.25f), u2), u2, mul2(mul2(C(13.f / 12.f), t2), t2));
  const float2 g0 = add2(C(WENO_EPS), b0), g1 = add2(C(WENO_EPS), b1), g2 = add2(C(WENO_EPS), b2);
  const float2 e0 = mul2(g0, g0), e1 = mul2(g1, g1), e2 = mul2(g2, g2);
  const float2 n0 = mul2(C(0.1f), mul2(e1, e2)), n1 = mul2(C(0.6f), mul2(e0, e2)), n2 = mul2(C(0.3f), mul2(e0, e1));
  const float2 num = fma2(n2, p2, fma2(n1, p1, mul2(n0, p0))), den = add2(add2(n0, n1), n2);
  return make_float2(fdiv(num.x, den.x), fdiv(num.y, den.y));
}
#endif
__device__ __forceinline__ void prim_floor_fast(Q &q) {  // :565-571
  q.r = fmaxf(q.r, RHO_P_FLOOR);
  q.p = fmaxf(q.p, RHO_P_FLOOR);
  q.ev = fmaxf(q.ev, 0.f);
}
__device__ __forceinline__ Q inflow_prim(const Par &P) {  // :611-622
  Q q;
  q.r = fmaxf(P.inflow_r, RHO_P_FLOOR);
  q.u = P.inflow_u;
  q.v = P.inflow_v;
  q.w = P.inflow_w;
  q.p = fmaxf(P.inflow_p, RHO_P_FLOOR);
  q.ev = evib_eq(P, fdiv(q.p, q.r * P.R));
  return q;
}
__device__ __forceinline__ void apply_wall(const Par &P, Q &q) {  // :511-521
  float p_keep = fmaxf(q.p, RHO_P_FLOOR);
  q.u = q.v = q.w = 0.f;
  q.p = p_keep;
  q.r = fmaxf(fdiv(q.p, P.R * fmaxf(P.Twall, NEWTON_TEMP_FLOOR)), RHO_P_FLOOR);
  q.ev = evib_eq(P, P.Twall);
}
__device__ __forceinline__ float sdf_sphere(const Par &P, float x, float y, float z) {  // :173-178
  float dx = x - P.sdf_cx, dy = y - P.sdf_cy, dz = z - P.sdf_cz;
  return sqrtf(dx * dx + dy * dy + dz * dz) - P.sdf_r;
}

// plane index (incl. the 3 ghost planes) of local plane lz: slabs read real ghost planes, a
// single-GPU handle wraps (z is periodic, :1031)
__device__ __forceinline__ int zplane(const Par &P, int lz) {
  return (P.slab ? min(max(lz, -T3_H), P.nz_local + T3_H - 1) : wrapi(lz, P.nz_local)) + T3_H;
}

// prim_at_xbc :724-749 for grid column gx (any value), row gy (already wrapped) and LOCAL plane
// glz: inflow state for x < 0, transmissive outflow (:691-722) for x >= nx, the decoded cell
// otherwise; solid cells (mask inside the grid, analytic sphere outside it, :186-188) are forced to
// the isothermal no-slip wall state.
__device__ __forceinline__ Q halo_prim(const Par &P, const float *__restrict__ in, const float4 *__restrict__ pin,
                                       const uint8_t *__restrict__ solid, int gx, int gy, int glz,
                                       bool &is_solid) {
  const size_t PL = P.plane;
  const int nxy = P.nx * P.ny;
  const int pz = zplane(P, glz);
  // Decoded state of cell gi.  `pin` (may be null) holds decode(in) for the slab's own planes, written by the step that
  // produced `in` (see the update tail): the same function of the same six numbers, evaluated once per cell instead of once
  // per tile that stages it (7.7 tiles).  Ghost planes of a slab arrive encoded from the neighbours and are decoded here.
  auto fetch = [&](size_t gi) -> Q {
    if (pin != nullptr && !(P.slab && (pz < T3_H || pz >= P.nz_local + T3_H))) {
      const float4 a = pin[2 * gi], b = pin[2 * gi + 1];
      return Q{a.x, a.y, a.z, a.w, b.x, b.y};
    }
    return decode(P, in[gi], in[PL + gi], in[2 * PL + gi], in[3 * PL + gi], in[4 * PL + gi], in[5 * PL + gi]);
  };
  Q q;
  if (gx < 0 || gx >= P.nx) {
    const int gz = wrapi(P.z_begin + glz, P.nz);
    is_solid = sdf_sphere(P, (gx + 0.5f) * P.dx, (gy + 0.5f) * P.dy, (gz + 0.5f) * P.dz) < 0.f;
    if (gx < 0) {
      q = inflow_prim(P);
    } else {
      q = fetch((size_t)pz * nxy + (size_t)gy * P.nx + (P.nx - 1));
      const float aR = soundspeed(P, q), un = q.u;
      if (un < 0.0f) {
        q = inflow_prim(P);
      } else {
        if (un < aR) {
          const float p_amb = fmaxf(P.inflow_p, RHO_P_FLOOR);
          q.p = fmaxf(q.p + 0.05f * (p_amb - q.p), RHO_P_FLOOR);
        }
        prim_floor_fast(q);
      }
    }
  } else {
    const size_t gi = (size_t)pz * nxy + (size_t)gy * P.nx + gx;
    is_solid = solid[gi] != 0;
    q = fetch(gi);
  }
  if (is_solid) apply_wall(P, q);
  return q;
}

// out of line: the step kernel's staging loop reaches it for x-boundary columns, slab ghost planes and the first step
// after an upload only, and its decode / boundary-state code would otherwise sit in the middle of the hot loop
struct QS { Q q; int solid; };
__device__ __noinline__ QS halo_prim_cold(const Par &P, const float *__restrict__ in, const float4 *__restrict__ pin,
                                          const uint8_t *__restrict__ solid, int gx, int gy, int glz) {
  bool is_solid;
  const Q q = halo_prim(P, in, pin, solid, gx, gy, glz, is_solid);
  return QS{q, is_solid ? 1 : 0};
}

__global__ void __launch_bounds__(T3_THREADS, T3_TZ == 4 ? 3 : 2)   // what shared memory admits per SM: keep the registers to that
hyp3d_step(const __grid_constant__ Par P, const float *__restrict__ in, float *__restrict__ out,
           const float4 *__restrict__ pin, float4 *__restrict__ pout, const uint8_t *__restrict__ solid,
           Clock *__restrict__ clk, int slot) {
  extern __shared__ __align__(16) unsigned char smem[];
  float *s_q = reinterpret_cast<float *>(smem);                 // [6][SVOL] r,u,v,w,p,ev
  float *s_f = s_q + 6 * T3_SVOL;                                // [6][NF]   face fluxes
  uint8_t *s_solid = reinterpret_cast<uint8_t *>(s_f + 6 * T3_NF);

  // ---- log-time clock, :1680-1683 -------------------------------------------------------------
  const float d_tau = clk->d_tau[slot];
  const float t = clk->t[slot] * expf(d_tau);
  const float dt = t * d_tau;
  const float inflow_gain = fminf(fmaxf(t / 0.02f, 0.f), 1.f);

  const int tid = (threadIdx.z * T3_TY + threadIdx.y) * T3_TX + threadIdx.x;
  const int bx0 = blockIdx.x * T3_TX, by0 = blockIdx.y * T3_TY, bz0 = blockIdx.z * T3_TZ;
  const size_t PL = P.plane;
  const int nxy = P.nx * P.ny;

  // ---- stage the halo tile as primitives (k_step :1019-1056) ------------------------------------
  // Half an (x, y) column of the tile per thread (T3_STAGE_BATCH planes): the column's grid position, its periodic wrap
  // in y and its x-boundary class are computed once, a plane is one z-plane index and two 16-byte loads.
  // (Round 2 measured the flat `for tt < T3_SVOL` form of this loop at 1069 of the kernel's 4373 warp-instructions per
  // 32 cells and 48 % of its stall samples — 140 instructions per staged cell of index arithmetic, with the loads of one
  // cell issued and consumed before the next cell's: profiles/hyp3d_step_r2b_ncu_full.txt.)
  for (int unit = tid; unit < T3_SXY * (T3_SZ / T3_STAGE_BATCH); unit += T3_THREADS) {
    const int part = unit / T3_SXY, col = unit - part * T3_SXY;  // (tile column, batch of planes)
    const int ly = col / T3_SX, lx = col - ly * T3_SX;
    const int gx = bx0 + lx - T3_H, gy = wrapi(by0 + ly - T3_H, P.ny);
    const bool col_in_grid = (unsigned)gx < (unsigned)P.nx && pin != nullptr;  // no x-boundary state, primitives at hand
    const size_t col_off = (size_t)gy * P.nx + (size_t)max(gx, 0);
    // planes in batches of T3_STAGE_BATCH: all loads of a batch are issued before the first is consumed (the staging
    // phase waits on memory, not on issue slots: measured 33 % of the kernel's stall samples with one plane in flight)
    {
      const int lz0 = part * T3_STAGE_BATCH;
      float4 a[T3_STAGE_BATCH], b[T3_STAGE_BATCH];
      bool hot[T3_STAGE_BATCH];
#pragma unroll
      for (int k = 0; k < T3_STAGE_BATCH; ++k) {
        const int pz = zplane(P, bz0 + lz0 + k - T3_H);
        hot[k] = col_in_grid && !(P.slab && (pz < T3_H || pz >= P.nz_local + T3_H));  // (== halo_prim's interior branch)
        if (hot[k]) {
          const size_t gi = (size_t)pz * nxy + col_off;
          a[k] = pin[2 * gi];      // (r, u, v, w)
          b[k] = pin[2 * gi + 1];  // (p, ev, solid flag, -)
        }
      }
#pragma unroll
      for (int k = 0; k < T3_STAGE_BATCH; ++k) {
        const int lz = lz0 + k;
        const int tt = lz * T3_SXY + col;
        bool is_solid;
        Q q;
        if (hot[k]) {
          q = Q{a[k].x, a[k].y, a[k].z, a[k].w, b[k].x, b[k].y};
          is_solid = b[k].z != 0.f;
          if (is_solid) apply_wall(P, q);
        } else {
          const QS c = halo_prim_cold(P, in, pin, solid, gx, gy, bz0 + lz - T3_H);
          q = c.q;
          is_solid = c.solid != 0;
        }
        s_q[tt] = q.r;
        s_q[T3_SVOL + tt] = q.u;
        s_q[2 * T3_SVOL + tt] = q.v;
        s_q[3 * T3_SVOL + tt] = q.w;
        s_q[4 * T3_SVOL + tt] = q.p;
        s_q[5 * T3_SVOL + tt] = q.ev;
        s_solid[tt] = is_solid ? 1 : 0;
      }
    }
  }
  __syncthreads();

  auto load_q = [&](int j) {
    return Q{s_q[j], s_q[T3_SVOL + j], s_q[2 * T3_SVOL + j], s_q[3 * T3_SVOL + j],
             s_q[4 * T3_SVOL + j], s_q[5 * T3_SVOL + j]};
  };

  // ---- every face of the tile once (k_step :1113-1264 evaluates each from both sides) ---------
  for (int f = tid; f < T3_NF; f += T3_THREADS) {
    int axis, fx, fy, fz;  // (fx,fy,fz): tile coordinates of the cell on the PLUS side of the face
    if (f < T3_NFX) {
      axis = 0;
      fx = f % (T3_TX + 1); fy = (f / (T3_TX + 1)) % T3_TY; fz = f / ((T3_TX + 1) * T3_TY);
    } else if (f < T3_NFX + T3_NFY) {
      const int g = f - T3_NFX;
      axis = 1;
      fx = g % T3_TX; fy = (g / T3_TX) % (T3_TY + 1); fz = g / (T3_TX * (T3_TY + 1));
    } else {
      const int g = f - T3_NFX - T3_NFY;
      axis = 2;
      fx = g % T3_TX; fy = (g / T3_TX) % T3_TY; fz = g / (T3_TX * T3_TY);
    }
    const int stride = axis == 0 ? 1 : (axis == 1 ? T3_SX : T3_SXY);
    const int jb = ((fz + T3_H) * T3_SY + (fy + T3_H)) * T3_SX + (fx + T3_H);  // plus-side cell
    const int ja = jb - stride;                                                // minus-side cell
    const bool sa = s_solid[ja] != 0, sb = s_solid[jb] != 0;
    Q L, R;
    if (sa || sb) {  // face touches a solid: mirror state of the fluid side (:1125-1128, :1146-1149)
      if (!sb) {
        R = load_q(jb);
        L = R;
        if (axis == 0) L.u = -L.u; else if (axis == 1) L.v = -L.v; else L.w = -L.w;
      } else {
        L = load_q(ja);  // both solid: flux unused, any finite state will do
        R = L;
        if (axis == 0) R.u = -R.u; else if (axis == 1) R.v = -R.v; else R.w = -R.w;
      }
    } else {
      const bool stencil_solid = s_solid[ja - 2 * stride] | s_solid[ja - stride] |
                                 s_solid[jb + stride] | s_solid[jb + 2 * stride];
      if (stencil_solid) {  // first order next to the body (:1129-1135)
        L = load_q(ja);
        R = load_q(jb);
      } else {  // weno_face_from_6 :578-598 on cells a-2 .. b+2
        const int j0 = ja - 2 * stride;
#ifdef T3_PACKED_WENO
#define WENO_LR(l, r) { const float2 lr = weno5_pair(v0, v1, v2, v3, v4, v5); l = lr.x; r = lr.y; }
#else
#define WENO_LR(l, r) { l = weno5_left(v0, v1, v2, v3, v4); r = weno5_left(v5, v4, v3, v2, v1); }
#endif
#define WENO_FIELD(k, fld)                                                                       \
  {                                                                                              \
    const float *s = s_q + (k) * T3_SVOL + j0;                                                   \
    const float v0 = s[0], v1 = s[stride], v2 = s[2 * stride], v3 = s[3 * stride],               \
                v4 = s[4 * stride], v5 = s[5 * stride];                                          \
    WENO_LR(L.fld, R.fld)                                                                        \
  }
        WENO_FIELD(0, r) WENO_FIELD(1, u) WENO_FIELD(2, v) WENO_FIELD(3, w) WENO_FIELD(4, p)
        WENO_FIELD(5, ev)
#undef WENO_FIELD
#undef WENO_LR
      }
      prim_floor_fast(L);
      prim_floor_fast(R);
    }
    const C6 F = hllc_flux_axis(P, L, R, axis);
    s_f[f] = F.r;
    s_f[T3_NF + f] = F.mx;
    s_f[2 * T3_NF + f] = F.my;
    s_f[3 * T3_NF + f] = F.mz;
    s_f[4 * T3_NF + f] = F.Et;
    s_f[5 * T3_NF + f] = F.Ev;
  }
  __syncthreads();

  // ---- conservative update, relaxation, sponges, re-encode (k_step :1266-1358) ----------------
  const int x = bx0 + threadIdx.x, y = by0 + threadIdx.y, lz = bz0 + threadIdx.z;
  float ssum = 0.f;
  if (x < P.nx && y < P.ny && lz < P.nz_local) {
    const size_t i = (size_t)(lz + T3_H) * nxy + (size_t)y * P.nx + x;
    if (solid[i]) {
      float e[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) out[k * PL + i] = e[k] = in[k * PL + i];
      if (pout != nullptr) {
        const Q d = decode(P, e[0], e[1], e[2], e[3], e[4], e[5]);
        pout[2 * i] = make_float4(d.r, d.u, d.v, d.w);
        pout[2 * i + 1] = make_float4(d.p, d.ev, 1.f, 0.f);
      }
    } else {
      const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
      const int fxm = (tz * T3_TY + ty) * (T3_TX + 1) + tx, fxp = fxm + 1;
      const int fym = T3_NFX + (tz * (T3_TY + 1) + ty) * T3_TX + tx, fyp = fym + T3_TX;
      const int fzm = T3_NFX + T3_NFY + (tz * T3_TY + ty) * T3_TX + tx, fzp = fzm + T3_TX * T3_TY;
      const int jc = ((tz + T3_H) * T3_SY + (ty + T3_H)) * T3_SX + (tx + T3_H);
      const Q q0 = load_q(jc);
      const C6 U0 = prim_to_cons(P, q0);
      float U1[6];
      const float U0a[6] = {U0.r, U0.mx, U0.my, U0.mz, U0.Et, U0.Ev};
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const float *sf = s_f + k * T3_NF;
        const float dU = -(fdiv(sf[fxp] - sf[fxm], P.dx) + fdiv(sf[fyp] - sf[fym], P.dy) +
                           fdiv(sf[fzp] - sf[fzm], P.dz));
        U1[k] = U0a[k] + dU * dt;
      }
      // cons_to_prim :247-262
      Q q1;
      q1.r = fmaxf(U1[0], RHO_P_FLOOR);
      q1.u = fdiv(U1[1], q1.r);
      q1.v = fdiv(U1[2], q1.r);
      q1.w = fdiv(U1[3], q1.r);
      {
        const float ke = 0.5f * (q1.u * q1.u + q1.v * q1.v + q1.w * q1.w);
        const float ev = fmaxf(fdiv(U1[5], q1.r), 0.f);
        const float e_th = fmaxf(fdiv(U1[4], q1.r) - ke - ev, THERMAL_ENERGY_FLOOR);
        q1.p = fmaxf((P.gamma_floor - 1.f) * q1.r * e_th, RHO_P_FLOOR);
        q1.ev = ev;
      }
      if (!isfinite(q1.r) || !isfinite(q1.p) || !isfinite(q1.u) || !isfinite(q1.v) ||
          !isfinite(q1.w) || !isfinite(q1.ev) || q1.r <= 0.f || q1.p <= 0.f || q1.ev < 0.f)
        q1 = inflow_prim(P);  // :1284-1289
      float T1 = fdiv(q1.p, q1.r * P.R);
      q1.ev = fmaxf(q1.ev + (evib_eq(P, T1) - q1.ev) * fdiv(dt, fmaxf(P.tau_vib, TAU_VIB_MIN)), 0.f);
      const float tr = fmaxf(P.inflow_r, RHO_P_FLOOR), tp = fmaxf(P.inflow_p, RHO_P_FLOOR);
      const int nsp = P.sponge_n > 0 ? P.sponge_n : 0;
      if (nsp > 0 && x < nsp) {  // inflow sponge :1295-1318
        float s = 1.0f - (float)x / (float)nsp;
        s = fminf(fmaxf(s, 0.0f), 1.0f);
        const float k = P.sponge_strength * (s * s);
        const float tev = evib_eq(P, fdiv(tp, tr * P.R));
        q1.r = fmaxf(q1.r + k * (tr - q1.r), RHO_P_FLOOR);
        q1.p = fmaxf(q1.p + k * (tp - q1.p), RHO_P_FLOOR);
        q1.u = q1.u + k * (inflow_gain * P.inflow_u - q1.u);
        q1.v = q1.v + k * (inflow_gain * P.inflow_v - q1.v);
        q1.w = q1.w + k * (inflow_gain * P.inflow_w - q1.w);
        q1.ev = fmaxf(q1.ev + k * (tev - q1.ev), 0.f);
      }
      const int nspo = P.sponge_out_n > 0 ? P.sponge_out_n : 0;
      if (nspo > 0 && x >= (P.nx - nspo)) {  // outflow sponge :1319-1343
        float s = (float)(x - (P.nx - nspo)) / (float)nspo;
        s = fminf(fmaxf(s, 0.0f), 1.0f);
        const float k = P.sponge_out_strength * (s * s);
        const float tev = evib_eq(P, fdiv(tp, tr * P.R));
        q1.r = fmaxf(q1.r + k * (tr - q1.r), RHO_P_FLOOR);
        q1.p = fmaxf(q1.p + k * (tp - q1.p), RHO_P_FLOOR);
        q1.u = q1.u + k * (0.0f - q1.u);
        q1.v = q1.v + k * (0.0f - q1.v);
        q1.w = q1.w + k * (0.0f - q1.w);
        q1.ev = fmaxf(q1.ev + k * (tev - q1.ev), 0.f);
      }
      const float a = soundspeed(P, q1);  // :1345-1351
      const float sw = fdiv(fabsf(q1.u) + a, P.dx) + fdiv(fabsf(q1.v) + a, P.dy) + fdiv(fabsf(q1.w) + a, P.dz);
      if (isfinite(sw) && sw > 0.f) ssum = sw;
      const float e0 = __logf(fmaxf(q1.r, RHO_P_FLOOR));  // :1353-1358
      const float e1 = asinhf_dev(fdiv(q1.u, P.u_ref));
      const float e2 = asinhf_dev(fdiv(q1.v, P.u_ref));
      const float e3 = asinhf_dev(fdiv(q1.w, P.u_ref));
      const float e4 = __logf(fmaxf(q1.p, RHO_P_FLOOR));
      const float e5 = __logf(fmaxf(q1.ev, RHO_P_FLOOR));
      out[i] = e0;
      out[PL + i] = e1;
      out[2 * PL + i] = e2;
      out[3 * PL + i] = e3;
      out[4 * PL + i] = e4;
      out[5 * PL + i] = e5;
      // what the NEXT step's tile builds would each recompute from these six numbers (the reference decodes the stored
      // logs, :213-225, so the round trip is part of the algorithm: q1 itself must not be handed on)
      if (pout != nullptr) {
        const Q d = decode(P, e0, e1, e2, e3, e4, e5);
        pout[2 * i] = make_float4(d.r, d.u, d.v, d.w);
        pout[2 * i + 1] = make_float4(d.p, d.ev, 0.f, 0.f);
      }
    }
  }
  ssum = tau::warp_max(ssum);
  if ((tid & 31) == 0 && ssum > 0.f) tau::atomic_max_nonneg(&clk->maxs, ssum);
}

// host controller :1680-1704 as a one-thread kernel: commits t, adapts d_tau from the max wavespeed
// the step just measured, arms the other clock slot, clears the accumulator.
__global__ void hyp3d_controller(const Par P, Clock *clk, int slot) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float d_tau = clk->d_tau[slot];
  const float t = clk->t[slot] * expf(d_tau);
  const float dt = t * d_tau;
  const float maxs = clk->maxs;
  const float dt_cfl = P.cfl / fmaxf(maxs, 1e-9f);
  if (dt > 1.10f * dt_cfl) d_tau *= 0.80f;
  else if (dt < 0.85f * dt_cfl) d_tau *= 1.10f;
  d_tau = fminf(fmaxf(d_tau, 1e-7f), 5e-2f);
  clk->t[slot ^ 1] = t;
  clk->d_tau[slot ^ 1] = d_tau;
  clk->dt_last = dt;
  clk->maxs_last = maxs;
  clk->maxs = 0.f;
}

// k_build_solid_mask :759-770 for every plane of the slab incl. the 3 ghost planes on either side
__global__ void hyp3d_build_solid(const Par P, uint8_t *solid) {
  const size_t n = P.plane;
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % P.nx), y = (int)((i / P.nx) % P.ny), pz = (int)(i / ((size_t)P.nx * P.ny));
  const int gz = wrapi(P.z_begin + pz - T3_H, P.nz);
  solid[i] = sdf_sphere(P, (x + 0.5f) * P.dx, (y + 0.5f) * P.dy, (gz + 0.5f) * P.dz) < 0.f ? 1 : 0;
}

// k_init :939-985 (ghost planes included: the initial state is z-periodic by construction)
__global__ void hyp3d_init(const Par P, float *st, const uint8_t *solid) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= P.plane) return;
  float r = fmaxf(P.inflow_r, RHO_P_FLOOR), p = fmaxf(P.inflow_p, RHO_P_FLOOR);
  float T = p / (r * P.R);
  float ev = evib_eq(P, T);
  if (solid[i]) {
    T = P.Twall;
    r = fmaxf(p / (P.R * fmaxf(T, NEWTON_TEMP_FLOOR)), RHO_P_FLOOR);
    ev = evib_eq(P, T);
  }
  st[i] = __logf(fmaxf(r, RHO_P_FLOOR));
  st[P.plane + i] = asinhf_dev(0.f / P.u_ref);
  st[2 * P.plane + i] = asinhf_dev(0.f / P.u_ref);
  st[3 * P.plane + i] = asinhf_dev(0.f / P.u_ref);
  st[4 * P.plane + i] = __logf(fmaxf(p, RHO_P_FLOOR));
  st[5 * P.plane + i] = __logf(fmaxf(ev, RHO_P_FLOOR));
}


// ---- diagnostic scalar field: k_vis :800-905 (the volume renderer's input) ---------------------------
// The reference decodes a cell and its six neighbours per thread (7 x 6 transcendentals).  Here a CTA
// decodes its 8x8x4 tile + a one-cell halo once into shared memory (600 decodes for 256 cells) with
// the same boundary rule as the step kernel's tile build (prim_at_xbc), then every thread evaluates
// the requested quantity from shared memory.  Modes: VisMode :784-794.
constexpr int V3_SX = T3_TX + 2, V3_SY = T3_TY + 2, V3_SZ = T3_TZ + 2;
constexpr int V3_SXY = V3_SX * V3_SY, V3_SVOL = V3_SXY * V3_SZ;
__global__ void __launch_bounds__(T3_THREADS)
hyp3d_vis(const Par P, const float *__restrict__ in, const uint8_t *__restrict__ solid,
          float *__restrict__ out, int mode) {
  __shared__ float s_r[V3_SVOL], s_u[V3_SVOL], s_v[V3_SVOL], s_w[V3_SVOL], s_p[V3_SVOL];
  const int tid = (threadIdx.z * T3_TY + threadIdx.y) * T3_TX + threadIdx.x;
  const int bx0 = blockIdx.x * T3_TX, by0 = blockIdx.y * T3_TY, bz0 = blockIdx.z * T3_TZ;
  for (int tt = tid; tt < V3_SVOL; tt += T3_THREADS) {
    const int lz = tt / V3_SXY, rem = tt - lz * V3_SXY, ly = rem / V3_SX, lx = rem - ly * V3_SX;
    bool is_solid;
    const Q q = halo_prim(P, in, static_cast<const float4 *>(nullptr), solid, bx0 + lx - 1, wrapi(by0 + ly - 1, P.ny), bz0 + lz - 1, is_solid);
    s_r[tt] = q.r; s_u[tt] = q.u; s_v[tt] = q.v; s_w[tt] = q.w; s_p[tt] = q.p;
  }
  __syncthreads();
  const int x = bx0 + threadIdx.x, y = by0 + threadIdx.y, lz = bz0 + threadIdx.z;
  if (x >= P.nx || y >= P.ny || lz >= P.nz_local) return;
  const size_t i = ((size_t)lz * P.ny + y) * P.nx + x;  // idx3 :152 within the slab
  if (solid[((size_t)(lz + T3_H) * P.ny + y) * P.nx + x]) {
    out[i] = 0.f;
    return;
  }
  const int c = ((threadIdx.z + 1) * V3_SY + threadIdx.y + 1) * V3_SX + threadIdx.x + 1;
  const float r0 = s_r[c], u0 = s_u[c], v0 = s_v[c], w0 = s_w[c], p0 = s_p[c];
  if (mode == 1) { out[i] = logf(1.0f + fmaxf(r0, 0.0f)); return; }   // VIS_LOG_RHO, safe_log1pf_dev :796
  if (mode == 2) { out[i] = logf(1.0f + fmaxf(p0, 0.0f)); return; }   // VIS_LOG_P
  const float sp = sqrtf(u0 * u0 + v0 * v0 + w0 * w0);
  if (mode == 3) { out[i] = sp; return; }                              // VIS_SPEED
  if (mode == 4) {                                                     // VIS_MACH
    const float a = sqrtf(fmaxf(P.gamma_floor * p0 / r0, DENOM_EPS));  // soundspeed :264
    out[i] = sp / fmaxf(a, DENOM_EPS);
    return;
  }
  const int xm = c - 1, xp = c + 1, ym = c - V3_SX, yp = c + V3_SX, zm = c - V3_SXY, zp = c + V3_SXY;
  const float inv2dx = 0.5f / P.dx, inv2dy = 0.5f / P.dy, inv2dz = 0.5f / P.dz;
  const float dudx = (s_u[xp] - s_u[xm]) * inv2dx, dudy = (s_u[yp] - s_u[ym]) * inv2dy, dudz = (s_u[zp] - s_u[zm]) * inv2dz;
  const float dvdx = (s_v[xp] - s_v[xm]) * inv2dx, dvdy = (s_v[yp] - s_v[ym]) * inv2dy, dvdz = (s_v[zp] - s_v[zm]) * inv2dz;
  const float dwdx = (s_w[xp] - s_w[xm]) * inv2dx, dwdy = (s_w[yp] - s_w[ym]) * inv2dy, dwdz = (s_w[zp] - s_w[zm]) * inv2dz;
  if (mode == 6) { out[i] = dudx + dvdy + dwdz; return; }              // VIS_DIV
  const float wx = dwdy - dvdz, wy = dudz - dwdx, wz = dvdx - dudy;
  if (mode == 5) { out[i] = sqrtf(wx * wx + wy * wy + wz * wz); return; }   // VIS_VORT_MAG
  if (mode == 7) {                                                     // VIS_Q_CRITERION :879-899
    const float O12 = 0.5f * (dudy - dvdx), O13 = 0.5f * (dudz - dwdx), O23 = 0.5f * (dvdz - dwdy);
    const float Om2 = 2.0f * (O12 * O12 + O13 * O13 + O23 * O23);
    const float S12 = 0.5f * (dudy + dvdx), S13 = 0.5f * (dudz + dwdx), S23 = 0.5f * (dvdz + dwdy);
    const float Sm2 = (dudx * dudx + dvdy * dvdy + dwdz * dwdz) + 2.0f * (S12 * S12 + S13 * S13 + S23 * S23);
    out[i] = 0.5f * (Om2 - Sm2);
    return;
  }
  if (mode == 8) {  // th3cs.cu k_schlieren_export :641-673: the same quantity, divisions instead of 0.5f/d factors
    const float ex = (s_r[xp] - s_r[xm]) / (2.0f * P.dx), ey = (s_r[yp] - s_r[ym]) / (2.0f * P.dy),
                ez = (s_r[zp] - s_r[zm]) / (2.0f * P.dz);
    out[i] = sqrtf(ex * ex + ey * ey + ez * ez);
    return;
  }
  const float drdx = (s_r[xp] - s_r[xm]) * inv2dx, drdy = (s_r[yp] - s_r[ym]) * inv2dy, drdz = (s_r[zp] - s_r[zm]) * inv2dz;
  out[i] = sqrtf(drdx * drdx + drdy * drdy + drdz * drdz);             // VIS_SCHLIEREN_RHO (mode 0)
}

// ---- .4spl frame export (th3cs.cu main :1193-1222): the reference copies the whole schlieren volume to the
// host (4 B/voxel), takes min/max there and quantises every voxel with powf on the host.  Here: min/max by
// warp shuffle + ordered-integer atomics, the 8-bit palette index on the device, 1 B/voxel over PCIe.
// (int)(powf(norm, 0.65f) * 255.0f) is a monotone step function of norm, so the index is "how many of its
// 255 steps lie at or below norm": the step positions come from the HOST's powf (export_thresholds below),
// which makes the device result identical to the reference's host loop by construction.
__device__ __forceinline__ unsigned f32_key(float f) {      // order-preserving map float -> unsigned
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_unkey(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__global__ void __launch_bounds__(256) hyp3d_export_minmax(const float *__restrict__ v, size_t n, unsigned *keys) {
  float lo = 1e30f, hi = -1e30f;                            // the reference's start values :1200
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    lo = fminf(lo, v[i]);
    hi = fmaxf(hi, v[i]);
  }
  hi = tau::warp_max(hi);
  lo = -tau::warp_max(-lo);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&keys[0], f32_key(lo));
    atomicMax(&keys[1], f32_key(hi));
  }
}
__global__ void __launch_bounds__(256)
hyp3d_export_quantize(const float *__restrict__ v, size_t n, const unsigned *__restrict__ keys,
                      const float *__restrict__ thr, uint8_t *__restrict__ out) {
  __shared__ float s_thr[256];
  s_thr[threadIdx.x] = threadIdx.x ? thr[threadIdx.x - 1] : 0.f;   // s_thr[k] = first norm with index >= k
  __syncthreads();
  const float vmin = f32_unkey(keys[0]), vmax = f32_unkey(keys[1]);
  const float range = fmaxf(vmax - vmin, 1e-12f);                   // :1205
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float norm = (v[i] - vmin) / range;                       // :1212 (IEEE division: this TU has no fast-math)
    int lo = 0, hi = 256;                                           // largest k with s_thr[k] <= norm
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_thr[mid] <= norm) lo = mid; else hi = mid;
    }
    out[i] = (uint8_t)lo;                                           // NaN compares false everywhere -> 0
  }
}

constexpr size_t T3_SMEM = (size_t)(6 * T3_SVOL + 6 * T3_NF) * sizeof(float) + ((T3_SVOL + 15) / 16) * 16;

}  // namespace

struct tau_hyp3d {
  tau_hyp3d_params prm;
  int device, z_begin, nz_local;
  bool slab;
  cudaStream_t stream;
  bool own_stream;
  float *st[2];      // 6 contiguous planes each, (nz_local+6) z-planes
  float4 *pr[2];     // decoded primitives of st[b]'s own planes: two float4 per cell, (r, u, v, w) (p, ev, solid flag, -),
                     // cell index as in st (null: TAU_HYP3D_PRIMS=0), see halo_prim
  bool pr_valid[2];  // pr[b] == decode(st[b]) on the slab's own planes
  uint8_t *solid;
  Clock *clk;
  int cur;
  long long steps, launches;
  size_t plane;
  cudaEvent_t ev0, ev1;
  bool timed, have_state;
  float *vis;        // diagnostic field (device), allocated on first use
  uint8_t *ex_idx;   // .4spl frame export: palette indices, min/max keys, quantiser steps (device)
  unsigned *ex_keys;
  float *ex_thr;
};

namespace {
Par make_par(const tau_hyp3d *h) {
  Par P;
  const tau_hyp3d_params &p = h->prm;
  P.nx = p.nx; P.ny = p.ny; P.nz = p.nz;
  P.dx = p.dx; P.dy = p.dy; P.dz = p.dz; P.cfl = p.cfl; P.u_ref = p.u_ref; P.R = p.R;
  P.gamma_floor = p.gamma_floor; P.Twall = p.Twall; P.tau_vib = p.tau_vib; P.theta_v = p.theta_v;
  P.sdf_cx = p.sdf_cx; P.sdf_cy = p.sdf_cy; P.sdf_cz = p.sdf_cz; P.sdf_r = p.sdf_r;
  P.inflow_r = p.inflow_r; P.inflow_p = p.inflow_p; P.inflow_u = p.inflow_u;
  P.inflow_v = p.inflow_v; P.inflow_w = p.inflow_w;
  P.sponge_n = p.sponge_n; P.sponge_strength = p.sponge_strength;
  P.sponge_out_n = p.sponge_out_n; P.sponge_out_strength = p.sponge_out_strength;
  P.z_begin = h->z_begin; P.nz_local = h->nz_local; P.slab = h->slab ? 1 : 0;
  P.plane = h->plane;
  return P;
}
}  // namespace

extern "C" {

// main()'s hard-coded Params :1531-1557 for an nx x ny x nz grid (the reference uses 64^3)
void tau_hyp3d_default_params(tau_hyp3d_params *p, int nx, int ny, int nz) {
  p->nx = nx; p->ny = ny; p->nz = nz;
  p->dx = 1.f / nx; p->dy = 1.f / ny; p->dz = 1.f / nz;
  p->cfl = 0.3333f; p->u_ref = 10.f; p->R = 10.f; p->gamma_floor = 1.1f; p->Twall = 0.02f;
  p->tau_vib = 2e-4f; p->theta_v = 0.2f;
  p->sdf_cx = 0.5f; p->sdf_cy = 0.5f; p->sdf_cz = 0.5f; p->sdf_r = 0.25f;
  p->inflow_r = 0.02f; p->inflow_p = 0.02f; p->inflow_u = 100.0f; p->inflow_v = 0.0f; p->inflow_w = 0.0f;
  p->sponge_n = 24; p->sponge_strength = 0.05f; p->sponge_out_n = 24; p->sponge_out_strength = 0.05f;
  p->t0 = 1e-5f;      // :1635
  p->d_tau0 = 1e-3f;  // :1636
}

int tau_hyp3d_create(const tau_hyp3d_params *p, int device, int z_begin, int nz_local, void *stream,
                     tau_hyp3d **out) {
  TAU_REQUIRE(p && out, "tau_hyp3d_create: null argument");
  TAU_REQUIRE(p->nx >= 1 && p->ny >= 1 && p->nz >= 1, "tau_hyp3d_create: bad grid %d x %d x %d",
              p->nx, p->ny, p->nz);
  TAU_REQUIRE(z_begin >= 0 && nz_local > 0 && z_begin + nz_local <= p->nz,
              "tau_hyp3d_create: slab planes [%d,%d) outside [0,%d)", z_begin, z_begin + nz_local, p->nz);
  TAU_REQUIRE(nz_local == p->nz || nz_local >= T3_H, "tau_hyp3d_create: a slab needs >= %d planes", T3_H);
  if (tau_device_count() <= 0) {
    tau_set_error("tau_hyp3d_create: no CUDA device (this library has no CPU fallback)");
    return TAU_ERR_NODEV;
  }
  TAU_CUDA(cudaSetDevice(device));
  tau_hyp3d *h = new (std::nothrow) tau_hyp3d();
  if (!h) return TAU_ERR_NOMEM;
  h->prm = *p;
  h->device = device;
  h->z_begin = z_begin;
  h->nz_local = nz_local;
  h->slab = (nz_local != p->nz);
  h->vis = nullptr;
  h->ex_idx = nullptr;
  h->ex_keys = nullptr;
  h->ex_thr = nullptr;
  h->cur = 0;
  h->steps = h->launches = 0;
  h->timed = h->have_state = false;
  if (stream) {
    h->stream = (cudaStream_t)stream;
    h->own_stream = false;
  } else {
    TAU_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  h->plane = (size_t)p->nx * p->ny * (nz_local + 2 * T3_H);
  for (int b = 0; b < 2; ++b) {
    TAU_CUDA(cudaMalloc(&h->st[b], 6 * h->plane * sizeof(float)));
    TAU_CUDA(cudaMemsetAsync(h->st[b], 0, 6 * h->plane * sizeof(float), h->stream));
  }
  h->pr[0] = h->pr[1] = nullptr;
  h->pr_valid[0] = h->pr_valid[1] = false;
  {
    const char *e = getenv("TAU_HYP3D_PRIMS");  // 0: every tile decodes its own halo (the round-1 kernel)
    if (!(e && atoi(e) == 0))
      for (int b = 0; b < 2; ++b) TAU_CUDA(cudaMalloc(&h->pr[b], 2 * h->plane * sizeof(float4)));
  }
  TAU_CUDA(cudaMalloc(&h->solid, h->plane));
  TAU_CUDA(cudaMalloc(&h->clk, sizeof(Clock)));
  TAU_CUDA(cudaEventCreate(&h->ev0));
  TAU_CUDA(cudaEventCreate(&h->ev1));
  TAU_CUDA(cudaFuncSetAttribute(hyp3d_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T3_SMEM));
  const Par P = make_par(h);
  hyp3d_build_solid<<<(unsigned)((h->plane + 255) / 256), 256, 0, h->stream>>>(P, h->solid);
  h->launches++;
  TAU_CUDA(cudaGetLastError());
  *out = h;
  return TAU_OK;
}

static int hyp3d_reset_clock(tau_hyp3d *h) {
  Clock c;
  memset(&c, 0, sizeof(c));
  c.t[0] = c.t[1] = h->prm.t0;
  c.d_tau[0] = c.d_tau[1] = h->prm.d_tau0;
  TAU_CUDA(cudaMemcpyAsync(h->clk, &c, sizeof(c), cudaMemcpyHostToDevice, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  h->steps = 0;
  return TAU_OK;
}

int tau_hyp3d_init(tau_hyp3d *h) {
  TAU_REQUIRE(h, "tau_hyp3d_init: null handle");
  TAU_CUDA(cudaSetDevice(h->device));
  const Par P = make_par(h);
  hyp3d_init<<<(unsigned)((h->plane + 255) / 256), 256, 0, h->stream>>>(P, h->st[h->cur], h->solid);
  h->pr_valid[h->cur] = false;
  h->launches++;
  TAU_CUDA(cudaGetLastError());
  h->have_state = true;
  return hyp3d_reset_clock(h);
}

int tau_hyp3d_upload(tau_hyp3d *h, const float *const planes[6], const float *clock2) {
  TAU_REQUIRE(h && planes, "tau_hyp3d_upload: null argument");
  TAU_CUDA(cudaSetDevice(h->device));
  const size_t nxy = (size_t)h->prm.nx * h->prm.ny, n = nxy * h->nz_local;
  for (int f = 0; f < 6; ++f) {
    TAU_REQUIRE(planes[f], "tau_hyp3d_upload: null plane %d", f);
    TAU_CUDA(cudaMemcpyAsync(h->st[h->cur] + f * h->plane + T3_H * nxy, planes[f], n * sizeof(float),
                             cudaMemcpyHostToDevice, h->stream));
  }
  h->have_state = true;
  h->pr_valid[h->cur] = false;
  int rc = hyp3d_reset_clock(h);
  if (rc) return rc;
  if (clock2) {
    Clock c;
    memset(&c, 0, sizeof(c));
    c.t[0] = c.t[1] = clock2[0];
    c.d_tau[0] = c.d_tau[1] = clock2[1];
    TAU_CUDA(cudaMemcpyAsync(h->clk, &c, sizeof(c), cudaMemcpyHostToDevice, h->stream));
    TAU_CUDA(cudaStreamSynchronize(h->stream));
  }
  return TAU_OK;
}

int tau_hyp3d_step(tau_hyp3d *h, int nsteps) {
  TAU_REQUIRE(h, "tau_hyp3d_step: null handle");
  TAU_REQUIRE(nsteps >= 0, "tau_hyp3d_step: nsteps must be >= 0");
  TAU_REQUIRE(h->have_state, "tau_hyp3d_step: no state (call tau_hyp3d_init or tau_hyp3d_upload)");
  TAU_REQUIRE(!h->slab, "tau_hyp3d_step: slab handles use tau_hyp3d_step_begin / tau_hyp3d_step_end "
                        "(ghost planes and the max wavespeed are exchanged in between)");
  TAU_CUDA(cudaSetDevice(h->device));
  TAU_CUDA(cudaEventRecord(h->ev0, h->stream));
  for (int s = 0; s < nsteps; ++s) {
    int rc = tau_hyp3d_step_begin(h);
    if (rc) return rc;
    rc = tau_hyp3d_step_end(h);
    if (rc) return rc;
  }
  TAU_CUDA(cudaEventRecord(h->ev1, h->stream));
  h->timed = true;
  return TAU_OK;
}

// k_step only (slab protocol: exchange ghost planes BEFORE, all-reduce MAX the wavespeed AFTER,
// then tau_hyp3d_step_end runs the controller)
int tau_hyp3d_step_begin(tau_hyp3d *h) {
  TAU_REQUIRE(h && h->have_state, "tau_hyp3d_step_begin: no state");
  const Par P = make_par(h);
  const int slot = (int)(h->steps & 1);
  dim3 block(T3_TX, T3_TY, T3_TZ);
  dim3 grid((P.nx + T3_TX - 1) / T3_TX, (P.ny + T3_TY - 1) / T3_TY, (h->nz_local + T3_TZ - 1) / T3_TZ);
  // the first step after init / upload decodes on the fly (no side buffer for that state yet) and writes the first one
  hyp3d_step<<<grid, block, T3_SMEM, h->stream>>>(P, h->st[h->cur], h->st[h->cur ^ 1],
                                                   h->pr_valid[h->cur] ? h->pr[h->cur] : nullptr, h->pr[h->cur ^ 1], h->solid,
                                                   h->clk, slot);
  h->pr_valid[h->cur ^ 1] = h->pr[h->cur ^ 1] != nullptr;
  h->launches++;
  TAU_CUDA(cudaGetLastError());
  return TAU_OK;
}

int tau_hyp3d_step_end(tau_hyp3d *h) {
  TAU_REQUIRE(h, "tau_hyp3d_step_end: null handle");
  const Par P = make_par(h);
  const int slot = (int)(h->steps & 1);
  hyp3d_controller<<<1, 32, 0, h->stream>>>(P, h->clk, slot);
  h->launches++;
  TAU_CUDA(cudaGetLastError());
  h->cur ^= 1;
  h->steps++;
  return TAU_OK;
}

int tau_hyp3d_clock(tau_hyp3d *h, float *t, float *d_tau, float *dt_last, float *maxs_last) {
  TAU_REQUIRE(h, "tau_hyp3d_clock: null handle");
  Clock c;
  TAU_CUDA(cudaMemcpyAsync(&c, h->clk, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  const int slot = (int)(h->steps & 1);
  if (t) *t = c.t[slot];
  if (d_tau) *d_tau = c.d_tau[slot];
  if (dt_last) *dt_last = c.dt_last;
  if (maxs_last) *maxs_last = c.maxs_last;
  return TAU_OK;
}

int tau_hyp3d_download(tau_hyp3d *h, float *const planes[6], uint8_t *solid) {
  TAU_REQUIRE(h && planes, "tau_hyp3d_download: null argument");
  TAU_CUDA(cudaSetDevice(h->device));
  const size_t nxy = (size_t)h->prm.nx * h->prm.ny, n = nxy * h->nz_local;
  for (int f = 0; f < 6; ++f)
    if (planes[f])
      TAU_CUDA(cudaMemcpyAsync(planes[f], h->st[h->cur] + f * h->plane + T3_H * nxy, n * sizeof(float),
                               cudaMemcpyDeviceToHost, h->stream));
  if (solid) TAU_CUDA(cudaMemcpyAsync(solid, h->solid + T3_H * nxy, n, cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

// k_vis :800-905: the scalar field the reference's volume renderer consumes (nz_local*ny*nx floats,
// index (z*ny+y)*nx+x; solid cells 0).  mode = VisMode :784-794.  Slab handles read their ghost planes
// for the z-neighbours: exchange them first.
int tau_hyp3d_vis(tau_hyp3d *h, int mode, float *out) {
  TAU_REQUIRE(h && out, "tau_hyp3d_vis: null argument");
  TAU_REQUIRE(mode >= 0 && mode <= 8, "tau_hyp3d_vis: mode %d not in [0, 8]", mode);
  TAU_REQUIRE(h->have_state, "tau_hyp3d_vis: no state (call tau_hyp3d_init or tau_hyp3d_upload)");
  TAU_CUDA(cudaSetDevice(h->device));
  const size_t n = (size_t)h->prm.nx * h->prm.ny * h->nz_local;
  if (!h->vis) TAU_CUDA(cudaMalloc(&h->vis, n * sizeof(float)));
  const Par P = make_par(h);
  dim3 grid((P.nx + T3_TX - 1) / T3_TX, (P.ny + T3_TY - 1) / T3_TY, (h->nz_local + T3_TZ - 1) / T3_TZ);
  hyp3d_vis<<<grid, dim3(T3_TX, T3_TY, T3_TZ), 0, h->stream>>>(P, h->st[h->cur], h->solid, h->vis, mode);
  h->launches++;
  TAU_CUDA(cudaGetLastError());
  TAU_CUDA(cudaMemcpyAsync(out, h->vis, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

// The palette index the reference computes on the host, as a function of norm :1215-1218
static int export_index_host(float norm) {
  const float g = powf(norm, 0.65f);
  int p = (int)(g * 255.0f);
  return p < 0 ? 0 : (p > 255 ? 255 : p);
}
// thr[k-1] = the smallest float norm in [0, 1] whose index is >= k (k = 1..255), by bisection over the
// float bit patterns (non-negative floats order like their bits) with the host's own powf.
void tau_4spl_index_thresholds(float thr[255]) {
  for (int k = 1; k <= 255; ++k) {
    uint32_t lo = 0u, hi = 0x3f800000u;  // index(0) = 0 < k <= index(1) = 255
    while (hi - lo > 1u) {
      const uint32_t mid = lo + (hi - lo) / 2u;
      float x;
      memcpy(&x, &mid, 4);
      if (export_index_host(x) >= k) hi = mid; else lo = mid;
    }
    memcpy(&thr[k - 1], &hi, 4);
  }
}

// one frame of th3cs.cu's export loop :1193-1222: schlieren field -> min/max -> 8-bit palette indices
int tau_hyp3d_export_frame(tau_hyp3d *h, uint8_t *indices, float minmax[2]) {
  TAU_REQUIRE(h && indices, "tau_hyp3d_export_frame: null argument");
  TAU_REQUIRE(h->have_state, "tau_hyp3d_export_frame: no state (call tau_hyp3d_init or tau_hyp3d_upload)");
  TAU_REQUIRE(!h->slab, "tau_hyp3d_export_frame: z-slab handles are not supported (the frame's min/max is global)");
  TAU_CUDA(cudaSetDevice(h->device));
  const size_t n = (size_t)h->prm.nx * h->prm.ny * h->nz_local;
  if (!h->vis) TAU_CUDA(cudaMalloc(&h->vis, n * sizeof(float)));
  if (!h->ex_idx) {
    TAU_CUDA(cudaMalloc(&h->ex_idx, n));
    TAU_CUDA(cudaMalloc(&h->ex_keys, 2 * sizeof(unsigned)));
    TAU_CUDA(cudaMalloc(&h->ex_thr, 255 * sizeof(float)));
    float thr[255];
    tau_4spl_index_thresholds(thr);
    TAU_CUDA(cudaMemcpyAsync(h->ex_thr, thr, sizeof(thr), cudaMemcpyHostToDevice, h->stream));
    TAU_CUDA(cudaStreamSynchronize(h->stream));
  }
  const Par P = make_par(h);
  dim3 grid((P.nx + T3_TX - 1) / T3_TX, (P.ny + T3_TY - 1) / T3_TY, (h->nz_local + T3_TZ - 1) / T3_TZ);
  hyp3d_vis<<<grid, dim3(T3_TX, T3_TY, T3_TZ), 0, h->stream>>>(P, h->st[h->cur], h->solid, h->vis, 8);
  const unsigned init[2] = {0xffffffffu, 0u};
  TAU_CUDA(cudaMemcpyAsync(h->ex_keys, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
  hyp3d_export_minmax<<<148 * 4, 256, 0, h->stream>>>(h->vis, n, h->ex_keys);
  hyp3d_export_quantize<<<148 * 8, 256, 0, h->stream>>>(h->vis, n, h->ex_keys, h->ex_thr, h->ex_idx);
  h->launches += 3;
  TAU_CUDA(cudaGetLastError());
  TAU_CUDA(cudaMemcpyAsync(indices, h->ex_idx, n, cudaMemcpyDeviceToHost, h->stream));
  unsigned k[2];
  TAU_CUDA(cudaMemcpyAsync(k, h->ex_keys, sizeof(k), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  if (minmax)
    for (int i = 0; i < 2; ++i) {
      const unsigned b = (k[i] & 0x80000000u) ? (k[i] & 0x7fffffffu) : ~k[i];
      memcpy(&minmax[i], &b, 4);
    }
  return TAU_OK;
}

int tau_hyp3d_sync(tau_hyp3d *h) {
  TAU_REQUIRE(h, "tau_hyp3d_sync: null handle");
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

int tau_hyp3d_device_state(tau_hyp3d *h, float **planes, float **maxs) {
  TAU_REQUIRE(h, "tau_hyp3d_device_state: null handle");
  if (planes) *planes = h->st[h->cur];
  if (maxs) *maxs = &h->clk->maxs;
  return TAU_OK;
}

long long tau_hyp3d_steps_done(tau_hyp3d *h) { return h ? h->steps : -1; }
long long tau_hyp3d_launch_count(tau_hyp3d *h) { return h ? h->launches : -1; }

int tau_hyp3d_last_step_ms(tau_hyp3d *h, float *ms) {
  TAU_REQUIRE(h && ms, "tau_hyp3d_last_step_ms: null argument");
  TAU_REQUIRE(h->timed, "tau_hyp3d_last_step_ms: no step has been timed yet");
  TAU_CUDA(cudaEventSynchronize(h->ev1));
  TAU_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return TAU_OK;
}

int tau_hyp3d_destroy(tau_hyp3d *h) {
  if (!h) return TAU_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->vis) cudaFree(h->vis);
  if (h->ex_idx) cudaFree(h->ex_idx);
  if (h->ex_keys) cudaFree(h->ex_keys);
  if (h->ex_thr) cudaFree(h->ex_thr);
  cudaFree(h->clk);
  cudaFree(h->solid);
  cudaFree(h->pr[1]);
  cudaFree(h->pr[0]);
  cudaFree(h->st[1]);
  cudaFree(h->st[0]);
  cudaEventDestroy(h->ev1);
  cudaEventDestroy(h->ev0);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return TAU_OK;
}

// ---- multi-GPU behind the C boundary: one process, one z-slab handle per device ---------------------------------------------
// (create(cfg, dims, ngpus); the reference has no multi-GPU path.)  z is periodic, so the slabs form a ring.  Per
// step: the three boundary planes of every field travel to the neighbours' ghost planes (cudaMemcpyPeerAsync on the
// destination's stream), every device runs its step kernel, the max wavespeed sums are folded on the host — one
// synchronisation per step, which the reference's own loop has twice (:1684-1700) — and every device's controller kernel
// commits the same clock.  Bit-identical to one handle over the whole grid (emulator test; tests/test_multi_gpu.py on >= 2 GPUs).
struct tau_hyp3d_group {
  int n;
  tau_hyp3d_params prm;
  tau_hyp3d *h[8];
  int dev[8], z0[8], nl[8];
  long long steps;
};

int tau_hyp3d_group_create(const tau_hyp3d_params *p, int ngpus, const int *devices, tau_hyp3d_group **out) {
  TAU_REQUIRE(p && out, "tau_hyp3d_group_create: null argument");
  TAU_REQUIRE(ngpus >= 1 && ngpus <= 8, "tau_hyp3d_group_create: ngpus must be in [1, 8] (got %d)", ngpus);
  TAU_REQUIRE(ngpus == 1 || p->nz >= ngpus * T3_H, "tau_hyp3d_group_create: %d planes cannot be split into %d slabs of >= %d",
              p->nz, ngpus, T3_H);
  const int have = tau_device_count();
  if (have <= 0) {
    tau_set_error("tau_hyp3d_group_create: no CUDA device (this library has no CPU fallback)");
    return TAU_ERR_NODEV;
  }
  TAU_REQUIRE(have >= ngpus, "tau_hyp3d_group_create: %d GPUs requested, %d visible", ngpus, have);
  tau_hyp3d_group *g = new (std::nothrow) tau_hyp3d_group();
  if (!g) return TAU_ERR_NOMEM;
  memset(g, 0, sizeof(*g));
  g->n = ngpus;
  g->prm = *p;
  const int base = p->nz / ngpus, rem = p->nz % ngpus;
  int z = 0, rc = TAU_OK;
  for (int i = 0; i < ngpus && !rc; ++i) {
    g->dev[i] = devices ? devices[i] : i;
    g->z0[i] = z;
    g->nl[i] = base + (i < rem ? 1 : 0);
    z += g->nl[i];
    rc = tau_hyp3d_create(p, g->dev[i], g->z0[i], g->nl[i], nullptr, &g->h[i]);
  }
  for (int i = 0; i < ngpus && !rc && ngpus > 1; ++i) {  // direct peer copies where the devices allow it
    if (cudaSetDevice(g->dev[i]) != cudaSuccess) rc = TAU_ERR_CUDA;
    for (int j = 0; j < ngpus && !rc; ++j) {
      if (j == i) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, g->dev[i], g->dev[j]);
      if (can) cudaDeviceEnablePeerAccess(g->dev[j], 0);
      cudaGetLastError();  // already enabled: fine; not possible: cudaMemcpyPeerAsync stages through the host
    }
  }
  if (rc) {
    for (int i = 0; i < ngpus; ++i) tau_hyp3d_destroy(g->h[i]);
    delete g;
    return rc;
  }
  *out = g;
  return TAU_OK;
}

int tau_hyp3d_group_size(tau_hyp3d_group *g) { return g ? g->n : -1; }
int tau_hyp3d_group_member(tau_hyp3d_group *g, int i, tau_hyp3d **h, int *z_begin, int *nz_local) {
  TAU_REQUIRE(g && i >= 0 && i < g->n, "tau_hyp3d_group_member: bad argument");
  if (h) *h = g->h[i];
  if (z_begin) *z_begin = g->z0[i];
  if (nz_local) *nz_local = g->nl[i];
  return TAU_OK;
}

int tau_hyp3d_group_init(tau_hyp3d_group *g) {
  TAU_REQUIRE(g, "tau_hyp3d_group_init: null handle");
  for (int i = 0; i < g->n; ++i) {
    const int rc = tau_hyp3d_init(g->h[i]);
    if (rc) return rc;
  }
  g->steps = 0;
  return TAU_OK;
}

// planes cover the WHOLE grid (nz x ny x nx, reference layout); clock2 = (t, d_tau) or null for the defaults
int tau_hyp3d_group_upload(tau_hyp3d_group *g, const float *const planes[6], const float *clock2) {
  TAU_REQUIRE(g && planes, "tau_hyp3d_group_upload: null argument");
  const size_t nxy = (size_t)g->prm.nx * g->prm.ny;
  for (int i = 0; i < g->n; ++i) {
    const float *pl[6];
    for (int f = 0; f < 6; ++f) {
      TAU_REQUIRE(planes[f], "tau_hyp3d_group_upload: null plane %d", f);
      pl[f] = planes[f] + (size_t)g->z0[i] * nxy;
    }
    const int rc = tau_hyp3d_upload(g->h[i], pl, clock2);
    if (rc) return rc;
  }
  g->steps = 0;
  return TAU_OK;
}

int tau_hyp3d_group_step(tau_hyp3d_group *g, int nsteps) {
  TAU_REQUIRE(g && nsteps >= 0, "tau_hyp3d_group_step: bad argument");
  if (g->n == 1) {
    const int rc = tau_hyp3d_step(g->h[0], nsteps);
    if (!rc) g->steps += nsteps;
    return rc;
  }
  const size_t nxy = (size_t)g->prm.nx * g->prm.ny, ghost = (size_t)T3_H * nxy * sizeof(float);
  for (int s = 0; s < nsteps; ++s) {
    // ghost planes of the CURRENT state: every source's step kernel has completed (synchronised below / by upload)
    for (int r = 0; r < g->n; ++r) {
      tau_hyp3d *a = g->h[r], *b = g->h[(r + 1) % g->n];  // b sits above a (ring)
      for (int f = 0; f < 6; ++f) {
        float *pa = a->st[a->cur] + (size_t)f * a->plane, *pb = b->st[b->cur] + (size_t)f * b->plane;
        // a's last T3_H own planes -> b's lower ghost planes; b's first T3_H own planes -> a's upper ghost planes
        TAU_CUDA(cudaMemcpyPeerAsync(pb, b->device, pa + (size_t)a->nz_local * nxy, a->device, ghost, b->stream));
        TAU_CUDA(cudaMemcpyPeerAsync(pa + (size_t)(a->nz_local + T3_H) * nxy, a->device, pb + (size_t)T3_H * nxy, b->device, ghost,
                                     a->stream));
      }
    }
    for (int r = 0; r < g->n; ++r) {
      TAU_CUDA(cudaSetDevice(g->dev[r]));
      const int rc = tau_hyp3d_step_begin(g->h[r]);
      if (rc) return rc;
    }
    float m = 0.f, mr[8];
    for (int r = 0; r < g->n; ++r)
      TAU_CUDA(cudaMemcpyAsync(&mr[r], &g->h[r]->clk->maxs, sizeof(float), cudaMemcpyDeviceToHost, g->h[r]->stream));
    for (int r = 0; r < g->n; ++r) {
      TAU_CUDA(cudaStreamSynchronize(g->h[r]->stream));
      m = mr[r] > m ? mr[r] : m;  // exact: max is associative (what the all-reduce of the torchrun path computes)
    }
    for (int r = 0; r < g->n; ++r) {
      TAU_CUDA(cudaSetDevice(g->dev[r]));
      TAU_CUDA(cudaMemcpyAsync(&g->h[r]->clk->maxs, &m, sizeof(float), cudaMemcpyHostToDevice, g->h[r]->stream));
      const int rc = tau_hyp3d_step_end(g->h[r]);
      if (rc) return rc;
    }
    for (int r = 0; r < g->n; ++r) TAU_CUDA(cudaStreamSynchronize(g->h[r]->stream));  // `m` is read by the copies above
    g->steps++;
  }
  return TAU_OK;
}

int tau_hyp3d_group_clock(tau_hyp3d_group *g, float *t, float *d_tau, float *dt_last, float *maxs_last) {
  TAU_REQUIRE(g, "tau_hyp3d_group_clock: null handle");
  TAU_CUDA(cudaSetDevice(g->dev[0]));
  return tau_hyp3d_clock(g->h[0], t, d_tau, dt_last, maxs_last);  // every slab carries the same clock
}

int tau_hyp3d_group_download(tau_hyp3d_group *g, float *const planes[6], uint8_t *solid) {
  TAU_REQUIRE(g && planes, "tau_hyp3d_group_download: null argument");
  const size_t nxy = (size_t)g->prm.nx * g->prm.ny;
  for (int i = 0; i < g->n; ++i) {
    float *pl[6];
    for (int f = 0; f < 6; ++f) pl[f] = planes[f] ? planes[f] + (size_t)g->z0[i] * nxy : nullptr;
    const int rc = tau_hyp3d_download(g->h[i], pl, solid ? solid + (size_t)g->z0[i] * nxy : nullptr);
    if (rc) return rc;
  }
  return TAU_OK;
}

long long tau_hyp3d_group_steps_done(tau_hyp3d_group *g) { return g ? g->steps : -1; }

int tau_hyp3d_group_destroy(tau_hyp3d_group *g) {
  if (!g) return TAU_OK;
  for (int i = 0; i < g->n; ++i) tau_hyp3d_destroy(g->h[i]);
  delete g;
  return TAU_OK;
}

}  // extern "C"
