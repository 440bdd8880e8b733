/* hyp3d_oracle.c — CPU restatement (fp32) of the reference 3-D hypersonic step.
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may call this; the product never links or imports it.
 *
 * Follows tau_hypersonic_3d_cuda.cu (th3cs.cu carries the same k_step): one cell at a time, six
 * face fluxes per cell exactly as k_step (:987-1359) does, state stored as
 * xi=ln rho, phi=asinh(u/u_ref) x3, lambda=ln p, zeta=ln e_vib in six planes indexed
 * (z*ny+y)*nx+x (:152).  The device code uses the fast intrinsics __expf/__logf by name
 * (:113-117,161-171); here they are libm expf/logf, so agreement with the GPU reference is at
 * float round-off (~1e-6 relative per step), not bit level.
 *
 * Pinning: the reference has no tests or golden vectors for this solver;
 * tests/golden/hyp3d_ref_*.npz hold outputs of the reference's own k_step run on a B200 through
 * oracle/_ref/libref_hyp3d.so (tests/golden/make_golden_gpu.py) and tests/test_oracle_cpu.py
 * compares this file against them.
 * Also pinned BIT FOR BIT on the reference's kernels executed on the CPU (oracle/_ref/libref_hyp3d_host.so and the
 * th3cs exporter, tests/test_oracle_cpu.py::test_hyp3d_oracle_equals_reference_kernel_run_on_the_cpu, ::test_th3cs_*).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { /* struct Params :21-42 */
  int nx, ny, nz;
  float dx, dy, dz;
  float cfl, u_ref, R, gamma_floor, Twall, tau_vib, theta_v;
  float sdf_cx, sdf_cy, sdf_cz, sdf_r;
  float inflow_r, inflow_p, inflow_u, inflow_v, inflow_w;
  int sponge_n;
  float sponge_strength;
  int sponge_out_n;
  float sponge_out_strength;
} oracle_hyp3d_params;

typedef struct { float r, mx, my, mz, Et, Ev; } Cons;
typedef struct { float r, u, v, w, p, T, ev, Tv; } Prim;

#define RHO_P_FLOOR 1e-30f          /* :52-58 */
#define THERMAL_ENERGY_FLOOR 1e-12f
#define DENOM_EPS 1e-12f
#define NEWTON_TEMP_FLOOR 1e-6f
#define WENO_EPS 1e-6f
#define TAU_VIB_MIN 1e-9f

static const oracle_hyp3d_params *P;

/* main()'s constants :1531-1557 for an n^3 grid (dx = 1/n) */
void oracle_hyp3d_default_params(oracle_hyp3d_params *p, int nx, int ny, int nz) {
  p->nx = nx; p->ny = ny; p->nz = nz;
  p->dx = 1.f / nx; p->dy = 1.f / ny; p->dz = 1.f / nz;
  p->cfl = 0.3333f; p->u_ref = 10.f; p->R = 10.f; p->gamma_floor = 1.1f; p->Twall = 0.02f;
  p->tau_vib = 2e-4f; p->theta_v = 0.2f;
  p->sdf_cx = 0.5f; p->sdf_cy = 0.5f; p->sdf_cz = 0.5f; p->sdf_r = 0.25f;
  p->inflow_r = 0.02f; p->inflow_p = 0.02f; p->inflow_u = 100.0f; p->inflow_v = 0.0f; p->inflow_w = 0.0f;
  p->sponge_n = 24; p->sponge_strength = 0.05f; p->sponge_out_n = 24; p->sponge_out_strength = 0.05f;
}

static inline float clampf(float x, float a, float b) { return fminf(fmaxf(x, a), b); }
static inline float signed_denom_guard(float x) { return copysignf(fmaxf(fabsf(x), DENOM_EPS), x); } /* :147 */
static inline int idx3(int x, int y, int z) { return (z * P->ny + y) * P->nx + x; }                /* :152 */
static inline int wrapi(int i, int n) { i %= n; return (i < 0) ? i + n : i; }                       /* :156 */
static inline float asinhf_dev(float x) {                                                           /* :121-125 */
  float ax = fabsf(x);
  return copysignf(logf(ax + sqrtf(ax * ax + 1.0f)), x);
}
static inline float vel_from_phi(float phi) { return P->u_ref * sinhf(phi); }
static inline float phi_from_vel(float u) { return asinhf_dev(u / P->u_ref); }
static inline float enc_log(float v) { return logf(fmaxf(v, RHO_P_FLOOR)); }                        /* :161-171 */

static inline float sdf_sphere(float x, float y, float z) {                                        /* :173-178 */
  float dx = x - P->sdf_cx, dy = y - P->sdf_cy, dz = z - P->sdf_cz;
  return sqrtf(dx * dx + dy * dy + dz * dz) - P->sdf_r;
}
static int cell_is_solid(const uint8_t *solid, int x, int y, int z) {                               /* :180-189 */
  if (x >= 0 && x < P->nx && y >= 0 && y < P->ny && z >= 0 && z < P->nz) return solid[idx3(x, y, z)] != 0;
  return sdf_sphere((x + 0.5f) * P->dx, (y + 0.5f) * P->dy, (z + 0.5f) * P->dz) < 0.f;
}
static float evib_eq(float T) {                                                                     /* :206-211 */
  float a = P->theta_v / fmaxf(T, NEWTON_TEMP_FLOOR);
  float ea = expf(a);
  float denom = fmaxf(ea - 1.f, NEWTON_TEMP_FLOOR);
  return (P->R * P->theta_v) / denom;
}
static Prim log_to_prim_fast(float xi, float phx, float phy, float phz, float lam, float zet) {     /* :213-225 */
  Prim q;
  q.r = expf(xi); q.u = vel_from_phi(phx); q.v = vel_from_phi(phy); q.w = vel_from_phi(phz);
  q.p = expf(lam); q.ev = expf(zet);
  q.T = q.p / (q.r * P->R);
  q.Tv = 0.f;
  return q;
}
static Cons prim_to_cons(const Prim *q) {                                                           /* :234-245 */
  Cons U;
  U.r = q->r; U.mx = q->r * q->u; U.my = q->r * q->v; U.mz = q->r * q->w;
  float ke = 0.5f * (q->u * q->u + q->v * q->v + q->w * q->w);
  float e_th = q->p / fmaxf((P->gamma_floor - 1.f) * q->r, RHO_P_FLOOR);
  U.Ev = q->r * q->ev;
  U.Et = q->r * (ke + e_th + q->ev);
  return U;
}
static Prim cons_to_prim(const Cons *U) {                                                           /* :247-262, Tv unused */
  Prim q;
  q.r = fmaxf(U->r, RHO_P_FLOOR);
  q.u = U->mx / q.r; q.v = U->my / q.r; q.w = U->mz / q.r;
  float ke = 0.5f * (q.u * q.u + q.v * q.v + q.w * q.w);
  float ev = fmaxf(U->Ev / q.r, 0.f);
  float e_tot = U->Et / q.r;
  float e_th = fmaxf(e_tot - ke - ev, THERMAL_ENERGY_FLOOR);
  q.p = fmaxf((P->gamma_floor - 1.f) * q.r * e_th, RHO_P_FLOOR);
  q.ev = ev;
  q.T = q.p / (q.r * P->R);
  q.Tv = 0.f;
  return q;
}
static float soundspeed(const Prim *q) { return sqrtf(fmaxf(P->gamma_floor * q->p / q->r, DENOM_EPS)); } /* :264 */

static Cons axis_flux(const Prim *q, int axis) {                                                    /* flux_x/y/z :268-308 */
  Cons F;
  float un = axis == 0 ? q->u : (axis == 1 ? q->v : q->w);
  float H = (q->p / q->r) + (0.5f * (q->u * q->u + q->v * q->v + q->w * q->w) + q->ev) +
            q->p / fmaxf((P->gamma_floor - 1.f) * q->r, RHO_P_FLOOR);
  F.r = q->r * un;
  F.mx = q->r * q->u * un + (axis == 0 ? q->p : 0.f);
  F.my = q->r * q->v * un + (axis == 1 ? q->p : 0.f);
  F.mz = q->r * q->w * un + (axis == 2 ? q->p : 0.f);
  F.Et = q->r * H * un;
  F.Ev = q->r * q->ev * un;
  return F;
}
static float axis_crossflow_speed(const Prim *L, const Prim *R, int axis) {                         /* :318-325 */
  if (axis == 0) return (fabsf(L->v) + fabsf(R->v) + fabsf(L->w) + fabsf(R->w)) * 0.5f;
  if (axis == 1) return (fabsf(L->u) + fabsf(R->u) + fabsf(L->w) + fabsf(R->w)) * 0.5f;
  return (fabsf(L->u) + fabsf(R->u) + fabsf(L->v) + fabsf(R->v)) * 0.5f;
}
static Cons addC(Cons a, Cons b) { Cons r = {a.r + b.r, a.mx + b.mx, a.my + b.my, a.mz + b.mz, a.Et + b.Et, a.Ev + b.Ev}; return r; }
static Cons subC(Cons a, Cons b) { Cons r = {a.r - b.r, a.mx - b.mx, a.my - b.my, a.mz - b.mz, a.Et - b.Et, a.Ev - b.Ev}; return r; }
static Cons mulC(Cons a, float s) { Cons r = {a.r * s, a.mx * s, a.my * s, a.mz * s, a.Et * s, a.Ev * s}; return r; }

static float entropy_fix_speed(float s, float a_ref) {                                              /* :366-374 */
  float d = 0.1f * a_ref, as = fabsf(s);
  if (as >= d) return s;
  float sgn = (s >= 0.f) ? 1.f : -1.f;
  return sgn * (0.5f * (as * as / fmaxf(d, DENOM_EPS) + d));
}
static float shock_sensor(const Prim *L, const Prim *R) {                                           /* :376-381 */
  float dp = fabsf(R->p - L->p) / fmaxf(R->p + L->p, DENOM_EPS);
  float dr = fabsf(R->r - L->r) / fmaxf(R->r + L->r, DENOM_EPS);
  return clampf(5.f * (0.5f * (dp + dr)), 0.f, 1.f);
}
static void fill_star_momentum(Cons *U, const Prim *q, float rStar, float sM, int axis) {           /* :335-350 */
  U->mx = rStar * (axis == 0 ? sM : q->u);
  U->my = rStar * (axis == 1 ? sM : q->v);
  U->mz = rStar * (axis == 2 ? sM : q->w);
}

/* hllc_flux_axis :383-460 */
static Cons hllc_flux_axis(const Prim *L, const Prim *R, int axis) {
  float aL = soundspeed(L), aR = soundspeed(R);
  float unL = axis == 0 ? L->u : (axis == 1 ? L->v : L->w);
  float unR = axis == 0 ? R->u : (axis == 1 ? R->v : R->w);
  float sL = fminf(unL - aL, unR - aR), sR = fmaxf(unL + aL, unR + aR);
  float aRef = fmaxf(aL, aR);
  sL = entropy_fix_speed(sL, aRef);
  sR = entropy_fix_speed(sR, aRef);
  Cons UL = prim_to_cons(L), UR = prim_to_cons(R);
  Cons FL = axis_flux(L, axis), FR = axis_flux(R, axis);
  if (sL >= 0.f) return FL;
  if (sR <= 0.f) return FR;
  float rL = L->r, rR = R->r, pL = L->p, pR = R->p;
  float denom = signed_denom_guard(rL * (sL - unL) - rR * (sR - unR));
  float sM = (pR - pL + rL * unL * (sL - unL) - rR * unR * (sR - unR)) / denom;
  float pStarL = pL + rL * (sL - unL) * (sM - unL);
  float pStarR = pR + rR * (sR - unR) * (sM - unR);
  float pStar = 0.5f * (pStarL + pStarR);
  float vCarb = axis_crossflow_speed(L, R, axis);
  float align = clampf(1.f - vCarb / fmaxf(aRef, DENOM_EPS), 0.f, 1.f);
  float alpha = shock_sensor(L, R) * align;
  Cons FHLL;
  {
    Cons num = subC(mulC(FL, sR), mulC(FR, sL));
    Cons corr = mulC(subC(UR, UL), sL * sR);
    FHLL = mulC(addC(num, corr), 1.f / signed_denom_guard(sR - sL));
  }
  const Prim *K = (sM >= 0.f) ? L : R;
  Cons UK = (sM >= 0.f) ? UL : UR, FK = (sM >= 0.f) ? FL : FR;
  float sK = (sM >= 0.f) ? sL : sR, unK = (sM >= 0.f) ? unL : unR, rK = K->r, pK = K->p;
  float starDenom = signed_denom_guard(sK - sM);
  float rStar = rK * (sK - unK) / starDenom;
  float EStar = ((sK - unK) * UK.Et - pK * unK + pStar * sM) / starDenom;
  float EvStar = UK.Ev * (sK - unK) / starDenom;
  Cons UStar;
  UStar.r = rStar;
  fill_star_momentum(&UStar, K, rStar, sM, axis);
  UStar.Et = EStar;
  UStar.Ev = EvStar;
  Cons FHLLC = addC(FK, mulC(subC(UStar, UK), sK));
  return addC(mulC(FHLLC, 1.f - alpha), mulC(FHLL, alpha));
}

static void apply_wall(Prim *q) {                                                                   /* :511-521 */
  float p_keep = fmaxf(q->p, RHO_P_FLOOR);
  q->u = q->v = q->w = 0.f;
  q->T = P->Twall;
  q->p = p_keep;
  q->r = fmaxf(q->p / (P->R * fmaxf(q->T, NEWTON_TEMP_FLOOR)), RHO_P_FLOOR);
  q->ev = evib_eq(P->Twall);
  q->Tv = P->Twall;
}
static float weno5_left(float v0, float v1, float v2, float v3, float v4) {                         /* :534-558 */
  float p0 = (2.f * v0 - 7.f * v1 + 11.f * v2) * (1.f / 6.f);
  float p1 = (-1.f * v1 + 5.f * v2 + 2.f * v3) * (1.f / 6.f);
  float p2 = (2.f * v2 + 5.f * v3 - 1.f * v4) * (1.f / 6.f);
  float b0 = (13.f / 12.f) * (v0 - 2.f * v1 + v2) * (v0 - 2.f * v1 + v2) +
             0.25f * (v0 - 4.f * v1 + 3.f * v2) * (v0 - 4.f * v1 + 3.f * v2);
  float b1 = (13.f / 12.f) * (v1 - 2.f * v2 + v3) * (v1 - 2.f * v2 + v3) + 0.25f * (v1 - v3) * (v1 - v3);
  float b2 = (13.f / 12.f) * (v2 - 2.f * v3 + v4) * (v2 - 2.f * v3 + v4) +
             0.25f * (3.f * v2 - 4.f * v3 + v4) * (3.f * v2 - 4.f * v3 + v4);
  float eps = WENO_EPS;
  float a0 = 0.1f / ((eps + b0) * (eps + b0));
  float a1 = 0.6f / ((eps + b1) * (eps + b1));
  float a2 = 0.3f / ((eps + b2) * (eps + b2));
  float s = a0 + a1 + a2;
  return (a0 / s) * p0 + (a1 / s) * p1 + (a2 / s) * p2;
}
static void prim_floor_fast(Prim *q) {                                                              /* :565-571 */
  q->r = fmaxf(q->r, RHO_P_FLOOR);
  q->p = fmaxf(q->p, RHO_P_FLOOR);
  q->ev = fmaxf(q->ev, 0.f);
  q->T = q->p / (q->r * P->R);
  q->Tv = 0.f;
}
static void weno_face_from_6(const Prim *q, Prim *L, Prim *R) {                                     /* :578-598 */
#define WL(f) weno5_left(q[0].f, q[1].f, q[2].f, q[3].f, q[4].f)
#define WR(f) weno5_left(q[5].f, q[4].f, q[3].f, q[2].f, q[1].f)
  L->r = WL(r); L->u = WL(u); L->v = WL(v); L->w = WL(w); L->p = WL(p); L->ev = WL(ev);
  R->r = WR(r); R->u = WR(u); R->v = WR(v); R->w = WR(w); R->p = WR(p); R->ev = WR(ev);
#undef WL
#undef WR
  prim_floor_fast(L);
  prim_floor_fast(R);
}
static Prim inflow_prim(void) {                                                                     /* :611-622 */
  Prim q;
  q.r = fmaxf(P->inflow_r, RHO_P_FLOOR);
  q.u = P->inflow_u; q.v = P->inflow_v; q.w = P->inflow_w;
  q.p = fmaxf(P->inflow_p, RHO_P_FLOOR);
  q.T = q.p / (q.r * P->R);
  q.ev = evib_eq(q.T);
  q.Tv = 0.f;
  return q;
}
typedef struct { const float *xi, *px, *py, *pz, *lam, *zet; } State;
static Prim outflow_prim_transmissive(const State *s, int y, int z) {                               /* :691-722 */
  int iR = idx3(P->nx - 1, y, z);
  Prim qR = log_to_prim_fast(s->xi[iR], s->px[iR], s->py[iR], s->pz[iR], s->lam[iR], s->zet[iR]);
  Prim q = qR;
  float aR = soundspeed(&qR), un = qR.u;
  if (un < 0.0f) return inflow_prim();
  if (un < aR) {
    float p_amb = fmaxf(P->inflow_p, RHO_P_FLOOR);
    q.p = fmaxf(q.p + 0.05f * (p_amb - q.p), RHO_P_FLOOR);
  }
  q.r = fmaxf(q.r, RHO_P_FLOOR);
  q.p = fmaxf(q.p, RHO_P_FLOOR);
  q.ev = fmaxf(q.ev, 0.f);
  q.T = q.p / (q.r * P->R);
  q.Tv = 0.f;
  return q;
}
/* the halo-tile fill of k_step :1019-1056 for one (possibly out-of-range) cell */
static Prim tile_prim(const State *s, const uint8_t *solid, int gx, int gy, int gz, int *is_solid) {
  int gyw = wrapi(gy, P->ny), gzw = wrapi(gz, P->nz);
  *is_solid = cell_is_solid(solid, gx, gyw, gzw);
  Prim q;
  if (gx < 0) q = inflow_prim();
  else if (gx >= P->nx) q = outflow_prim_transmissive(s, gyw, gzw);
  else {
    int gi = idx3(gx, gyw, gzw);
    q = log_to_prim_fast(s->xi[gi], s->px[gi], s->py[gi], s->pz[gi], s->lam[gi], s->zet[gi]);
  }
  if (*is_solid) apply_wall(&q);
  return q;
}

/* one face of one cell, minus (side=-1) or plus (side=+1) along axis — k_step :1113-1264 */
static Cons cell_face_flux(const Prim *line /* cells -3..+3 along the axis */, const int *sol, int side,
                           int axis) {
  const Prim *q0 = &line[3];
  if (side < 0) {
    int face_solid = sol[2] || sol[3];
    int stencil_solid = sol[0] || sol[1] || sol[2] || sol[3] || sol[4] || sol[5];
    if (face_solid) {
      Prim R = *q0, L = R;
      if (axis == 0) L.u = -L.u; else if (axis == 1) L.v = -L.v; else L.w = -L.w;
      return hllc_flux_axis(&L, &R, axis);
    } else if (stencil_solid) {
      Prim L = line[2], R = line[3];
      prim_floor_fast(&L); prim_floor_fast(&R);
      return hllc_flux_axis(&L, &R, axis);
    }
    Prim L, R;
    weno_face_from_6(&line[0], &L, &R);
    return hllc_flux_axis(&L, &R, axis);
  }
  int face_solid = sol[3] || sol[4];
  int stencil_solid = sol[1] || sol[2] || sol[3] || sol[4] || sol[5] || sol[6];
  if (face_solid) {
    Prim L = *q0, R = L;
    if (axis == 0) R.u = -R.u; else if (axis == 1) R.v = -R.v; else R.w = -R.w;
    return hllc_flux_axis(&L, &R, axis);
  } else if (stencil_solid) {
    Prim L = line[3], R = line[4];
    prim_floor_fast(&L); prim_floor_fast(&R);
    return hllc_flux_axis(&L, &R, axis);
  }
  Prim L, R;
  weno_face_from_6(&line[1], &L, &R);
  return hllc_flux_axis(&L, &R, axis);
}

/* k_build_solid_mask :759-770 */
void oracle_hyp3d_build_solid(const oracle_hyp3d_params *p, uint8_t *solid) {
  P = p;
  for (int z = 0; z < p->nz; ++z)
    for (int y = 0; y < p->ny; ++y)
      for (int x = 0; x < p->nx; ++x)
        solid[idx3(x, y, z)] = sdf_sphere((x + 0.5f) * p->dx, (y + 0.5f) * p->dy, (z + 0.5f) * p->dz) < 0.f;
}

/* k_init :939-985 */
void oracle_hyp3d_init(const oracle_hyp3d_params *p, float *xi, float *px, float *py, float *pz, float *lam,
                       float *zet, const uint8_t *solid) {
  P = p;
  const int N = p->nx * p->ny * p->nz;
  for (int i = 0; i < N; ++i) {
    Prim q;
    q.r = fmaxf(p->inflow_r, RHO_P_FLOOR);
    q.p = fmaxf(p->inflow_p, RHO_P_FLOOR);
    q.u = q.v = q.w = 0.f;
    q.T = q.p / (q.r * p->R);
    q.ev = evib_eq(q.T);
    if (solid[i]) {
      float pk = q.p;
      q.T = p->Twall;
      q.p = pk;
      q.r = fmaxf(q.p / (p->R * fmaxf(q.T, NEWTON_TEMP_FLOOR)), RHO_P_FLOOR);
      q.ev = evib_eq(q.T);
    }
    xi[i] = enc_log(q.r);
    px[i] = phi_from_vel(q.u); py[i] = phi_from_vel(q.v); pz[i] = phi_from_vel(q.w);
    lam[i] = enc_log(q.p);
    zet[i] = enc_log(q.ev);
  }
}

/* k_step :987-1359 — one step in -> out; returns the max wavespeed sum the kernel atomically
 * accumulates (:1345-1351).  z0..z1 restricts the computed planes (threaded callers). */
float oracle_hyp3d_step_planes(const oracle_hyp3d_params *p, const float *const in[6], float *const out[6],
                               const uint8_t *solid, float dt, float inflow_gain, int z0, int z1) {
  P = p;
  State s = {in[0], in[1], in[2], in[3], in[4], in[5]};
  float maxs = 0.f;
  for (int z = z0; z < z1; ++z)
    for (int y = 0; y < p->ny; ++y)
      for (int x = 0; x < p->nx; ++x) {
        const int i = idx3(x, y, z);
        if (solid[i]) {
          for (int f = 0; f < 6; ++f) out[f][i] = in[f][i];
          continue;
        }
        Prim lx[7], ly[7], lz[7];
        int sx[7], sy[7], sz[7];
        for (int k = -3; k <= 3; ++k) {
          lx[k + 3] = tile_prim(&s, solid, x + k, y, z, &sx[k + 3]);
          ly[k + 3] = tile_prim(&s, solid, x, y + k, z, &sy[k + 3]);
          lz[k + 3] = tile_prim(&s, solid, x, y, z + k, &sz[k + 3]);
        }
        Prim q0 = lx[3];
        Cons Fx_m = cell_face_flux(lx, sx, -1, 0), Fx_p = cell_face_flux(lx, sx, +1, 0);
        Cons Fy_m = cell_face_flux(ly, sy, -1, 1), Fy_p = cell_face_flux(ly, sy, +1, 1);
        Cons Fz_m = cell_face_flux(lz, sz, -1, 2), Fz_p = cell_face_flux(lz, sz, +1, 2);
        Cons U0 = prim_to_cons(&q0), dU;
#define DIV(c) (-((Fx_p.c - Fx_m.c) / p->dx + (Fy_p.c - Fy_m.c) / p->dy + (Fz_p.c - Fz_m.c) / p->dz))
        dU.r = DIV(r); dU.mx = DIV(mx); dU.my = DIV(my); dU.mz = DIV(mz); dU.Et = DIV(Et); dU.Ev = DIV(Ev);
#undef DIV
        Cons U1 = addC(U0, mulC(dU, dt));
        Prim q1 = cons_to_prim(&U1);
        if (!isfinite(q1.r) || !isfinite(q1.p) || !isfinite(q1.u) || !isfinite(q1.v) || !isfinite(q1.w) ||
            !isfinite(q1.ev) || q1.r <= 0.f || q1.p <= 0.f || q1.ev < 0.f)
          q1 = inflow_prim();
        float ev_eq = evib_eq(q1.T);
        q1.ev = fmaxf(q1.ev + (ev_eq - q1.ev) * (dt / fmaxf(p->tau_vib, TAU_VIB_MIN)), 0.f);
        int nsp = p->sponge_n > 0 ? p->sponge_n : 0;
        if (nsp > 0 && x < nsp) {                                               /* :1295-1318 */
          float sgm = 1.0f - (float)x / (float)nsp;
          sgm = fminf(fmaxf(sgm, 0.0f), 1.0f);
          float k = p->sponge_strength * (sgm * sgm);
          float tr = fmaxf(p->inflow_r, RHO_P_FLOOR), tp = fmaxf(p->inflow_p, RHO_P_FLOOR);
          float tT = tp / (tr * p->R), tev = evib_eq(tT);
          q1.r = fmaxf(q1.r + k * (tr - q1.r), RHO_P_FLOOR);
          q1.p = fmaxf(q1.p + k * (tp - q1.p), RHO_P_FLOOR);
          q1.u = q1.u + k * (inflow_gain * p->inflow_u - q1.u);
          q1.v = q1.v + k * (inflow_gain * p->inflow_v - q1.v);
          q1.w = q1.w + k * (inflow_gain * p->inflow_w - q1.w);
          q1.T = q1.p / (q1.r * p->R);
          q1.ev = fmaxf(q1.ev + k * (tev - q1.ev), 0.f);
        }
        int nspo = p->sponge_out_n > 0 ? p->sponge_out_n : 0;
        if (nspo > 0 && x >= (p->nx - nspo)) {                                  /* :1319-1343 */
          int xo = x - (p->nx - nspo);
          float sgm = (float)xo / (float)nspo;
          sgm = fminf(fmaxf(sgm, 0.0f), 1.0f);
          float k = p->sponge_out_strength * (sgm * sgm);
          float tr = fmaxf(p->inflow_r, RHO_P_FLOOR), tp = fmaxf(p->inflow_p, RHO_P_FLOOR);
          float tT = tp / (tr * p->R), tev = evib_eq(tT);
          q1.r = fmaxf(q1.r + k * (tr - q1.r), RHO_P_FLOOR);
          q1.p = fmaxf(q1.p + k * (tp - q1.p), RHO_P_FLOOR);
          q1.u = q1.u + k * (0.0f - q1.u);
          q1.v = q1.v + k * (0.0f - q1.v);
          q1.w = q1.w + k * (0.0f - q1.w);
          q1.T = q1.p / (q1.r * p->R);
          q1.ev = fmaxf(q1.ev + k * (tev - q1.ev), 0.f);
        }
        float a = soundspeed(&q1);
        float ssum = (fabsf(q1.u) + a) / p->dx + (fabsf(q1.v) + a) / p->dy + (fabsf(q1.w) + a) / p->dz;
        if (isfinite(ssum) && ssum > 0.f && ssum > maxs) maxs = ssum;
        out[0][i] = enc_log(q1.r);
        out[1][i] = phi_from_vel(q1.u);
        out[2][i] = phi_from_vel(q1.v);
        out[3][i] = phi_from_vel(q1.w);
        out[4][i] = enc_log(q1.p);
        out[5][i] = enc_log(q1.ev);
      }
  return maxs;
}

/* Host clock/controller of the reference loop :1680-1704 around nsteps calls of k_step.
 * clock = {t, d_tau} in/out; maxs_hist/dt_hist (optional, nsteps floats each). */
void oracle_hyp3d_run(const oracle_hyp3d_params *p, float *const planes[6], const uint8_t *solid, int nsteps,
                      float clock[2], float *dt_hist, float *maxs_hist) {
  const size_t N = (size_t)p->nx * p->ny * p->nz;
  float *buf = (float *)malloc(6 * N * sizeof(float));
  float *a[6], *b[6];
  for (int f = 0; f < 6; ++f) { a[f] = planes[f]; b[f] = buf + f * N; }
  float t = clock[0], d_tau = clock[1];
  for (int s = 0; s < nsteps; ++s) {
    t *= expf(d_tau);
    float dt = t * d_tau;
    float inflow_gain = fminf(fmaxf(t / 0.02f, 0.f), 1.f);
    const float *in[6] = {a[0], a[1], a[2], a[3], a[4], a[5]};
    float maxs = oracle_hyp3d_step_planes(p, in, b, solid, dt, inflow_gain, 0, p->nz);
    float dt_cfl = p->cfl / fmaxf(maxs, 1e-9f);
    if (dt > 1.10f * dt_cfl) d_tau *= 0.80f;
    else if (dt < 0.85f * dt_cfl) d_tau *= 1.10f;
    d_tau = fminf(fmaxf(d_tau, 1e-7f), 5e-2f);
    if (dt_hist) dt_hist[s] = dt;
    if (maxs_hist) maxs_hist[s] = maxs;
    for (int f = 0; f < 6; ++f) { float *tmp = a[f]; a[f] = b[f]; b[f] = tmp; }
  }
  if (a[0] != planes[0])
    for (int f = 0; f < 6; ++f) memcpy(planes[f], a[f], N * sizeof(float));
  clock[0] = t;
  clock[1] = d_tau;
  free(buf);
}

/* k_vis :800-905 (VisMode :784-794) through prim_at_xbc :724-749 (== tile_prim above) */
void oracle_hyp3d_vis(const oracle_hyp3d_params *p, const float *const in[6], const uint8_t *solid, int mode,
                      float *out) {
  P = p;
  State s = {in[0], in[1], in[2], in[3], in[4], in[5]};
  for (int z = 0; z < p->nz; ++z)
    for (int y = 0; y < p->ny; ++y)
      for (int x = 0; x < p->nx; ++x) {
        const int i = idx3(x, y, z);
        if (solid[i]) { out[i] = 0.f; continue; }
        int sol;
        Prim q0 = tile_prim(&s, solid, x, y, z, &sol);
        if (mode == 1) { out[i] = logf(1.0f + fmaxf(q0.r, 0.0f)); continue; }
        if (mode == 2) { out[i] = logf(1.0f + fmaxf(q0.p, 0.0f)); continue; }
        float sp = sqrtf(q0.u * q0.u + q0.v * q0.v + q0.w * q0.w);
        if (mode == 3) { out[i] = sp; continue; }
        if (mode == 4) { out[i] = sp / fmaxf(soundspeed(&q0), DENOM_EPS); continue; }
        Prim qxm = tile_prim(&s, solid, x - 1, y, z, &sol), qxp = tile_prim(&s, solid, x + 1, y, z, &sol);
        Prim qym = tile_prim(&s, solid, x, y - 1, z, &sol), qyp = tile_prim(&s, solid, x, y + 1, z, &sol);
        Prim qzm = tile_prim(&s, solid, x, y, z - 1, &sol), qzp = tile_prim(&s, solid, x, y, z + 1, &sol);
        float inv2dx = 0.5f / p->dx, inv2dy = 0.5f / p->dy, inv2dz = 0.5f / p->dz;
        float dudx = (qxp.u - qxm.u) * inv2dx, dudy = (qyp.u - qym.u) * inv2dy, dudz = (qzp.u - qzm.u) * inv2dz;
        float dvdx = (qxp.v - qxm.v) * inv2dx, dvdy = (qyp.v - qym.v) * inv2dy, dvdz = (qzp.v - qzm.v) * inv2dz;
        float dwdx = (qxp.w - qxm.w) * inv2dx, dwdy = (qyp.w - qym.w) * inv2dy, dwdz = (qzp.w - qzm.w) * inv2dz;
        if (mode == 6) { out[i] = dudx + dvdy + dwdz; continue; }
        float wx = dwdy - dvdz, wy = dudz - dwdx, wz = dvdx - dudy;
        if (mode == 5) { out[i] = sqrtf(wx * wx + wy * wy + wz * wz); continue; }
        if (mode == 7) {
          float O12 = 0.5f * (dudy - dvdx), O13 = 0.5f * (dudz - dwdx), O23 = 0.5f * (dvdz - dwdy);
          float Om2 = 2.0f * (O12 * O12 + O13 * O13 + O23 * O23);
          float S12 = 0.5f * (dudy + dvdx), S13 = 0.5f * (dudz + dwdx), S23 = 0.5f * (dvdz + dwdy);
          float Sm2 = (dudx * dudx + dvdy * dvdy + dwdz * dwdz) + 2.0f * (S12 * S12 + S13 * S13 + S23 * S23);
          out[i] = 0.5f * (Om2 - Sm2);
          continue;
        }
        if (mode == 8) { /* th3cs.cu k_schlieren_export :669-672 */
          float ex = (qxp.r - qxm.r) / (2.0f * p->dx), ey = (qyp.r - qym.r) / (2.0f * p->dy),
                ez = (qzp.r - qzm.r) / (2.0f * p->dz);
          out[i] = sqrtf(ex * ex + ey * ey + ez * ez);
          continue;
        }
        float drdx = (qxp.r - qxm.r) * inv2dx, drdy = (qyp.r - qym.r) * inv2dy, drdz = (qzp.r - qzm.r) * inv2dz;
        out[i] = sqrtf(drdx * drdx + drdy * drdy + drdz * drdz);
      }
}
