// hypersonic2d.cu — 2-D compressible-flow (MUSCL-Hancock + HLLC + 4th-order diffusion) update path
// for sm_100a.  Replaces the per-step host sequence of the reference `tau_2d_hypersonic_cuda`
// binary (tau_hypersonic_cuda.cu:1833-1889):
//     k_apply_inflow_left -> k_max_wavespeed_blocks -> k_reduce_block_max -> 8-byte D2H + host dt
//     -> k_predict_face_states -> k_compute_xface_flux -> k_compute_yface_flux -> k_step -> swap
// with ONE kernel per step and no host round trip.  The arithmetic of every stage follows the
// reference device helpers (cited at each function); what changes is where intermediate data
// lives: the reference writes 16 predicted-state planes and 8 flux planes to HBM and re-reads them
// (517 B/cell in fp64); here they never leave registers.
//
// Decomposition ("column marching"): one WARP owns a strip of 30 columns (+1 halo lane each side)
// and marches down a segment of rows.  Segments ("work items") come from a host-built table —
// body-touching items first, then tall ones, tapering to short ones — and the warps of a persistent
// grid (one resident set of CTAs per SM) claim them from a device-side counter, so that every SM
// runs out of work at the same time (the step ends with a global max-wavespeed reduction, so the
// tail of one step cannot be overlapped with the next).
//   * x-direction neighbours (reconstruction stencils, face states, face fluxes) move between
//     lanes with warp shuffles;
//   * y-direction state is carried in registers from row to row: the prims of rows r..r+2, the
//     predicted top state of row r and the flux through the face below it — every y-face flux is
//     computed exactly once;
//   * rows of the conserved state are staged through a per-warp shared-memory ring, 4 rows x 40
//     columns x 4 fields per TMA box (cp.async.bulk.tensor.3d completing on a per-slot mbarrier),
//     prefetched >= 4 rows ahead, so global latency is hidden without occupancy;
//   * boundary conditions are folded in at load time: ghost rows of the planes hold the y-clamp
//     (the kernel refreshes them in its output), edge strips patch inflow / outflow-copy columns in
//     shared memory, the column-0 inflow overwrite (k_apply_inflow_left) is applied as the tile is
//     read; so the marching code itself is branch-uniform;
//   * the max-wavespeed reduction for the NEXT step's dt is fused into the epilogue (warp-shuffle
//     max + one integer atomicMax per warp; max is exactly associative, so dt is bit-identical to a
//     standalone reduction) and dt / sim_t live in device memory.
//
// Templated on the storage/arithmetic type: double reproduces the reference (fp64) to round-off;
// float is the BASELINE configuration ("4096x4096 fp32").
#include "common.cuh"
#include "../../include/tau_b200.h"

#include <math.h>
#include <new>
#include <type_traits>
#include <vector>

namespace {

constexpr int H2_OWN = 30;    // columns owned per warp strip
constexpr int H2_BOXW = 40;   // staged columns: bx .. bx+39, bx = (x0-2) rounded down to a
                              // multiple of 4 (TMA wants a 16-byte aligned box origin in fp32)
constexpr int H2_RB = 4;      // rows per TMA box / ring slot
#ifndef H2_NS_SLOTS
#define H2_NS_SLOTS 3
#endif
constexpr int H2_NS = H2_NS_SLOTS;  // ring slots per warp
#ifndef H2_WARPS_PER_CTA
#define H2_WARPS_PER_CTA 4
#endif
constexpr int H2_WARPS = H2_WARPS_PER_CTA;   // warps per CTA
constexpr int H2_GHOST = 2;   // ghost rows above/below each plane
#ifndef H2_MIN_CTAS
#define H2_MIN_CTAS 5          // fp32: resident CTAs per SM the register allocation must allow (measured: 5 > 4 > 6)
#endif

struct Ctrl {          // device-resident step control (replaces the host dt logic :1852-1869)
  double maxspeed[3];  // rotating slots: step s reads [s%3], reduces into [(s+1)%3], clears [(s+2)%3]
  double sim_t;
  double dt_last;
  unsigned int done_blocks; // CTAs of the running step kernel that have finished
  unsigned int next_item[3]; // work-item claim counters, rotating like maxspeed[]
  // multi-GPU: inbox[slot][src] = bit pattern of rank src's max wavespeed for the step that reads
  // `slot`, written by src with ONE release store when its previous step is complete (its boundary
  // rows are then in our ghost rows).  Zero = not arrived yet (a max wavespeed is >= 1e-12 > 0), so
  // the value is its own flag: one one-way NVLink latency per step, no atomics, no second message.
  unsigned long long inbox[3][8];
  // multi-GPU diagnostics (globaltimer ns, accumulated): waiting for peers, start -> last CTA out,
  // last CTA out -> next start; t_prev_end = end of the previous step
  unsigned long long t_wait, t_busy, t_gap, t_prev_end, t_steps;
};

// Multi-GPU (one process per GPU): peer-memory views of the two slab neighbours' output planes and
// of every rank's Ctrl block, opened through CUDA IPC.  The step kernel pushes its boundary rows
// straight into the neighbours' ghost rows over NVLink; the all-reduce(max) of the wavespeed and
// the step barrier are one message per peer per step, sent by the step kernel itself (last CTA out:
// a release store of this rank's max into every peer's inbox; first thing in the next step: wait
// until every peer's inbox entry is non-zero and fold them into dt).
struct PeerCtrls {
  Ctrl *ctrl[8];
  int world, rank;
};
struct PeerPush {
  void *up_out, *dn_out;       // neighbour planes for the CURRENT output buffer (or null)
  size_t up_plane, dn_plane;   // their plane strides (elements)
  int up_hl;                   // rows owned by the upper neighbour
  PeerCtrls pc;                // world == 1: single GPU or host-driven exchange
};

template <typename R>
struct Params {
  R gamma, gm1, inv_gm1;
  R visc_nu, visc_rho, visc_e;
  R infl_cons[4];  // prim_to_cons(inflow_state())
  R eps_rho, eps_p;
  double cfl, nu_max, infl_speed;
  int W, H_local, H_global, y_begin;
  int nstrips;    // 30-column strips per row
  int nitems;     // entries of the work-item table
  size_t plane;   // elements per plane incl. ghost rows
};

template <typename R> struct Cons4 { R rho, mx, my, E; };
template <typename R> struct Prim4 { R rho, u, v, p; };

// d_fmax/d_fmin :131-136.  (a > b ? a : b) and fmax differ only when an argument is NaN; every
// call site either floors against a finite constant or feeds the guarded HLLC fall-back logic.
template <typename R> __device__ __forceinline__ R rmax(R a, R b) { return fmax(a, b); }
template <typename R> __device__ __forceinline__ R rmin(R a, R b) { return fmin(a, b); }
template <typename R> __device__ __forceinline__ R rabs(R a) { return fabs(a); }  // :137

// reciprocal / sound speed: IEEE in fp64 (parity with the fp64 reference to ~1e-13), single
// MUFU approximations (<= 1 ulp / 2 ulp) in fp32 where they are below the fp32 noise floor.  The
// .ftz forms are single MUFU instructions; without .ftz ptxas wraps each in 6 range-scaling
// instructions (operands here are never subnormal: EPS_RHO = EPS_P = 1e-25).
__device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
__device__ __forceinline__ float rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ double sqrt_pos(double x) { return sqrt(x); }
__device__ __forceinline__ float sqrt_pos(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// tau_hypersonic_cuda.cu:143-159
template <typename R>
__device__ __forceinline__ Prim4<R> cons_to_prim(const Params<R> &P, Cons4<R> c) {
  Prim4<R> p;
  R rho = rmax(c.rho, P.eps_rho);
  R inv = rcp(rho);
  R u = c.mx * inv;
  R v = c.my * inv;
  // contractions written out: the fused step epilogue and the standalone wavespeed scan (first
  // step after init / upload / checkpoint load) must round alike, or a resumed run's first dt
  // differs from the uninterrupted run's in the last bit
  R q2 = fma(v, v, u * u);
  R eint = fma(-(R(0.5) * rho), q2, c.E);
  p.rho = rho;
  p.u = u;
  p.v = v;
  p.p = P.gm1 * rmax(eint, P.eps_p);
  return p;
}
// :161-170
template <typename R>
__device__ __forceinline__ Cons4<R> prim_to_cons(const Params<R> &P, Prim4<R> p) {
  Cons4<R> c;
  R rho = rmax(p.rho, P.eps_rho);
  R pr = rmax(p.p, P.eps_p);
  c.rho = rho;
  c.mx = rho * p.u;
  c.my = rho * p.v;
  c.E = pr * P.inv_gm1 + R(0.5) * rho * (p.u * p.u + p.v * p.v);
  return c;
}
// :172-174
template <typename R>
__device__ __forceinline__ R sound_speed(const Params<R> &P, Prim4<R> p) {
  return sqrt_pos(P.gamma * rmax(p.p, P.eps_p) * rcp(rmax(p.rho, P.eps_rho)));
}
// A face state: primitives plus total energy (what the Riemann solver consumes).  The reference
// round-trips every face state prim -> cons -> prim (prim_to_cons :161, then cons_to_prim again
// inside flux_axis :195 and hllc_axis :520-521); algebraically those are identities, so the
// fused kernel carries (rho,u,v,p,E) once instead.  Differences are at round-off level.
template <typename R> struct Face { R rho, u, v, p, E, a; };  // a = sound speed sqrt(gamma p / rho)

// mc_limiter :217-228.  With dl*dr > 0 the nested minmods of the reference reduce to "the
// argument of smallest magnitude among dl, dr, dc" (2dl and 2dr can never be the smallest) and to 0
// otherwise.  Since dc = (dl+dr)/2 lies between dl and dr it can only be the smallest by a rounding
// ulp, so the limited slope is min-magnitude(dl, dr): identical to the reference except for
// last-bit ties.
template <typename R> __device__ __forceinline__ R mc_limiter(R dl, R dr) {
  // min-magnitude(dl, dr) if they agree in sign, else 0 == median(dl, dr, 0): four min/max, no
  // product, compare, sign transfer or select.  (Differs from the product test only when dl*dr
  // underflows, i.e. for slopes below 1e-19 in fp32 / 1e-154 in fp64.)
  return fmax(fmin(dl, dr), fmin(fmax(dl, dr), R(0)));
}
// :217-221, kept for the known-answer tests
template <typename R> __device__ __forceinline__ R minmod(R a, R b) {
  if (a * b <= R(0)) return R(0);
  return (rabs(a) < rabs(b)) ? a : b;
}
// wall_ghost_prim :262-264
template <typename R>
__device__ __forceinline__ Prim4<R> ghost_prim(Prim4<R> in) {
  return Prim4<R>{in.rho, -in.u, -in.v, in.p};
}
// prim_to_cons(wall_ghost_prim(centre)) :286-287 as a conserved state (diffusion taps)
template <typename R>
__device__ __forceinline__ Cons4<R> ghost_cons(const Params<R> &P, Prim4<R> q) {
  const R rho = rmax(q.rho, P.eps_rho), pr = rmax(q.p, P.eps_p);
  return Cons4<R>{rho, -(rho * q.u), -(rho * q.v),
                  pr * P.inv_gm1 + R(0.5) * rho * (q.u * q.u + q.v * q.v)};
}
// ... and as a face state (Riemann-solver input)
template <typename R>
__device__ __forceinline__ Face<R> ghost_face(const Params<R> &P, Prim4<R> q) {
  const R rho = rmax(q.rho, P.eps_rho), pr = rmax(q.p, P.eps_p);
  return Face<R>{rho, -q.u, -q.v, q.p, pr * P.inv_gm1 + R(0.5) * rho * (q.u * q.u + q.v * q.v),
                 sqrt_pos(P.gamma * q.p * rcp(rho))};
}
// an unpredicted conserved state used directly as a face state (inflow, outflow copy, y-clamp)
template <typename R>
__device__ __forceinline__ Face<R> face_from_cons(const Params<R> &P, Cons4<R> c) {
  const Prim4<R> q = cons_to_prim(P, c);
  return Face<R>{q.rho, q.u, q.v, q.p, c.E, sqrt_pos(P.gamma * q.p * rcp(q.rho))};
}

// enforce_positive_faces :373-398 (rarely taken: only when a limited face state is non-positive).
// Out of line and fed by value so that it does not force the kernel parameters into local memory.
template <typename R> struct PrimPair { Prim4<R> m, p; };
template <typename R>
__device__ __noinline__ PrimPair<R> enforce_positive_faces(R eps_rho, R eps_p, Prim4<R> qm,
                                                           Prim4<R> qc, Prim4<R> qp) {
  for (int it = 0; it < 8; it++) {
    bool bad = (qm.rho <= eps_rho || qp.rho <= eps_rho) || (qm.p <= eps_p || qp.p <= eps_p);
    if (!bad) return PrimPair<R>{qm, qp};
    qm.rho = R(0.5) * (qm.rho + qc.rho);
    qm.u = R(0.5) * (qm.u + qc.u);
    qm.v = R(0.5) * (qm.v + qc.v);
    qm.p = R(0.5) * (qm.p + qc.p);
    qp.rho = R(0.5) * (qp.rho + qc.rho);
    qp.u = R(0.5) * (qp.u + qc.u);
    qp.v = R(0.5) * (qp.v + qc.v);
    qp.p = R(0.5) * (qp.p + qc.p);
  }
  qm.rho = rmax(qm.rho, eps_rho);
  qp.rho = rmax(qp.rho, eps_rho);
  qm.p = rmax(qm.p, eps_p);
  qp.p = rmax(qp.p, eps_p);
  return PrimPair<R>{qm, qp};
}

// Limited reconstruction (reconstruct_limited_faces :400-425) + Hancock half-step predictor
// (k_predict_face_states :923-961, half_step_predict_axis :442-455) for one axis of one cell.
// lo/hi = predicted states on the cell's low ("L"/"B") and high ("R"/"T") face.
template <int AX, typename R>
__device__ __forceinline__ void reconstruct_predict(const Params<R> &P, const Prim4<R> &qm,
                                                    const Prim4<R> &qc, const Prim4<R> &qp,
                                                    R half_dt, Face<R> &lo, Face<R> &hi) {
  const R s_rho = mc_limiter(qc.rho - qm.rho, qp.rho - qc.rho);
  const R s_u = mc_limiter(qc.u - qm.u, qp.u - qc.u);
  const R s_v = mc_limiter(qc.v - qm.v, qp.v - qc.v);
  const R s_p = mc_limiter(qc.p - qm.p, qp.p - qc.p);
  Prim4<R> qL{qc.rho - R(0.5) * s_rho, qc.u - R(0.5) * s_u, qc.v - R(0.5) * s_v,
              qc.p - R(0.5) * s_p};
  Prim4<R> qR{qc.rho + R(0.5) * s_rho, qc.u + R(0.5) * s_u, qc.v + R(0.5) * s_v,
              qc.p + R(0.5) * s_p};
  if ((qL.rho <= P.eps_rho || qR.rho <= P.eps_rho) || (qL.p <= P.eps_p || qR.p <= P.eps_p)) {
    const PrimPair<R> fixed = enforce_positive_faces(P.eps_rho, P.eps_p, qL, qc, qR);
    qL = fixed.m;
    qR = fixed.p;
  }
  // conserved variables and physical flux of both face states (flux_axis :194-203)
  const R mxL = qL.rho * qL.u, myL = qL.rho * qL.v;
  const R mxR = qR.rho * qR.u, myR = qR.rho * qR.v;
  const R EL = qL.p * P.inv_gm1 + R(0.5) * qL.rho * (qL.u * qL.u + qL.v * qL.v);
  const R ER = qR.p * P.inv_gm1 + R(0.5) * qR.rho * (qR.u * qR.u + qR.v * qR.v);
  const R unL = AX == 0 ? qL.u : qL.v, unR = AX == 0 ? qR.u : qR.v;
  const R d_rho = (AX == 0 ? mxR : myR) - (AX == 0 ? mxL : myL);
  const R d_mx = (mxR * unR + (AX == 0 ? qR.p : R(0))) - (mxL * unL + (AX == 0 ? qL.p : R(0)));
  const R d_my = (myR * unR + (AX == 1 ? qR.p : R(0))) - (myL * unL + (AX == 1 ? qL.p : R(0)));
  const R d_E = (ER + qR.p) * unR - (EL + qL.p) * unL;
  // half-step predictor on each face state, back to primitives with the reference's floors
  {
    const R rho = rmax(qL.rho - half_dt * d_rho, P.eps_rho);
    const R inv = rcp(rho);
    const R u = (mxL - half_dt * d_mx) * inv, v = (myL - half_dt * d_my) * inv;
    const R kin = R(0.5) * rho * (u * u + v * v);
    const R pr = rmax(P.gm1 * rmax((EL - half_dt * d_E) - kin, P.eps_p), P.eps_p);
    lo = Face<R>{rho, u, v, pr, pr * P.inv_gm1 + kin, sqrt_pos(P.gamma * pr * inv)};
  }
  {
    const R rho = rmax(qR.rho - half_dt * d_rho, P.eps_rho);
    const R inv = rcp(rho);
    const R u = (mxR - half_dt * d_mx) * inv, v = (myR - half_dt * d_my) * inv;
    const R kin = R(0.5) * rho * (u * u + v * v);
    const R pr = rmax(P.gm1 * rmax((ER - half_dt * d_E) - kin, P.eps_p), P.eps_p);
    hi = Face<R>{rho, u, v, pr, pr * P.inv_gm1 + kin, sqrt_pos(P.gamma * pr * inv)};
  }
}

template <int AX, typename R>
__device__ __forceinline__ Cons4<R> phys_flux(const Face<R> &f) {
  const R un = AX == 0 ? f.u : f.v;
  const R m = f.rho * un;
  return Cons4<R>{m, m * f.u + (AX == 0 ? f.p : R(0)), m * f.v + (AX == 1 ? f.p : R(0)),
                  (f.E + f.p) * un};
}

// hlle_axis :483-509 — only reached through the guarded fall-backs of HLLC
template <int AX, typename R>
__device__ __noinline__ Cons4<R> hlle_flux(Face<R> L, Face<R> Rr, R SL, R SR) {
  const Cons4<R> FL = phys_flux<AX>(L), FR = phys_flux<AX>(Rr);
  if (SL >= R(0)) return FL;
  if (SR <= R(0)) return FR;
  const R denom = SR - SL;
  if (rabs(denom) < R(1e-14))
    return Cons4<R>{R(0.5) * (FL.rho + FR.rho), R(0.5) * (FL.mx + FR.mx), R(0.5) * (FL.my + FR.my),
                    R(0.5) * (FL.E + FR.E)};
  const R inv = R(1) / denom, ss = SL * SR;
  Cons4<R> o;
  o.rho = inv * ((SR * FL.rho + (-SL) * FR.rho) + ss * (Rr.rho - L.rho));
  o.mx = inv * ((SR * FL.mx + (-SL) * FR.mx) + ss * (Rr.rho * Rr.u - L.rho * L.u));
  o.my = inv * ((SR * FL.my + (-SL) * FR.my) + ss * (Rr.rho * Rr.v - L.rho * L.v));
  o.E = inv * ((SR * FL.E + (-SL) * FR.E) + ss * (Rr.E - L.E));
  return o;
}

// hllc_axis :519-606: Davis wave speeds, Toro contact speed, the reference's guarded fall-backs to
// HLLE.  Only the star state on the upwind side of the contact is evaluated; the positivity
// guards of BOTH sides (:568-571) are kept as sign tests.
template <int AX, typename R>
__device__ __forceinline__ Cons4<R> hllc_flux(const Params<R> &P, const Face<R> &L,
                                              const Face<R> &Rr) {
  const R unL = AX == 0 ? L.u : L.v, unR = AX == 0 ? Rr.u : Rr.v;
  const R aL = L.a, aR = Rr.a;
  const R SL = rmin(unL - aL, unR - aR), SR = rmax(unL + aL, unR + aR);
  // supersonic shortcut (:537-540); warp-uniform so whole free-stream strips skip the star state
  if (__all_sync(0xffffffffu, SL >= R(0))) return phys_flux<AX>(L);
  const R qL = L.rho * (SL - unL), qR = Rr.rho * (SR - unR);
  const R num = Rr.p - L.p + qL * unL - qR * unR;
  const R den = qL - qR;
  const R SM = num * rcp(den);
  const R dLS = SL - SM, dRS = SR - SM;
  const bool left = (SL >= R(0)) || (!(SR <= R(0)) && SM >= R(0));
  const Face<R> &K = left ? L : Rr;
  const R SK = left ? SL : SR, unK = left ? unL : unR, qK = left ? qL : qR;
  const R dKS = left ? dLS : dRS;
  const Cons4<R> FK = phys_flux<AX>(K);
  const R pStar = rmax(L.p + qL * (SM - unL), P.eps_p);
  const R invd = rcp(dKS);
  const R rhoStar = qK * invd;
  const R EStar = ((SK - unK) * K.E - K.p * unK + pStar * SM) * invd;
  // guarded fall-backs :548-587.  The reference tests num, den, SM, rho* and E* for finiteness one
  // by one; with finite inputs they can only fail together, so one test on their sum stands in.
  const bool fallback = (rabs(den) < R(1e-14)) || (rabs(dLS) < R(1e-14)) || (rabs(dRS) < R(1e-14)) ||
                        !(qL * dLS > R(0)) || !(qR * dRS > R(0)) ||
                        !isfinite((num + den) + (SM + rhoStar) + EStar);
  const bool supersonic = (SL >= R(0)) || (SR <= R(0));
  if (!supersonic && fallback) return hlle_flux<AX>(L, Rr, SL, SR);
  if (supersonic) return FK;
  const R sn = rhoStar * SM, st = rhoStar * (AX == 0 ? K.v : K.u);
  Cons4<R> F;
  F.rho = FK.rho + SK * (rhoStar - K.rho);
  F.mx = FK.mx + SK * ((AX == 0 ? sn : st) - K.rho * K.u);
  F.my = FK.my + SK * ((AX == 0 ? st : sn) - K.rho * K.v);
  F.E = FK.E + SK * (EStar - K.E);
  return F;
}

template <typename R> __device__ __forceinline__ Face<R> shfl_down_face(Face<R> f) {
  return Face<R>{__shfl_down_sync(0xffffffffu, f.rho, 1), __shfl_down_sync(0xffffffffu, f.u, 1),
                 __shfl_down_sync(0xffffffffu, f.v, 1), __shfl_down_sync(0xffffffffu, f.p, 1),
                 __shfl_down_sync(0xffffffffu, f.E, 1), __shfl_down_sync(0xffffffffu, f.a, 1)};
}
template <typename R> __device__ __forceinline__ Prim4<R> shfl_up_prim(Prim4<R> p) {
  return Prim4<R>{__shfl_up_sync(0xffffffffu, p.rho, 1), __shfl_up_sync(0xffffffffu, p.u, 1),
                  __shfl_up_sync(0xffffffffu, p.v, 1), __shfl_up_sync(0xffffffffu, p.p, 1)};
}
template <typename R> __device__ __forceinline__ Prim4<R> shfl_down_prim(Prim4<R> p) {
  return Prim4<R>{__shfl_down_sync(0xffffffffu, p.rho, 1), __shfl_down_sync(0xffffffffu, p.u, 1),
                  __shfl_down_sync(0xffffffffu, p.v, 1), __shfl_down_sync(0xffffffffu, p.p, 1)};
}
template <typename R> __device__ __forceinline__ Cons4<R> shfl_up_cons(Cons4<R> c) {
  return Cons4<R>{__shfl_up_sync(0xffffffffu, c.rho, 1), __shfl_up_sync(0xffffffffu, c.mx, 1),
                  __shfl_up_sync(0xffffffffu, c.my, 1), __shfl_up_sync(0xffffffffu, c.E, 1)};
}
template <typename R> __device__ __forceinline__ Cons4<R> shfl_down_cons(Cons4<R> c) {
  return Cons4<R>{__shfl_down_sync(0xffffffffu, c.rho, 1), __shfl_down_sync(0xffffffffu, c.mx, 1),
                  __shfl_down_sync(0xffffffffu, c.my, 1), __shfl_down_sync(0xffffffffu, c.E, 1)};
}

// Per-warp view of the shared-memory row ring.  A slot is laid out [field][row-in-box][BOXW columns]
// as the TMA box lands; `off` is the element offset of (field 0, staged row, column 0) — see row_off
// in the step kernel — and fields are FSTRIDE apart.
template <typename R>
struct Ring {
  R *base;
  static constexpr int FSTRIDE = H2_RB * H2_BOXW;
  __device__ __forceinline__ Cons4<R> at(int off, int c) const {
    const R *p = base + off + c;
    return Cons4<R>{p[0], p[FSTRIDE], p[2 * FSTRIDE], p[3 * FSTRIDE]};
  }
};

// 5-tap second derivative (-1, 16, -30, 16, -1)/12 of k_step :1126-1153
template <typename R>
__device__ __forceinline__ R d2(R m2, R m1, R c, R p1, R p2) {
  return (-m2 + R(16) * m1 - R(30) * c + R(16) * p1 - p2) * (R(1) / R(12));
}

template <typename R, bool USE_TMA>
#ifdef H2_MAXNREG   // experiments with other CTA shapes: an explicit register cap instead of a resident-CTA count
__global__ void __maxnreg__(H2_MAXNREG)
#else
__global__ void __launch_bounds__(H2_WARPS * 32, sizeof(R) == 4 ? H2_MIN_CTAS : 2)
#endif
hyp2d_step(const __grid_constant__ CUtensorMap tmU, const Params<R> P, const R *__restrict__ Uin,
           R *__restrict__ Uout, const uint8_t *__restrict__ mask,
           const uint2 *__restrict__ items, Ctrl *__restrict__ ctrl, int step_slot,
           const PeerPush peer) {
#define H2_WARP_RING_ELEMS (H2_NS * (4 * H2_RB * H2_BOXW))
#include "hypersonic2d_prologue.inc"
#undef H2_WARP_RING_ELEMS
  while (item < (unsigned)P.nitems) {
#include "hypersonic2d_item.inc"
  }  // work-item loop
#include "hypersonic2d_epilogue.inc"
}

// One flag per (strip, owned row): does the march read a body cell when it produces this row, i.e. is there one in
// rows r-2 .. r+2 of the strip's staged columns?  The flag depends on the mask alone (ghost rows hold the
// neighbour's rows or the y-clamp images), NOT on how rows are cut into work items or slabs: items are cut where
// it changes, so a cell is always produced by the same march variant — which is what makes results bit-identical
// for every segment height and every slab decomposition (the variants differ in FMA contraction).
template <typename R>
__global__ void hyp2d_flag_rows(const Params<R> P, const uint8_t *__restrict__ mask, uint8_t *__restrict__ flags) {
  const int strip = blockIdx.x, r = blockIdx.y * blockDim.x + threadIdx.x;
  if (r >= P.H_local) return;
  const int x0 = strip * H2_OWN, bx = (x0 - 2) & ~3;
  const int c0 = max(bx, 0), c1 = min(bx + H2_BOXW, P.W);
  int any = 0;
  for (int pr = r; pr <= r + 2 * H2_GHOST; ++pr)  // plane rows r .. r+4 = local rows r-2 .. r+2
    for (int c = c0; c < c1; ++c) any |= mask[(size_t)pr * P.W + c];
  flags[(size_t)strip * P.H_local + r] = any ? 1 : 0;
}

// Standalone max-wavespeed scan of the CURRENT state (first step after init/upload) —
// k_apply_inflow_left + k_max_wavespeed_blocks + k_reduce_block_max :772-847.
template <typename R>
__global__ void hyp2d_wavespeed(const Params<R> P, const R *__restrict__ U,
                                const uint8_t *__restrict__ mask, Ctrl *ctrl, int slot) {
  const size_t n = (size_t)P.W * P.H_local;
  R wmax = R(0);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t o = i + (size_t)H2_GHOST * P.W;
    if (mask[o]) continue;
    R ws;
    if (i % P.W == 0) {
      ws = (R)P.infl_speed;
    } else {
      Prim4<R> p = cons_to_prim(P, Cons4<R>{U[o], U[P.plane + o], U[2 * P.plane + o],
                                            U[3 * P.plane + o]});
      const R a = sound_speed(P, p);
      const R sx = rabs(p.u) + a, sy = rabs(p.v) + a;
      ws = sx > sy ? sx : sy;
      if (!isfinite(ws)) ws = R(1e-12);
    }
    wmax = ws > wmax ? ws : wmax;
  }
  wmax = tau::warp_max(wmax);
  if ((threadIdx.x & 31) == 0 && wmax > R(0)) tau::atomic_max_nonneg(&ctrl->maxspeed[slot], (double)wmax);
}

// ---- multi-GPU frame hand-over without the host (tau_hyp2d_upload_peers_async) --------------------------
// A frame upload is a pseudo-step of the control-slot rotation.  hyp2d_frame_wait: every peer's "finished
// my last step" message is in inbox[slot] (its pushes into our ghost rows are done) -> the H2D copies of the
// new frame may overwrite the planes; it also does the step kernel's clear-two-ahead duty.
__global__ void hyp2d_frame_wait(Ctrl *ctrl, unsigned int *pair_ctr, int slot, const PeerCtrls pc) {
  const int lane = threadIdx.x;
  if (lane < pc.world && lane != pc.rank) {
    const unsigned long long *a = &ctrl->inbox[slot][lane];
    unsigned long long v;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
      if (v != 0ull) break;
      __nanosleep(40);
    }
  }
  __syncwarp();
  const int clr = (slot + 2) % 3;
  if (lane == 0) {
    ctrl->maxspeed[clr] = 0.0;
    ctrl->next_item[clr] = 0u;
    if (pair_ctr) pair_ctr[clr] = 0u;
  }
  if (lane < 8) ctrl->inbox[clr][lane] = 0ull;
}
// hyp2d_frame_announce: push the first / last two owned rows of the uploaded state into the neighbours'
// ghost rows (what the step kernel does for the rows it computes), then one release store per peer of this
// rank's max wavespeed into inbox[next]: the neighbours' next step kernel finds ghost rows and dt inputs.
template <typename R>
__global__ void hyp2d_frame_announce(const Params<R> P, const R *__restrict__ U, const PeerPush peer, Ctrl *ctrl,
                                     int next) {
  const int W = P.W;
  const size_t n = (size_t)4 * H2_GHOST * W;  // 4 planes x 2 rows x W
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), g = (int)((i / W) % H2_GHOST), f = (int)(i / ((size_t)W * H2_GHOST));
    if (peer.up_out != nullptr) {  // my rows 0, 1 -> the upper neighbour's bottom ghost rows
      R *o = static_cast<R *>(peer.up_out);
      o[f * peer.up_plane + (size_t)(peer.up_hl + H2_GHOST + g) * W + x] = U[f * P.plane + (size_t)(H2_GHOST + g) * W + x];
    }
    if (peer.dn_out != nullptr) {  // my rows H_local-2, H_local-1 -> the lower neighbour's top ghost rows
      R *o = static_cast<R *>(peer.dn_out);
      o[f * peer.dn_plane + (size_t)g * W + x] = U[f * P.plane + (size_t)(P.H_local + g) * W + x];
    }
  }
  // last CTA out sends the messages (after every CTA's peer stores)
  __threadfence_system();
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&ctrl->done_blocks, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x == 0) ctrl->done_blocks = 0;
  Ctrl *pc = nullptr;
#pragma unroll
  for (int p = 0; p < 8; ++p)
    if (p == (int)threadIdx.x && p < peer.pc.world && p != peer.pc.rank) pc = peer.pc.ctrl[p];
  if (pc != nullptr) {
    const double m = *reinterpret_cast<volatile double *>(&ctrl->maxspeed[next]);
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(fmax(m, 1e-12)));
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&pc->inbox[next][peer.pc.rank]), "l"(bits) : "memory");
  }
}

// Fill the ghost rows of planes and mask at GLOBAL y-edges with the clamp images (rows 0 / H-1).
template <typename R>
__global__ void hyp2d_fill_ghost(const Params<R> P, R *U, uint8_t *mask, int do_mask) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= P.W) return;
  const bool top = (P.y_begin == 0), bot = (P.y_begin + P.H_local == P.H_global);
  for (int g = 0; g < H2_GHOST; ++g) {
    if (top) {
      const size_t dst = (size_t)g * P.W + x, src = (size_t)H2_GHOST * P.W + x;
      for (int f = 0; f < 4; ++f) U[f * P.plane + dst] = U[f * P.plane + src];
      if (do_mask) mask[dst] = mask[src];
    }
    if (bot) {
      const size_t dst = (size_t)(P.H_local + H2_GHOST + g) * P.W + x;
      const size_t src = (size_t)(P.H_local + H2_GHOST - 1) * P.W + x;
      for (int f = 0; f < 4; ++f) U[f * P.plane + dst] = U[f * P.plane + src];
      if (do_mask) mask[dst] = mask[src];
    }
  }
}

// ---- geometry + initial condition: k_init :740-770 (always evaluated in fp64) -------------------
__device__ __forceinline__ double g_clamp01(double t) { return t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t); }
__device__ __forceinline__ double g_len2(double x, double y) { return sqrt(x * x + y * y); }
__device__ double g_sd_segment(double px, double py, double ax, double ay, double bx, double by) {
  double abx = bx - ax, aby = by - ay, apx = px - ax, apy = py - ay;
  double denom = abx * abx + aby * aby + 1e-30;
  double t = g_clamp01((apx * abx + apy * aby) / denom);
  double qx = ax + t * abx, qy = ay + t * aby;
  return g_len2(px - qx, py - qy);
}
// sdSphereConeCapsule :644-686
__device__ double g_sd_sphere_cone(double x, double y, double Rb, double Rn, double theta) {
  double r = y < 0 ? -y : y;
  double st = sin(theta), ct = cos(theta), tt = tan(theta);
  double xt = Rn * (1.0 - st), rt = Rn * ct;
  double xb = xt + (Rb - rt) / (tt > 1e-30 ? tt : 1e-30);
  double rprof;
  if (x < 0.0) rprof = -1.0;
  else if (x <= xt) {
    double dx = x - Rn, inside = Rn * Rn - dx * dx;
    rprof = inside > 0.0 ? sqrt(inside) : 0.0;
  } else if (x <= xb) rprof = rt + (x - xt) * tt;
  else rprof = -1.0;
  int inside = (x >= 0.0 && x <= xb && r <= rprof);
  double d = fabs(g_len2(x - Rn, r) - Rn);
  double d_cone = g_sd_segment(x, r, xt, rt, xb, Rb);
  double d_base = g_sd_segment(x, y, xb, -Rb, xb, +Rb);
  double d_rim = g_len2(x - xb, r - Rb);
  if (d_cone < d) d = d_cone;
  if (d_base < d) d = d_base;
  if (d_rim < d) d = d_rim;
  return inside ? -d : d;
}


// ================================================================================================
// Render / diagnostic pass — the step on the other side of the hot path in the reference's frame
// loop (tau_hypersonic_cuda.cu:1892-1926): k_render_vals (:1178-1248) -> k_reduce_minmax x n
// (:1273-1320) -> k_compute_inv_range (:1322-1326) -> k_render_pixels (:1250-1271).
// The reference writes a value plane (8 B/cell), 2 x N/256 block extrema, reduces them in a chain of
// launches and re-reads the value plane.  Here: pass 1 evaluates the view value and reduces min/max
// with warp shuffles + one pair of integer atomics per warp (order-preserving key of the double, so
// the result is exact and order independent); pass 2 re-evaluates the value (cheaper than 16 B/cell
// of HBM traffic) and writes RGBA8.  All arithmetic in fp64 on the handle's state, like the reference.
// ================================================================================================
struct RenderPar {
  double gamma, gm1, eps_rho, eps_p, inflow_u;  // inflow_state() :230-238 = (1, inflow_u, 0, 1)
  int W, H_local;
  size_t plane;
};
struct PrimD { double rho, u, v, p; };

__device__ __forceinline__ unsigned long long f64_key(double v) {  // monotonic double -> u64
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_unkey(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}
// cons_to_prim :143-159 in fp64 on plane element `o`
template <typename R>
__device__ __forceinline__ PrimD render_prim(const RenderPar &P, const R *__restrict__ U, size_t o) {
  const double c_rho = (double)U[o], c_mx = (double)U[P.plane + o], c_my = (double)U[2 * P.plane + o],
               c_E = (double)U[3 * P.plane + o];
  PrimD p;
  const double rho = fmax(c_rho, P.eps_rho);
  const double inv = 1.0 / rho;
  p.rho = rho;
  p.u = c_mx * inv;
  p.v = c_my * inv;
  const double kin = 0.5 * rho * (p.u * p.u + p.v * p.v);
  p.p = P.gm1 * fmax(c_E - kin, P.eps_p);
  return p;
}
// sample_prim_bc :706-727 for the neighbour (x+dx, row+dy) of the fluid cell (x, row); rows are
// plane rows (ghost rows hold the y-clamp images at the global edges, the neighbours' rows inside a
// slab decomposition — callers keep them current, see tau_hyp2d_render_minmax).
template <typename R>
__device__ __forceinline__ PrimD render_neighbour(const RenderPar &P, const R *__restrict__ U,
                                                  const uint8_t *__restrict__ mask, const PrimD &centre,
                                                  int x, int prow, int dx, int dy) {
  const int xn = x + dx, rn = prow + dy;
  if (xn < 0) return PrimD{1.0, P.inflow_u, 0.0, 1.0};
  if (xn >= P.W) return render_prim(P, U, (size_t)rn * P.W + (P.W - 1));
  const size_t o = (size_t)rn * P.W + xn;
  if (mask[o]) return PrimD{centre.rho, -centre.u, -centre.v, centre.p};
  return render_prim(P, U, o);
}
template <typename R>
__device__ __forceinline__ double render_value(const RenderPar &P, const R *__restrict__ U,
                                               const uint8_t *__restrict__ mask, int mode, int x, int prow) {
  const PrimD p = render_prim(P, U, (size_t)prow * P.W + x);
  double v;
  if (mode == 0) {
    v = log(p.rho);
  } else if (mode == 1) {
    v = log(p.p);
  } else if (mode == 2) {
    v = sqrt(p.u * p.u + p.v * p.v);
  } else if (mode == 3) {
    const double rhoL = render_neighbour(P, U, mask, p, x, prow, -1, 0).rho;
    const double rhoR = render_neighbour(P, U, mask, p, x, prow, +1, 0).rho;
    const double rhoB = render_neighbour(P, U, mask, p, x, prow, 0, -1).rho;
    const double rhoT = render_neighbour(P, U, mask, p, x, prow, 0, +1).rho;
    const double gx = 0.5 * (rhoR - rhoL), gy = 0.5 * (rhoT - rhoB);
    v = log(1e-12 + sqrt(gx * gx + gy * gy));
  } else if (mode == 4) {
    const PrimD pL = render_neighbour(P, U, mask, p, x, prow, -1, 0);
    const PrimD pR = render_neighbour(P, U, mask, p, x, prow, +1, 0);
    const PrimD pB = render_neighbour(P, U, mask, p, x, prow, 0, -1);
    const PrimD pT = render_neighbour(P, U, mask, p, x, prow, 0, +1);
    v = asinh(0.5 * (pR.v - pL.v) - 0.5 * (pT.u - pB.u));
  } else if (mode == 5) {
    const double a = sqrt(P.gamma * fmax(p.p, P.eps_p) / fmax(p.rho, P.eps_rho));  // sound_speed :172
    v = sqrt(p.u * p.u + p.v * p.v) / fmax(a, 1e-30);
  } else {
    v = log(fmax(p.p / fmax(p.rho, P.eps_rho), 1e-30));
  }
  return isfinite(v) ? v : 0.0;
}

// pass 1: mm[0] = key(min), mm[1] = key(max) over the fluid cells of the slab
template <typename R>
__global__ void __launch_bounds__(256)
hyp2d_render_minmax(const RenderPar P, const R *__restrict__ U, const uint8_t *__restrict__ mask,
                    int mode, unsigned long long *__restrict__ mm) {
  const size_t n = (size_t)P.W * P.H_local;
  double mn = 1e300, mx = -1e300;  // the reference's identities (:1190-1191)
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % P.W), prow = (int)(i / P.W) + H2_GHOST;
    if (mask[(size_t)prow * P.W + x]) continue;
    const double v = render_value(P, U, mask, mode, x, prow);
    mn = fmin(mn, v);
    mx = fmax(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&mm[0], f64_key(mn));
    atomicMax(&mm[1], f64_key(mx));
  }
}
// get_color :692-704
__device__ __forceinline__ uchar4 render_color(double t) {
  t = t < 0 ? 0 : t;
  t = t > 1 ? 1 : t;
  const double rr = 255.0 * fmin(1.0, fmax(0.0, 3.0 * t - 1.0));
  const double gg = 255.0 * fmin(1.0, fmax(0.0, 2.0 - 4.0 * fabs(t - 0.5)));
  const double bb = 255.0 * fmin(1.0, fmax(0.0, 2.0 - 3.0 * t));
  return uchar4{(unsigned char)rr, (unsigned char)gg, (unsigned char)bb, 255};
}
// pass 2: pixels (k_compute_inv_range + k_render_pixels)
template <typename R>
__global__ void __launch_bounds__(256)
hyp2d_render_pixels(const RenderPar P, const R *__restrict__ U, const uint8_t *__restrict__ mask,
                    int mode, double vmin, double vmax, uchar4 *__restrict__ out) {
  const size_t n = (size_t)P.W * P.H_local;
  const double inv_range = 1.0 / fmax(vmax - vmin, 1e-30);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % P.W), prow = (int)(i / P.W) + H2_GHOST;
    if (mask[(size_t)prow * P.W + x]) {
      out[i] = uchar4{110, 110, 110, 255};
      continue;
    }
    const double v = render_value(P, U, mask, mode, x, prow);
    out[i] = render_color((v - vmin) * inv_range);
  }
}

struct Geom { double x0, cy, Rb, Rn, theta; };

template <typename R>
__global__ void hyp2d_init(const Params<R> P, Geom G, R rest_E, R *U, uint8_t *mask) {
  const size_t n = (size_t)P.W * P.H_local;
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % P.W), y = (int)(i / P.W) + P.y_begin;
  const double X = (double)x - G.x0, Y = (double)y - G.cy;
  const double st = sin(G.theta), ct = cos(G.theta), tt = tan(G.theta);
  const double xb = G.Rn * (1.0 - st) + (G.Rb - G.Rn * ct) / (tt > 1e-30 ? tt : 1e-30);  // :729-737
  double sd = g_sd_sphere_cone(X, Y, G.Rb, G.Rn, G.theta) - G.Rb;
  sd = sd > (X - xb) ? sd : (X - xb);
  const uint8_t m = (sd < 0.0) ? 1 : 0;
  const size_t o = i + (size_t)H2_GHOST * P.W;
  mask[o] = m;
  U[o] = P.infl_cons[0];
  U[P.plane + o] = m ? R(0) : P.infl_cons[1];
  U[2 * P.plane + o] = m ? R(0) : P.infl_cons[2];
  U[3 * P.plane + o] = m ? rest_E : P.infl_cons[3];
}

}  // namespace

#include "hypersonic2d_pair.cuh"   // experimental packed two-column kernel (TAU_HYP2D_PAIR=1 | 2 only)
#include "hypersonic2d_fused.cuh"  // experimental: pair + production items in one kernel (TAU_HYP2D_PAIR=2)

// ================================================================================================
// host side
// ================================================================================================
struct tau_hyp2d {
  tau_hyp2d_config cfg;
  int W, H, dtype;  // dtype: 0 = f32, 1 = f64
  int device, y_begin, h_local;
  bool slab, use_tma;
  cudaStream_t stream;
  bool own_stream;
  void *U[2];        // 4 planes each, contiguous, incl. ghost rows
  uint8_t *mask;
  // multi-GPU peer views (CUDA IPC); null/0 when single-GPU or exchange is host-driven
  void *peer_up[2], *peer_dn[2];
  int peer_up_hl, peer_dn_hl;
  PeerCtrls pctrl;
  bool peers_attached;
  uint2 *items;          // work-item table (device): {strip | masked<<31, ys | rows<<20}
  size_t items_cap;
  int nitems;
  int grid_ctas;         // persistent grid: resident CTAs per SM x SMs (capped by the item count)
  bool items_dirty;
  Ctrl *ctrl;
  CUtensorMap tm[2];
  int cur;
  long long steps, launches;
  bool speed_valid;  // ctrl->maxspeed[ctl_slot(h)] holds the max wavespeed of the current state
  long long slot_bias; // control-slot rotations that were not solver steps (tau_hyp2d_upload_peers_async)
  int seg_rows;          // tallest segment of the table (the only height when set by the caller)
  bool seg_auto;         // tapered schedule chosen by build_items (not set by the caller)
  int taper_k, min_rows, max_rows; // schedule tuning (TAU_HYP2D_TAPER_K / _MIN_ROWS / _MAX_ROWS)
  bool max_rows_auto;    // cap of a segment = half a warp's fair share of the slab's rows (not set by the environment)
  cudaEvent_t ev0, ev1;
  bool timed;
  size_t plane_elems;
  // experimental pair mode (TAU_HYP2D_PAIR=1, fp32 + TMA only; see hypersonic2d_pair.cuh)
  bool pair_mode;
  bool fused_mode;                  // TAU_HYP2D_PAIR=2: one kernel per step claims both kinds of item
  uint2 *items_fused;
  int nitems_fused, grid_fused;
  CUtensorMap tm_pair[2];
  uint2 *items_pair, *items_rest;   // interior body-free 60-column items / everything else (30-column)
  int nitems_pair, nitems_rest, grid_pair, grid_rest;
  unsigned int *pair_ctr;           // 3 rotating claim counters of the pair kernel
  bool peers_local;                 // peers are handles of THIS process (tau_hyp2d_group): plain pointers, no IPC mappings
  uchar4 *pixels;             // render target (device), allocated on first use
  unsigned long long *mmkeys; // render min/max keys (device)
};

namespace {

// rotating control slot (maxspeed / next_item / inbox) the NEXT step kernel reads
inline int ctl_slot(const tau_hyp2d *h) { return (int)((h->steps + h->slot_bias) % 3); }

template <typename R>
Params<R> make_params(const tau_hyp2d *h) {
  Params<R> P;
  const tau_hyp2d_config &c = h->cfg;
  P.gamma = (R)c.gamma;
  P.gm1 = (R)(c.gamma - 1.0);
  P.inv_gm1 = (R)(1.0 / (c.gamma - 1.0));
  P.visc_nu = (R)c.visc_nu;
  P.visc_rho = (R)c.visc_rho;
  P.visc_e = (R)c.visc_e;
  P.eps_rho = (R)1e-25;
  P.eps_p = (R)1e-25;
  // inflow_state() :230-238 then prim_to_cons :161-170, in the handle's arithmetic type
  const R rho = R(1), p = R(1);
  const R a = (R)sqrt((double)((R)c.gamma * p / rho));
  const R u = (R)c.inflow_mach * a;
  P.infl_cons[0] = rho;
  P.infl_cons[1] = rho * u;
  P.infl_cons[2] = rho * R(0);
  P.infl_cons[3] = p / P.gm1 + R(0.5) * rho * (u * u + R(0) * R(0));
  // wavespeed of a column-0 inflow cell as k_max_wavespeed_blocks would compute it from those cons
  {
    R r = P.infl_cons[0], inv = R(1) / r, uu = P.infl_cons[1] * inv, vv = P.infl_cons[2] * inv;
    R kin = R(0.5) * r * (uu * uu + vv * vv);
    R eint = P.infl_cons[3] - kin;
    R pr = P.gm1 * (eint > P.eps_p ? eint : P.eps_p);
    R aa = (R)sqrt((double)(P.gamma * pr / r));
    if (sizeof(R) == 4) aa = sqrtf((float)(P.gamma * pr / r));
    R sx = (uu < 0 ? -uu : uu) + aa, sy = (vv < 0 ? -vv : vv) + aa;
    P.infl_speed = (double)(sx > sy ? sx : sy);
  }
  P.cfl = c.cfl;
  P.nu_max = fmax(c.visc_nu, fmax(c.visc_rho, c.visc_e));
  P.W = h->W;
  P.H_local = h->h_local;
  P.H_global = h->H;
  P.y_begin = h->y_begin;
  P.nstrips = (h->W + H2_OWN - 1) / H2_OWN;
  P.nitems = h->nitems;
  P.plane = h->plane_elems;
  return P;
}

template <typename R>
size_t step_smem_bytes() {
  return (size_t)H2_WARPS * H2_NS * 4 * H2_RB * H2_BOXW * sizeof(R) + H2_WARPS * H2_NS * sizeof(uint64_t);
}

template <typename R>
int launch_wavespeed(tau_hyp2d *h, int slot) {
  Params<R> P = make_params<R>(h);
  TAU_CUDA(cudaMemsetAsync(&h->ctrl->maxspeed[slot], 0, sizeof(double), h->stream));
  hyp2d_wavespeed<R><<<148 * 8, 256, 0, h->stream>>>(P, (const R *)h->U[h->cur], h->mask, h->ctrl, slot);
  h->launches++;
  TAU_CUDA(cudaGetLastError());
  return TAU_OK;
}

template <typename R>
int launch_fill_ghost(tau_hyp2d *h, int do_mask) {
  Params<R> P = make_params<R>(h);
  hyp2d_fill_ghost<R><<<(h->W + 255) / 256, 256, 0, h->stream>>>(P, (R *)h->U[h->cur], h->mask, do_mask);
  h->launches++;
  TAU_CUDA(cudaGetLastError());
  return TAU_OK;
}

template <typename R>
int launch_init(tau_hyp2d *h) {
  Params<R> P = make_params<R>(h);
  Geom G{h->cfg.geom_x0, h->cfg.geom_cy, h->cfg.geom_Rb, h->cfg.geom_Rn, h->cfg.geom_theta};
  // body cells start at rest: prim_to_cons({rho, 0, 0, p}) -> E = p/(gamma-1)
  const R rest_E = R(1) / P.gm1 + R(0.5) * R(1) * (R(0) * R(0) + R(0) * R(0));
  const size_t n = (size_t)h->W * h->h_local;
  hyp2d_init<R><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(P, G, rest_E, (R *)h->U[h->cur], h->mask);
  h->launches++;
  TAU_CUDA(cudaGetLastError());
  return TAU_OK;
}

}  // namespace

// The row schedule of the work-item table as a pure host function (unit-tested without a GPU):
// cuts rows [0, h_local) into layers of layer_h[i] rows starting at layer_y[i].  seg_rows > 0:
// uniform layers of that height; seg_rows == 0: guided self-scheduling (see build_items).  A layer
// is at most 4095 rows (12-bit field of the item descriptor).  Returns the number of layers.
extern "C" int tau_hyp2d_plan_layers(int h_local, int nstrips, int resident_warps, int seg_rows, int taper_k,
                                     int min_rows, int max_rows, int *layer_y, int *layer_h, int cap) {
  TAU_REQUIRE(h_local >= 1 && nstrips >= 1 && resident_warps >= 1 && layer_y && layer_h,
              "tau_hyp2d_plan_layers: bad argument");
  TAU_REQUIRE(seg_rows >= 0 && taper_k >= 1 && min_rows >= 1 && max_rows >= min_rows,
              "tau_hyp2d_plan_layers: need seg_rows >= 0, taper_k >= 1, 1 <= min_rows <= max_rows");
  int n = 0, y = 0;
  while (y < h_local) {
    int hgt;
    if (seg_rows == 0) {
      const long long remaining = (long long)(h_local - y) * nstrips;
      hgt = (int)(remaining / ((long long)taper_k * resident_warps));
      if (hgt > max_rows) hgt = max_rows;
      if (hgt < min_rows) hgt = min_rows;
    } else {
      hgt = seg_rows;
    }
    if (hgt > 4095) hgt = 4095;
    if (hgt > h_local - y) hgt = h_local - y;
    TAU_REQUIRE(n < cap, "tau_hyp2d_plan_layers: more than %d layers", cap);
    layer_y[n] = y;
    layer_h[n] = hgt;
    ++n;
    y += hgt;
  }
  return n;
}

namespace {

// Work-item table.  The rows of the slab are cut into layers; a layer is one row segment of every
// strip.  Caller-chosen height (tau_hyp2d_set_seg_rows): uniform layers.  Default: guided
// self-scheduling — a layer's height is (remaining item-rows) / (taper_k x resident warps), clamped
// to [min_rows, max_rows]: tall segments first (the two warm-up rows of a segment are amortised over up to 96
// rows), short ones last (so that all SMs run dry together; the step ends in a global reduction and
// its tail cannot be overlapped with the next step).  Items whose window touches the body run the
// costlier masked march and go to the front of the table (longest-processing-time-first).
template <typename R>
int build_items(tau_hyp2d *h, size_t smem) {
  int dev_sms = 148, per_sm = 1;
  TAU_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, h->device));
  if (h->use_tma)
    TAU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hyp2d_step<R, true>, H2_WARPS * 32, smem));
  else
    TAU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hyp2d_step<R, false>, H2_WARPS * 32, smem));
  if (per_sm < 1) per_sm = 1;
  const int resident_warps = dev_sms * per_sm * H2_WARPS;
  const int nstrips = (h->W + H2_OWN - 1) / H2_OWN;
  std::vector<int> layer_y(h->h_local), layer_h(h->h_local);
  // The tallest segment: half of what a warp marches in the whole step (rows x strips / resident warps), within [12, 96].
  // Measured (profiles/r2_second_session.md; K = 1): 4096 rows 96 > 64 > 128 >> 192; 2048 rows 48 > 64 >> 96; a first
  // layer as tall as the whole fair share leaves nothing to balance with.
  int max_rows = h->max_rows;
  if (h->max_rows_auto) {
    const long long fair = (long long)h->h_local * nstrips / (resident_warps > 0 ? resident_warps : 1);
    max_rows = (int)(fair / 2);
    if (max_rows > 96) max_rows = 96;
    if (max_rows < 12) max_rows = 12;
    if (max_rows < h->min_rows) max_rows = h->min_rows;
  }
  const int nl = tau_hyp2d_plan_layers(h->h_local, nstrips, resident_warps, h->seg_auto ? 0 : h->seg_rows,
                                       h->taper_k, h->min_rows, max_rows, layer_y.data(), layer_h.data(),
                                       h->h_local);
  if (nl < 0) return nl;
  layer_y.resize(nl);
  layer_h.resize(nl);
  if (h->seg_auto) h->seg_rows = layer_h.empty() ? 0 : layer_h[0];
  TAU_REQUIRE(nstrips <= 0xffff && h->h_local < (1 << 20), "tau_hyp2d: grid too large for the item table");
  Params<R> P = make_params<R>(h);
  // per-(strip, row) variant flags: a property of the mask, see hyp2d_flag_rows
  std::vector<uint8_t> flags((size_t)nstrips * h->h_local);
  {
    uint8_t *dflags = nullptr;
    TAU_CUDA(cudaMalloc(&dflags, flags.size()));
    hyp2d_flag_rows<R><<<dim3((unsigned)nstrips, (unsigned)((h->h_local + 127) / 128)), 128, 0, h->stream>>>(P, h->mask, dflags);
    h->launches++;
    TAU_CUDA(cudaGetLastError());
    TAU_CUDA(cudaMemcpyAsync(flags.data(), dflags, flags.size(), cudaMemcpyDeviceToHost, h->stream));
    TAU_CUDA(cudaStreamSynchronize(h->stream));
    TAU_CUDA(cudaFree(dflags));
  }
  // The guided schedule puts its SHORT layers at the end of the slab.  Where those rows hold the body, the masked
  // march (the costliest variant) is cut into many 4-row items, each paying its two warm-up rows: measured at N = 2,
  // the rank with the body at the bottom of its slab took 254 us per step against 233 us for the mirror-image rank.
  // So: if the masked rows sit in the lower half of the slab, lay the layers out bottom-up instead.
  if (h->seg_auto && !getenv("TAU_HYP2D_NO_MIRROR")) {
    double rows = 0.0, centre = 0.0;
    for (int s = 0; s < nstrips; ++s)
      for (int r = 0; r < h->h_local; ++r)
        if (flags[(size_t)s * h->h_local + r]) {
          rows += 1.0;
          centre += r + 0.5;
        }
    if (rows > 0.0 && centre / rows > 0.5 * h->h_local)
      for (size_t l = 0; l < layer_y.size(); ++l) layer_y[l] = h->h_local - layer_y[l] - layer_h[l];
  }
  // layers x strips, every item cut where its rows' flag changes
  std::vector<uint2> tab;
  tab.reserve(layer_y.size() * (size_t)nstrips + 64);
  for (size_t l = 0; l < layer_y.size(); ++l)
    for (int s = 0; s < nstrips; ++s) {
      const uint8_t *f = &flags[(size_t)s * h->h_local];
      int y = layer_y[l];
      const int ye = layer_y[l] + layer_h[l];
      while (y < ye) {
        int e = y + 1;
        while (e < ye && f[e] == f[y]) ++e;
        tab.push_back(make_uint2((unsigned)s | (f[y] ? 0x80000000u : 0u), (unsigned)y | ((unsigned)(e - y) << 20)));
        y = e;
      }
    }
  const size_t n = tab.size();
  if (h->items_cap < n) {
    if (h->items) TAU_CUDA(cudaFree(h->items));
    TAU_CUDA(cudaMalloc(&h->items, n * sizeof(uint2)));
    h->items_cap = n;
  }
  h->nitems = (int)n;
  // masked items first; within each class the table order (tall -> short) is kept
  std::vector<uint2> sorted;
  sorted.reserve(n);
  // (... then the two edge strips, whose rows fold the inflow column / clamp at the outflow edge: the costlier items go first)
  auto edge = [&](const uint2 &d) { const int st = (int)(d.x & 0xffffu); return st == 0 || st == nstrips - 1; };
  for (size_t i = 0; i < n; ++i) if (tab[i].x >> 31) sorted.push_back(tab[i]);
  for (size_t i = 0; i < n; ++i) if (!(tab[i].x >> 31) && edge(tab[i])) sorted.push_back(tab[i]);
  for (size_t i = 0; i < n; ++i) if (!(tab[i].x >> 31) && !edge(tab[i])) sorted.push_back(tab[i]);
  TAU_CUDA(cudaMemcpyAsync(h->items, sorted.data(), n * sizeof(uint2), cudaMemcpyHostToDevice, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));  // `sorted` goes out of scope
  const int need = (int)((n + H2_WARPS - 1) / H2_WARPS);
  h->grid_ctas = dev_sms * per_sm < need ? dev_sms * per_sm : need;
  if (h->grid_ctas < 1) h->grid_ctas = 1;
  h->items_dirty = false;
  return TAU_OK;
}

// ---- experimental pair mode: split the item table, launch pair kernel + production kernel -------------
size_t pair_smem_bytes() {
  return (size_t)H2_WARPS * H2_NS * HP_SLOT * sizeof(float) + H2_WARPS * H2_NS * sizeof(uint64_t);
}
// After build_items<float>: a 60-column super-strip s (30-column strips 2s, 2s+1) of a layer becomes
// one pair item iff both strips exist, neither touches the body, and its 68-column box lies inside
// the grid; every other 30-column item stays with the production kernel.
int build_items_pair(tau_hyp2d *h) {
  const int n = h->nitems;
  std::vector<uint2> all((size_t)n);
  TAU_CUDA(cudaMemcpyAsync(all.data(), h->items, (size_t)n * sizeof(uint2), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  const int nstrips = (h->W + H2_OWN - 1) / H2_OWN;
  // index the 30-column items by (layer start row, strip)
  std::vector<uint2> pair, rest;
  // the fused kernel's single table: masked items first (the costliest), then layer by layer — tall layers
  // first, short ones last, as the guided schedule built them — pair and plain items in strip order
  std::vector<uint2> fused_head, fused_tail;
  std::vector<int> pos((size_t)nstrips);
  std::vector<unsigned> layers;
  for (int i = 0; i < n; ++i) {
    bool seen = false;
    for (unsigned y : layers) seen |= (y == all[i].y);
    if (!seen) layers.push_back(all[i].y);
  }
  for (unsigned ly : layers) {
    for (int s = 0; s < nstrips; ++s) pos[s] = -1;
    for (int i = 0; i < n; ++i)
      if (all[i].y == ly) pos[all[i].x & 0xffffu] = i;
    for (int ss = 0; 2 * ss < nstrips; ++ss) {
      const int a = pos[2 * ss], b = (2 * ss + 1 < nstrips) ? pos[2 * ss + 1] : -1;
      const int x0 = ss * HP_OWN, bx = x0 - 4;
      const bool inside = bx >= 0 && bx + HP_BOXW <= h->W && x0 + HP_OWN <= h->W;
      const bool clean = a >= 0 && b >= 0 && !(all[a].x >> 31) && !(all[b].x >> 31);
      if (inside && clean) {
        pair.push_back(make_uint2((unsigned)ss, ly));
        fused_tail.push_back(make_uint2((unsigned)ss | HF_PAIR_BIT, ly));
      } else {
        for (int i : {a, b})
          if (i >= 0) {
            rest.push_back(all[i]);
            ((all[i].x >> 31) ? fused_head : fused_tail).push_back(all[i]);
          }
      }
    }
  }
  if (h->fused_mode) {
    std::vector<uint2> fused(fused_head);
    fused.insert(fused.end(), fused_tail.begin(), fused_tail.end());
    if (h->items_fused) TAU_CUDA(cudaFree(h->items_fused));
    h->items_fused = nullptr;
    TAU_CUDA(cudaMalloc(&h->items_fused, (fused.size() + 1) * sizeof(uint2)));
    TAU_CUDA(cudaMemcpyAsync(h->items_fused, fused.data(), fused.size() * sizeof(uint2), cudaMemcpyHostToDevice,
                             h->stream));
    TAU_CUDA(cudaStreamSynchronize(h->stream));
    h->nitems_fused = (int)fused.size();
    int sms = 148, per = 1;
    TAU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    TAU_CUDA(cudaFuncSetAttribute(hyp2d_step_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair_smem_bytes()));
    TAU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, hyp2d_step_fused, H2_WARPS * 32, pair_smem_bytes()));
    if (per < 1) per = 1;
    int g = (h->nitems_fused + H2_WARPS - 1) / H2_WARPS;
    if (g > sms * per) g = sms * per;
    h->grid_fused = g < 1 ? 1 : g;
  }
  // The production kernel runs second and cannot overlap the pair kernel (both are ordered behind the
  // previous step), so its duration is its longest item chain.  It has only a few per cent of the work but
  // the whole device: cut its items into short pieces (TAU_HYP2D_REST_ROWS, default 8 rows; 0 = leave them)
  // — a 48-row item alone takes tens of microseconds, an 8-row piece a few; the two warm-up rows per piece
  // are paid on ~3 % of the cells.  A piece keeps its parent's masked flag (conservative: the masked march
  // is correct everywhere).
  int rest_rows = 8;
  if (const char *e = getenv("TAU_HYP2D_REST_ROWS")) rest_rows = atoi(e);
  if (rest_rows > 0) {
    std::vector<uint2> cut;
    for (const uint2 &d : rest) {
      const int ys = (int)(d.y & 0xfffffu), rows = (int)(d.y >> 20);
      for (int off = 0; off < rows; off += rest_rows) {
        const int r = rows - off < rest_rows ? rows - off : rest_rows;
        cut.push_back(make_uint2(d.x, (unsigned)(ys + off) | ((unsigned)r << 20)));
      }
    }
    rest.swap(cut);
  }
  // masked items first in the production table (as before); pair items tall-to-short = layer order
  std::vector<uint2> rest_sorted;
  for (const uint2 &d : rest) if (d.x >> 31) rest_sorted.push_back(d);
  for (const uint2 &d : rest) if (!(d.x >> 31)) rest_sorted.push_back(d);
  if (h->items_pair) TAU_CUDA(cudaFree(h->items_pair));
  if (h->items_rest) TAU_CUDA(cudaFree(h->items_rest));
  h->items_pair = h->items_rest = nullptr;
  TAU_CUDA(cudaMalloc(&h->items_pair, (pair.size() + 1) * sizeof(uint2)));
  TAU_CUDA(cudaMalloc(&h->items_rest, (rest_sorted.size() + 1) * sizeof(uint2)));
  TAU_CUDA(cudaMemcpyAsync(h->items_pair, pair.data(), pair.size() * sizeof(uint2), cudaMemcpyHostToDevice, h->stream));
  TAU_CUDA(cudaMemcpyAsync(h->items_rest, rest_sorted.data(), rest_sorted.size() * sizeof(uint2),
                           cudaMemcpyHostToDevice, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  h->nitems_pair = (int)pair.size();
  h->nitems_rest = (int)rest_sorted.size();
  int dev_sms = 148, per_sm = 1;
  TAU_CUDA(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, h->device));
  TAU_CUDA(cudaFuncSetAttribute(hyp2d_step_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair_smem_bytes()));
  TAU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hyp2d_step_pair, H2_WARPS * 32, pair_smem_bytes()));
  if (per_sm < 1) per_sm = 1;
  auto grid_for = [&](int items, int cap) {
    int g = (items + H2_WARPS - 1) / H2_WARPS;
    if (g > cap) g = cap;
    return g < 1 ? 1 : g;
  };
  h->grid_pair = grid_for(h->nitems_pair, dev_sms * per_sm);
  h->grid_rest = grid_for(h->nitems_rest, h->grid_ctas);
  return TAU_OK;
}

int launch_steps_pair(tau_hyp2d *h, int nsteps, size_t smem) {
  Params<float> P = make_params<float>(h);
  P.nitems = h->nitems_rest;
  for (int s = 0; s < nsteps; ++s) {
    const int a = h->cur, b = a ^ 1;
    const int slot = ctl_slot(h);
    PeerPush peer;
    memset(&peer, 0, sizeof(peer));
    peer.pc.world = 1;
    if (h->peers_attached) {
      peer.up_out = h->peer_up[b];
      peer.dn_out = h->peer_dn[b];
      peer.up_hl = h->peer_up_hl;
      peer.up_plane = (size_t)h->W * (h->peer_up_hl + 2 * H2_GHOST);
      peer.dn_plane = (size_t)h->W * (h->peer_dn_hl + 2 * H2_GHOST);
      peer.pc = h->pctrl;
    }
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t lc = {};
    lc.blockDim = dim3(H2_WARPS * 32);
    lc.stream = h->stream;
    lc.attrs = attr;
    lc.numAttrs = 1;
    if (h->fused_mode) {  // both kinds of item from one table, one kernel (hypersonic2d_fused.cuh)
      Params<float> PF = P;
      PF.nitems = h->nitems_fused;
      lc.gridDim = dim3((unsigned)h->grid_fused);
      lc.dynamicSmemBytes = pair_smem_bytes();
      TAU_CUDA(cudaLaunchKernelEx(&lc, hyp2d_step_fused, h->tm[a], h->tm_pair[a], PF, (const float *)h->U[a],
                                  (float *)h->U[b], (const uint8_t *)h->mask, (const uint2 *)h->items_fused,
                                  h->ctrl, slot, peer));
      h->launches++;
      h->cur = b;
      h->steps++;
      continue;
    }
    if (h->nitems_pair > 0) {  // interior body-free items first
      lc.gridDim = dim3((unsigned)h->grid_pair);
      lc.dynamicSmemBytes = pair_smem_bytes();
      TAU_CUDA(cudaLaunchKernelEx(&lc, hyp2d_step_pair, h->tm_pair[a], P, (float *)h->U[b],
                                  (const uint2 *)h->items_pair, h->nitems_pair, h->ctrl, h->pair_ctr, slot, peer));
      h->launches++;
    }
    // everything else + the step's bookkeeping + the multi-GPU message: the production kernel
    lc.gridDim = dim3((unsigned)h->grid_rest);
    lc.dynamicSmemBytes = smem;
    TAU_CUDA(cudaLaunchKernelEx(&lc, hyp2d_step<float, true>, h->tm[a], P, (const float *)h->U[a], (float *)h->U[b],
                                (const uint8_t *)h->mask, (const uint2 *)h->items_rest, h->ctrl, slot, peer));
    h->launches++;
    h->cur = b;
    h->steps++;
  }
  TAU_CUDA(cudaGetLastError());
  return TAU_OK;
}

template <typename R>
int launch_steps(tau_hyp2d *h, int nsteps) {
  const size_t smem = step_smem_bytes<R>();
  static bool attr_done[2][64] = {};  // per arithmetic type and device (the attribute is per context)
  auto kern_tma = hyp2d_step<R, true>;
  auto kern_gen = hyp2d_step<R, false>;
  const int ti = sizeof(R) == 8, di = h->device & 63;
  if (!attr_done[ti][di]) {
    TAU_CUDA(cudaFuncSetAttribute(kern_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TAU_CUDA(cudaFuncSetAttribute(kern_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[ti][di] = true;
  }
  if (h->items_dirty) {
    int rc = build_items<R>(h, smem);
    if (rc) return rc;
    if (h->pair_mode) {
      rc = build_items_pair(h);
      if (rc) return rc;
    }
  }
  if constexpr (std::is_same<R, float>::value) {
    if (h->pair_mode) return launch_steps_pair(h, nsteps, smem);
  }
  Params<R> P = make_params<R>(h);
  const int grid = h->grid_ctas;
  for (int s = 0; s < nsteps; ++s) {
    const int a = h->cur, b = a ^ 1;
    const int slot = ctl_slot(h);
    PeerPush peer;
    memset(&peer, 0, sizeof(peer));
    peer.pc.world = 1;
    if (h->peers_attached) {
      peer.up_out = h->peer_up[b];
      peer.dn_out = h->peer_dn[b];
      peer.up_hl = h->peer_up_hl;
      peer.up_plane = (size_t)h->W * (h->peer_up_hl + 2 * H2_GHOST);
      peer.dn_plane = (size_t)h->W * (h->peer_dn_hl + 2 * H2_GHOST);
      peer.pc = h->pctrl;
    }
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)grid);
    lc.blockDim = dim3(H2_WARPS * 32);
    lc.dynamicSmemBytes = smem;
    lc.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = 1;
    TAU_CUDA(cudaLaunchKernelEx(&lc, h->use_tma ? kern_tma : kern_gen, h->tm[a], P, (const R *)h->U[a],
                                (R *)h->U[b], (const uint8_t *)h->mask, (const uint2 *)h->items,
                                h->ctrl, slot, peer));
    h->launches++;
    h->cur = b;
    h->steps++;
  }
  TAU_CUDA(cudaGetLastError());
  return TAU_OK;
}

int elem_size(const tau_hyp2d *h) { return h->dtype ? 8 : 4; }

}  // namespace

extern "C" {

// default_config() tau_hypersonic_cuda.cu:1394-1409 with the H-derived geometry evaluated for the
// runtime grid height.
void tau_hyp2d_default_config(tau_hyp2d_config *c, int W, int H) {
  (void)W;
  c->gamma = 1.1;
  c->cfl = 0.25;
  c->visc_nu = 5e-2;
  c->visc_rho = 5e-2;
  c->visc_e = 2e-2;
  c->inflow_mach = 25.0;
  c->geom_x0 = 125.0;
  c->geom_cy = (double)H / 2.0;
  c->geom_Rb = (double)H / 12.0;
  c->geom_Rn = (double)H / 24.0;
  c->geom_theta = 3.14159265358979323846 / 4.0;
  c->steps_per_frame = 2;
}

// The validation block of parse_args() :1545-1637, same messages.
int tau_hyp2d_validate_config(const tau_hyp2d_config *cfg) {
  TAU_REQUIRE(cfg, "tau_hyp2d_validate_config: null config");
  TAU_REQUIRE(cfg->gamma > 1.0, "Invalid --gamma: %.8g (must be > 1).", cfg->gamma);
  TAU_REQUIRE(cfg->cfl > 0.0, "Invalid --cfl: %.8g (must be > 0).", cfg->cfl);
  TAU_REQUIRE(cfg->visc_nu >= 0.0, "Invalid --visc-nu: %.8g (must be >= 0).", cfg->visc_nu);
  TAU_REQUIRE(cfg->visc_rho >= 0.0, "Invalid --visc-rho: %.8g (must be >= 0).", cfg->visc_rho);
  TAU_REQUIRE(cfg->visc_e >= 0.0, "Invalid --visc-e: %.8g (must be >= 0).", cfg->visc_e);
  TAU_REQUIRE(cfg->inflow_mach > 0.0, "Invalid --mach: %.8g (must be > 0).", cfg->inflow_mach);
  TAU_REQUIRE(cfg->steps_per_frame > 0 && cfg->steps_per_frame <= 1024,
              "Invalid --steps-per-frame: %d (must be in [1, 1024]).", cfg->steps_per_frame);
  TAU_REQUIRE(cfg->geom_Rb > 0.0, "Invalid --geom-rb: %.8g (must be > 0).", cfg->geom_Rb);
  TAU_REQUIRE(cfg->geom_Rn > 0.0, "Invalid --geom-rn: %.8g (must be > 0).", cfg->geom_Rn);
  TAU_REQUIRE(cfg->geom_theta > 0.0 && cfg->geom_theta < 3.14159265358979323846 / 2.0,
              "Invalid --geom-theta: %.8g (must be in (0, pi/2)).", cfg->geom_theta);
  const double st = sin(cfg->geom_theta), ct = cos(cfg->geom_theta), tt = tan(cfg->geom_theta);
  const double xt = cfg->geom_Rn * (1.0 - st), rt = cfg->geom_Rn * ct;
  TAU_REQUIRE(cfg->geom_Rb >= rt,
              "Invalid geometry: --geom-rb %.8g is smaller than the tangent radius %.8g implied by "
              "--geom-rn %.8g and --geom-theta %.8g. Require geom-rb >= geom-rn*cos(theta).",
              cfg->geom_Rb, rt, cfg->geom_Rn, cfg->geom_theta);
  TAU_REQUIRE(isfinite(tt) && tt > 0.0,
              "Invalid geometry: tan(theta)=%.8g for --geom-theta %.8g must be finite and positive.",
              tt, cfg->geom_theta);
  const double xb = xt + (cfg->geom_Rb - rt) / tt;
  TAU_REQUIRE(isfinite(xb), "Invalid geometry: computed xb is non-finite (xb=%.8g).", xb);
  TAU_REQUIRE(xb >= xt, "Invalid geometry: computed xb %.8g is behind cone tangent point xt %.8g.",
              xb, xt);
  return TAU_OK;
}

int tau_hyp2d_create(const tau_hyp2d_config *cfg, int W, int H, int dtype, int device, int y_begin,
                     int h_local, void *stream, tau_hyp2d **out) {
  TAU_REQUIRE(cfg && out, "tau_hyp2d_create: null argument");
  TAU_REQUIRE(W >= 4 && H >= 1, "tau_hyp2d_create: grid %d x %d too small (W >= 4, H >= 1)", W, H);
  TAU_REQUIRE(dtype == 0 || dtype == 1, "tau_hyp2d_create: dtype must be 0 (f32) or 1 (f64)");
  TAU_REQUIRE(y_begin >= 0 && h_local > 0 && y_begin + h_local <= H,
              "tau_hyp2d_create: slab rows [%d,%d) outside [0,%d)", y_begin, y_begin + h_local, H);
  TAU_REQUIRE(h_local == H || h_local >= H2_GHOST,
              "tau_hyp2d_create: a slab needs at least %d rows", H2_GHOST);
  int rc = tau_hyp2d_validate_config(cfg);
  if (rc) return rc;
  if (tau_device_count() <= 0) {
    tau_set_error("tau_hyp2d_create: no CUDA device (this library has no CPU fallback)");
    return TAU_ERR_NODEV;
  }
  TAU_CUDA(cudaSetDevice(device));
  tau_hyp2d *h = new (std::nothrow) tau_hyp2d();
  if (!h) return TAU_ERR_NOMEM;
  h->cfg = *cfg;
  h->W = W;
  h->H = H;
  h->dtype = dtype;
  h->device = device;
  h->y_begin = y_begin;
  h->h_local = h_local;
  h->slab = (h_local != H);
  const int es = dtype ? 8 : 4;
  h->use_tma = ((size_t)W * es) % 16 == 0;
  h->cur = 0;
  h->steps = 0;
  h->launches = 0;
  h->speed_valid = false;
  h->timed = false;
  h->peer_up[0] = h->peer_up[1] = h->peer_dn[0] = h->peer_dn[1] = nullptr;
  h->peer_up_hl = h->peer_dn_hl = 0;
  memset(&h->pctrl, 0, sizeof(h->pctrl));
  h->pctrl.world = 1;
  h->peers_attached = false;
  h->peers_local = false;
  h->pair_mode = false;
  h->items_pair = h->items_rest = nullptr;
  h->nitems_pair = h->nitems_rest = 0;
  h->grid_pair = h->grid_rest = 1;
  h->pair_ctr = nullptr;
  memset(h->tm_pair, 0, sizeof(h->tm_pair));
  h->pixels = nullptr;
  h->mmkeys = nullptr;
  h->items = nullptr;
  h->items_cap = 0;
  h->nitems = 0;
  h->grid_ctas = 1;
  h->items_dirty = true;
  h->seg_rows = 24;
  h->seg_auto = true;
  h->taper_k = 1;   // (2 / 4 / 48 until the edge strips stopped being the tail of every step: profiles/r2_second_session.md)
  h->min_rows = 6;
  h->max_rows = 96;
  h->max_rows_auto = true;
  if (const char *e = getenv("TAU_HYP2D_MAX_ROWS")) {
    int v = atoi(e);
    if (v >= 4) {
      h->max_rows = v;
      h->max_rows_auto = false;
    }
  }
  if (const char *e = getenv("TAU_HYP2D_SEG_ROWS")) {
    int v = atoi(e);
    if (v >= 4) {
      h->seg_rows = v;
      h->seg_auto = false;
    }
  }
  if (const char *e = getenv("TAU_HYP2D_TAPER_K")) {
    int v = atoi(e);
    if (v >= 1) h->taper_k = v;
  }
  if (const char *e = getenv("TAU_HYP2D_MIN_ROWS")) {
    int v = atoi(e);
    if (v >= 2) h->min_rows = v;
  }
  if (stream) {
    h->stream = (cudaStream_t)stream;
    h->own_stream = false;
  } else {
    TAU_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  h->plane_elems = (size_t)W * (h_local + 2 * H2_GHOST);
  const size_t bytes = 4 * h->plane_elems * es;
  for (int b = 0; b < 2; ++b) {
    TAU_CUDA(cudaMalloc(&h->U[b], bytes));
    TAU_CUDA(cudaMemsetAsync(h->U[b], 0, bytes, h->stream));
  }
  TAU_CUDA(cudaMalloc(&h->mask, h->plane_elems));
  TAU_CUDA(cudaMemsetAsync(h->mask, 0, h->plane_elems, h->stream));
  TAU_CUDA(cudaMalloc(&h->ctrl, sizeof(Ctrl)));
  TAU_CUDA(cudaMemsetAsync(h->ctrl, 0, sizeof(Ctrl), h->stream));
  if (h->use_tma) {
    const uint64_t dims[3] = {(uint64_t)W, (uint64_t)(h_local + 2 * H2_GHOST), 4};
    const uint64_t strides[2] = {(uint64_t)W * es, (uint64_t)h->plane_elems * es};
    const uint32_t box[3] = {H2_BOXW, H2_RB, 4};
    for (int b = 0; b < 2; ++b) {
      rc = tau_make_tensor_map(&h->tm[b], h->U[b], es, 3, dims, strides, box);
      if (rc) return rc;
    }
  } else {
    memset(h->tm, 0, sizeof(h->tm));
  }
  if (const char *e = getenv("TAU_HYP2D_PAIR")) {  // experimental, see hypersonic2d_pair.cuh
    if ((atoi(e) == 1 || atoi(e) == 2) && h->use_tma && dtype == 0 && W >= 2 * HP_BOXW) {
      const uint64_t dims[3] = {(uint64_t)W, (uint64_t)(h_local + 2 * H2_GHOST), 4};
      const uint64_t strides[2] = {(uint64_t)W * es, (uint64_t)h->plane_elems * es};
      const uint32_t box[3] = {HP_BOXW, H2_RB, 4};
      for (int b = 0; b < 2; ++b) {
        rc = tau_make_tensor_map(&h->tm_pair[b], h->U[b], es, 3, dims, strides, box);
        if (rc) return rc;
      }
      TAU_CUDA(cudaMalloc(&h->pair_ctr, 3 * sizeof(unsigned int)));
      TAU_CUDA(cudaMemsetAsync(h->pair_ctr, 0, 3 * sizeof(unsigned int), h->stream));
      h->pair_mode = true;
      h->fused_mode = atoi(e) == 2;
    }
  }
  TAU_CUDA(cudaEventCreate(&h->ev0));
  TAU_CUDA(cudaEventCreate(&h->ev1));
  *out = h;
  return TAU_OK;
}

static int hyp2d_state_changed(tau_hyp2d *h, int fill_mask) {
  int rc = h->dtype ? launch_fill_ghost<double>(h, fill_mask) : launch_fill_ghost<float>(h, fill_mask);
  if (rc) return rc;
  const int slot = ctl_slot(h);
  rc = h->dtype ? launch_wavespeed<double>(h, slot) : launch_wavespeed<float>(h, slot);
  if (rc) return rc;
  h->speed_valid = true;
  return TAU_OK;
}

int tau_hyp2d_init(tau_hyp2d *h) {
  TAU_REQUIRE(h, "tau_hyp2d_init: null handle");
  TAU_CUDA(cudaSetDevice(h->device));
  h->steps = 0;
  h->slot_bias = 0;
  h->items_dirty = true;
  TAU_CUDA(cudaMemsetAsync(h->ctrl, 0, sizeof(Ctrl), h->stream));
  // the pair kernel's claim counters rotate with the step count like the control slots: a stale one would make
  // the first step after a re-init claim from nwarps_grid + stale and skip work items
  if (h->pair_ctr) TAU_CUDA(cudaMemsetAsync(h->pair_ctr, 0, 3 * sizeof(unsigned int), h->stream));
  int rc = h->dtype ? launch_init<double>(h) : launch_init<float>(h);
  if (rc) return rc;
  rc = hyp2d_state_changed(h, 1);
  if (rc) return rc;
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

static int hyp2d_upload_impl(tau_hyp2d *h, const void *const planes[4], const uint8_t *mask, bool sync);
int tau_hyp2d_upload(tau_hyp2d *h, const void *const planes[4], const uint8_t *mask) {
  return hyp2d_upload_impl(h, planes, mask, true);
}
// Enqueue-only forms for pipelined frame loops (pinned host buffers; they must stay valid until
// tau_hyp2d_sync): with two handles on two streams the upload of frame i+1 overlaps the download of
// frame i on the two PCIe copy engines.
int tau_hyp2d_upload_async(tau_hyp2d *h, const void *const planes[4], const uint8_t *mask) {
  return hyp2d_upload_impl(h, planes, mask, false);
}
static int hyp2d_upload_impl(tau_hyp2d *h, const void *const planes[4], const uint8_t *mask, bool sync) {
  TAU_REQUIRE(h && planes, "tau_hyp2d_upload: null argument");
  TAU_CUDA(cudaSetDevice(h->device));
  const int es = elem_size(h);
  const size_t row0 = (size_t)H2_GHOST * h->W;
  const size_t n = (size_t)h->W * h->h_local;
  for (int f = 0; f < 4; ++f) {
    TAU_REQUIRE(planes[f], "tau_hyp2d_upload: null plane %d", f);
    TAU_CUDA(cudaMemcpyAsync((char *)h->U[h->cur] + (f * h->plane_elems + row0) * es, planes[f], n * es,
                             cudaMemcpyHostToDevice, h->stream));
  }
  if (mask) {
    TAU_CUDA(cudaMemcpyAsync(h->mask + row0, mask, n, cudaMemcpyHostToDevice, h->stream));
    h->items_dirty = true;
  }
  int rc = hyp2d_state_changed(h, mask ? 1 : 0);
  if (rc) return rc;
  if (sync) TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

// Multi-GPU frame upload with no host exchange (peers attached; pinned host planes; enqueue-only).
// Stream order: wait until every peer finished its previous step (their pushes into our ghost rows are
// done) -> H2D of the owned rows -> global-edge ghosts -> max-wavespeed scan -> push the boundary rows into
// the neighbours' ghost rows + one message per peer.  The next tau_hyp2d_step finds everything it needs, as
// after hyp2d_sync_state + tau_hyp2d_peers_ready, but without NCCL, host synchronisation or a barrier.
// EVERY rank must call it for the same frame (the control slots rotate in lock step).  The body mask is
// static across such frames (distribute it once with tau_hyp2d_upload + the host-driven exchange).
int tau_hyp2d_upload_peers_async(tau_hyp2d *h, const void *const planes[4]) {
  TAU_REQUIRE(h && planes, "tau_hyp2d_upload_peers_async: null argument");
  TAU_REQUIRE(h->peers_attached, "tau_hyp2d_upload_peers_async: no peers attached (use tau_hyp2d_upload_async)");
  TAU_REQUIRE(h->speed_valid, "tau_hyp2d_upload_peers_async: no state yet (first frame: tau_hyp2d_upload + exchange)");
  for (int f = 0; f < 4; ++f) TAU_REQUIRE(planes[f], "tau_hyp2d_upload_peers_async: null plane %d", f);
  TAU_CUDA(cudaSetDevice(h->device));
  if (h->items_dirty) {  // pair mode needs its claim counters allocated before the first clear
    const int rc = h->dtype ? launch_steps<double>(h, 0) : launch_steps<float>(h, 0);
    if (rc) return rc;
  }
  const int slot = ctl_slot(h), next = (slot + 1) % 3;
  hyp2d_frame_wait<<<1, 32, 0, h->stream>>>(h->ctrl, h->pair_mode ? h->pair_ctr : nullptr, slot, h->pctrl);
  h->launches++;
  const int es = elem_size(h);
  const size_t row0 = (size_t)H2_GHOST * h->W, n = (size_t)h->W * h->h_local;
  for (int f = 0; f < 4; ++f)
    TAU_CUDA(cudaMemcpyAsync((char *)h->U[h->cur] + (f * h->plane_elems + row0) * es, planes[f], n * es,
                             cudaMemcpyHostToDevice, h->stream));
  int rc = h->dtype ? launch_fill_ghost<double>(h, 0) : launch_fill_ghost<float>(h, 0);
  if (rc) return rc;
  rc = h->dtype ? launch_wavespeed<double>(h, next) : launch_wavespeed<float>(h, next);
  if (rc) return rc;
  PeerPush peer;
  memset(&peer, 0, sizeof(peer));
  peer.up_out = h->peer_up[h->cur];
  peer.dn_out = h->peer_dn[h->cur];
  peer.up_hl = h->peer_up_hl;
  peer.up_plane = (size_t)h->W * (h->peer_up_hl + 2 * H2_GHOST);
  peer.dn_plane = (size_t)h->W * (h->peer_dn_hl + 2 * H2_GHOST);
  peer.pc = h->pctrl;
  const int grid = (int)((4 * H2_GHOST * (size_t)h->W + 255) / 256);
  if (h->dtype)
    hyp2d_frame_announce<double><<<grid, 256, 0, h->stream>>>(make_params<double>(h), (const double *)h->U[h->cur], peer, h->ctrl, next);
  else
    hyp2d_frame_announce<float><<<grid, 256, 0, h->stream>>>(make_params<float>(h), (const float *)h->U[h->cur], peer, h->ctrl, next);
  h->launches++;
  TAU_CUDA(cudaGetLastError());
  h->slot_bias++;  // the upload consumed control slot `slot`
  return TAU_OK;
}

int tau_hyp2d_step(tau_hyp2d *h, int nsteps) {
  TAU_REQUIRE(h, "tau_hyp2d_step: null handle");
  TAU_REQUIRE(nsteps >= 0, "tau_hyp2d_step: nsteps must be >= 0");
  TAU_REQUIRE(!(h->slab && !h->peers_attached && nsteps > 1),
              "tau_hyp2d_step: a slab handle without attached peers advances one step per call "
              "(ghost rows and the max wavespeed must be exchanged in between)");
  TAU_REQUIRE(h->speed_valid, "tau_hyp2d_step: no state (call tau_hyp2d_init or tau_hyp2d_upload)");
  TAU_CUDA(cudaSetDevice(h->device));
  TAU_CUDA(cudaEventRecord(h->ev0, h->stream));
  int rc = h->dtype ? launch_steps<double>(h, nsteps) : launch_steps<float>(h, nsteps);
  if (rc) return rc;
  TAU_CUDA(cudaEventRecord(h->ev1, h->stream));
  h->timed = true;
  return TAU_OK;
}

int tau_hyp2d_clock(tau_hyp2d *h, double *sim_t, double *dt_last) {
  TAU_REQUIRE(h, "tau_hyp2d_clock: null handle");
  Ctrl c;
  TAU_CUDA(cudaMemcpyAsync(&c, h->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  if (sim_t) *sim_t = c.sim_t;
  if (dt_last) *dt_last = c.dt_last;
  return TAU_OK;
}

static int hyp2d_download_impl(tau_hyp2d *h, void *const planes[4], uint8_t *mask, bool sync);
int tau_hyp2d_download(tau_hyp2d *h, void *const planes[4], uint8_t *mask) {
  return hyp2d_download_impl(h, planes, mask, true);
}
int tau_hyp2d_download_async(tau_hyp2d *h, void *const planes[4], uint8_t *mask) {
  return hyp2d_download_impl(h, planes, mask, false);
}
static int hyp2d_download_impl(tau_hyp2d *h, void *const planes[4], uint8_t *mask, bool sync) {
  TAU_REQUIRE(h && planes, "tau_hyp2d_download: null argument");
  TAU_CUDA(cudaSetDevice(h->device));
  const int es = elem_size(h);
  const size_t row0 = (size_t)H2_GHOST * h->W;
  const size_t n = (size_t)h->W * h->h_local;
  for (int f = 0; f < 4; ++f)
    if (planes[f])
      TAU_CUDA(cudaMemcpyAsync(planes[f], (char *)h->U[h->cur] + (f * h->plane_elems + row0) * es, n * es,
                               cudaMemcpyDeviceToHost, h->stream));
  if (mask) TAU_CUDA(cudaMemcpyAsync(mask, h->mask + row0, n, cudaMemcpyDeviceToHost, h->stream));
  if (sync) TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

int tau_hyp2d_sync(tau_hyp2d *h) {
  TAU_REQUIRE(h, "tau_hyp2d_sync: null handle");
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

int tau_hyp2d_device_state(tau_hyp2d *h, void **planes, uint8_t **mask, double **maxspeed_slot) {
  TAU_REQUIRE(h, "tau_hyp2d_device_state: null handle");
  if (planes) *planes = h->U[h->cur];
  if (mask) *mask = h->mask;
  if (maxspeed_slot) *maxspeed_slot = &h->ctrl->maxspeed[ctl_slot(h)];
  return TAU_OK;
}

// ---- multi-GPU peer plumbing (CUDA IPC; one process per GPU on one box) ----------------------
int tau_hyp2d_ipc_export(tau_hyp2d *h, void *out, size_t out_bytes) {
  TAU_REQUIRE(h && out, "tau_hyp2d_ipc_export: null argument");
  TAU_REQUIRE(out_bytes >= 3 * sizeof(cudaIpcMemHandle_t),
              "tau_hyp2d_ipc_export: buffer must hold %zu bytes", 3 * sizeof(cudaIpcMemHandle_t));
  TAU_CUDA(cudaSetDevice(h->device));
  cudaIpcMemHandle_t *hd = static_cast<cudaIpcMemHandle_t *>(out);
  TAU_CUDA(cudaIpcGetMemHandle(&hd[0], h->U[0]));
  TAU_CUDA(cudaIpcGetMemHandle(&hd[1], h->U[1]));
  TAU_CUDA(cudaIpcGetMemHandle(&hd[2], h->ctrl));
  return TAU_OK;
}

int tau_hyp2d_ipc_attach(tau_hyp2d *h, int rank, int world, const void *all_handles,
                         const int *h_locals) {
  TAU_REQUIRE(h && all_handles && h_locals, "tau_hyp2d_ipc_attach: null argument");
  TAU_REQUIRE(world >= 2 && world <= 8 && rank >= 0 && rank < world,
              "tau_hyp2d_ipc_attach: need 2 <= world <= 8 (got rank %d of %d)", rank, world);
  TAU_REQUIRE(h_locals[rank] == h->h_local, "tau_hyp2d_ipc_attach: h_locals[rank] mismatch");
  TAU_CUDA(cudaSetDevice(h->device));
  const cudaIpcMemHandle_t *hd = static_cast<const cudaIpcMemHandle_t *>(all_handles);
  h->pctrl.world = world;
  h->pctrl.rank = rank;
  for (int p = 0; p < world; ++p) {
    if (p == rank) {
      h->pctrl.ctrl[p] = h->ctrl;
      continue;
    }
    void *ptr = nullptr;
    TAU_CUDA(cudaIpcOpenMemHandle(&ptr, hd[3 * p + 2], cudaIpcMemLazyEnablePeerAccess));
    h->pctrl.ctrl[p] = static_cast<Ctrl *>(ptr);
  }
  // y is clamped, not periodic: the slabs form a chain
  if (rank > 0) {
    for (int b = 0; b < 2; ++b)
      TAU_CUDA(cudaIpcOpenMemHandle(&h->peer_up[b], hd[3 * (rank - 1) + b], cudaIpcMemLazyEnablePeerAccess));
    h->peer_up_hl = h_locals[rank - 1];
  }
  if (rank < world - 1) {
    for (int b = 0; b < 2; ++b)
      TAU_CUDA(cudaIpcOpenMemHandle(&h->peer_dn[b], hd[3 * (rank + 1) + b], cudaIpcMemLazyEnablePeerAccess));
    h->peer_dn_hl = h_locals[rank + 1];
  }
  h->peers_attached = true;
  return TAU_OK;
}

// Unmap the peers' planes and control blocks.  CUDA wants an importer to close its mapping before the
// exporter frees the memory: every rank detaches, THEN (after a barrier between the processes,
// slab.hyp2d_detach_peers) the handles are destroyed.  tau_hyp2d_destroy calls it as a fallback.
int tau_hyp2d_ipc_detach(tau_hyp2d *h) {
  TAU_REQUIRE(h, "tau_hyp2d_ipc_detach: null handle");
  if (!h->peers_attached) return TAU_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (int b = 0; b < 2; ++b) {
    if (!h->peers_local) {
      if (h->peer_up[b]) cudaIpcCloseMemHandle(h->peer_up[b]);
      if (h->peer_dn[b]) cudaIpcCloseMemHandle(h->peer_dn[b]);
    }
    h->peer_up[b] = h->peer_dn[b] = nullptr;
  }
  for (int p = 0; p < h->pctrl.world; ++p) {
    if (!h->peers_local && p != h->pctrl.rank && h->pctrl.ctrl[p]) cudaIpcCloseMemHandle(h->pctrl.ctrl[p]);
    h->pctrl.ctrl[p] = nullptr;
  }
  h->peers_local = false;
  h->pctrl.world = 1;
  h->peers_attached = false;
  return TAU_OK;
}

// Call once after the caller has exchanged the ghost rows of the current state and all-reduced
// the wavespeed slot on the host side (init / upload): arms the first device-side barrier.
int tau_hyp2d_peers_ready(tau_hyp2d *h) {
  TAU_REQUIRE(h && h->peers_attached, "tau_hyp2d_peers_ready: no peers attached");
  TAU_CUDA(cudaSetDevice(h->device));
  // the caller has all-reduced maxspeed[steps%3] already: every peer "arrives" with a tiny value
  unsigned long long box[3][8];
  memset(box, 0, sizeof(box));
  const double tiny = 1e-12;
  for (int p = 0; p < h->pctrl.world; ++p)
    if (p != h->pctrl.rank) memcpy(&box[ctl_slot(h)][p], &tiny, sizeof(double));
  TAU_CUDA(cudaMemsetAsync(&h->ctrl->done_blocks, 0, sizeof(unsigned int), h->stream));
  TAU_CUDA(cudaMemcpyAsync(h->ctrl->inbox, box, sizeof(box), cudaMemcpyHostToDevice, h->stream));
  TAU_CUDA(cudaMemsetAsync(&h->ctrl->t_wait, 0, 5 * sizeof(unsigned long long), h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}


// ---- render pass (reference frame loop :1892-1926) ---------------------------------------------
static RenderPar make_render_par(const tau_hyp2d *h) {
  RenderPar P;
  P.gamma = h->cfg.gamma;
  P.gm1 = h->cfg.gamma - 1.0;
  P.eps_rho = 1e-25;
  P.eps_p = 1e-25;
  P.inflow_u = h->cfg.inflow_mach * sqrt(h->cfg.gamma * 1.0 / 1.0);  // inflow_state() :230-238
  P.W = h->W;
  P.H_local = h->h_local;
  P.plane = h->plane_elems;
  return P;
}

int tau_hyp2d_render_minmax(tau_hyp2d *h, int view_mode, double minmax[2]) {
  TAU_REQUIRE(h && minmax, "tau_hyp2d_render_minmax: null argument");
  TAU_REQUIRE(view_mode >= 0 && view_mode <= 6, "tau_hyp2d_render_minmax: view mode %d not in [0, 6]", view_mode);
  TAU_REQUIRE(h->speed_valid, "tau_hyp2d_render_minmax: no state (call tau_hyp2d_init or tau_hyp2d_upload)");
  TAU_CUDA(cudaSetDevice(h->device));
  if (!h->mmkeys) TAU_CUDA(cudaMalloc(&h->mmkeys, 2 * sizeof(unsigned long long)));
  const unsigned long long init[2] = {~0ull, 0ull};
  TAU_CUDA(cudaMemcpyAsync(h->mmkeys, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
  const RenderPar P = make_render_par(h);
  const int grid = 148 * 8;
  if (h->dtype)
    hyp2d_render_minmax<double><<<grid, 256, 0, h->stream>>>(P, (const double *)h->U[h->cur], h->mask, view_mode, h->mmkeys);
  else
    hyp2d_render_minmax<float><<<grid, 256, 0, h->stream>>>(P, (const float *)h->U[h->cur], h->mask, view_mode, h->mmkeys);
  h->launches++;
  TAU_CUDA(cudaGetLastError());
  unsigned long long k[2];
  TAU_CUDA(cudaMemcpyAsync(k, h->mmkeys, sizeof(k), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < 2; ++i) {  // f64_unkey on the host
    const unsigned long long b = (k[i] >> 63) ? (k[i] & 0x7fffffffffffffffull) : ~k[i];
    memcpy(&minmax[i], &b, sizeof(double));
  }
  return TAU_OK;
}

int tau_hyp2d_render_pixels(tau_hyp2d *h, int view_mode, const double minmax[2], uint32_t *rgba) {
  TAU_REQUIRE(h && minmax && rgba, "tau_hyp2d_render_pixels: null argument");
  TAU_REQUIRE(view_mode >= 0 && view_mode <= 6, "tau_hyp2d_render_pixels: view mode %d not in [0, 6]", view_mode);
  TAU_REQUIRE(h->speed_valid, "tau_hyp2d_render_pixels: no state (call tau_hyp2d_init or tau_hyp2d_upload)");
  TAU_CUDA(cudaSetDevice(h->device));
  const size_t n = (size_t)h->W * h->h_local;
  if (!h->pixels) TAU_CUDA(cudaMalloc(&h->pixels, n * sizeof(uchar4)));
  const RenderPar P = make_render_par(h);
  const int grid = 148 * 8;
  if (h->dtype)
    hyp2d_render_pixels<double><<<grid, 256, 0, h->stream>>>(P, (const double *)h->U[h->cur], h->mask, view_mode,
                                                              minmax[0], minmax[1], h->pixels);
  else
    hyp2d_render_pixels<float><<<grid, 256, 0, h->stream>>>(P, (const float *)h->U[h->cur], h->mask, view_mode,
                                                             minmax[0], minmax[1], h->pixels);
  h->launches++;
  TAU_CUDA(cudaGetLastError());
  TAU_CUDA(cudaMemcpyAsync(rgba, h->pixels, n * sizeof(uchar4), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

int tau_hyp2d_render(tau_hyp2d *h, int view_mode, uint32_t *rgba, double minmax_out[2]) {
  double mm[2];
  int rc = tau_hyp2d_render_minmax(h, view_mode, mm);
  if (rc) return rc;
  rc = tau_hyp2d_render_pixels(h, view_mode, mm, rgba);
  if (rc) return rc;
  if (minmax_out) {
    minmax_out[0] = mm[0];
    minmax_out[1] = mm[1];
  }
  return TAU_OK;
}

// Multi-GPU diagnostics: average per-step times in microseconds since tau_hyp2d_peers_ready():
// out[0] waiting for the peers' messages, out[1] peers seen -> last CTA out, out[2] last CTA out ->
// next step's start (launch gap), out[3] = steps counted.
int tau_hyp2d_peer_timing(tau_hyp2d *h, double out[4]) {
  TAU_REQUIRE(h && out, "tau_hyp2d_peer_timing: null argument");
  Ctrl c;
  TAU_CUDA(cudaMemcpyAsync(&c, h->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  const double n = c.t_steps ? (double)c.t_steps : 1.0;
  out[0] = (double)c.t_wait / n * 1e-3;
  out[1] = (double)c.t_busy / n * 1e-3;
  out[2] = (double)c.t_gap / n * 1e-3;
  out[3] = (double)c.t_steps;
  return TAU_OK;
}

int tau_hyp2d_set_seg_rows(tau_hyp2d *h, int rows) {
  TAU_REQUIRE(h && rows >= 4, "tau_hyp2d_set_seg_rows: rows must be >= 4");
  h->seg_rows = rows;
  h->seg_auto = false;
  h->items_dirty = true;
  return TAU_OK;
}

int tau_hyp2d_get_seg_rows(tau_hyp2d *h) { return h ? h->seg_rows : -1; }

int tau_hyp2d_describe(tau_hyp2d *h, int *W, int *H, int *dtype, int *y_begin, int *h_local,
                       tau_hyp2d_config *cfg) {
  TAU_REQUIRE(h, "tau_hyp2d_describe: null handle");
  if (W) *W = h->W;
  if (H) *H = h->H;
  if (dtype) *dtype = h->dtype;
  if (y_begin) *y_begin = h->y_begin;
  if (h_local) *h_local = h->h_local;
  if (cfg) *cfg = h->cfg;
  return TAU_OK;
}

// Sizes of the work-item table(s) the next step will use (builds them if the schedule is stale):
// out[0] = items of the production kernel's table, out[1] = its persistent grid (CTAs); pair mode:
// out[2] = 60-column items of the pair kernel, out[3] = items left to the production kernel,
// out[4] / out[5] = the two grids.  Host bookkeeping only.
int tau_hyp2d_work_items(tau_hyp2d *h, int out[6]) {
  TAU_REQUIRE(h && out, "tau_hyp2d_work_items: null argument");
  TAU_CUDA(cudaSetDevice(h->device));
  if (h->items_dirty) {  // a zero-step launch sets the kernel attributes and builds the tables
    const int rc = h->dtype ? launch_steps<double>(h, 0) : launch_steps<float>(h, 0);
    if (rc) return rc;
  }
  out[0] = h->nitems;
  out[1] = h->grid_ctas;
  out[2] = h->pair_mode ? h->nitems_pair : 0;
  out[3] = h->pair_mode ? h->nitems_rest : 0;
  out[4] = h->pair_mode ? h->grid_pair : 0;
  out[5] = h->pair_mode ? h->grid_rest : 0;
  return TAU_OK;
}

// Restore the device-resident clock after tau_hyp2d_upload (checkpoint/resume).  The step counter
// only selects the rotating control slot; the max wavespeed of the restored state was re-scanned by
// the upload, so the next dt equals the one the original run would have taken.
int tau_hyp2d_set_clock(tau_hyp2d *h, double sim_t, long long steps_done) {
  TAU_REQUIRE(h && steps_done >= 0, "tau_hyp2d_set_clock: bad argument");
  TAU_REQUIRE(h->speed_valid, "tau_hyp2d_set_clock: no state (call tau_hyp2d_upload first)");
  TAU_CUDA(cudaSetDevice(h->device));
  Ctrl c;
  TAU_CUDA(cudaMemcpyAsync(&c, h->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  const double ms = c.maxspeed[ctl_slot(h)];
  memset(&c, 0, sizeof(c));
  h->steps = steps_done;
  h->slot_bias = 0;
  c.maxspeed[ctl_slot(h)] = ms;
  c.sim_t = sim_t;
  TAU_CUDA(cudaMemcpyAsync(h->ctrl, &c, sizeof(Ctrl), cudaMemcpyHostToDevice, h->stream));
  if (h->pair_ctr) TAU_CUDA(cudaMemsetAsync(h->pair_ctr, 0, 3 * sizeof(unsigned int), h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}
long long tau_hyp2d_steps_done(tau_hyp2d *h) { return h ? h->steps : -1; }
// which step kernel this handle launches: 0 hyp2d_step, 1 hyp2d_step_pair + hyp2d_step, 2 hyp2d_step_fused
int tau_hyp2d_kernel_mode(tau_hyp2d *h) { return !h ? -1 : (h->pair_mode ? (h->fused_mode ? 2 : 1) : 0); }
long long tau_hyp2d_launch_count(tau_hyp2d *h) { return h ? h->launches : -1; }

int tau_hyp2d_last_step_ms(tau_hyp2d *h, float *ms) {
  TAU_REQUIRE(h && ms, "tau_hyp2d_last_step_ms: null argument");
  TAU_REQUIRE(h->timed, "tau_hyp2d_last_step_ms: no step has been timed yet");
  TAU_CUDA(cudaEventSynchronize(h->ev1));
  TAU_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return TAU_OK;
}

int tau_hyp2d_destroy(tau_hyp2d *h) {
  if (!h) return TAU_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  tau_hyp2d_ipc_detach(h);  // unmap what tau_hyp2d_ipc_attach opened (the peers free their own memory)
  cudaFree(h->ctrl);
  if (h->items) cudaFree(h->items);
  if (h->items_pair) cudaFree(h->items_pair);
  if (h->items_rest) cudaFree(h->items_rest);
  if (h->items_fused) cudaFree(h->items_fused);
  if (h->pair_ctr) cudaFree(h->pair_ctr);
  if (h->pixels) cudaFree(h->pixels);
  if (h->mmkeys) cudaFree(h->mmkeys);
  cudaFree(h->mask);
  cudaFree(h->U[1]);
  cudaFree(h->U[0]);
  cudaEventDestroy(h->ev1);
  cudaEventDestroy(h->ev0);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return TAU_OK;
}


// ---- multi-GPU behind the C boundary: ONE process, one slab handle per device ---------------------------------
// tau_<solver>_create(cfg, dims, ngpus, &handle).  The slab handles are the ones torchrun's ranks
// use (y-slabs, chain, 2 ghost rows); here their peers live in the same process, so the neighbours' planes and
// every rank's control block are plain device pointers made reachable with cudaDeviceEnablePeerAccess instead of
// CUDA-IPC mappings.  Everything per step stays on the devices (boundary rows pushed over NVLink by the step
// kernel, one release store per peer as all-reduce(max) + barrier); the host only enqueues.  The hand-over after
// init / upload (ghost rows, mask ghost rows, max wavespeed) is cudaMemcpyPeer + a host max: no NCCL needed.
struct tau_hyp2d_group {
  int n, W, H, dtype;
  tau_hyp2d *h[8];
  int dev[8], y0[8], hl[8];
  int chunk;  // steps enqueued per device before moving on to the next one
};

namespace {

int group_sync_state(tau_hyp2d_group *g, bool with_mask) {
  if (g->n == 1) return TAU_OK;
  const int es = g->dtype ? 8 : 4;
  const size_t rowb = (size_t)g->W * es, ghostb = H2_GHOST * rowb;
  for (int i = 0; i < g->n; ++i) TAU_CUDA(cudaStreamSynchronize(g->h[i]->stream));
  for (int i = 0; i + 1 < g->n; ++i) {
    tau_hyp2d *a = g->h[i], *b = g->h[i + 1];  // a above b
    for (int f = 0; f < 4; ++f) {
      char *pa = (char *)a->U[a->cur] + (size_t)f * a->plane_elems * es, *pb = (char *)b->U[b->cur] + (size_t)f * b->plane_elems * es;
      // a's last two owned rows -> b's upper ghost rows; b's first two owned rows -> a's lower ghost rows
      TAU_CUDA(cudaMemcpyPeerAsync(pb, b->device, pa + (size_t)a->h_local * rowb, a->device, ghostb, a->stream));
      TAU_CUDA(cudaMemcpyPeerAsync(pa + (size_t)(a->h_local + H2_GHOST) * rowb, a->device, pb + ghostb, b->device, ghostb, b->stream));
    }
    if (with_mask) {
      const size_t mrow = (size_t)g->W;
      TAU_CUDA(cudaMemcpyPeerAsync(b->mask, b->device, a->mask + (size_t)a->h_local * mrow, a->device, H2_GHOST * mrow, a->stream));
      TAU_CUDA(cudaMemcpyPeerAsync(a->mask + (size_t)(a->h_local + H2_GHOST) * mrow, a->device, b->mask + H2_GHOST * mrow, b->device,
                                   H2_GHOST * mrow, b->stream));
    }
  }
  // all-reduce(max) of the wavespeed slot the next step reads (exact: max is associative)
  double m = 0.0;
  for (int i = 0; i < g->n; ++i) {
    tau_hyp2d *h = g->h[i];
    TAU_CUDA(cudaSetDevice(h->device));
    TAU_CUDA(cudaStreamSynchronize(h->stream));
    double v = 0.0;
    TAU_CUDA(cudaMemcpy(&v, &h->ctrl->maxspeed[ctl_slot(h)], sizeof(double), cudaMemcpyDeviceToHost));
    if (v > m) m = v;
  }
  for (int i = 0; i < g->n; ++i) {
    tau_hyp2d *h = g->h[i];
    TAU_CUDA(cudaSetDevice(h->device));
    TAU_CUDA(cudaMemcpy(&h->ctrl->maxspeed[ctl_slot(h)], &m, sizeof(double), cudaMemcpyHostToDevice));
    if (with_mask) h->items_dirty = true;  // the ghost rows of the mask decide which items take the masked march
    const int rc = tau_hyp2d_peers_ready(h);
    if (rc) return rc;
  }
  return TAU_OK;
}

}  // namespace

int tau_hyp2d_group_create(const tau_hyp2d_config *cfg, int W, int H, int dtype, int ngpus, const int *devices,
                           tau_hyp2d_group **out) {
  TAU_REQUIRE(cfg && out, "tau_hyp2d_group_create: null argument");
  TAU_REQUIRE(ngpus >= 1 && ngpus <= 8, "tau_hyp2d_group_create: ngpus must be in [1, 8] (got %d)", ngpus);
  TAU_REQUIRE(H >= ngpus * 2 * H2_GHOST, "tau_hyp2d_group_create: %d rows cannot be split over %d GPUs", H, ngpus);
  const int have = tau_device_count();
  if (have <= 0) {
    tau_set_error("tau_hyp2d_group_create: no CUDA device (this library has no CPU fallback)");
    return TAU_ERR_NODEV;
  }
  TAU_REQUIRE(have >= ngpus, "tau_hyp2d_group_create: %d GPUs requested, %d visible", ngpus, have);
  tau_hyp2d_group *g = new (std::nothrow) tau_hyp2d_group();
  if (!g) return TAU_ERR_NOMEM;
  memset(g, 0, sizeof(*g));
  g->n = ngpus; g->W = W; g->H = H; g->dtype = dtype;
  g->chunk = 16;
  if (const char *e = getenv("TAU_HYP2D_GROUP_CHUNK")) {
    const int v = atoi(e);
    if (v >= 1) g->chunk = v;
  }
  const int base = H / ngpus, rem = H % ngpus;  // balanced contiguous partition, earlier slabs take the remainder
  int y = 0, rc = TAU_OK;
  for (int i = 0; i < ngpus && !rc; ++i) {
    g->dev[i] = devices ? devices[i] : i;
    g->y0[i] = y;
    g->hl[i] = base + (i < rem ? 1 : 0);
    y += g->hl[i];
    rc = tau_hyp2d_create(cfg, W, H, dtype, g->dev[i], g->y0[i], g->hl[i], nullptr, &g->h[i]);
  }
  if (!rc && ngpus > 1) {
    for (int i = 0; i < ngpus && !rc; ++i) {
      if (cudaSetDevice(g->dev[i]) != cudaSuccess) rc = TAU_ERR_CUDA;
      for (int j = 0; j < ngpus && !rc; ++j) {
        if (j == i) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, g->dev[i], g->dev[j]);
        if (!can) {
          tau_set_error("tau_hyp2d_group_create: device %d cannot access device %d (no NVLink/P2P path)", g->dev[i], g->dev[j]);
          rc = TAU_ERR_CUDA;
          break;
        }
        const cudaError_t e = cudaDeviceEnablePeerAccess(g->dev[j], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          tau_set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", g->dev[i], g->dev[j], cudaGetErrorString(e));
          rc = TAU_ERR_CUDA;
        }
        cudaGetLastError();
      }
    }
    for (int i = 0; i < ngpus && !rc; ++i) {  // what tau_hyp2d_ipc_attach does, with in-process pointers
      tau_hyp2d *h = g->h[i];
      h->pctrl.world = ngpus;
      h->pctrl.rank = i;
      for (int p = 0; p < ngpus; ++p) h->pctrl.ctrl[p] = g->h[p]->ctrl;
      if (i > 0) {
        for (int b = 0; b < 2; ++b) h->peer_up[b] = g->h[i - 1]->U[b];
        h->peer_up_hl = g->hl[i - 1];
      }
      if (i < ngpus - 1) {
        for (int b = 0; b < 2; ++b) h->peer_dn[b] = g->h[i + 1]->U[b];
        h->peer_dn_hl = g->hl[i + 1];
      }
      h->peers_attached = true;
      h->peers_local = true;
    }
  }
  if (rc) {
    for (int i = 0; i < ngpus; ++i)
      if (g->h[i]) {
        g->h[i]->peers_attached = false;
        tau_hyp2d_destroy(g->h[i]);
      }
    delete g;
    return rc;
  }
  *out = g;
  return TAU_OK;
}

int tau_hyp2d_group_size(tau_hyp2d_group *g) { return g ? g->n : -1; }
int tau_hyp2d_group_member(tau_hyp2d_group *g, int i, tau_hyp2d **h, int *y_begin, int *h_local) {
  TAU_REQUIRE(g && i >= 0 && i < g->n, "tau_hyp2d_group_member: bad argument");
  if (h) *h = g->h[i];
  if (y_begin) *y_begin = g->y0[i];
  if (h_local) *h_local = g->hl[i];
  return TAU_OK;
}

int tau_hyp2d_group_init(tau_hyp2d_group *g) {  // k_init on every slab, then the hand-over
  TAU_REQUIRE(g, "tau_hyp2d_group_init: null handle");
  for (int i = 0; i < g->n; ++i) {
    const int rc = tau_hyp2d_init(g->h[i]);
    if (rc) return rc;
  }
  return group_sync_state(g, true);
}

// planes / mask cover the WHOLE grid (H x W, reference layout); every device receives its rows
int tau_hyp2d_group_upload(tau_hyp2d_group *g, const void *const planes[4], const uint8_t *mask) {
  TAU_REQUIRE(g && planes, "tau_hyp2d_group_upload: null argument");
  const int es = g->dtype ? 8 : 4;
  for (int i = 0; i < g->n; ++i) {
    const void *pl[4];
    for (int f = 0; f < 4; ++f) {
      TAU_REQUIRE(planes[f], "tau_hyp2d_group_upload: null plane %d", f);
      pl[f] = (const char *)planes[f] + (size_t)g->y0[i] * g->W * es;
    }
    const int rc = tau_hyp2d_upload(g->h[i], pl, mask ? mask + (size_t)g->y0[i] * g->W : nullptr);
    if (rc) return rc;
  }
  return group_sync_state(g, mask != nullptr);
}

// THE hot path on ngpus devices: one step kernel per device per step, nothing else; no host synchronisation.
int tau_hyp2d_group_step(tau_hyp2d_group *g, int nsteps) {
  TAU_REQUIRE(g && nsteps >= 0, "tau_hyp2d_group_step: bad argument");
  for (int done = 0; done < nsteps; done += g->chunk) {
    const int k = nsteps - done < g->chunk ? nsteps - done : g->chunk;
    for (int i = 0; i < g->n; ++i) {
      const int rc = tau_hyp2d_step(g->h[i], k);
      if (rc) return rc;
    }
  }
  return TAU_OK;
}

int tau_hyp2d_group_sync(tau_hyp2d_group *g) {
  TAU_REQUIRE(g, "tau_hyp2d_group_sync: null handle");
  for (int i = 0; i < g->n; ++i) {
    const int rc = tau_hyp2d_sync(g->h[i]);
    if (rc) return rc;
  }
  return TAU_OK;
}

int tau_hyp2d_group_clock(tau_hyp2d_group *g, double *sim_t, double *dt_last) {
  TAU_REQUIRE(g, "tau_hyp2d_group_clock: null handle");
  return tau_hyp2d_clock(g->h[0], sim_t, dt_last);  // every slab carries the same clock (same dt sequence)
}

int tau_hyp2d_group_download(tau_hyp2d_group *g, void *const planes[4], uint8_t *mask) {
  TAU_REQUIRE(g && planes, "tau_hyp2d_group_download: null argument");
  const int es = g->dtype ? 8 : 4;
  for (int i = 0; i < g->n; ++i) {
    void *pl[4];
    for (int f = 0; f < 4; ++f) pl[f] = planes[f] ? (char *)planes[f] + (size_t)g->y0[i] * g->W * es : nullptr;
    for (int f = 0; f < 4; ++f) TAU_REQUIRE(pl[f], "tau_hyp2d_group_download: null plane %d", f);
    const int rc = tau_hyp2d_download(g->h[i], pl, mask ? mask + (size_t)g->y0[i] * g->W : nullptr);
    if (rc) return rc;
  }
  return TAU_OK;
}

// render pass over the whole grid: per-slab min/max, host fold, per-slab pixels with the global range
int tau_hyp2d_group_render(tau_hyp2d_group *g, int view_mode, uint32_t *rgba, double minmax_out[2]) {
  TAU_REQUIRE(g && rgba, "tau_hyp2d_group_render: null argument");
  double mm[2] = {1e300, -1e300};
  for (int i = 0; i < g->n; ++i) {
    double m[2];
    const int rc = tau_hyp2d_render_minmax(g->h[i], view_mode, m);
    if (rc) return rc;
    if (m[0] < mm[0]) mm[0] = m[0];
    if (m[1] > mm[1]) mm[1] = m[1];
  }
  for (int i = 0; i < g->n; ++i) {
    const int rc = tau_hyp2d_render_pixels(g->h[i], view_mode, mm, rgba + (size_t)g->y0[i] * g->W);
    if (rc) return rc;
  }
  if (minmax_out) {
    minmax_out[0] = mm[0];
    minmax_out[1] = mm[1];
  }
  return TAU_OK;
}

long long tau_hyp2d_group_launch_count(tau_hyp2d_group *g) {
  if (!g) return -1;
  long long n = 0;
  for (int i = 0; i < g->n; ++i) n += g->h[i]->launches;
  return n;
}

int tau_hyp2d_group_destroy(tau_hyp2d_group *g) {
  if (!g) return TAU_OK;
  for (int i = 0; i < g->n; ++i) tau_hyp2d_sync(g->h[i]);
  for (int i = 0; i < g->n; ++i) tau_hyp2d_ipc_detach(g->h[i]);  // forget the peers' pointers before anything is freed
  for (int i = g->n - 1; i >= 0; --i) tau_hyp2d_destroy(g->h[i]);
  delete g;
  return TAU_OK;
}

}  // extern "C"
