// hypersonic2d_pair.cuh — EXPERIMENTAL step kernel, compiled into the library but only launched when
// TAU_HYP2D_PAIR=1 is set at handle creation.  NOT YET RUN ON HARDWARE (written when the round-1 GPU
// budget was spent); the default path does not touch it.  Its results are verified on the CPU emulator
// (tests/hostemu, tests/test_hostemu_cpu.py: equal to the production kernel to FMA rounding on random
// fields with walls, in slab mode, at the benchmarked grid's extents, under UBSan alignment checks and in
// a 400-case fuzz sweep); speed, register pressure at run time and memory ordering are what hardware
// still has to show.  Included by hypersonic2d.cu (same translation unit: it uses Params, Ctrl, PeerPush
// and the scalar helpers).
//
// The "two adjacent columns per lane" formulation of the 2-D hypersonic step for the interior,
// body-free work items (97 % of the items at 4096^2): a warp owns a strip of 60 columns; lane l = 1..30
// owns columns x0 + 2(l-1) and +1 as the two halves of a float2 (lanes 0 / 31 hold the halo pairs);
// arithmetic uses the packed FADD2/FMUL2/FFMA2 of sm_100, everything without a packed form (min/max,
// compares, selects, MUFU) runs per half.  In pair mode this kernel takes the interior unmasked items
// and the production kernel (hyp2d_step) is launched right after it on the masked / edge items (cut into
// 8-row pieces, build_items_pair), does the step's bookkeeping (sim_t, slot clearing) and sends the multi-GPU message.  Numerics follow
// hypersonic2d.cu (same expression trees, FMA contractions spelled out).  Static facts and the
// projection: profiles/hyp2d_pair_probe_r1.md.
#pragma once


namespace {

#ifndef HP_MIN_CTAS
#define HP_MIN_CTAS 3   // resident CTAs per SM the register allocation must allow: 3 -> 168 registers, no spills;
#endif                  // 4 -> 128 registers with ~190 bytes of spill stores per thread (compile-time fact; untried)
constexpr int HP_OWN = 60;   // columns owned per warp strip
constexpr int HP_BOXW = 68;  // staged columns: bx = x0 - 4 (x0 = 60 s is a multiple of 4), bx .. bx+67
constexpr int HP_SLOT = 4 * H2_RB * HP_BOXW;  // floats per ring slot

// ---- a pair of cells: packed add / mul / fma, per-half everything else ---------------------------------
struct f2 {
  float2 v;
  __device__ __forceinline__ f2() {}
  __device__ __forceinline__ f2(float a, float b) : v(make_float2(a, b)) {}
  __device__ __forceinline__ explicit f2(float a) : v(make_float2(a, a)) {}
  __device__ __forceinline__ explicit f2(float2 a) : v(a) {}
};
struct m2 { bool x, y; };
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { f2 r; r.v = __fadd2_rn(a.v, b.v); return r; }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { f2 r; r.v = __fmul2_rn(a.v, b.v); return r; }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { f2 r; r.v = __ffma2_rn(b.v, make_float2(-1.f, -1.f), a.v); return r; }
__device__ __forceinline__ f2 operator-(f2 a) { return f2(-a.v.x, -a.v.y); }
__device__ __forceinline__ f2 fma_(f2 a, f2 b, f2 c) { f2 r; r.v = __ffma2_rn(a.v, b.v, c.v); return r; }
__device__ __forceinline__ f2 max_(f2 a, f2 b) { return f2(fmaxf(a.v.x, b.v.x), fmaxf(a.v.y, b.v.y)); }
__device__ __forceinline__ f2 min_(f2 a, f2 b) { return f2(fminf(a.v.x, b.v.x), fminf(a.v.y, b.v.y)); }
__device__ __forceinline__ f2 abs_(f2 a) { return f2(fabsf(a.v.x), fabsf(a.v.y)); }
__device__ __forceinline__ f2 rcp_(f2 a) { return f2(rcp(a.v.x), rcp(a.v.y)); }
__device__ __forceinline__ f2 sqrt_(f2 a) { return f2(sqrt_pos(a.v.x), sqrt_pos(a.v.y)); }
__device__ __forceinline__ m2 ge0(f2 a) { return m2{a.v.x >= 0.f, a.v.y >= 0.f}; }
__device__ __forceinline__ m2 le0(f2 a) { return m2{a.v.x <= 0.f, a.v.y <= 0.f}; }
__device__ __forceinline__ m2 operator|(m2 a, m2 b) { return m2{a.x || b.x, a.y || b.y}; }
__device__ __forceinline__ m2 operator&(m2 a, m2 b) { return m2{a.x && b.x, a.y && b.y}; }
__device__ __forceinline__ m2 operator!(m2 a) { return m2{!a.x, !a.y}; }
__device__ __forceinline__ bool any(m2 a) { return a.x || a.y; }
__device__ __forceinline__ f2 sel(m2 m, f2 a, f2 b) { return f2(m.x ? a.v.x : b.v.x, m.y ? a.v.y : b.v.y); }
__device__ __forceinline__ f2 shfl_up_y_to_x(f2 a, float own_x_src) {  // (lane-1's .y, own_x_src)
  return f2(__shfl_up_sync(0xffffffffu, a.v.y, 1), own_x_src);
}

struct Prim2 { f2 rho, u, v, p; };
struct Face2 { f2 rho, u, v, p, E, a; };
struct Cons2 { f2 rho, mx, my, E; };

__device__ __forceinline__ Prim4<float> half(const Prim2 &q, int k) {
  return k ? Prim4<float>{q.rho.v.y, q.u.v.y, q.v.v.y, q.p.v.y} : Prim4<float>{q.rho.v.x, q.u.v.x, q.v.v.x, q.p.v.x};
}
__device__ __forceinline__ Face<float> half(const Face2 &q, int k) {
  return k ? Face<float>{q.rho.v.y, q.u.v.y, q.v.v.y, q.p.v.y, q.E.v.y, q.a.v.y}
           : Face<float>{q.rho.v.x, q.u.v.x, q.v.v.x, q.p.v.x, q.E.v.x, q.a.v.x};
}

// cons_to_prim :143-159 (contractions as in hypersonic2d.cu)
__device__ __forceinline__ Prim2 cons_to_prim2(const Params<float> &P, const Cons2 &c) {
  const f2 rho = max_(c.rho, f2(P.eps_rho)), inv = rcp_(rho);
  const f2 u = c.mx * inv, v = c.my * inv;
  const f2 eint = fma_(-(f2(0.5f) * rho), fma_(v, v, u * u), c.E);
  return Prim2{rho, u, v, f2(P.gm1) * max_(eint, f2(P.eps_p))};
}
__device__ __forceinline__ f2 limiter2(f2 dl, f2 dr) {  // median(dl, dr, 0), see mc_limiter
  return max_(min_(dl, dr), min_(max_(dl, dr), f2(0.f)));
}

// reconstruct_predict<AX> of hypersonic2d.cu for a pair of cells
template <int AX>
__device__ __forceinline__ void reconstruct_predict2(const Params<float> &P, const Prim2 &qm, const Prim2 &qc,
                                                     const Prim2 &qp, f2 half_dt, Face2 &lo, Face2 &hi) {
  const f2 h(0.5f);
  const f2 s_rho = limiter2(qc.rho - qm.rho, qp.rho - qc.rho), s_u = limiter2(qc.u - qm.u, qp.u - qc.u);
  const f2 s_v = limiter2(qc.v - qm.v, qp.v - qc.v), s_p = limiter2(qc.p - qm.p, qp.p - qc.p);
  Prim2 qL{fma_(-h, s_rho, qc.rho), fma_(-h, s_u, qc.u), fma_(-h, s_v, qc.v), fma_(-h, s_p, qc.p)};
  Prim2 qR{fma_(h, s_rho, qc.rho), fma_(h, s_u, qc.u), fma_(h, s_v, qc.v), fma_(h, s_p, qc.p)};
  // enforce_positive_faces :373-398 — rare: done per half with the scalar routine
  const f2 er(P.eps_rho), ep(P.eps_p);
  const m2 bad = le0(qL.rho - er) | le0(qR.rho - er) | le0(qL.p - ep) | le0(qR.p - ep);
  if (any(bad)) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k ? bad.y : bad.x) {
        const PrimPair<float> fx = enforce_positive_faces(P.eps_rho, P.eps_p, half(qL, k), half(qc, k), half(qR, k));
        if (k) {
          qL.rho.v.y = fx.m.rho; qL.u.v.y = fx.m.u; qL.v.v.y = fx.m.v; qL.p.v.y = fx.m.p;
          qR.rho.v.y = fx.p.rho; qR.u.v.y = fx.p.u; qR.v.v.y = fx.p.v; qR.p.v.y = fx.p.p;
        } else {
          qL.rho.v.x = fx.m.rho; qL.u.v.x = fx.m.u; qL.v.v.x = fx.m.v; qL.p.v.x = fx.m.p;
          qR.rho.v.x = fx.p.rho; qR.u.v.x = fx.p.u; qR.v.v.x = fx.p.v; qR.p.v.x = fx.p.p;
        }
      }
    }
  }
  const f2 mxL = qL.rho * qL.u, myL = qL.rho * qL.v, mxR = qR.rho * qR.u, myR = qR.rho * qR.v;
  const f2 ig(P.inv_gm1);
  const f2 EL = fma_(qL.p, ig, (h * qL.rho) * fma_(qL.v, qL.v, qL.u * qL.u));
  const f2 ER = fma_(qR.p, ig, (h * qR.rho) * fma_(qR.v, qR.v, qR.u * qR.u));
  const f2 unL = AX == 0 ? qL.u : qL.v, unR = AX == 0 ? qR.u : qR.v;
  const f2 d_rho = (AX == 0 ? mxR : myR) - (AX == 0 ? mxL : myL);
  const f2 d_mx = AX == 0 ? (fma_(mxR, unR, qR.p) - fma_(mxL, unL, qL.p)) : (mxR * unR - mxL * unL);
  const f2 d_my = AX == 1 ? (fma_(myR, unR, qR.p) - fma_(myL, unL, qL.p)) : (myR * unR - myL * unL);
  const f2 d_E = (ER + qR.p) * unR - (EL + qL.p) * unL;
  const f2 nh = -half_dt;
  {
    const f2 rho = max_(fma_(nh, d_rho, qL.rho), er), inv = rcp_(rho);
    const f2 u = fma_(nh, d_mx, mxL) * inv, v = fma_(nh, d_my, myL) * inv;
    const f2 kin = (h * rho) * fma_(v, v, u * u);
    const f2 pr = max_(f2(P.gm1) * max_(fma_(nh, d_E, EL) - kin, ep), ep);
    lo = Face2{rho, u, v, pr, fma_(pr, ig, kin), sqrt_((f2(P.gamma) * pr) * inv)};
  }
  {
    const f2 rho = max_(fma_(nh, d_rho, qR.rho), er), inv = rcp_(rho);
    const f2 u = fma_(nh, d_mx, mxR) * inv, v = fma_(nh, d_my, myR) * inv;
    const f2 kin = (h * rho) * fma_(v, v, u * u);
    const f2 pr = max_(f2(P.gm1) * max_(fma_(nh, d_E, ER) - kin, ep), ep);
    hi = Face2{rho, u, v, pr, fma_(pr, ig, kin), sqrt_((f2(P.gamma) * pr) * inv)};
  }
}

template <int AX> __device__ __forceinline__ Cons2 phys_flux2(const Face2 &f) {
  const f2 un = AX == 0 ? f.u : f.v;
  const f2 m = f.rho * un;
  return Cons2{m, AX == 0 ? fma_(m, f.u, f.p) : m * f.u, AX == 1 ? fma_(m, f.v, f.p) : m * f.v, (f.E + f.p) * un};
}

// hllc_flux<AX> of hypersonic2d.cu for a pair of faces; the guarded HLLE fall-back runs per half
template <int AX>
__device__ __forceinline__ Cons2 hllc_flux2(const Params<float> &P, const Face2 &L, const Face2 &Rr) {
  const f2 unL = AX == 0 ? L.u : L.v, unR = AX == 0 ? Rr.u : Rr.v;
  const f2 SL = min_(unL - L.a, unR - Rr.a), SR = max_(unL + L.a, unR + Rr.a);
  if (__all_sync(0xffffffffu, SL.v.x >= 0.f && SL.v.y >= 0.f)) return phys_flux2<AX>(L);  // :537-540
  const f2 qL = L.rho * (SL - unL), qR = Rr.rho * (SR - unR);
  const f2 num = fma_(qL, unL, Rr.p - L.p) - qR * unR;
  const f2 den = qL - qR;
  const f2 SM = num * rcp_(den);
  const f2 dLS = SL - SM, dRS = SR - SM;
  const m2 left = ge0(SL) | (!le0(SR) & ge0(SM));
  const Face2 K{sel(left, L.rho, Rr.rho), sel(left, L.u, Rr.u), sel(left, L.v, Rr.v),
                sel(left, L.p, Rr.p),     sel(left, L.E, Rr.E), f2(0.f)};
  const f2 SK = sel(left, SL, SR), unK = sel(left, unL, unR), qK = sel(left, qL, qR), dKS = sel(left, dLS, dRS);
  const Cons2 FK = phys_flux2<AX>(K);
  const f2 pStar = max_(fma_(qL, SM - unL, L.p), f2(P.eps_p));
  const f2 invd = rcp_(dKS);
  const f2 rhoStar = qK * invd;
  const f2 EStar = fma_(pStar, SM, fma_(SK - unK, K.E, -(K.p * unK))) * invd;
  const f2 chk = (num + den) + (SM + rhoStar) + EStar, pL = qL * dLS, pR = qR * dRS;
  const f2 aden = abs_(den), aL = abs_(dLS), aR = abs_(dRS);
  const m2 fallback{(aden.v.x < 1e-14f) || (aL.v.x < 1e-14f) || (aR.v.x < 1e-14f) || !(pL.v.x > 0.f) ||
                        !(pR.v.x > 0.f) || !isfinite(chk.v.x),
                    (aden.v.y < 1e-14f) || (aL.v.y < 1e-14f) || (aR.v.y < 1e-14f) || !(pL.v.y > 0.f) ||
                        !(pR.v.y > 0.f) || !isfinite(chk.v.y)};
  const m2 supersonic = ge0(SL) | le0(SR);
  const f2 sn = rhoStar * SM, st = rhoStar * (AX == 0 ? K.v : K.u);
  Cons2 F;
  F.rho = sel(supersonic, FK.rho, fma_(SK, rhoStar - K.rho, FK.rho));
  F.mx = sel(supersonic, FK.mx, fma_(SK, (AX == 0 ? sn : st) - K.rho * K.u, FK.mx));
  F.my = sel(supersonic, FK.my, fma_(SK, (AX == 0 ? st : sn) - K.rho * K.v, FK.my));
  F.E = sel(supersonic, FK.E, fma_(SK, EStar - K.E, FK.E));
  const m2 use_hlle = !supersonic & fallback;
  if (any(use_hlle)) {  // rare
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k ? use_hlle.y : use_hlle.x) {
        const Cons4<float> e = hlle_flux<AX>(half(L, k), half(Rr, k), k ? SL.v.y : SL.v.x, k ? SR.v.y : SR.v.x);
        if (k) { F.rho.v.y = e.rho; F.mx.v.y = e.mx; F.my.v.y = e.my; F.E.v.y = e.E; }
        else { F.rho.v.x = e.rho; F.mx.v.x = e.mx; F.my.v.x = e.my; F.E.v.x = e.E; }
      }
    }
  }
  return F;
}

__device__ __forceinline__ Face2 face_from_cons2(const Params<float> &P, const Cons2 &c) {
  const Prim2 q = cons_to_prim2(P, c);
  return Face2{q.rho, q.u, q.v, q.p, c.E, sqrt_((f2(P.gamma) * q.p) * rcp_(q.rho))};
}

struct Ring2 {
  const float *base;
  static constexpr int FSTRIDE = H2_RB * HP_BOXW;
  __device__ __forceinline__ Cons2 at(int off, int c) const {  // c even: 8-byte aligned pair loads
    const float *p = base + off + c;
    return Cons2{f2(*reinterpret_cast<const float2 *>(p)), f2(*reinterpret_cast<const float2 *>(p + FSTRIDE)),
                 f2(*reinterpret_cast<const float2 *>(p + 2 * FSTRIDE)),
                 f2(*reinterpret_cast<const float2 *>(p + 3 * FSTRIDE))};
  }
  __device__ __forceinline__ Cons4<float> at1(int off, int c) const {
    const float *p = base + off + c;
    return Cons4<float>{p[0], p[FSTRIDE], p[2 * FSTRIDE], p[3 * FSTRIDE]};
  }
};

// Interior, body-free work items only.  `ctrl->next_item` is NOT used: the pair kernel claims from its own
// counter `claim_ctr[step_slot]` (cleared two steps ahead like the others).  No bookkeeping of sim_t here.
__global__ void __launch_bounds__(H2_WARPS * 32, HP_MIN_CTAS)
hyp2d_step_pair(const __grid_constant__ CUtensorMap tmU, const Params<float> P, float *__restrict__ Uout,
                const uint2 *__restrict__ items, int nitems, Ctrl *__restrict__ ctrl,
                unsigned int *__restrict__ claim_ctr, int step_slot, const PeerPush peer) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  int lane, warp;
  {
    unsigned t;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
    lane = (int)(t & 31u);
    warp = (int)(t >> 5);
  }
  float *ring_base = reinterpret_cast<float *>(smem_raw) + (size_t)warp * H2_NS * HP_SLOT;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)H2_WARPS * H2_NS * HP_SLOT * sizeof(float)) +
                   warp * H2_NS;
  Ring2 ring{ring_base};

  asm volatile("griddepcontrol.launch_dependents;");
  if (lane == 0) {
    for (int s = 0; s < H2_NS; ++s) tau::mbar_init(&bars[s], 1);
    tau::mbar_fence_init();
  }
  __syncwarp();
  const unsigned nwarps_grid = gridDim.x * H2_WARPS;
  unsigned item = blockIdx.x * H2_WARPS + warp;
  uint2 desc = make_uint2(0u, 0u);
  if (item < (unsigned)nitems) desc = items[item];
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // multi-GPU: this kernel runs FIRST in the step, so it is the one that must see the peers' messages
  __shared__ unsigned long long s_peer_max;
  if (peer.pc.world > 1) {
    if (warp == 0) {
      unsigned long long v = 0ull;
      if (lane < peer.pc.world && lane != peer.pc.rank) {
        const unsigned long long *a = &ctrl->inbox[step_slot][lane];
        for (;;) {
          asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
          if (v != 0ull) break;
          __nanosleep(40);
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
      }
      if (lane == 0) s_peer_max = v;
    }
    __syncthreads();
  }

  float dt = 0.f;
  f2 half_dt(0.f);
  auto compute_dt = [&]() {
    double maxs = *reinterpret_cast<volatile double *>(&ctrl->maxspeed[step_slot]);
    if (peer.pc.world > 1) maxs = fmax(maxs, __longlong_as_double((long long)s_peer_max));
    if (!isfinite(maxs) || maxs < 1e-12) maxs = 1e-12;
    const double dt_conv = P.cfl * 1.0 / maxs;
    double dt_diff = dt_conv;
    if (isfinite(P.nu_max) && P.nu_max > 1e-12) dt_diff = 0.25 / P.nu_max;
    const double dt_d = fmin(dt_conv, dt_diff);
    dt = (float)dt_d;
    half_dt = f2((float)(0.5 * dt_d));
    if (blockIdx.x == 0 && threadIdx.x == 0) claim_ctr[(step_slot + 2) % 3] = 0u;
  };

  bool pushed = false;
  float wmax = 0.f;
  const int W = P.W;
  const size_t PL = P.plane;
  unsigned kb = 0, claim = 0;
  bool first = true;
  while (item < (unsigned)nitems) {
#define HP_TM tmU
#define HP_NITEMS nitems
#define HP_CLAIM_CTR claim_ctr
#define HP_HALF_DT half_dt
#define HP_RING ring
#define HP_RING_BASE ring_base
#include "hypersonic2d_pair_item.inc"
#undef HP_TM
#undef HP_NITEMS
#undef HP_CLAIM_CTR
#undef HP_HALF_DT
#undef HP_RING
#undef HP_RING_BASE
  }
  if (first) compute_dt();
  wmax = tau::warp_max(wmax);
  if (lane == 0 && wmax > 0.f) tau::atomic_max_nonneg(&ctrl->maxspeed[(step_slot + 1) % 3], (double)wmax);
  if (peer.pc.world > 1 && pushed) __threadfence_system();  // the production kernel (launched next) signals
}

}  // namespace
