// pair_probe.cu — offline feasibility probe (no GPU needed: static SASS counts of straight-line code).
// Question: if one lane of the 2-D hypersonic step kernel carried TWO cells (two adjacent strips) as
// float2, how many issue slots would the arithmetic core — limited reconstruction + Hancock predictor
// + HLLC, ~60 % of the kernel's instructions — need per cell, given that sm_100 has packed
// FADD2/FMUL2/FFMA2 but no packed min/max, compare, select or MUFU?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -cubin pair_probe.cu -o pair_probe.cubin
//   cuobjdump -sass pair_probe.cubin | python ../sass_count.py        (see profiles/hyp2d_pair_probe_r1.md)
// The two kernels evaluate exactly the same expression trees (those of hypersonic2d.cu) on the same
// inputs; k_scalar processes the two cells one after the other, k_packed side by side.
#include <cuda_runtime.h>

namespace {

// ---- a 2-wide value: packed add/mul/fma, per-half everything else ------------------------------------
struct f2 {
  float2 v;
  __device__ __forceinline__ f2() {}
  __device__ __forceinline__ f2(float a, float b) : v(make_float2(a, b)) {}
  __device__ __forceinline__ explicit f2(float a) : v(make_float2(a, a)) {}
};
struct m2 { bool x, y; };
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { f2 r; r.v = __fadd2_rn(a.v, b.v); return r; }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { f2 r; r.v = __fmul2_rn(a.v, b.v); return r; }
__device__ __forceinline__ f2 operator-(f2 a) { return f2(-a.v.x, -a.v.y); }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { f2 r; r.v = __ffma2_rn(b.v, make_float2(-1.f, -1.f), a.v); return r; }
__device__ __forceinline__ f2 fma_(f2 a, f2 b, f2 c) { f2 r; r.v = __ffma2_rn(a.v, b.v, c.v); return r; }
__device__ __forceinline__ f2 max_(f2 a, f2 b) { return f2(fmaxf(a.v.x, b.v.x), fmaxf(a.v.y, b.v.y)); }
__device__ __forceinline__ f2 min_(f2 a, f2 b) { return f2(fminf(a.v.x, b.v.x), fminf(a.v.y, b.v.y)); }
__device__ __forceinline__ f2 abs_(f2 a) { return f2(fabsf(a.v.x), fabsf(a.v.y)); }
__device__ __forceinline__ float rcp1(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt1(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ f2 rcp_(f2 a) { return f2(rcp1(a.v.x), rcp1(a.v.y)); }
__device__ __forceinline__ f2 sqrt_(f2 a) { return f2(sqrt1(a.v.x), sqrt1(a.v.y)); }
__device__ __forceinline__ m2 ge0(f2 a) { return m2{a.v.x >= 0.f, a.v.y >= 0.f}; }
__device__ __forceinline__ m2 le0(f2 a) { return m2{a.v.x <= 0.f, a.v.y <= 0.f}; }
__device__ __forceinline__ m2 lt_(f2 a, float b) { return m2{a.v.x < b, a.v.y < b}; }
__device__ __forceinline__ m2 gt0(f2 a) { return m2{a.v.x > 0.f, a.v.y > 0.f}; }
__device__ __forceinline__ m2 operator|(m2 a, m2 b) { return m2{a.x || b.x, a.y || b.y}; }
__device__ __forceinline__ m2 operator&(m2 a, m2 b) { return m2{a.x && b.x, a.y && b.y}; }
__device__ __forceinline__ m2 operator!(m2 a) { return m2{!a.x, !a.y}; }
__device__ __forceinline__ f2 sel(m2 m, f2 a, f2 b) { return f2(m.x ? a.v.x : b.v.x, m.y ? a.v.y : b.v.y); }

// ---- the same vocabulary for a scalar ---------------------------------------------------------------------
__device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ float abs_(float a) { return fabsf(a); }
__device__ __forceinline__ float rcp_(float a) { return rcp1(a); }
__device__ __forceinline__ float sqrt_(float a) { return sqrt1(a); }
__device__ __forceinline__ bool ge0(float a) { return a >= 0.f; }
__device__ __forceinline__ bool le0(float a) { return a <= 0.f; }
__device__ __forceinline__ bool lt_(float a, float b) { return a < b; }
__device__ __forceinline__ bool gt0(float a) { return a > 0.f; }
__device__ __forceinline__ float sel(bool m, float a, float b) { return m ? a : b; }
template <typename T> __device__ __forceinline__ T K(float c);
template <> __device__ __forceinline__ float K<float>(float c) { return c; }
template <> __device__ __forceinline__ f2 K<f2>(float c) { return f2(c); }

template <typename T> struct Prim { T rho, u, v, p; };
template <typename T> struct Face { T rho, u, v, p, E, a; };
template <typename T> struct Cons { T rho, mx, my, E; };
struct Par { float gamma, gm1, inv_gm1, eps; };

template <typename T> __device__ __forceinline__ T limiter(T dl, T dr) {  // median(dl, dr, 0)
  return max_(min_(dl, dr), min_(max_(dl, dr), K<T>(0.f)));
}

// reconstruct_predict<0> of hypersonic2d.cu (positivity fix-up branch left out: rarely taken)
template <typename T>
__device__ __forceinline__ void reconstruct_predict(const Par &P, const Prim<T> &qm, const Prim<T> &qc,
                                                    const Prim<T> &qp, T half_dt, Face<T> &lo, Face<T> &hi) {
  const T h = K<T>(0.5f);
  const T s_rho = limiter(qc.rho - qm.rho, qp.rho - qc.rho), s_u = limiter(qc.u - qm.u, qp.u - qc.u);
  const T s_v = limiter(qc.v - qm.v, qp.v - qc.v), s_p = limiter(qc.p - qm.p, qp.p - qc.p);
  const Prim<T> qL{fma_(-h, s_rho, qc.rho), fma_(-h, s_u, qc.u), fma_(-h, s_v, qc.v), fma_(-h, s_p, qc.p)};
  const Prim<T> qR{fma_(h, s_rho, qc.rho), fma_(h, s_u, qc.u), fma_(h, s_v, qc.v), fma_(h, s_p, qc.p)};
  const T mxL = qL.rho * qL.u, myL = qL.rho * qL.v, mxR = qR.rho * qR.u, myR = qR.rho * qR.v;
  const T ig = K<T>(P.inv_gm1);
  const T EL = fma_(qL.p, ig, (h * qL.rho) * fma_(qL.v, qL.v, qL.u * qL.u));
  const T ER = fma_(qR.p, ig, (h * qR.rho) * fma_(qR.v, qR.v, qR.u * qR.u));
  const T d_rho = mxR - mxL;
  const T d_mx = fma_(mxR, qR.u, qR.p) - fma_(mxL, qL.u, qL.p);
  const T d_my = myR * qR.u - myL * qL.u;
  const T d_E = (ER + qR.p) * qR.u - (EL + qL.p) * qL.u;
  const T eps = K<T>(P.eps), nh = -half_dt;
  {
    const T rho = max_(fma_(nh, d_rho, qL.rho), eps), inv = rcp_(rho);
    const T u = fma_(nh, d_mx, mxL) * inv, v = fma_(nh, d_my, myL) * inv;
    const T kin = (h * rho) * fma_(v, v, u * u);
    const T pr = max_(K<T>(P.gm1) * max_(fma_(nh, d_E, EL) - kin, eps), eps);
    lo = Face<T>{rho, u, v, pr, fma_(pr, ig, kin), sqrt_((K<T>(P.gamma) * pr) * inv)};
  }
  {
    const T rho = max_(fma_(nh, d_rho, qR.rho), eps), inv = rcp_(rho);
    const T u = fma_(nh, d_mx, mxR) * inv, v = fma_(nh, d_my, myR) * inv;
    const T kin = (h * rho) * fma_(v, v, u * u);
    const T pr = max_(K<T>(P.gm1) * max_(fma_(nh, d_E, ER) - kin, eps), eps);
    hi = Face<T>{rho, u, v, pr, fma_(pr, ig, kin), sqrt_((K<T>(P.gamma) * pr) * inv)};
  }
}

// hllc_flux<0> of hypersonic2d.cu, branch-free (the guarded HLLE fall-back is a rare call there)
template <typename T>
__device__ __forceinline__ Cons<T> hllc(const Par &P, const Face<T> &L, const Face<T> &R) {
  const T SL = min_(L.u - L.a, R.u - R.a), SR = max_(L.u + L.a, R.u + R.a);
  const T qL = L.rho * (SL - L.u), qR = R.rho * (SR - R.u);
  const T num = fma_(qL, L.u, R.p - L.p) - qR * R.u;
  const T SM = num * rcp_(qL - qR);
  const auto left = ge0(SL) | (!le0(SR) & ge0(SM));
  const T Krho = sel(left, L.rho, R.rho), Ku = sel(left, L.u, R.u), Kv = sel(left, L.v, R.v);
  const T Kp = sel(left, L.p, R.p), KE = sel(left, L.E, R.E);
  const T SK = sel(left, SL, SR), qK = sel(left, qL, qR);
  const T m = Krho * Ku;
  const Cons<T> FK{m, fma_(m, Ku, Kp), m * Kv, (KE + Kp) * Ku};
  const T pStar = max_(fma_(qL, SM - L.u, L.p), K<T>(P.eps));
  const T invd = rcp_(SK - SM);
  const T rhoStar = qK * invd;
  const T EStar = fma_(pStar, SM, fma_(SK - Ku, KE, -(Kp * Ku))) * invd;
  const auto supersonic = ge0(SL) | le0(SR);
  const T sn = rhoStar * SM, st = rhoStar * Kv;
  Cons<T> F;
  F.rho = sel(supersonic, FK.rho, fma_(SK, rhoStar - Krho, FK.rho));
  F.mx = sel(supersonic, FK.mx, fma_(SK, sn - Krho * Ku, FK.mx));
  F.my = sel(supersonic, FK.my, fma_(SK, st - Krho * Kv, FK.my));
  F.E = sel(supersonic, FK.E, fma_(SK, EStar - KE, FK.E));
  return F;
}

template <typename T> __device__ __forceinline__ Prim<T> cons_to_prim(const Par &P, const Cons<T> &c) {
  const T rho = max_(c.rho, K<T>(P.eps)), inv = rcp_(rho);
  const T u = c.mx * inv, v = c.my * inv;
  const T eint = fma_(-(K<T>(0.5f) * rho), fma_(v, v, u * u), c.E);
  return Prim<T>{rho, u, v, K<T>(P.gm1) * max_(eint, K<T>(P.eps))};
}

// one "x-sweep + y-sweep" worth of arithmetic for one cell: 3 conversions, 2 reconstructions, 2 Riemann solves
template <typename T>
__device__ __forceinline__ Cons<T> cell_work(const Par &P, const Cons<T> *c, T half_dt) {
  const Prim<T> a = cons_to_prim(P, c[0]), b = cons_to_prim(P, c[1]), d = cons_to_prim(P, c[2]);
  Face<T> lo, hi, lo2, hi2;
  reconstruct_predict(P, a, b, d, half_dt, lo, hi);
  reconstruct_predict(P, d, b, a, half_dt, lo2, hi2);
  const Cons<T> F = hllc(P, hi, lo2), G = hllc(P, hi2, lo);
  return Cons<T>{F.rho + G.rho, F.mx + G.mx, F.my + G.my, F.E + G.E};
}

}  // namespace

// two cells per thread, one after the other
__global__ void k_scalar(const float4 *in, float4 *out, Par P, float half_dt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  Cons<float> r[2];
  for (int s = 0; s < 2; ++s) {
    Cons<float> c[3];
    for (int k = 0; k < 3; ++k) {
      const float4 q = in[(i * 2 + s) * 3 + k];
      c[k] = Cons<float>{q.x, q.y, q.z, q.w};
    }
    r[s] = cell_work<float>(P, c, half_dt);
  }
  out[i * 2] = make_float4(r[0].rho, r[0].mx, r[0].my, r[0].E);
  out[i * 2 + 1] = make_float4(r[1].rho, r[1].mx, r[1].my, r[1].E);
}

// two cells per thread, side by side in float2
__global__ void k_packed(const float4 *in, float4 *out, Par P, float half_dt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  Cons<f2> c[3];
  for (int k = 0; k < 3; ++k) {
    const float4 q0 = in[(i * 2) * 3 + k], q1 = in[(i * 2 + 1) * 3 + k];
    c[k] = Cons<f2>{f2(q0.x, q1.x), f2(q0.y, q1.y), f2(q0.z, q1.z), f2(q0.w, q1.w)};
  }
  const Cons<f2> r = cell_work<f2>(P, c, f2(half_dt));
  out[i * 2] = make_float4(r.rho.v.x, r.mx.v.x, r.my.v.x, r.E.v.x);
  out[i * 2 + 1] = make_float4(r.rho.v.y, r.mx.v.y, r.my.v.y, r.E.v.y);
}
