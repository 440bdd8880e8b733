/* hyp2d_oracle.c — CPU restatement (fp64) of the reference 2-D hypersonic step.
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may call this; the product never links or imports it.
 *
 * Follows tau_hypersonic_cuda.cu pass by pass (line numbers cited at each function): the same
 * kernel sequence as the reference step loop (:1833-1889) executed as plain loops over a runtime
 * W x H grid, fp64, SoA planes rho,mx,my,E indexed y*W+x (:130).
 *
 * Pinning: (1) the reference's own known-answer tests (tau_hypersonic_cuda_tests.cu:245-371,
 * expected values :386-484, :613-631) are replayed against the helpers exported here by
 * tests/test_oracle_hyp2d.py; (2) full-field outputs of the reference kernels run on a B200
 * (tests/golden/hyp2d_*.npz, made by tests/golden/make_golden_gpu.py via oracle/_ref) are compared
 * in the same test file.  The reference's 12-scalar regression snapshot (:143-176) is restated in
 * oracle_hyp2d_snapshot().
 * Also pinned BIT FOR BIT on the reference's kernels executed on the CPU (oracle/_ref/libref_hyp2d_host_*.so,
 * tests/test_oracle_cpu.py::test_hyp2d_oracle_equals_reference_kernels_run_on_the_cpu).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EPS_RHO 1e-25 /* :32 */
#define EPS_P 1e-25   /* :33 */

typedef struct {
  double gamma, cfl, visc_nu, visc_rho, visc_e, inflow_mach;
  double geom_x0, geom_cy, geom_Rb, geom_Rn, geom_theta;
  int W, H;
} oracle_hyp2d_cfg;

typedef struct { double rho, mx, my, E; } Cons;
typedef struct { double rho, u, v, p; } Prim;

typedef struct {
  const oracle_hyp2d_cfg *c;
  const double *rho, *mx, *my, *E;
  const uint8_t *mask;
} Field;

static inline double dmax(double a, double b) { return a > b ? a : b; } /* :131 */
static inline double dmin(double a, double b) { return a < b ? a : b; } /* :134 */
static inline double dabs(double a) { return a < 0 ? -a : a; }          /* :137 */

/* default_config() :1394-1409 with H-derived geometry */
void oracle_hyp2d_default_cfg(oracle_hyp2d_cfg *c, int W, int H) {
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
  c->W = W;
  c->H = H;
}

/* :143-159 */
static Prim cons_to_prim(const oracle_hyp2d_cfg *c, Cons q) {
  Prim p;
  double rho = dmax(q.rho, EPS_RHO);
  double inv = 1.0 / rho;
  double u = q.mx * inv, v = q.my * inv;
  double kin = 0.5 * rho * (u * u + v * v);
  double eint = q.E - kin;
  p.rho = rho;
  p.u = u;
  p.v = v;
  p.p = (c->gamma - 1.0) * dmax(eint, EPS_P);
  return p;
}
/* :161-170 */
static Cons prim_to_cons(const oracle_hyp2d_cfg *c, Prim p) {
  Cons q;
  double rho = dmax(p.rho, EPS_RHO), pr = dmax(p.p, EPS_P);
  q.rho = rho;
  q.mx = rho * p.u;
  q.my = rho * p.v;
  q.E = pr / (c->gamma - 1.0) + 0.5 * rho * (p.u * p.u + p.v * p.v);
  return q;
}
/* :172-174 */
static double sound_speed(const oracle_hyp2d_cfg *c, Prim p) {
  return sqrt(c->gamma * dmax(p.p, EPS_P) / dmax(p.rho, EPS_RHO));
}
/* flux_axis<AX>(Cons) :194-203; ax 0 = x, 1 = y */
static Cons flux_axis(const oracle_hyp2d_cfg *c, int ax, Cons q) {
  Prim p = cons_to_prim(c, q);
  double un = ax == 0 ? p.u : p.v;
  Cons f;
  f.rho = ax == 0 ? q.mx : q.my;
  f.mx = ax == 0 ? (q.mx * un + p.p) : (q.mx * un);
  f.my = ax == 0 ? (q.my * un) : (q.my * un + p.p);
  f.E = (q.E + p.p) * un;
  return f;
}
/* :217-228 */
static double minmod(double a, double b) {
  if (a * b <= 0.0) return 0.0;
  return (dabs(a) < dabs(b)) ? a : b;
}
static double mc_limiter(double dl, double dc, double dr) {
  double mm1 = minmod(dl, dr), mm2 = minmod(dc, 2.0 * dl), mm3 = minmod(dc, 2.0 * dr);
  return minmod(mm1, minmod(mm2, mm3));
}
/* :230-238 */
static Prim inflow_state(const oracle_hyp2d_cfg *c) {
  Prim p;
  p.rho = 1.0;
  p.p = 1.0;
  p.u = c->inflow_mach * sqrt(c->gamma * 1.0 / 1.0);
  p.v = 0.0;
  return p;
}
static Prim wall_ghost(Prim in) { /* :262-264 */
  Prim g = {in.rho, -in.u, -in.v, in.p};
  return g;
}
static Cons load_cons(const Field *f, int i) {
  Cons q = {f->rho[i], f->mx[i], f->my[i], f->E[i]};
  return q;
}

/* neighbor_or_wall :266-290, neighbor_for_diff :292-313 and load_neighbor_or_wall_tiled :349-371
 * are the same rule: clamp y; x<0 -> inflow; x>=W -> raw column W-1; masked -> no-slip ghost of the
 * CENTRE cell's primitive state; else the cell itself. */
static Cons neighbor_rule(const Field *f, Prim centre, int xn, int yn) {
  const int W = f->c->W, H = f->c->H;
  if (yn < 0) yn = 0;
  if (yn >= H) yn = H - 1;
  if (xn < 0) return prim_to_cons(f->c, inflow_state(f->c));
  if (xn >= W) return load_cons(f, yn * W + (W - 1));
  int j = yn * W + xn;
  if (f->mask[j]) return prim_to_cons(f->c, wall_ghost(centre));
  return load_cons(f, j);
}
static Cons neighbor_or_wall(const Field *f, int x, int y, int dx, int dy) {
  Prim centre = cons_to_prim(f->c, load_cons(f, y * f->c->W + x));
  return neighbor_rule(f, centre, x + dx, y + dy);
}

/* :373-398 */
static void enforce_positive_faces(Prim *qm, const Prim *qc, Prim *qp) {
  for (int it = 0; it < 8; it++) {
    int bad = 0;
    if (qm->rho <= EPS_RHO || qp->rho <= EPS_RHO) bad = 1;
    if (qm->p <= EPS_P || qp->p <= EPS_P) bad = 1;
    if (!bad) return;
    qm->rho = 0.5 * (qm->rho + qc->rho);
    qm->u = 0.5 * (qm->u + qc->u);
    qm->v = 0.5 * (qm->v + qc->v);
    qm->p = 0.5 * (qm->p + qc->p);
    qp->rho = 0.5 * (qp->rho + qc->rho);
    qp->u = 0.5 * (qp->u + qc->u);
    qp->v = 0.5 * (qp->v + qc->v);
    qp->p = 0.5 * (qp->p + qc->p);
  }
  qm->rho = dmax(qm->rho, EPS_RHO);
  qp->rho = dmax(qp->rho, EPS_RHO);
  qm->p = dmax(qm->p, EPS_P);
  qp->p = dmax(qp->p, EPS_P);
}

/* reconstruct_limited_faces :400-425 */
static void reconstruct_faces(Prim qm, Prim qc, Prim qp, Prim *qL, Prim *qR) {
  double s_rho = mc_limiter(qc.rho - qm.rho, 0.5 * (qp.rho - qm.rho), qp.rho - qc.rho);
  double s_u = mc_limiter(qc.u - qm.u, 0.5 * (qp.u - qm.u), qp.u - qc.u);
  double s_v = mc_limiter(qc.v - qm.v, 0.5 * (qp.v - qm.v), qp.v - qc.v);
  double s_p = mc_limiter(qc.p - qm.p, 0.5 * (qp.p - qm.p), qp.p - qc.p);
  qL->rho = qc.rho - 0.5 * s_rho;
  qL->u = qc.u - 0.5 * s_u;
  qL->v = qc.v - 0.5 * s_v;
  qL->p = qc.p - 0.5 * s_p;
  qR->rho = qc.rho + 0.5 * s_rho;
  qR->u = qc.u + 0.5 * s_u;
  qR->v = qc.v + 0.5 * s_v;
  qR->p = qc.p + 0.5 * s_p;
  enforce_positive_faces(qL, &qc, qR);
}

/* half_step_predict_axis :442-455 */
static Prim half_step_predict(const oracle_hyp2d_cfg *c, Prim q, Cons dF, double half_dt) {
  Cons w = prim_to_cons(c, q);
  w.rho -= half_dt * dF.rho;
  w.mx -= half_dt * dF.mx;
  w.my -= half_dt * dF.my;
  w.E -= half_dt * dF.E;
  Prim o = cons_to_prim(c, w);
  o.rho = dmax(o.rho, EPS_RHO);
  o.p = dmax(o.p, EPS_P);
  return o;
}

static Cons c_sub(Cons a, Cons b) { Cons r = {a.rho - b.rho, a.mx - b.mx, a.my - b.my, a.E - b.E}; return r; }
static Cons c_add(Cons a, Cons b) { Cons r = {a.rho + b.rho, a.mx + b.mx, a.my + b.my, a.E + b.E}; return r; }
static Cons c_mul(double s, Cons a) { Cons r = {s * a.rho, s * a.mx, s * a.my, s * a.E}; return r; }

/* hlle_axis :483-509 */
static Cons hlle_axis(const oracle_hyp2d_cfg *c, int ax, Cons UL, Cons UR) {
  Prim L = cons_to_prim(c, UL), R = cons_to_prim(c, UR);
  double uL = ax == 0 ? L.u : L.v, uR = ax == 0 ? R.u : R.v;
  double aL = sound_speed(c, L), aR = sound_speed(c, R);
  double SL = dmin(uL - aL, uR - aR), SR = dmax(uL + aL, uR + aR);
  Cons FL = flux_axis(c, ax, UL), FR = flux_axis(c, ax, UR);
  if (SL >= 0.0) return FL;
  if (SR <= 0.0) return FR;
  double denom = SR - SL;
  if (dabs(denom) < 1e-14) return c_mul(0.5, c_add(FL, FR));
  Cons t1 = c_mul(SR, FL), t2 = c_mul(-SL, FR), t3 = c_mul(SL * SR, c_sub(UR, UL));
  return c_mul(1.0 / denom, c_add(c_add(t1, t2), t3));
}

/* hllc_axis :519-606 */
static Cons hllc_axis(const oracle_hyp2d_cfg *c, int ax, Cons UL, Cons UR) {
  Prim L = cons_to_prim(c, UL), R = cons_to_prim(c, UR);
  double unL = ax == 0 ? L.u : L.v, unR = ax == 0 ? R.u : R.v;
  double utL = ax == 0 ? L.v : L.u, utR = ax == 0 ? R.v : R.u;
  double aL = sound_speed(c, L), aR = sound_speed(c, R);
  double SL = dmin(unL - aL, unR - aR), SR = dmax(unL + aL, unR + aR);
  Cons FL = flux_axis(c, ax, UL), FR = flux_axis(c, ax, UR);
  if (SL >= 0.0) return FL;
  if (SR <= 0.0) return FR;
  double rhoL = L.rho, rhoR = R.rho, pL = L.p, pR = R.p;
  double num = pR - pL + rhoL * unL * (SL - unL) - rhoR * unR * (SR - unR);
  double den = rhoL * (SL - unL) - rhoR * (SR - unR);
  if (dabs(den) < 1e-14 || !isfinite(num) || !isfinite(den)) return hlle_axis(c, ax, UL, UR);
  double SM = num / den;
  if (!isfinite(SM)) return hlle_axis(c, ax, UL, UR);
  double pStar = pL + rhoL * (SL - unL) * (SM - unL);
  pStar = dmax(pStar, EPS_P);
  double dLS = SL - SM, dRS = SR - SM;
  if (dabs(dLS) < 1e-14 || dabs(dRS) < 1e-14) return hlle_axis(c, ax, UL, UR);
  double rhoStarL = rhoL * (SL - unL) / dLS, rhoStarR = rhoR * (SR - unR) / dRS;
  if (!(rhoStarL > 0.0) || !(rhoStarR > 0.0) || !isfinite(rhoStarL) || !isfinite(rhoStarR))
    return hlle_axis(c, ax, UL, UR);
  double EStarL = ((SL - unL) * UL.E - pL * unL + pStar * SM) / dLS;
  if (!isfinite(EStarL)) return hlle_axis(c, ax, UL, UR);
  double EStarR = ((SR - unR) * UR.E - pR * unR + pStar * SM) / dRS;
  if (!isfinite(EStarR)) return hlle_axis(c, ax, UL, UR);
  Cons SLs, SRs;
  SLs.rho = rhoStarL;
  SLs.mx = ax == 0 ? rhoStarL * SM : rhoStarL * utL;
  SLs.my = ax == 0 ? rhoStarL * utL : rhoStarL * SM;
  SLs.E = EStarL;
  SRs.rho = rhoStarR;
  SRs.mx = ax == 0 ? rhoStarR * SM : rhoStarR * utR;
  SRs.my = ax == 0 ? rhoStarR * utR : rhoStarR * SM;
  SRs.E = EStarR;
  Cons F;
  if (SM >= 0.0) {
    F.rho = FL.rho + SL * (SLs.rho - UL.rho);
    F.mx = FL.mx + SL * (SLs.mx - UL.mx);
    F.my = FL.my + SL * (SLs.my - UL.my);
    F.E = FL.E + SL * (SLs.E - UL.E);
    return F;
  }
  F.rho = FR.rho + SR * (SRs.rho - UR.rho);
  F.mx = FR.mx + SR * (SRs.mx - UR.mx);
  F.my = FR.my + SR * (SRs.my - UR.my);
  F.E = FR.E + SR * (SRs.E - UR.E);
  return F;
}

/* geometry :625-686, :729-737 */
static double clamp01(double t) { return t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t); }
static double len2(double x, double y) { return sqrt(x * x + y * y); }
static double sd_segment(double px, double py, double ax, double ay, double bx, double by) {
  double abx = bx - ax, aby = by - ay, apx = px - ax, apy = py - ay;
  double denom = abx * abx + aby * aby + 1e-30;
  double t = clamp01((apx * abx + apy * aby) / denom);
  double qx = ax + t * abx, qy = ay + t * aby;
  return len2(px - qx, py - qy);
}
double oracle_hyp2d_sdf(double x, double y, double Rb, double Rn, double theta) {
  double r = dabs(y);
  double st = sin(theta), ct = cos(theta), tt = tan(theta);
  double xt = Rn * (1.0 - st), rt = Rn * ct;
  double xb = xt + (Rb - rt) / dmax(tt, 1e-30);
  double rprof;
  if (x < 0.0) rprof = -1.0;
  else if (x <= xt) {
    double dx = x - Rn, inside = Rn * Rn - dx * dx;
    rprof = inside > 0.0 ? sqrt(inside) : 0.0;
  } else if (x <= xb) rprof = rt + (x - xt) * tt;
  else rprof = -1.0;
  int inside = (x >= 0.0 && x <= xb && r <= rprof);
  double d = dabs(len2(x - Rn, r) - Rn);
  double d_cone = sd_segment(x, r, xt, rt, xb, Rb);
  double d_base = sd_segment(x, y, xb, -Rb, xb, +Rb);
  double d_rim = len2(x - xb, r - Rb);
  if (d_cone < d) d = d_cone;
  if (d_base < d) d = d_base;
  if (d_rim < d) d = d_rim;
  return inside ? -d : d;
}
static double spherecone_xb(double Rb, double Rn, double theta) {
  double st = sin(theta), ct = cos(theta), tt = tan(theta);
  double xt = Rn * (1.0 - st), rt = Rn * ct;
  return xt + (Rb - rt) / dmax(tt, 1e-30);
}

/* k_init :740-770 */
void oracle_hyp2d_init(const oracle_hyp2d_cfg *c, double *rho, double *mx, double *my, double *E,
                       uint8_t *mask) {
  const int W = c->W, H = c->H;
  const double xb = spherecone_xb(c->geom_Rb, c->geom_Rn, c->geom_theta);
  Prim infl = inflow_state(c);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int i = y * W + x;
      double X = (double)x - c->geom_x0, Y = (double)y - c->geom_cy;
      double sd = oracle_hyp2d_sdf(X, Y, c->geom_Rb, c->geom_Rn, c->geom_theta) - c->geom_Rb;
      sd = dmax(sd, X - xb);
      uint8_t m = (sd < 0.0) ? 1 : 0;
      mask[i] = m;
      Prim s = infl;
      if (m) { s.u = 0.0; s.v = 0.0; }
      Cons q = prim_to_cons(c, s);
      rho[i] = q.rho; mx[i] = q.mx; my[i] = q.my; E[i] = q.E;
    }
}

/* k_apply_inflow_left :772-784 */
static void apply_inflow_left(const oracle_hyp2d_cfg *c, double *rho, double *mx, double *my,
                              double *E, const uint8_t *mask) {
  Cons q = prim_to_cons(c, inflow_state(c));
  for (int y = 0; y < c->H; ++y) {
    int i = y * c->W;
    if (mask[i]) continue;
    rho[i] = q.rho; mx[i] = q.mx; my[i] = q.my; E[i] = q.E;
  }
}

/* k_max_wavespeed_blocks + k_reduce_block_max :786-847 (max is order independent) */
double oracle_hyp2d_max_wavespeed(const oracle_hyp2d_cfg *c, const double *rho, const double *mx,
                                  const double *my, const double *E, const uint8_t *mask) {
  double vmax = 1e-12;
  const int N = c->W * c->H;
  for (int i = 0; i < N; ++i) {
    if (mask[i]) continue;
    Cons q = {rho[i], mx[i], my[i], E[i]};
    Prim p = cons_to_prim(c, q);
    double a = sound_speed(c, p);
    double sx = dabs(p.u) + a, sy = dabs(p.v) + a;
    double v = sx > sy ? sx : sy;
    if (!isfinite(v)) v = 1e-12;
    if (v > vmax) vmax = v;
  }
  return vmax;
}

/* host dt rule :1852-1869 */
double oracle_hyp2d_dt(const oracle_hyp2d_cfg *c, double maxs) {
  if (!isfinite(maxs) || maxs < 1e-12) maxs = 1e-12;
  double dt_conv = c->cfl * 1.0 / maxs;
  double nu_max = fmax(c->visc_nu, fmax(c->visc_rho, c->visc_e));
  double dt_diff = dt_conv;
  if (isfinite(nu_max) && nu_max > 1e-12) dt_diff = 0.25 / nu_max;
  return fmin(dt_conv, dt_diff);
}

typedef struct { double *rho, *mx, *my, *E; } Planes;
static void store(Planes *p, int i, Cons q) { p->rho[i] = q.rho; p->mx[i] = q.mx; p->my[i] = q.my; p->E[i] = q.E; }
static Cons loadp(const Planes *p, int i) { Cons q = {p->rho[i], p->mx[i], p->my[i], p->E[i]}; return q; }

/* One full step of the reference loop body :1833-1889 on planes U (updated in place).
 * Returns dt.  Scratch planes are allocated per call (oracle: clarity over speed). */
double oracle_hyp2d_step(const oracle_hyp2d_cfg *c, double *rho, double *mx, double *my, double *E,
                         const uint8_t *mask) {
  const int W = c->W, H = c->H, N = W * H;
  apply_inflow_left(c, rho, mx, my, E, mask);
  const double maxs = oracle_hyp2d_max_wavespeed(c, rho, mx, my, E, mask);
  const double dt = oracle_hyp2d_dt(c, maxs);
  const double half_dt = 0.5 * dt;
  Field f = {c, rho, mx, my, E, mask};

  double *buf = (double *)malloc(sizeof(double) * ((size_t)N * 20 + (size_t)(W + 1) * H * 4 +
                                                   (size_t)W * (H + 1) * 4));
  Planes xL, xR, yL, yR, xF, yF, out;
  double *b = buf;
#define TAKE(P, n) P.rho = b; b += (n); P.mx = b; b += (n); P.my = b; b += (n); P.E = b; b += (n)
  TAKE(xL, N); TAKE(xR, N); TAKE(yL, N); TAKE(yR, N); TAKE(out, N);
  TAKE(xF, (size_t)(W + 1) * H); TAKE(yF, (size_t)W * (H + 1));
#undef TAKE

  /* k_predict_face_states :849-962 */
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int i = y * W + x;
      Cons Uc = load_cons(&f, i);
      if (mask[i]) { store(&xL, i, Uc); store(&xR, i, Uc); store(&yL, i, Uc); store(&yR, i, Uc); continue; }
      Prim qc = cons_to_prim(c, Uc);
      for (int ax = 0; ax < 2; ++ax) {
        int dx = ax == 0, dy = ax == 1;
        Prim qm = cons_to_prim(c, neighbor_rule(&f, qc, x - dx, y - dy));
        Prim qp = cons_to_prim(c, neighbor_rule(&f, qc, x + dx, y + dy));
        Prim fL, fR;
        reconstruct_faces(qm, qc, qp, &fL, &fR);
        Cons FL = flux_axis(c, ax, prim_to_cons(c, fL));
        Cons FR = flux_axis(c, ax, prim_to_cons(c, fR));
        Cons dF = c_sub(FR, FL);
        Prim pL = half_step_predict(c, fL, dF, half_dt);
        Prim pR = half_step_predict(c, fR, dF, half_dt);
        pL.rho = dmax(pL.rho, EPS_RHO); pL.p = dmax(pL.p, EPS_P);
        pR.rho = dmax(pR.rho, EPS_RHO); pR.p = dmax(pR.p, EPS_P);
        if (ax == 0) { store(&xL, i, prim_to_cons(c, pL)); store(&xR, i, prim_to_cons(c, pR)); }
        else { store(&yL, i, prim_to_cons(c, pL)); store(&yR, i, prim_to_cons(c, pR)); }
      }
    }

  /* k_compute_xface_flux :964-996 */
  for (int y = 0; y < H; ++y)
    for (int fx = 0; fx <= W; ++fx) {
      int i = y * (W + 1) + fx, xl = fx - 1, xr = fx;
      int hasL = (xl >= 0) && !mask[y * W + xl];
      int hasR = (xr < W) && !mask[y * W + xr];
      Cons UL, UR, z = {0, 0, 0, 0};
      if (hasL && hasR) { UL = loadp(&xR, y * W + xl); UR = loadp(&xL, y * W + xr); }
      else if (hasR) { UL = neighbor_or_wall(&f, xr, y, -1, 0); UR = loadp(&xL, y * W + xr); }
      else if (hasL) { UL = loadp(&xR, y * W + xl); UR = neighbor_or_wall(&f, xl, y, +1, 0); }
      else { store(&xF, i, z); continue; }
      store(&xF, i, hllc_axis(c, 0, UL, UR));
    }
  /* k_compute_yface_flux :998-1030 */
  for (int fy = 0; fy <= H; ++fy)
    for (int x = 0; x < W; ++x) {
      int i = fy * W + x, yb = fy - 1, yt = fy;
      int hasB = (yb >= 0) && !mask[yb * W + x];
      int hasT = (yt < H) && !mask[yt * W + x];
      Cons UB, UT, z = {0, 0, 0, 0};
      if (hasB && hasT) { UB = loadp(&yR, yb * W + x); UT = loadp(&yL, yt * W + x); }
      else if (hasT) { UB = neighbor_or_wall(&f, x, yt, 0, -1); UT = loadp(&yL, yt * W + x); }
      else if (hasB) { UB = loadp(&yR, yb * W + x); UT = neighbor_or_wall(&f, x, yb, 0, +1); }
      else { store(&yF, i, z); continue; }
      store(&yF, i, hllc_axis(c, 1, UB, UT));
    }

  /* k_step :1032-1176 */
  const double inv12 = 1.0 / 12.0;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int i = y * W + x;
      Cons Uc = load_cons(&f, i);
      if (mask[i]) { store(&out, i, Uc); continue; }
      Cons FxL = loadp(&xF, y * (W + 1) + x), FxR = loadp(&xF, y * (W + 1) + x + 1);
      Cons GyB = loadp(&yF, y * W + x), GyT = loadp(&yF, (y + 1) * W + x);
      Prim centre = cons_to_prim(c, Uc);
      Cons Un = Uc;
      Un.rho -= dt * (FxR.rho - FxL.rho);
      Un.mx -= dt * (FxR.mx - FxL.mx);
      Un.my -= dt * (FxR.my - FxL.my);
      Un.E -= dt * (FxR.E - FxL.E);
      Un.rho -= dt * (GyT.rho - GyB.rho);
      Un.mx -= dt * (GyT.mx - GyB.mx);
      Un.my -= dt * (GyT.my - GyB.my);
      Un.E -= dt * (GyT.E - GyB.E);
      Cons xm2 = neighbor_rule(&f, centre, x - 2, y), xm1 = neighbor_rule(&f, centre, x - 1, y);
      Cons xp1 = neighbor_rule(&f, centre, x + 1, y), xp2 = neighbor_rule(&f, centre, x + 2, y);
      Cons ym2 = neighbor_rule(&f, centre, x, y - 2), ym1 = neighbor_rule(&f, centre, x, y - 1);
      Cons yp1 = neighbor_rule(&f, centre, x, y + 1), yp2 = neighbor_rule(&f, centre, x, y + 2);
#define D2(m2, m1, cc, p1, p2) ((-(m2) + 16.0 * (m1) - 30.0 * (cc) + 16.0 * (p1) - (p2)) * inv12)
      double lap_rho = D2(xm2.rho, xm1.rho, Uc.rho, xp1.rho, xp2.rho) + D2(ym2.rho, ym1.rho, Uc.rho, yp1.rho, yp2.rho);
      double lap_mx = D2(xm2.mx, xm1.mx, Uc.mx, xp1.mx, xp2.mx) + D2(ym2.mx, ym1.mx, Uc.mx, yp1.mx, yp2.mx);
      double lap_my = D2(xm2.my, xm1.my, Uc.my, xp1.my, xp2.my) + D2(ym2.my, ym1.my, Uc.my, yp1.my, yp2.my);
      double lap_E = D2(xm2.E, xm1.E, Uc.E, xp1.E, xp2.E) + D2(ym2.E, ym1.E, Uc.E, yp1.E, yp2.E);
#undef D2
      Un.rho += (c->visc_rho * dt) * lap_rho;
      Un.mx += (c->visc_nu * dt) * lap_mx;
      Un.my += (c->visc_nu * dt) * lap_my;
      Un.E += (c->visc_e * dt) * lap_E;
      Un.rho = dmax(Un.rho, EPS_RHO);
      Prim pp = cons_to_prim(c, Un);
      if (pp.p <= EPS_P || !isfinite(pp.p) || !isfinite(pp.rho) || !isfinite(pp.u) || !isfinite(pp.v)) {
        pp.rho = dmax(pp.rho, EPS_RHO);
        pp.p = dmax(pp.p, EPS_P);
        Un = prim_to_cons(c, pp);
      }
      store(&out, i, Un);
    }

  memcpy(rho, out.rho, sizeof(double) * N);
  memcpy(mx, out.mx, sizeof(double) * N);
  memcpy(my, out.my, sizeof(double) * N);
  memcpy(E, out.E, sizeof(double) * N);
  free(buf);
  return dt;
}

/* nsteps steps; returns accumulated sim_t (:1888); dts (optional) receives each step's dt */
double oracle_hyp2d_run(const oracle_hyp2d_cfg *c, double *rho, double *mx, double *my, double *E,
                        const uint8_t *mask, int nsteps, double *dts) {
  double t = 0.0;
  for (int s = 0; s < nsteps; ++s) {
    double dt = oracle_hyp2d_step(c, rho, mx, my, E, mask);
    if (dts) dts[s] = dt;
    t += dt;
  }
  return t;
}

/* compute_snapshot tau_hypersonic_cuda_tests.cu:143-176 -> out[12] in the file's line order */
void oracle_hyp2d_snapshot(const oracle_hyp2d_cfg *c, int steps, const double *rho, const double *mx,
                           const double *my, const double *E, const uint8_t *mask, double out[12]) {
  double s_cells = 0, sum_rho = 0, sum_mx = 0, sum_my = 0, sum_E = 0, min_rho = 1e300, min_p = 1e300;
  double max_mach = 0, ck_rho = 0, ck_mx = 0, ck_E = 0;
  const int N = c->W * c->H;
  for (int i = 0; i < N; ++i) {
    if (mask[i]) continue;
    Cons q = {rho[i], mx[i], my[i], E[i]};
    Prim p = cons_to_prim(c, q);
    double a = sqrt(c->gamma * fmax(p.p, EPS_P) / fmax(p.rho, EPS_RHO));
    double mach = sqrt(p.u * p.u + p.v * p.v) / fmax(a, 1e-30);
    double w = (double)((i % 8191) + 1);
    s_cells += 1;
    sum_rho += p.rho; sum_mx += q.mx; sum_my += q.my; sum_E += q.E;
    min_rho = fmin(min_rho, p.rho); min_p = fmin(min_p, p.p); max_mach = fmax(max_mach, mach);
    ck_rho += w * p.rho; ck_mx += w * q.mx; ck_E += w * q.E;
  }
  out[0] = steps; out[1] = s_cells; out[2] = sum_rho; out[3] = sum_mx; out[4] = sum_my; out[5] = sum_E;
  out[6] = min_rho; out[7] = min_p; out[8] = max_mach; out[9] = ck_rho; out[10] = ck_mx; out[11] = ck_E;
}

/* ---- render pass: k_render_vals :1178-1248, the min/max reduction :1273-1320, k_compute_inv_range
 * :1322-1326, k_render_pixels :1250-1271, get_color :692-704, sample_prim_bc :706-727 ------------- */
static Prim sample_prim_bc(const Field *f, int xc, int yc, int x, int y) {
  const oracle_hyp2d_cfg *c = f->c;
  if (y < 0) y = 0;
  if (y >= c->H) y = c->H - 1;
  if (x < 0) return inflow_state(c);
  if (x >= c->W) return cons_to_prim(c, load_cons(f, y * c->W + (c->W - 1)));
  int i = y * c->W + x;
  if (f->mask[i]) return wall_ghost(cons_to_prim(c, load_cons(f, yc * c->W + xc)));
  return cons_to_prim(c, load_cons(f, i));
}

double oracle_hyp2d_render_value(const oracle_hyp2d_cfg *c, const double *rho, const double *mx,
                                 const double *my, const double *E, const uint8_t *mask, int view_mode,
                                 int x, int y) {
  Field f = {c, rho, mx, my, E, mask};
  Prim p = cons_to_prim(c, load_cons(&f, y * c->W + x));
  double v;
  if (view_mode == 0) v = log(p.rho);
  else if (view_mode == 1) v = log(p.p);
  else if (view_mode == 2) v = sqrt(p.u * p.u + p.v * p.v);
  else if (view_mode == 3) {
    double rhoL = sample_prim_bc(&f, x, y, x - 1, y).rho, rhoR = sample_prim_bc(&f, x, y, x + 1, y).rho;
    double rhoB = sample_prim_bc(&f, x, y, x, y - 1).rho, rhoT = sample_prim_bc(&f, x, y, x, y + 1).rho;
    double gx = 0.5 * (rhoR - rhoL), gy = 0.5 * (rhoT - rhoB);
    v = log(1e-12 + sqrt(gx * gx + gy * gy));
  } else if (view_mode == 4) {
    Prim pL = sample_prim_bc(&f, x, y, x - 1, y), pR = sample_prim_bc(&f, x, y, x + 1, y);
    Prim pB = sample_prim_bc(&f, x, y, x, y - 1), pT = sample_prim_bc(&f, x, y, x, y + 1);
    double dv_dx = 0.5 * (pR.v - pL.v), du_dy = 0.5 * (pT.u - pB.u);
    v = asinh(dv_dx - du_dy);
  } else if (view_mode == 5) {
    double a = sound_speed(c, p), sp = sqrt(p.u * p.u + p.v * p.v);
    v = sp / dmax(a, 1e-30);
  } else {
    v = log(dmax(p.p / dmax(p.rho, EPS_RHO), 1e-30));
  }
  if (!isfinite(v)) v = 0.0;
  return v;
}

/* rgba: W*H pixels, bytes R,G,B,A.  vals (optional, W*H doubles) receives tmpVal; minmax[2]. */
void oracle_hyp2d_render(const oracle_hyp2d_cfg *c, const double *rho, const double *mx, const double *my,
                         const double *E, const uint8_t *mask, int view_mode, uint8_t *rgba,
                         double *vals, double *minmax) {
  const int N = c->W * c->H;
  double mn = 1e300, mxv = -1e300;
  double *tmp = vals ? vals : (double *)malloc((size_t)N * sizeof(double));
  for (int i = 0; i < N; i++) {
    if (mask[i]) { tmp[i] = 0.0; continue; }
    double v = oracle_hyp2d_render_value(c, rho, mx, my, E, mask, view_mode, i % c->W, i / c->W);
    tmp[i] = v;
    if (v < mn) mn = v;
    if (v > mxv) mxv = v;
  }
  const double inv_range = 1.0 / dmax(mxv - mn, 1e-30);
  for (int i = 0; i < N; i++) {
    uint8_t *px = rgba + 4 * (size_t)i;
    if (mask[i]) { px[0] = px[1] = px[2] = 110; px[3] = 255; continue; }
    double t = (tmp[i] - mn) * inv_range;
    if (t < 0) t = 0;
    if (t > 1) t = 1;
    double rr = 255.0 * dmin(1.0, dmax(0.0, 3.0 * t - 1.0));
    double gg = 255.0 * dmin(1.0, dmax(0.0, 2.0 - 4.0 * dabs(t - 0.5)));
    double bb = 255.0 * dmin(1.0, dmax(0.0, 2.0 - 3.0 * t));
    px[0] = (uint8_t)rr; px[1] = (uint8_t)gg; px[2] = (uint8_t)bb; px[3] = 255;
  }
  if (minmax) { minmax[0] = mn; minmax[1] = mxv; }
  if (!vals) free(tmp);
}


/* ---- helper exports for the known-answer tests (tau_hypersonic_cuda_tests.cu:245-371) ---------- */
void oracle_hyp2d_kat_cons_to_prim(const oracle_hyp2d_cfg *c, const double q[4], double p[4]) {
  Cons a = {q[0], q[1], q[2], q[3]};
  Prim r = cons_to_prim(c, a);
  p[0] = r.rho; p[1] = r.u; p[2] = r.v; p[3] = r.p;
}
void oracle_hyp2d_kat_prim_to_cons(const oracle_hyp2d_cfg *c, const double p[4], double q[4]) {
  Prim a = {p[0], p[1], p[2], p[3]};
  Cons r = prim_to_cons(c, a);
  q[0] = r.rho; q[1] = r.mx; q[2] = r.my; q[3] = r.E;
}
double oracle_hyp2d_kat_minmod(double a, double b) { return minmod(a, b); }
double oracle_hyp2d_kat_mc(double dl, double dc, double dr) { return mc_limiter(dl, dc, dr); }
void oracle_hyp2d_kat_flux(const oracle_hyp2d_cfg *c, int ax, const double q[4], double f[4]) {
  Cons a = {q[0], q[1], q[2], q[3]};
  Cons r = flux_axis(c, ax, a);
  f[0] = r.rho; f[1] = r.mx; f[2] = r.my; f[3] = r.E;
}
double oracle_hyp2d_kat_sound(const oracle_hyp2d_cfg *c, const double p[4]) {
  Prim a = {p[0], p[1], p[2], p[3]};
  return sound_speed(c, a);
}
void oracle_hyp2d_kat_inflow(const oracle_hyp2d_cfg *c, double p[4]) {
  Prim r = inflow_state(c);
  p[0] = r.rho; p[1] = r.u; p[2] = r.v; p[3] = r.p;
}
void oracle_hyp2d_kat_hllc(const oracle_hyp2d_cfg *c, int ax, const double ul[4], const double ur[4],
                           double f[4]) {
  Cons a = {ul[0], ul[1], ul[2], ul[3]}, b = {ur[0], ur[1], ur[2], ur[3]};
  Cons r = hllc_axis(c, ax, a, b);
  f[0] = r.rho; f[1] = r.mx; f[2] = r.my; f[3] = r.E;
}
void oracle_hyp2d_kat_hlle(const oracle_hyp2d_cfg *c, int ax, const double ul[4], const double ur[4],
                           double f[4]) {
  Cons a = {ul[0], ul[1], ul[2], ul[3]}, b = {ur[0], ur[1], ur[2], ur[3]};
  Cons r = hlle_axis(c, ax, a, b);
  f[0] = r.rho; f[1] = r.mx; f[2] = r.my; f[3] = r.E;
}
void oracle_hyp2d_kat_enforce_positive(double qm[4], const double qc[4], double qp[4]) {
  Prim m = {qm[0], qm[1], qm[2], qm[3]}, cc = {qc[0], qc[1], qc[2], qc[3]}, p = {qp[0], qp[1], qp[2], qp[3]};
  enforce_positive_faces(&m, &cc, &p);
  qm[0] = m.rho; qm[1] = m.u; qm[2] = m.v; qm[3] = m.p;
  qp[0] = p.rho; qp[1] = p.u; qp[2] = p.v; qp[3] = p.p;
}
void oracle_hyp2d_kat_neighbor(const oracle_hyp2d_cfg *c, const double *rho, const double *mx,
                               const double *my, const double *E, const uint8_t *mask, int x, int y,
                               int dx, int dy, double out[4]) {
  Field f = {c, rho, mx, my, E, mask};
  Cons r = neighbor_or_wall(&f, x, y, dx, dy);
  out[0] = r.rho; out[1] = r.mx; out[2] = r.my; out[3] = r.E;
}
/* neighbor_for_diff :292-313 takes absolute neighbour coordinates */
void oracle_hyp2d_kat_neighbor_for_diff(const oracle_hyp2d_cfg *c, const double *rho, const double *mx,
                                        const double *my, const double *E, const uint8_t *mask, int xc,
                                        int yc, int xn, int yn, double out[4]) {
  Field f = {c, rho, mx, my, E, mask};
  Prim centre = cons_to_prim(c, load_cons(&f, yc * c->W + xc));
  Cons r = neighbor_rule(&f, centre, xn, yn);
  out[0] = r.rho; out[1] = r.mx; out[2] = r.my; out[3] = r.E;
}
