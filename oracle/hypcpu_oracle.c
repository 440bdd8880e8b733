/* hypcpu_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into, imported or executed by the product).
 *
 * Plain-C restatement of the reference's CPU solver tau_hypersonic.c (BASELINE config 1: 256 x 256,
 * "speed mode" = view_mode 2) with run-time grid extents and SoA planes (rho, mx, my, E; index y*W+x).
 * gamma 1.4, CFL 0.3, Mach-15 inflow, slip wall at a circular body, MUSCL-Hancock + HLLC, no diffusion.
 * Pinned: 0 ulp against the reference's own object code (oracle/_ref/libref_hypcpu_<W>x<H>.so, built by
 * oracle/Makefile from /root/reference/tau_hypersonic.c with the reference's `gcc -O3`) after 10 + 200
 * steps at 256 x 256 — tests/test_oracle_cpu.py::test_hypcpu_oracle_is_the_reference_bit_for_bit.
 * Built with -ffp-contract=off: the reference's flags (no -mfma) cannot contract either.
 *
 * Every function cites the tau_hypersonic.c lines it follows.  x and y sweeps share one body here:
 * `n` is the face-normal velocity / momentum, `t` the tangential one; the operations and their order
 * are those of the reference's hllc_x / hllc_y, flux_x / flux_y, reconstruct_x / reconstruct_y. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HC_GAMMA 1.4   /* :15 */
#define HC_CFL 0.3     /* :16 */
#define HC_EPS 1e-10   /* EPS_RHO == EPS_P, :20-21 */

typedef struct { double rho, mx, my, E; } hc_cons;
typedef struct { double rho, u, v, p; } hc_prim;

/* :65-81 */
static hc_prim hc_c2p(hc_cons c) {
  hc_prim q;
  const double rho = fmax(c.rho, HC_EPS), inv = 1.0 / rho;
  const double u = c.mx * inv, v = c.my * inv;
  const double kin = 0.5 * rho * (u * u + v * v);
  q.rho = rho; q.u = u; q.v = v;
  q.p = (HC_GAMMA - 1.0) * fmax(c.E - kin, HC_EPS);
  return q;
}
/* :83-93 */
static hc_cons hc_p2c(hc_prim q) {
  hc_cons c;
  const double rho = fmax(q.rho, HC_EPS), pr = fmax(q.p, HC_EPS);
  c.rho = rho; c.mx = rho * q.u; c.my = rho * q.v;
  c.E = pr / (HC_GAMMA - 1.0) + 0.5 * rho * (q.u * q.u + q.v * q.v);
  return c;
}
/* :95-97 */
static double hc_sound(hc_prim q) { return sqrt(HC_GAMMA * fmax(q.p, HC_EPS) / fmax(q.rho, HC_EPS)); }

/* flux_x :99-107 (ax 0), flux_y :109-117 (ax 1) */
static hc_cons hc_flux(hc_cons c, int ax) {
  const hc_prim q = hc_c2p(c);
  const double w = ax ? q.v : q.u;
  hc_cons f;
  f.rho = ax ? c.my : c.mx;
  f.mx = ax ? c.mx * w : c.mx * w + q.p;
  f.my = ax ? c.my * w + q.p : c.my * w;
  f.E = (c.E + q.p) * w;
  return f;
}

/* hllc_x :119-180, hllc_y :182-243 */
static hc_cons hc_hllc(hc_cons UL, hc_cons UR, int ax) {
  const hc_prim L = hc_c2p(UL), R = hc_c2p(UR);
  const double aL = hc_sound(L), aR = hc_sound(R);
  const double nL = ax ? L.v : L.u, nR = ax ? R.v : R.u;   /* normal velocities */
  const double tL = ax ? L.u : L.v, tR = ax ? R.u : R.v;   /* tangential velocities */
  const double SL = fmin(nL - aL, nR - aR), SR = fmax(nL + aL, nR + aR);
  const hc_cons FL = hc_flux(UL, ax), FR = hc_flux(UR, ax);
  if (SL >= 0.0) return FL;
  if (SR <= 0.0) return FR;
  const double num = R.p - L.p + L.rho * nL * (SL - nL) - R.rho * nR * (SR - nR);
  const double den = L.rho * (SL - nL) - R.rho * (SR - nR);
  const double SM = num / den;
  double pStar = L.p + L.rho * (SL - nL) * (SM - nL);
  pStar = fmax(pStar, HC_EPS);
  hc_cons S, F;
  if (SM >= 0.0) {   /* both star states are evaluated by the reference; only the selected one is used */
    const double rs = L.rho * (SL - nL) / (SL - SM);
    const double ms_n = rs * SM, ms_t = rs * tL;
    S.rho = rs; S.mx = ax ? ms_t : ms_n; S.my = ax ? ms_n : ms_t;
    S.E = ((SL - nL) * UL.E - L.p * nL + pStar * SM) / (SL - SM);
    F.rho = FL.rho + SL * (S.rho - UL.rho);
    F.mx = FL.mx + SL * (S.mx - UL.mx);
    F.my = FL.my + SL * (S.my - UL.my);
    F.E = FL.E + SL * (S.E - UL.E);
  } else {
    const double rs = R.rho * (SR - nR) / (SR - SM);
    const double ms_n = rs * SM, ms_t = rs * tR;
    S.rho = rs; S.mx = ax ? ms_t : ms_n; S.my = ax ? ms_n : ms_t;
    S.E = ((SR - nR) * UR.E - R.p * nR + pStar * SM) / (SR - SM);
    F.rho = FR.rho + SR * (S.rho - UR.rho);
    F.mx = FR.mx + SR * (S.mx - UR.mx);
    F.my = FR.my + SR * (S.my - UR.my);
    F.E = FR.E + SR * (S.E - UR.E);
  }
  return F;
}

/* :245-254 */
static hc_prim hc_inflow(void) {
  hc_prim s;
  s.rho = 1.0; s.p = 1.0; s.v = 0.0;
  s.u = 15.0 * sqrt(HC_GAMMA * 1.0 / 1.0);
  return s;
}

/* :279-294; (nx, ny) = (1, 0) or (0, 1) enter the arithmetic as the reference has them */
static hc_cons hc_reflect(hc_cons inside, double nx, double ny) {
  const hc_prim q = hc_c2p(inside);
  double vn = q.u * nx + q.v * ny;
  const double ut = -q.u * ny + q.v * nx;
  vn = -vn;
  hc_prim g;
  g.rho = q.rho; g.p = q.p;
  g.u = vn * nx - ut * ny;
  g.v = vn * ny + ut * nx;
  return hc_p2c(g);
}

typedef struct {
  int W, H;
  const double *rho, *mx, *my, *E;   /* the state being read (U) */
  const uint8_t *mask;
} hc_grid;

static hc_cons hc_at(const hc_grid *g, int i) {
  hc_cons c = {g->rho[i], g->mx[i], g->my[i], g->E[i]};
  return c;
}

/* :295-315 */
static hc_cons hc_neighbor_or_wall(const hc_grid *g, int x, int y, int dxc, int dyc, double nx, double ny) {
  const int xn = x + dxc;
  int yn = y + dyc;
  if (xn < 0) return hc_p2c(hc_inflow());
  if (xn >= g->W) return hc_at(g, y * g->W + g->W - 1);
  if (yn < 0) yn = 0;
  if (yn >= g->H) yn = g->H - 1;
  const int j = yn * g->W + xn;
  if (g->mask[j]) return hc_reflect(hc_at(g, y * g->W + x), nx, ny);
  return hc_at(g, j);
}

/* :35-63 */
static double hc_minmod(double a, double b) {
  if (a * b <= 0.0) return 0.0;
  return fabs(a) < fabs(b) ? a : b;
}
static double hc_mc(double dl, double dc, double dr) {
  const double m1 = hc_minmod(dl, dr), m2 = hc_minmod(dc, 2.0 * dl), m3 = hc_minmod(dc, 2.0 * dr);
  return hc_minmod(m1, hc_minmod(m2, m3));
}

/* :320-346 */
static void hc_positive_faces(hc_prim *qm, hc_prim qc, hc_prim *qp) {
  for (int it = 0; it < 8; ++it) {
    const int bad = qm->rho <= HC_EPS || qp->rho <= HC_EPS || qm->p <= HC_EPS || qp->p <= HC_EPS;
    if (!bad) return;
    qm->rho = 0.5 * (qm->rho + qc.rho); qm->u = 0.5 * (qm->u + qc.u);
    qm->v = 0.5 * (qm->v + qc.v);       qm->p = 0.5 * (qm->p + qc.p);
    qp->rho = 0.5 * (qp->rho + qc.rho); qp->u = 0.5 * (qp->u + qc.u);
    qp->v = 0.5 * (qp->v + qc.v);       qp->p = 0.5 * (qp->p + qc.p);
  }
  qm->rho = fmax(qm->rho, HC_EPS); qp->rho = fmax(qp->rho, HC_EPS);
  qm->p = fmax(qm->p, HC_EPS);     qp->p = fmax(qp->p, HC_EPS);
}

static double hc_slope(double m, double c, double p) { return hc_mc(c - m, 0.5 * (p - m), p - c); }

/* reconstruct_x :348-382 / reconstruct_y :384-418, then the Hancock half step of both face states
 * (half_step_predict_x/y :420-448 as step_physics calls them, :551-574 / :617-632): lo = the predicted
 * state on the cell's low face (left / bottom), hi = on its high face (right / top). */
static void hc_predict_cell(const hc_grid *g, int x, int y, int ax, double half_dt, hc_prim *lo, hc_prim *hi) {
  const double nx = ax ? 0.0 : 1.0, ny = ax ? 1.0 : 0.0;
  const hc_prim qc = hc_c2p(hc_at(g, y * g->W + x));
  const hc_prim qm = hc_c2p(hc_neighbor_or_wall(g, x, y, ax ? 0 : -1, ax ? -1 : 0, nx, ny));
  const hc_prim qp = hc_c2p(hc_neighbor_or_wall(g, x, y, ax ? 0 : +1, ax ? +1 : 0, nx, ny));
  const double s_rho = hc_slope(qm.rho, qc.rho, qp.rho), s_u = hc_slope(qm.u, qc.u, qp.u);
  const double s_v = hc_slope(qm.v, qc.v, qp.v), s_p = hc_slope(qm.p, qc.p, qp.p);
  hc_prim qL = {qc.rho - 0.5 * s_rho, qc.u - 0.5 * s_u, qc.v - 0.5 * s_v, qc.p - 0.5 * s_p};
  hc_prim qR = {qc.rho + 0.5 * s_rho, qc.u + 0.5 * s_u, qc.v + 0.5 * s_v, qc.p + 0.5 * s_p};
  hc_positive_faces(&qL, qc, &qR);
  const hc_cons Ff = hc_flux(hc_p2c(qR), ax), Fb = hc_flux(hc_p2c(qL), ax);
  const double d_rho = Ff.rho - Fb.rho, d_mx = Ff.mx - Fb.mx, d_my = Ff.my - Fb.my, d_E = Ff.E - Fb.E;
  for (int side = 0; side < 2; ++side) {
    hc_cons c = hc_p2c(side ? qR : qL);
    c.rho -= half_dt * d_rho; c.mx -= half_dt * d_mx; c.my -= half_dt * d_my; c.E -= half_dt * d_E;
    hc_prim o = hc_c2p(c);
    o.rho = fmax(o.rho, HC_EPS); o.p = fmax(o.p, HC_EPS);
    *(side ? hi : lo) = o;
  }
}

/* init_sim :450-475 */
void hypcpu_init(int W, int H, double *rho, double *mx, double *my, double *E, uint8_t *mask) {
  const int cx = W / 3, cy = H / 2, r = H / 6;
  const hc_prim in = hc_inflow();
  hc_prim rest = {in.rho, 0.0, 0.0, in.p};
  const hc_cons cin = hc_p2c(in), crest = hc_p2c(rest);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int i = y * W + x, dx = x - cx, dy = y - cy;
      mask[i] = (dx * dx + dy * dy < r * r) ? 1 : 0;
      const hc_cons c = mask[i] ? crest : cin;
      rho[i] = c.rho; mx[i] = c.mx; my[i] = c.my; E[i] = c.E;
    }
}

/* compute_dt :477-498 */
double hypcpu_dt(int W, int H, const double *rho, const double *mx, const double *my, const double *E,
                 const uint8_t *mask) {
  hc_grid g = {W, H, rho, mx, my, E, mask};
  double maxs = 1e-12;
  for (int i = 0; i < W * H; ++i) {
    if (mask[i]) continue;
    const hc_prim q = hc_c2p(hc_at(&g, i));
    const double a = hc_sound(q), sx = fabs(q.u) + a, sy = fabs(q.v) + a;
    if (sx > maxs) maxs = sx;
    if (sy > maxs) maxs = sy;
  }
  return HC_CFL * fmin(1.0, 1.0) / maxs;
}

/* one face of step_physics' sweeps (:524-586 x, :588-650 y): lo/hi = the cells on the face's low / high side */
static hc_cons hc_face_flux(const hc_grid *g, int xl, int yl, int xh, int yh, int ax, double half_dt) {
  const double nx = ax ? 0.0 : 1.0, ny = ax ? 1.0 : 0.0;
  const int il = yl * g->W + xl, ih = yh * g->W + xh;
  hc_prim ql, qh, other;
  if (!g->mask[il]) hc_predict_cell(g, xl, yl, ax, half_dt, &other, &ql);
  else ql = hc_c2p(hc_reflect(hc_at(g, ih), nx, ny));
  if (!g->mask[ih]) hc_predict_cell(g, xh, yh, ax, half_dt, &qh, &other);
  else qh = hc_c2p(hc_reflect(hc_at(g, il), nx, ny));
  ql.rho = fmax(ql.rho, HC_EPS); ql.p = fmax(ql.p, HC_EPS);
  qh.rho = fmax(qh.rho, HC_EPS); qh.p = fmax(qh.p, HC_EPS);
  return hc_hllc(hc_p2c(ql), hc_p2c(qh), ax);
}

/* step_physics :500-674, nsteps times; *sim_t advances by each dt (:673).  dts (may be NULL) receives every dt. */
void hypcpu_step(int W, int H, double *rho, double *mx, double *my, double *E, const uint8_t *mask, int nsteps,
                 double *sim_t, double *dts) {
  const size_t N = (size_t)W * H;
  double *n_rho = (double *)malloc(4 * N * sizeof(double));
  double *n_mx = n_rho + N, *n_my = n_mx + N, *n_E = n_my + N;
  const hc_cons cin = hc_p2c(hc_inflow());
  for (int s = 0; s < nsteps; ++s) {
    const double dt = hypcpu_dt(W, H, rho, mx, my, E, mask);   /* before the inflow overwrite, :502 */
    const double half_dt = 0.5 * (dt / 1.0);
    for (int y = 0; y < H; ++y)                                 /* :509-515 */
      if (!mask[y * W]) { rho[y * W] = cin.rho; mx[y * W] = cin.mx; my[y * W] = cin.my; E[y * W] = cin.E; }
    memcpy(n_rho, rho, N * sizeof(double)); memcpy(n_mx, mx, N * sizeof(double));
    memcpy(n_my, my, N * sizeof(double));   memcpy(n_E, E, N * sizeof(double));
    hc_grid g = {W, H, rho, mx, my, E, mask};
    for (int ax = 0; ax < 2; ++ax)
      for (int y = ax; y < H; ++y)
        for (int x = 1 - ax; x < W; ++x) {
          const int xl = x - (1 - ax), yl = y - ax, il = yl * W + xl, ih = y * W + x;
          if (mask[il] && mask[ih]) continue;
          const hc_cons F = hc_face_flux(&g, xl, yl, x, y, ax, half_dt);
          if (!mask[il]) { n_rho[il] -= dt * F.rho; n_mx[il] -= dt * F.mx; n_my[il] -= dt * F.my; n_E[il] -= dt * F.E; }
          if (!mask[ih]) { n_rho[ih] += dt * F.rho; n_mx[ih] += dt * F.mx; n_my[ih] += dt * F.my; n_E[ih] += dt * F.E; }
        }
    for (size_t i = 0; i < N; ++i) {                            /* :652-671 */
      if (mask[i]) continue;
      hc_cons c = {fmax(n_rho[i], HC_EPS), n_mx[i], n_my[i], n_E[i]};
      hc_prim q = hc_c2p(c);
      if (q.p <= HC_EPS) { q.p = HC_EPS; c = hc_p2c(q); }
      rho[i] = c.rho; mx[i] = c.mx; my[i] = c.my; E[i] = c.E;
    }
    *sim_t += dt;
    if (dts) dts[s] = dt;
  }
  free(n_rho);
}

/* get_cell_with_bc :256-277 (density only) */
static double hc_rho_bc(const hc_grid *g, int x, int y) {
  if (y < 0) y = 0;
  if (y >= g->H) y = g->H - 1;
  if (x < 0) return hc_c2p(hc_p2c(hc_inflow())).rho;
  if (x >= g->W) x = g->W - 1;
  return hc_c2p(hc_at(g, y * g->W + x)).rho;
}

/* the view value of main()'s render loop, :722-742: 0 log rho, 1 log p, 2 speed ("speed mode"), 3 schlieren */
static double hc_view(const hc_grid *g, int x, int y, int mode) {
  const hc_prim q = hc_c2p(hc_at(g, y * g->W + x));
  if (mode == 0) return log(q.rho);
  if (mode == 1) return log(q.p);
  if (mode == 2) return sqrt(q.u * q.u + q.v * q.v);
  const double gx = 0.5 * (hc_rho_bc(g, x + 1, y) - hc_rho_bc(g, x - 1, y));
  const double gy = 0.5 * (hc_rho_bc(g, x, y + 1) - hc_rho_bc(g, x, y - 1));
  return log(1e-12 + sqrt(gx * gx + gy * gy));
}

/* main()'s min/max scan and pixel loop (:713-786) with get_color (:676-686); rgba = 4 bytes per cell.
 * values (may be NULL) receives the view value of every fluid cell (0 for body cells). */
void hypcpu_render(int W, int H, const double *rho, const double *mx, const double *my, const double *E,
                   const uint8_t *mask, int mode, uint8_t *rgba, double minmax[2], double *values) {
  hc_grid g = {W, H, rho, mx, my, E, mask};
  double minv = 1e300, maxv = -1e300;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int i = y * W + x;
      if (values) values[i] = 0.0;
      if (mask[i]) continue;
      const double v = hc_view(&g, x, y, mode);
      if (values) values[i] = v;
      if (v < minv) minv = v;
      if (v > maxv) maxv = v;
    }
  const double inv = 1.0 / fmax(maxv - minv, 1e-30);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int i = y * W + x;
      uint8_t *px = rgba + 4 * (size_t)i;
      px[3] = 255;
      if (mask[i]) { px[0] = px[1] = px[2] = 110; continue; }
      double t = (hc_view(&g, x, y, mode) - minv) * inv;
      if (t < 0) t = 0;
      if (t > 1) t = 1;
      px[0] = (uint8_t)(255 * fmin(1, fmax(0, 3 * t - 1)));
      px[1] = (uint8_t)(255 * fmin(1, fmax(0, 2 - 4 * fabs(t - 0.5))));
      px[2] = (uint8_t)(255 * fmin(1, fmax(0, 2 - 3 * t)));
    }
  if (minmax) { minmax[0] = minv; minmax[1] = maxv; }
}
