// snapshot.cu — the on-disk artefacts either side of the 2-D hypersonic hot path (SURVEY.md 8(f)
// rank 2), host code over the public C-ABI only:
//   * the reference's 12-scalar regression snapshot (tau_hypersonic_cuda_tests.cu:20-36 struct,
//     compute_snapshot :143-176, write/read :84-125, verification tolerances :527-557) — same text
//     format, same summation order (sequential over y*W+x on the host, like the reference), so a
//     baseline file written by either side verifies against the other;
//   * a raw SoA checkpoint ("TAUCKPT1": config, grid, slab, step counter, sim_t, planes, mask) from
//     which a run — single GPU or one file per slab — resumes bit-identically.  The reference has
//     no state output.
#include "common.cuh"
#include "../../include/tau_b200.h"

#include <math.h>
#include <vector>

namespace {

struct HostState {
  int W, H, dtype, y_begin, h_local;
  tau_hyp2d_config cfg;
  std::vector<double> rho, mx, my, E;  // always widened to fp64 (exact for an fp32 handle)
  std::vector<uint8_t> mask;
};

int fetch(tau_hyp2d *h, HostState &s, bool widen) {
  int rc = tau_hyp2d_describe(h, &s.W, &s.H, &s.dtype, &s.y_begin, &s.h_local, &s.cfg);
  if (rc) return rc;
  const size_t n = (size_t)s.W * s.h_local;
  s.mask.resize(n);
  s.rho.resize(n); s.mx.resize(n); s.my.resize(n); s.E.resize(n);
  if (s.dtype == TAU_F64) {
    void *pl[4] = {s.rho.data(), s.mx.data(), s.my.data(), s.E.data()};
    return tau_hyp2d_download(h, pl, s.mask.data());
  }
  std::vector<float> f(4 * n);
  void *pl[4] = {f.data(), f.data() + n, f.data() + 2 * n, f.data() + 3 * n};
  rc = tau_hyp2d_download(h, pl, s.mask.data());
  if (rc) return rc;
  if (widen)
    for (size_t i = 0; i < n; ++i) {
      s.rho[i] = f[i]; s.mx[i] = f[n + i]; s.my[i] = f[2 * n + i]; s.E[i] = f[3 * n + i];
    }
  return TAU_OK;
}

}  // namespace

extern "C" {

// compute_snapshot :143-176 on the handle's current state.  For a slab handle the sums cover the
// slab's rows with GLOBAL cell indices in the checksum weights, so per-rank results combine by
// addition (counts, sums, checksums) and min/max.
int tau_hyp2d_snapshot(tau_hyp2d *h, tau_hyp2d_snapshot_t *out) {
  TAU_REQUIRE(h && out, "tau_hyp2d_snapshot: null argument");
  HostState s;
  int rc = fetch(h, s, true);
  if (rc) return rc;
  tau_hyp2d_snapshot_t r;
  memset(&r, 0, sizeof(r));
  r.steps = (int)tau_hyp2d_steps_done(h);
  r.min_rho = 1e300;
  r.min_p = 1e300;
  const double eps = 1e-25, gamma = s.cfg.gamma;
  const size_t n = (size_t)s.W * s.h_local, i0 = (size_t)s.y_begin * s.W;
  for (size_t k = 0; k < n; ++k) {
    if (s.mask[k]) continue;
    const double rho = fmax(s.rho[k], eps);  // host_cons_to_prim :127-141
    const double inv = 1.0 / rho;
    const double u = s.mx[k] * inv, v = s.my[k] * inv;
    const double kin = 0.5 * rho * (u * u + v * v);
    const double p = (gamma - 1.0) * fmax(s.E[k] - kin, eps);
    const double a = sqrt(gamma * fmax(p, eps) / fmax(rho, eps));
    const double mach = sqrt(u * u + v * v) / fmax(a, 1e-30);
    const double w = (double)(((i0 + k) % 8191) + 1);
    r.fluid_cells++;
    r.sum_rho += rho;
    r.sum_mx += s.mx[k];
    r.sum_my += s.my[k];
    r.sum_E += s.E[k];
    r.min_rho = fmin(r.min_rho, rho);
    r.min_p = fmin(r.min_p, p);
    r.max_mach = fmax(r.max_mach, mach);
    r.checksum_rho += w * rho;
    r.checksum_mx += w * s.mx[k];
    r.checksum_E += w * s.E[k];
  }
  *out = r;
  return TAU_OK;
}

// write_snapshot :84-105 — byte-for-byte the reference's text format
int tau_hyp2d_snapshot_write(const char *path, const tau_hyp2d_snapshot_t *s) {
  TAU_REQUIRE(path && s, "tau_hyp2d_snapshot_write: null argument");
  FILE *f = fopen(path, "w");
  TAU_REQUIRE(f, "Failed to open baseline for write: %s", path);
  fprintf(f, "steps %d\n", s->steps);
  fprintf(f, "fluid_cells %d\n", s->fluid_cells);
  fprintf(f, "sum_rho %.17g\n", s->sum_rho);
  fprintf(f, "sum_mx %.17g\n", s->sum_mx);
  fprintf(f, "sum_my %.17g\n", s->sum_my);
  fprintf(f, "sum_E %.17g\n", s->sum_E);
  fprintf(f, "min_rho %.17g\n", s->min_rho);
  fprintf(f, "min_p %.17g\n", s->min_p);
  fprintf(f, "max_mach %.17g\n", s->max_mach);
  fprintf(f, "checksum_rho %.17g\n", s->checksum_rho);
  fprintf(f, "checksum_mx %.17g\n", s->checksum_mx);
  fprintf(f, "checksum_E %.17g\n", s->checksum_E);
  fclose(f);
  return TAU_OK;
}

// read_snapshot :107-125
int tau_hyp2d_snapshot_read(const char *path, tau_hyp2d_snapshot_t *s) {
  TAU_REQUIRE(path && s, "tau_hyp2d_snapshot_read: null argument");
  FILE *f = fopen(path, "r");
  TAU_REQUIRE(f, "Failed to open baseline for read: %s", path);
  const int fields = fscanf(
      f,
      "steps %d\nfluid_cells %d\nsum_rho %lf\nsum_mx %lf\nsum_my %lf\nsum_E %lf\n"
      "min_rho %lf\nmin_p %lf\nmax_mach %lf\nchecksum_rho %lf\nchecksum_mx %lf\nchecksum_E %lf\n",
      &s->steps, &s->fluid_cells, &s->sum_rho, &s->sum_mx, &s->sum_my, &s->sum_E, &s->min_rho, &s->min_p,
      &s->max_mach, &s->checksum_rho, &s->checksum_mx, &s->checksum_E);
  fclose(f);
  TAU_REQUIRE(fields == 12, "read regression baseline: %s holds %d of 12 fields", path, fields);
  return TAU_OK;
}

// The verification block :527-557 with its tolerances.  Returns 0 when every check passes, else
// the number of failed checks (their names, the reference's messages, are in tau_last_error()).
int tau_hyp2d_snapshot_compare(const tau_hyp2d_snapshot_t *cur, const tau_hyp2d_snapshot_t *exp) {
  TAU_REQUIRE(cur && exp, "tau_hyp2d_snapshot_compare: null argument");
  char msg[900] = "";
  int failed = 0;
  auto fail = [&](const char *what) {
    ++failed;
    const size_t l = strlen(msg);
    snprintf(msg + l, sizeof(msg) - l, "%sFAIL: %s", l ? "; " : "", what);
  };
  auto near = [&](double a, double b, double tol, const char *what) {
    if (!(fabs(a - b) <= tol)) fail(what);
  };
  if (cur->steps != exp->steps) fail("steps match baseline");
  if (cur->fluid_cells != exp->fluid_cells) fail("fluid cell count matches baseline");
  near(cur->sum_rho, exp->sum_rho, 5e-8 * fabs(exp->sum_rho) + 1e-8, "sum_rho matches baseline");
  near(cur->sum_mx, exp->sum_mx, 5e-8 * fabs(exp->sum_mx) + 1e-8, "sum_mx matches baseline");
  near(cur->sum_my, exp->sum_my, 5e-8 * fabs(exp->sum_my) + 1e-8, "sum_my matches baseline");
  near(cur->sum_E, exp->sum_E, 5e-8 * fabs(exp->sum_E) + 1e-8, "sum_E matches baseline");
  near(cur->min_rho, exp->min_rho, 1e-9, "min_rho matches baseline");
  near(cur->min_p, exp->min_p, 1e-9, "min_p matches baseline");
  near(cur->max_mach, exp->max_mach, 5e-8 * fabs(exp->max_mach) + 1e-8, "max_mach matches baseline");
  near(cur->checksum_rho, exp->checksum_rho, 5e-8 * fabs(exp->checksum_rho) + 1e-8, "checksum_rho matches baseline");
  near(cur->checksum_mx, exp->checksum_mx, 5e-8 * fabs(exp->checksum_mx) + 1e-8, "checksum_mx matches baseline");
  near(cur->checksum_E, exp->checksum_E, 5e-8 * fabs(exp->checksum_E) + 1e-8, "checksum_E matches baseline");
  if (failed) tau_set_error("%s", msg);
  return failed;
}

// ---- checkpoint / resume ------------------------------------------------------------------------
// char magic[8] = "TAUCKPT1"; int32 W, H, dtype, y_begin, h_local, reserved; int64 steps;
// double sim_t; tau_hyp2d_config cfg; 4 planes of h_local*W elements in the handle's dtype; mask.
struct CkptHeader {
  char magic[8];
  int32_t W, H, dtype, y_begin, h_local, reserved;
  int64_t steps;
  double sim_t;
  tau_hyp2d_config cfg;
};

int tau_hyp2d_checkpoint_save(tau_hyp2d *h, const char *path) {
  TAU_REQUIRE(h && path, "tau_hyp2d_checkpoint_save: null argument");
  CkptHeader hd;
  memset(&hd, 0, sizeof(hd));
  memcpy(hd.magic, "TAUCKPT1", 8);
  int W, H, dtype, y0, hl;
  int rc = tau_hyp2d_describe(h, &W, &H, &dtype, &y0, &hl, &hd.cfg);
  if (rc) return rc;
  hd.W = W; hd.H = H; hd.dtype = dtype; hd.y_begin = y0; hd.h_local = hl;
  hd.steps = tau_hyp2d_steps_done(h);
  rc = tau_hyp2d_clock(h, &hd.sim_t, nullptr);
  if (rc) return rc;
  const size_t n = (size_t)W * hl, es = dtype ? 8 : 4;
  std::vector<unsigned char> buf(4 * n * es);
  std::vector<uint8_t> mask(n);
  void *pl[4] = {buf.data(), buf.data() + n * es, buf.data() + 2 * n * es, buf.data() + 3 * n * es};
  rc = tau_hyp2d_download(h, pl, mask.data());
  if (rc) return rc;
  FILE *f = fopen(path, "wb");
  TAU_REQUIRE(f, "tau_hyp2d_checkpoint_save: cannot open %s for writing", path);
  const bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1 && fwrite(buf.data(), 1, buf.size(), f) == buf.size() &&
                  fwrite(mask.data(), 1, n, f) == n;
  fclose(f);
  TAU_REQUIRE(ok, "tau_hyp2d_checkpoint_save: short write to %s", path);
  return TAU_OK;
}

// Reads only the header (to create a matching handle): any out pointer may be NULL.
int tau_hyp2d_checkpoint_info(const char *path, int *W, int *H, int *dtype, int *y_begin, int *h_local,
                              long long *steps, double *sim_t, tau_hyp2d_config *cfg) {
  TAU_REQUIRE(path, "tau_hyp2d_checkpoint_info: null path");
  FILE *f = fopen(path, "rb");
  TAU_REQUIRE(f, "tau_hyp2d_checkpoint_info: cannot open %s", path);
  CkptHeader hd;
  const bool ok = fread(&hd, sizeof(hd), 1, f) == 1 && !memcmp(hd.magic, "TAUCKPT1", 8);
  fclose(f);
  TAU_REQUIRE(ok, "tau_hyp2d_checkpoint_info: %s is not a TAUCKPT1 file", path);
  if (W) *W = hd.W;
  if (H) *H = hd.H;
  if (dtype) *dtype = hd.dtype;
  if (y_begin) *y_begin = hd.y_begin;
  if (h_local) *h_local = hd.h_local;
  if (steps) *steps = hd.steps;
  if (sim_t) *sim_t = hd.sim_t;
  if (cfg) *cfg = hd.cfg;
  return TAU_OK;
}

int tau_hyp2d_checkpoint_load(tau_hyp2d *h, const char *path) {
  TAU_REQUIRE(h && path, "tau_hyp2d_checkpoint_load: null argument");
  int W, H, dtype, y0, hl;
  int rc = tau_hyp2d_describe(h, &W, &H, &dtype, &y0, &hl, nullptr);
  if (rc) return rc;
  FILE *f = fopen(path, "rb");
  TAU_REQUIRE(f, "tau_hyp2d_checkpoint_load: cannot open %s", path);
  CkptHeader hd;
  if (fread(&hd, sizeof(hd), 1, f) != 1 || memcmp(hd.magic, "TAUCKPT1", 8)) {
    fclose(f);
    tau_set_error("tau_hyp2d_checkpoint_load: %s is not a TAUCKPT1 file", path);
    return TAU_ERR_INVALID;
  }
  if (hd.W != W || hd.H != H || hd.dtype != dtype || hd.y_begin != y0 || hd.h_local != hl) {
    fclose(f);
    tau_set_error("tau_hyp2d_checkpoint_load: %s holds a %dx%d %s slab [%d,%d), the handle is %dx%d %s [%d,%d)",
                  path, hd.W, hd.H, hd.dtype ? "f64" : "f32", hd.y_begin, hd.y_begin + hd.h_local, W, H,
                  dtype ? "f64" : "f32", y0, y0 + hl);
    return TAU_ERR_INVALID;
  }
  const size_t n = (size_t)W * hl, es = dtype ? 8 : 4;
  std::vector<unsigned char> buf(4 * n * es);
  std::vector<uint8_t> mask(n);
  const bool ok = fread(buf.data(), 1, buf.size(), f) == buf.size() && fread(mask.data(), 1, n, f) == n;
  fclose(f);
  TAU_REQUIRE(ok, "tau_hyp2d_checkpoint_load: %s is truncated", path);
  const void *pl[4] = {buf.data(), buf.data() + n * es, buf.data() + 2 * n * es, buf.data() + 3 * n * es};
  rc = tau_hyp2d_upload(h, pl, mask.data());
  if (rc) return rc;
  return tau_hyp2d_set_clock(h, hd.sim_t, hd.steps);
}

}  // extern "C"
