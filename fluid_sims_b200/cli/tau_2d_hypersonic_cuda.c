/* tau_2d_hypersonic_cuda — C host of the 2-D hypersonic solver over libtau_b200.so.
 * Keeps the reference binary's flags (parse_args, tau_hypersonic_cuda.cu:1482-1639: --mach --gamma
 * --cfl --visc-nu --visc-rho --visc-e --steps-per-frame --geom-x0 --geom-cy --geom-rb --geom-rn
 * --geom-theta --tile-bx --tile-by) with the same validation messages, and replaces the raylib
 * frame loop (:1824-1947) by a headless one.  Additive flags: --nx/--ny (the reference's grid is a
 * compile-time #define), --frames, --dtype f32|f64, --dump FILE; --view MODE --ppm FILE (the render
 * pass of the frame loop :1892-1926 written as a binary PPM instead of a raylib texture; MODE as
 * the reference's M key cycles it, 0..6); --write-baseline / --verify-baseline FILE (the regression
 * snapshot of tau_hypersonic_cuda_tests.cu:84-176, same file format and tolerances);
 * --checkpoint FILE / --resume FILE (raw SoA state, bit-identical resume);
 * --gpus N (additive): the grid is split into N y-slabs, one per device of this box, all driven from this
 * one process through tau_hyp2d_group_* — results are bit-identical to --gpus 1. */
#include <errno.h>
#include <limits.h>
#include <math.h>

#include "cli_common.h"

static void usage(const char *a0) {
  fprintf(stderr,
          "Usage: %s [--mach M] [--gamma G] [--cfl C] [--visc-nu NU]\n"
          "          [--visc-rho MU] [--visc-e K] [--steps-per-frame N]\n"
          "          [--geom-x0 X0] [--geom-cy CY] [--geom-rb RB]\n"
          "          [--geom-rn RN] [--geom-theta THETA]\n"
          "          [--tile-bx BX] [--tile-by BY]\n"
          "          [--nx W] [--ny H] [--frames N] [--dtype f32|f64] [--dump FILE]\n"
          "          [--view MODE] [--ppm FILE] [--write-baseline FILE | --verify-baseline FILE]\n"
          "          [--checkpoint FILE] [--resume FILE] [--gpus N]\n",
          a0);
}
static int parse_d(const char *name, const char *v, double *out) { /* parse_double_flag :1462 */
  char *end = NULL;
  double x = strtod(v, &end);
  if (!end || *end != '\0' || !isfinite(x)) {
    fprintf(stderr, "Invalid value for %s: %s\n", name, v);
    return 0;
  }
  *out = x;
  return 1;
}
static int parse_i(const char *name, const char *v, int *out) { /* parse_int_flag :1472 */
  char *end = NULL;
  errno = 0;
  long x = strtol(v, &end, 10);
  if (!end || *end != '\0' || errno == ERANGE || x < INT_MIN || x > INT_MAX) {
    fprintf(stderr, "Invalid value for %s: %s\n", name, v);
    return 0;
  }
  *out = (int)x;
  return 1;
}

int main(int argc, char **argv) {
  int W = 8192, H = 1024, frames = 100, dtype = TAU_F64, tile_bx = -1, tile_by = -1;
  const char *dump = NULL, *ppm = NULL, *wbase = NULL, *vbase = NULL, *ckpt = NULL, *resume = NULL;
  int view = 0, gpus = 1;
  /* first pass: grid size, because default_config derives the geometry from H (:1401-1405) */
  for (int i = 1; i + 1 < argc; i++) {
    if (!strcmp(argv[i], "--nx") && !parse_i("--nx", argv[i + 1], &W)) return 1;
    if (!strcmp(argv[i], "--ny") && !parse_i("--ny", argv[i + 1], &H)) return 1;
  }
  tau_hyp2d_config c;
  tau_hyp2d_default_config(&c, W, H);
  for (int i = 1; i < argc; i++) {
    const char *a = argv[i];
    const int has = i + 1 < argc;
#define DFLAG(flag, field) \
  if (!strcmp(a, flag) && has) { if (!parse_d(a, argv[++i], &c.field)) { usage(argv[0]); return 1; } continue; }
    DFLAG("--mach", inflow_mach) DFLAG("--gamma", gamma) DFLAG("--cfl", cfl) DFLAG("--visc-nu", visc_nu)
    DFLAG("--visc-rho", visc_rho) DFLAG("--visc-e", visc_e) DFLAG("--geom-x0", geom_x0)
    DFLAG("--geom-cy", geom_cy) DFLAG("--geom-rb", geom_Rb) DFLAG("--geom-rn", geom_Rn)
    DFLAG("--geom-theta", geom_theta)
#undef DFLAG
    if (!strcmp(a, "--steps-per-frame") && has) { if (!parse_i(a, argv[++i], &c.steps_per_frame)) return 1; continue; }
    if (!strcmp(a, "--tile-bx") && has) { if (!parse_i(a, argv[++i], &tile_bx)) return 1; continue; }
    if (!strcmp(a, "--tile-by") && has) { if (!parse_i(a, argv[++i], &tile_by)) return 1; continue; }
    if ((!strcmp(a, "--nx") || !strcmp(a, "--ny")) && has) { ++i; continue; }
    if (!strcmp(a, "--frames") && has) { if (!parse_i(a, argv[++i], &frames)) return 1; continue; }
    if (!strcmp(a, "--dtype") && has) { dtype = !strcmp(argv[++i], "f32") ? TAU_F32 : TAU_F64; continue; }
    if (!strcmp(a, "--dump") && has) { dump = argv[++i]; continue; }
    if (!strcmp(a, "--view") && has) { if (!parse_i(a, argv[++i], &view)) return 1; continue; }
    if (!strcmp(a, "--ppm") && has) { ppm = argv[++i]; continue; }
    if (!strcmp(a, "--write-baseline") && has) { wbase = argv[++i]; continue; }
    if (!strcmp(a, "--verify-baseline") && has) { vbase = argv[++i]; continue; }
    if (!strcmp(a, "--checkpoint") && has) { ckpt = argv[++i]; continue; }
    if (!strcmp(a, "--resume") && has) { resume = argv[++i]; continue; }
    if (!strcmp(a, "--gpus") && has) { if (!parse_i(a, argv[++i], &gpus)) return 1; continue; }
    fprintf(stderr, "Unknown or incomplete argument: %s\n", a);
    usage(argv[0]);
    return 1;
  }
  if (tau_hyp2d_validate_config(&c) != 0) { /* same checks/messages as parse_args :1545-1637 */
    fprintf(stderr, "%s\n", tau_last_error());
    usage(argv[0]);
    return 1;
  }
  printf("SimConfig:\n  gamma=%.8g\n  cfl=%.8g\n  visc_nu=%.8g\n  visc_rho=%.8g\n  visc_e=%.8g\n"
         "  inflow_mach=%.8g\n  steps_per_frame=%d\n"
         "  geom_x0=%.8g geom_cy=%.8g geom_Rb=%.8g geom_Rn=%.8g geom_theta=%.8g\n",
         c.gamma, c.cfl, c.visc_nu, c.visc_rho, c.visc_e, c.inflow_mach, c.steps_per_frame, c.geom_x0,
         c.geom_cy, c.geom_Rb, c.geom_Rn, c.geom_theta); /* print_config :1687-1709 */
  if (gpus != 1) { /* y-slabs over several devices of this box, one process (tau_hyp2d_group_*) */
    if (wbase || vbase || ckpt || resume) {
      fprintf(stderr, "--gpus %d: baselines and checkpoints are per-handle operations (run with --gpus 1)\n", gpus);
      return 1;
    }
    tau_hyp2d_group *g;
    TAU_OR_DIE(tau_hyp2d_group_create(&c, W, H, dtype, gpus, NULL, &g));
    TAU_OR_DIE(tau_hyp2d_group_init(g));
    printf("LaunchConfig:\n  grid=%dx%d dtype=%s y-slabs over %d GPUs:", W, H, dtype ? "f64" : "f32", gpus);
    for (int i = 0; i < gpus; ++i) {
      int y0, hl;
      TAU_OR_DIE(tau_hyp2d_group_member(g, i, NULL, &y0, &hl));
      printf(" [%d,%d)", y0, y0 + hl);
    }
    printf("\n");
    const double t0 = cli_now();
    double sim_t = 0, dt = 0;
    for (int f = 0; f < frames; f++) {
      TAU_OR_DIE(tau_hyp2d_group_step(g, c.steps_per_frame));
      if ((f + 1) % 50 == 0 || f + 1 == frames) {
        TAU_OR_DIE(tau_hyp2d_group_clock(g, &sim_t, &dt));
        printf("frame %d  t = %.6f  dt = %.3e\n", f + 1, sim_t, dt);
      }
    }
    TAU_OR_DIE(tau_hyp2d_group_sync(g));
    const double secs = cli_now() - t0;
    const double steps = (double)frames * c.steps_per_frame;
    printf("%.0f steps in %.3f s: %.1f Mcell-updates/s\n", steps, secs, steps * W * H / secs / 1e6);
    if (dump) {
      const size_t n = (size_t)W * H, es = dtype ? 8 : 4;
      void *planes[4];
      for (int p = 0; p < 4; ++p) planes[p] = malloc(n * es);
      TAU_OR_DIE(tau_hyp2d_group_download(g, planes, NULL));
      cli_dump(dump, 4, (int)es, W, H, 1, (long long)steps, sim_t, planes);
      for (int p = 0; p < 4; ++p) free(planes[p]);
    }
    if (ppm) {
      uint32_t *px = (uint32_t *)malloc((size_t)W * H * 4);
      double mm[2];
      TAU_OR_DIE(tau_hyp2d_group_render(g, view, px, mm));
      FILE *f = fopen(ppm, "wb");
      if (!f) { fprintf(stderr, "cannot open %s for writing\n", ppm); return 1; }
      fprintf(f, "P6\n%d %d\n255\n", W, H);
      for (size_t i = 0; i < (size_t)W * H; ++i) fwrite(&px[i], 1, 3, f);
      fclose(f);
      free(px);
      printf("view %d: value range [%.6g, %.6g] -> %s\n", view, mm[0], mm[1], ppm);
    }
    TAU_OR_DIE(tau_hyp2d_group_destroy(g));
    return 0;
  }
  tau_hyp2d *sim;
  TAU_OR_DIE(tau_hyp2d_create(&c, W, H, dtype, 0, 0, H, NULL, &sim));
  if (tile_by > 0) TAU_OR_DIE(tau_hyp2d_set_seg_rows(sim, tile_by < 4 ? 4 : tile_by));
  if (resume) TAU_OR_DIE(tau_hyp2d_checkpoint_load(sim, resume));
  else TAU_OR_DIE(tau_hyp2d_init(sim));
  printf("LaunchConfig:\n  grid=%dx%d dtype=%s marching segment=%d rows\n", W, H, dtype ? "f64" : "f32",
         tau_hyp2d_get_seg_rows(sim));
  const double t0 = cli_now();
  double sim_t = 0, dt = 0;
  for (int f = 0; f < frames; f++) {
    TAU_OR_DIE(tau_hyp2d_step(sim, c.steps_per_frame)); /* the loop body :1833-1889 */
    if ((f + 1) % 50 == 0 || f + 1 == frames) {
      TAU_OR_DIE(tau_hyp2d_clock(sim, &sim_t, &dt));
      printf("frame %d  t = %.6f  dt = %.3e\n", f + 1, sim_t, dt);
    }
  }
  TAU_OR_DIE(tau_hyp2d_sync(sim));
  const double secs = cli_now() - t0;
  const double steps = (double)frames * c.steps_per_frame;
  printf("%.0f steps in %.3f s: %.1f Mcell-updates/s\n", steps, secs, steps * W * H / secs / 1e6);
  if (dump) {
    const size_t n = (size_t)W * H, es = dtype ? 8 : 4;
    void *planes[4];
    for (int p = 0; p < 4; ++p) planes[p] = malloc(n * es);
    TAU_OR_DIE(tau_hyp2d_download(sim, planes, NULL));
    cli_dump(dump, 4, (int)es, W, H, 1, tau_hyp2d_steps_done(sim), sim_t, planes);
    for (int p = 0; p < 4; ++p) free(planes[p]);
  }
  if (ppm) { /* render pass A + B of the frame loop, then the texture upload becomes a file */
    uint32_t *px = (uint32_t *)malloc((size_t)W * H * 4);
    double mm[2];
    TAU_OR_DIE(tau_hyp2d_render(sim, view, px, mm));
    FILE *f = fopen(ppm, "wb");
    if (!f) { fprintf(stderr, "cannot open %s for writing\n", ppm); return 1; }
    fprintf(f, "P6\n%d %d\n255\n", W, H);
    for (size_t i = 0; i < (size_t)W * H; ++i) fwrite(&px[i], 1, 3, f); /* R,G,B of uchar4 */
    fclose(f);
    free(px);
    printf("view %d: value range [%.6g, %.6g] -> %s\n", view, mm[0], mm[1], ppm);
  }
  if (wbase || vbase) {
    tau_hyp2d_snapshot_t cur, exp;
    TAU_OR_DIE(tau_hyp2d_snapshot(sim, &cur));
    if (wbase) TAU_OR_DIE(tau_hyp2d_snapshot_write(wbase, &cur));
    if (vbase) {
      TAU_OR_DIE(tau_hyp2d_snapshot_read(vbase, &exp));
      const int failed = tau_hyp2d_snapshot_compare(&cur, &exp);
      if (failed) {
        fprintf(stderr, "%s\n%d baseline check(s) failed\n", tau_last_error(), failed);
        return 1;
      }
      printf("baseline %s verified\n", vbase);
    }
  }
  if (ckpt) TAU_OR_DIE(tau_hyp2d_checkpoint_save(sim, ckpt));
  TAU_OR_DIE(tau_hyp2d_destroy(sim));
  return 0;
}
