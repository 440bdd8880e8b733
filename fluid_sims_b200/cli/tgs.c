/* tgs — C host of the Gray-Scott solver over libtau_b200.so.  Keeps the reference's long options
 * (tau_gray_scott.cu:82-135: --nx --ny --dx --dt --Du --Dv --F --k --steps --headless --stride
 * --fps --seed --halfblocks, -h) and the loop of main() (:321-346).  The ncurses renderer is not
 * part of the update path: without --headless the field is summarised on stdout every `stride`
 * steps instead.  Additive: --dump FILE. */
#include <getopt.h>

#include "cli_common.h"

int main(int argc, char **argv) {
  tau_gs_params p;
  tau_gs_default_params(&p);
  p.nx = p.ny = 0;
  int steps = 0, headless = 0, stride = 4;
  const char *dump = NULL;
  static const struct option lo[] = {
      {"nx", required_argument, 0, 0},   {"ny", required_argument, 0, 0},    {"dx", required_argument, 0, 0},
      {"dt", required_argument, 0, 0},   {"Du", required_argument, 0, 0},    {"Dv", required_argument, 0, 0},
      {"F", required_argument, 0, 0},    {"k", required_argument, 0, 0},     {"steps", required_argument, 0, 0},
      {"headless", no_argument, 0, 0},   {"stride", required_argument, 0, 0}, {"fps", required_argument, 0, 0},
      {"seed", required_argument, 0, 0}, {"halfblocks", no_argument, 0, 0},  {"dump", required_argument, 0, 0},
      {"help", no_argument, 0, 'h'},     {0, 0, 0, 0}};
  for (;;) {
    int idx = 0, c = getopt_long(argc, argv, "h", lo, &idx);
    if (c == -1) break;
    if (c == 'h') {
      printf("Usage: %s [--nx N] [--ny N] [--dx DX] [--dt DT] [--Du D] [--Dv D] [--F F] [--k K]\n"
             "          [--steps K] [--headless] [--stride N] [--fps N] [--seed S] [--halfblocks] [--dump FILE]\n",
             argv[0]);
      return 0;
    }
    if (c) continue;
    const char *o = lo[idx].name;
    if (!strcmp(o, "nx")) p.nx = atoi(optarg);
    else if (!strcmp(o, "ny")) p.ny = atoi(optarg);
    else if (!strcmp(o, "dx")) p.dx = (float)atof(optarg);
    else if (!strcmp(o, "dt")) p.dt = (float)atof(optarg);
    else if (!strcmp(o, "Du")) p.Du = (float)atof(optarg);
    else if (!strcmp(o, "Dv")) p.Dv = (float)atof(optarg);
    else if (!strcmp(o, "F")) p.feed = (float)atof(optarg);
    else if (!strcmp(o, "k")) p.kill = (float)atof(optarg);
    else if (!strcmp(o, "steps")) steps = atoi(optarg);
    else if (!strcmp(o, "headless")) headless = 1;
    else if (!strcmp(o, "stride")) { stride = atoi(optarg); if (stride < 1) stride = 1; }
    else if (!strcmp(o, "seed")) p.seed = (unsigned)strtoul(optarg, NULL, 10);
    else if (!strcmp(o, "dump")) dump = optarg;
  }
  if (p.nx == 0) p.nx = 128; /* headless default :293-296 (no terminal to size from) */
  if (p.ny == 0) p.ny = 128;
  if (steps == 0) steps = 1000; /* the reference runs until 'q'; a host without a UI needs an end */
  tau_gs *gs;
  TAU_OR_DIE(tau_gs_create(&p, 0, 0, p.ny, NULL, &gs));
  TAU_OR_DIE(tau_gs_init(gs));
  const size_t n = (size_t)p.nx * p.ny;
  float *u = (float *)malloc(n * 4), *v = (float *)malloc(n * 4);
  const double t0 = cli_now();
  for (int step = 1; step <= steps; ++step) {
    TAU_OR_DIE(tau_gs_step(gs, 1)); /* step_kernel + swaps :323-329 */
    if (!headless && step % stride == 0 && (step % (stride * 250) == 0 || step == steps)) {
      TAU_OR_DIE(tau_gs_download(gs, u, v));
      float vmin = 1e9f, vmax = -1e9f;
      for (size_t i = 0; i < n; ++i) { if (v[i] < vmin) vmin = v[i]; if (v[i] > vmax) vmax = v[i]; }
      printf("step=%d dt=%.3f F=%.4f k=%.4f Du=%.3f Dv=%.3f v in [%.4f, %.4f]\n", step, p.dt, p.feed,
             p.kill, p.Du, p.Dv, vmin, vmax);
    }
  }
  TAU_OR_DIE(tau_gs_sync(gs));
  const double secs = cli_now() - t0;
  printf("%d steps of %dx%d in %.3f s: %.1f Mcell-updates/s\n", steps, p.nx, p.ny, secs,
         (double)steps * n / secs / 1e6);
  if (dump) {
    TAU_OR_DIE(tau_gs_download(gs, u, v));
    void *planes[2] = {u, v};
    cli_dump(dump, 2, 4, p.nx, p.ny, 1, tau_gs_steps_done(gs), (double)steps * p.dt, planes);
  }
  free(u);
  free(v);
  TAU_OR_DIE(tau_gs_destroy(gs));
  return 0;
}
