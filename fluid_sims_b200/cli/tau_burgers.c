/* tau_burgers — C host of the Burgers solver over libtau_b200.so.  Keeps the reference's long
 * options (parse_args, tau_burgers.cu:142-243: --nx --ny --dx --dy --nu --u0 --amp --bsig --swirl
 * --rc --offx --offy --asym --CFL --steps --tau0 --t0 --dtau --headless --stride --fps --halfblocks
 * --muscl --visc_substeps --colehopf --ck --ca, -h) and the headless loop of main() (:790-816) with
 * its report; the ncurses renderer is not part of the update path (without --headless a one-line
 * summary is printed every `stride` steps, with the Cole-Hopf error in --colehopf mode, as the
 * reference's status line shows it :568-580).  Additive: --dump FILE. */
#include <getopt.h>
#include <math.h>

#include "cli_common.h"

int main(int argc, char **argv) {
  tau_burgers_params p;
  tau_burgers_default_params(&p);
  int steps = 0, headless = 0, stride = 5;
  const char *dump = NULL;
  static const struct option lo[] = {
      {"nx", required_argument, 0, 0},     {"ny", required_argument, 0, 0},    {"dx", required_argument, 0, 0},
      {"dy", required_argument, 0, 0},     {"nu", required_argument, 0, 0},    {"u0", required_argument, 0, 0},
      {"amp", required_argument, 0, 0},    {"bsig", required_argument, 0, 0},  {"swirl", required_argument, 0, 0},
      {"rc", required_argument, 0, 0},     {"offx", required_argument, 0, 0},  {"offy", required_argument, 0, 0},
      {"asym", required_argument, 0, 0},   {"CFL", required_argument, 0, 0},   {"steps", required_argument, 0, 0},
      {"tau0", required_argument, 0, 0},   {"t0", required_argument, 0, 0},    {"dtau", required_argument, 0, 0},
      {"headless", no_argument, 0, 'H'},   {"stride", required_argument, 0, 'r'}, {"fps", required_argument, 0, 'f'},
      {"halfblocks", no_argument, 0, 0},   {"muscl", no_argument, 0, 0},       {"visc_substeps", required_argument, 0, 0},
      {"colehopf", no_argument, 0, 0},     {"ck", required_argument, 0, 0},    {"ca", required_argument, 0, 0},
      {"dump", required_argument, 0, 0},   {"help", no_argument, 0, 'h'},      {0, 0, 0, 0}};
  for (;;) {
    int idx = 0, c = getopt_long(argc, argv, "Hr:f:h", lo, &idx);
    if (c == -1) break;
    if (c == 'h') {
      printf("Usage: %s [options]  (options of tau_burgers.cu:104-139, plus --dump FILE)\n", argv[0]);
      return 0;
    }
    if (c == 'H') { headless = 1; continue; }
    if (c == 'r') { stride = atoi(optarg); if (stride < 1) stride = 1; continue; }
    if (c == 'f') continue;
    if (c) continue;
    const char *o = lo[idx].name;
#define F(name, field) else if (!strcmp(o, name)) p.field = (float)atof(optarg)
    if (!strcmp(o, "nx")) p.nx = atoi(optarg);
    else if (!strcmp(o, "ny")) p.ny = atoi(optarg);
    F("dx", dx); F("dy", dy); F("nu", nu); F("u0", u0); F("amp", amp); F("bsig", bsig); F("swirl", swirl);
    F("rc", rc); F("offx", offx); F("offy", offy); F("asym", asym); F("CFL", CFL); F("tau0", tau0); F("t0", t0);
    F("dtau", dtau); F("ca", ca);
#undef F
    else if (!strcmp(o, "steps")) steps = atoi(optarg);
    else if (!strcmp(o, "muscl")) p.muscl = 1;
    else if (!strcmp(o, "visc_substeps")) p.visc_substeps = atoi(optarg);
    else if (!strcmp(o, "colehopf")) p.colehopf = 1;
    else if (!strcmp(o, "ck")) p.ck = atoi(optarg);
    else if (!strcmp(o, "dump")) dump = optarg;
  }
  if (p.colehopf) p.ny = 1;          /* :649-650 */
  if (steps == 0) steps = 2000;      /* headless default :800; a host without a UI needs an end */
  tau_burgers *b;
  TAU_OR_DIE(tau_burgers_create(&p, 0, NULL, &b));
  TAU_OR_DIE(tau_burgers_init(b));
  const double w0 = cli_now();
  int frames = 0;
  for (int step = 0; step < steps; ++step) {
    TAU_OR_DIE(tau_burgers_step(b, 1)); /* do_step + clock :677-718, :768-769 */
    if (step % stride == 0) {
      frames++;
      if (!headless && (frames % 100 == 1 || step + 1 == steps)) {
        float t, tau, dt;
        TAU_OR_DIE(tau_burgers_clock(b, &t, &tau, &dt));
        if (p.colehopf) {
          double e;
          TAU_OR_DIE(tau_burgers_colehopf_error(b, &e));
          printf("step=%d t=%.4g tau=%.4g dt=%.3g relL2=%.3e\n", step, t, tau, dt, e);
        } else {
          printf("step=%d t=%.4g tau=%.4g dt=%.3g\n", step, t, tau, dt);
        }
      }
    }
  }
  TAU_OR_DIE(tau_burgers_sync(b));
  const double secs = cli_now() - w0;
  printf("Headless (stride=%d):\n  Steps: %d\n  Wall:  %d frames in %.3f s -> %.1f FPS\n", stride, steps, frames,
         secs, frames > 0 ? frames / secs : 0.0); /* the reference's report :818-823 */
  printf("  %.1f Mcell-updates/s\n", (double)steps * p.nx * p.ny / secs / 1e6);
  if (p.colehopf) { /* the harness's figure of merit :720-737, on the final state */
    double e;
    float t;
    TAU_OR_DIE(tau_burgers_colehopf_error(b, &e));
    TAU_OR_DIE(tau_burgers_clock(b, &t, NULL, NULL));
    printf("  Cole-Hopf: t=%.6g relL2=%.3e\n", t, e);
  }
  if (dump) {
    const size_t n = (size_t)p.nx * p.ny;
    float *u = (float *)malloc(n * 4), *v = (float *)malloc(n * 4), t;
    TAU_OR_DIE(tau_burgers_download(b, u, v));
    TAU_OR_DIE(tau_burgers_clock(b, &t, NULL, NULL));
    void *planes[2] = {u, v};
    cli_dump(dump, 2, 4, p.nx, p.ny, 1, tau_burgers_steps_done(b), (double)t, planes);
    free(u);
    free(v);
  }
  TAU_OR_DIE(tau_burgers_destroy(b));
  return 0;
}
