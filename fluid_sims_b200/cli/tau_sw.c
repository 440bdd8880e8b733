/* tau_sw — C host of the shallow-water solver over libtau_b200.so.  Keeps the reference's long options
 * (parse_args, tau_shallow_water.cu:142-236: --nx --ny --dx --dy --g --f0 --nu --H0 --amp --bsig --CFL
 * --steps --tau0 --t0 --dtau --headless --stride --fps --offx --offy --asym --swirl --rc, -h) and the
 * headless loop of main() (:759-790) with its report; the ncurses renderer is not part of the update
 * path (without --headless a one-line summary is printed instead of a frame).  Additive: --dump FILE.
 * Not yet run on hardware (see NEXT.md). */
#include <getopt.h>
#include <math.h>

#include "cli_common.h"

int main(int argc, char **argv) {
  tau_sw_params p;
  tau_sw_default_params(&p);
  int steps = 0, headless = 0, stride = 5;
  const char *dump = NULL;
  static const struct option lo[] = {
      {"nx", required_argument, 0, 0},   {"ny", required_argument, 0, 0},   {"dx", required_argument, 0, 0},
      {"dy", required_argument, 0, 0},   {"g", required_argument, 0, 0},    {"f0", required_argument, 0, 0},
      {"nu", required_argument, 0, 0},   {"H0", required_argument, 0, 0},   {"amp", required_argument, 0, 0},
      {"bsig", required_argument, 0, 0}, {"CFL", required_argument, 0, 0},  {"steps", required_argument, 0, 0},
      {"tau0", required_argument, 0, 0}, {"t0", required_argument, 0, 0},   {"dtau", required_argument, 0, 0},
      {"headless", no_argument, 0, 'H'}, {"stride", required_argument, 0, 'r'}, {"fps", required_argument, 0, 'f'},
      {"offx", required_argument, 0, 0}, {"offy", required_argument, 0, 0}, {"asym", required_argument, 0, 0},
      {"swirl", required_argument, 0, 0}, {"rc", required_argument, 0, 0},  {"dump", required_argument, 0, 0},
      {"help", no_argument, 0, 'h'},     {0, 0, 0, 0}};
  for (;;) {
    int idx = 0, c = getopt_long(argc, argv, "Hr:f:h", lo, &idx);
    if (c == -1) break;
    if (c == 'h') {
      printf("Usage: %s [options]  (options of tau_shallow_water.cu:118-141, plus --dump FILE)\n", argv[0]);
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
    F("dx", dx); F("dy", dy); F("g", g); F("f0", f0); F("nu", nu); F("H0", H0); F("amp", bumpAmp);
    F("bsig", bumpSigma); F("CFL", CFL); F("tau0", tau0); F("t0", t0); F("dtau", dtau); F("offx", offx);
    F("offy", offy); F("asym", asym); F("swirl", swirl); F("rc", swirlRc);
#undef F
    else if (!strcmp(o, "steps")) steps = atoi(optarg);
    else if (!strcmp(o, "dump")) dump = optarg;
  }
  if (steps == 0) steps = 2000; /* headless default :766; a host without a UI needs an end */
  tau_sw *s;
  TAU_OR_DIE(tau_sw_create(&p, 0, NULL, &s));
  TAU_OR_DIE(tau_sw_init(s));
  const double w0 = cli_now();
  int frames = 0;
  for (int step = 0; step < steps; ++step) {
    TAU_OR_DIE(tau_sw_step(s, 1)); /* do_step + clock :669-705, :767-768 */
    if (step % stride == 0) {
      frames++;
      if (!headless && (frames % 100 == 1 || step + 1 == steps)) {
        float t, tau, dt;
        TAU_OR_DIE(tau_sw_clock(s, &t, &tau, &dt));
        printf("step=%d t=%.4g tau=%.4g dt=%.3g\n", step, t, tau, dt);
      }
    }
  }
  TAU_OR_DIE(tau_sw_sync(s));
  const double secs = cli_now() - w0;
  printf("Headless benchmark (stride=%d):\n  Simulated steps: %d\n  Wall-clock: %d frames in %.3f s -> %.1f FPS\n",
         stride, steps, frames, secs, frames > 0 ? frames / secs : 0.0); /* the reference's report :781-786 */
  printf("  %.1f Mcell-updates/s\n", (double)steps * p.nx * p.ny / secs / 1e6);
  if (dump) {
    const size_t n = (size_t)p.nx * p.ny;
    float *a = (float *)malloc(n * 4), *u = (float *)malloc(n * 4), *v = (float *)malloc(n * 4), t;
    TAU_OR_DIE(tau_sw_download(s, a, u, v));
    TAU_OR_DIE(tau_sw_clock(s, &t, NULL, NULL));
    void *planes[3] = {a, u, v};
    cli_dump(dump, 3, 4, p.nx, p.ny, 1, tau_sw_steps_done(s), (double)t, planes);
    free(a);
    free(u);
    free(v);
  }
  TAU_OR_DIE(tau_sw_destroy(s));
  return 0;
}
