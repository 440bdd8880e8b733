/* tau3d — C host of the 3-D hypersonic solver over libtau_b200.so.  The reference binary has no
 * CLI at all (everything is hard-coded in main(), tau_hypersonic_3d_cuda.cu:1531-1557); this host
 * keeps those values as defaults and adds --n N (grid n^3), --frames N (2 steps per frame like the
 * reference loop :1679), --dump FILE and --gpus N (z-slabs on N devices from this one process, tau_hyp3d_group_*: the dump is
 * identical to --gpus 1).  The raylib volume viewer is not part of the update path; the
 * HUD line (:1763-1768) is printed instead. */
#include "cli_common.h"

int main(int argc, char **argv) {
  int n = 64, frames = 100, gpus = 1;
  const char *dump = NULL;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--n") && i + 1 < argc) n = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--frames") && i + 1 < argc) frames = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--dump") && i + 1 < argc) dump = argv[++i];
    else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
    else { fprintf(stderr, "Usage: %s [--n N] [--frames N] [--dump FILE] [--gpus N]\n", argv[0]); return 1; }
  }
  tau_hyp3d_params p;
  tau_hyp3d_default_params(&p, n, n, n);
  tau_hyp3d_group *sim;   /* one z-slab handle per device; --gpus 1 is the single-GPU handle */
  TAU_OR_DIE(tau_hyp3d_group_create(&p, gpus, NULL, &sim));
  TAU_OR_DIE(tau_hyp3d_group_init(sim));
  const int steps_per_frame = 2;
  float t = 0, d_tau = 0, dt = 0, maxs = 0;
  const double t0 = cli_now();
  for (int f = 0; f < frames; ++f) {
    TAU_OR_DIE(tau_hyp3d_group_step(sim, steps_per_frame));
    if ((f + 1) % 50 == 0 || f + 1 == frames) {
      TAU_OR_DIE(tau_hyp3d_group_clock(sim, &t, &d_tau, &dt, &maxs));   /* synchronises */
      printf("t=%.6g dt=%.3e d_tau=%.3e maxs=%.4g\n", t, dt, d_tau, maxs);
    }
  }
  const double secs = cli_now() - t0, steps = (double)frames * steps_per_frame;
  printf("%.0f steps of %d^3 on %d GPU%s in %.3f s: %.1f Mcell-updates/s\n", steps, n, gpus, gpus == 1 ? "" : "s", secs,
         steps * n * n * n / secs / 1e6);
  if (dump) {
    const size_t N = (size_t)n * n * n;
    float *planes[6];
    for (int k = 0; k < 6; ++k) planes[k] = (float *)malloc(N * 4);
    TAU_OR_DIE(tau_hyp3d_group_download(sim, planes, NULL));
    cli_dump(dump, 6, 4, n, n, n, tau_hyp3d_group_steps_done(sim), (double)t, (void *const *)planes);
    for (int k = 0; k < 6; ++k) free(planes[k]);
  }
  TAU_OR_DIE(tau_hyp3d_group_destroy(sim));
  return 0;
}
