/* tau_sph — C host of the SPH solver over libtau_b200.so.  Keeps the reference's options
 * (parse_args, tau_sph.cu:394-491: -n/--n, -b/--box WxH, -t/--dTau, -r/--rho0, -c/--c0, -g/--gamma,
 * -f/--CFL, -h/--hMul, -v/--visc, -y/--gravity, -p/--fps, -F/--fpscap, -S/--stride, -s/--seed,
 * -R/--rain, -H/--headless, -B/--halfblocks, -k/--visc_substeps, -m/--muscl (= XSPH), -x/--xsph_eps)
 * and the frame loop (:663-722).  The reference's headless mode never terminates; this host adds
 * --frames N (and --dump FILE) and prints the reference's status line (:785-790). */
#include <getopt.h>

#include "cli_common.h"

int main(int argc, char **argv) {
  tau_sph_params P;
  tau_sph_default_params(&P);
  int frames = 1000, stride = 1;
  const char *dump = NULL;
  const struct option lo[] = {
      {"n", required_argument, 0, 'n'},     {"box", required_argument, 0, 'b'},   {"dTau", required_argument, 0, 't'},
      {"rho0", required_argument, 0, 'r'},  {"c0", required_argument, 0, 'c'},    {"gamma", required_argument, 0, 'g'},
      {"CFL", required_argument, 0, 'f'},   {"hMul", required_argument, 0, 'h'},  {"visc", required_argument, 0, 'v'},
      {"gravity", required_argument, 0, 'y'}, {"fps", required_argument, 0, 'p'}, {"fpscap", required_argument, 0, 'F'},
      {"stride", required_argument, 0, 'S'}, {"seed", required_argument, 0, 's'}, {"rain", no_argument, 0, 'R'},
      {"headless", no_argument, 0, 'H'},    {"halfblocks", no_argument, 0, 'B'},  {"visc_substeps", required_argument, 0, 'k'},
      {"muscl", no_argument, 0, 'm'},       {"xsph_eps", required_argument, 0, 'x'}, {"frames", required_argument, 0, 1},
      {"dump", required_argument, 0, 2},    {0, 0, 0, 0}};
  int c;
  while ((c = getopt_long(argc, argv, "n:b:t:r:c:g:f:h:v:y:p:F:S:s:RHBk:mx:", lo, NULL)) != -1) {
    switch (c) {
      case 'n': P.N = atoi(optarg); break;
      case 'b': sscanf(optarg, "%fx%f", &P.boxX, &P.boxY); break;
      case 't': P.dTau = (float)atof(optarg); break;
      case 'r': P.rho0 = (float)atof(optarg); break;
      case 'c': P.c0 = (float)atof(optarg); break;
      case 'g': P.gammaEOS = (float)atof(optarg); break;
      case 'f': P.CFL = (float)atof(optarg); break;
      case 'h': P.hMul = (float)atof(optarg); break;
      case 'v': P.viscAlpha = (float)atof(optarg); break;
      case 'y': P.gravity = (float)atof(optarg); P.useGrav = (P.gravity != 0.f); break;
      case 'S': stride = atoi(optarg); if (stride < 1) stride = 1; break;
      case 's': P.seed = atoi(optarg); break;
      case 'R': P.rain = 1; break;
      case 'k': P.viscSub = atoi(optarg); if (P.viscSub < 1) P.viscSub = 1; break;
      case 'm': P.useXSPH = 1; break;
      case 'x': P.xsphEps = (float)atof(optarg); P.useXSPH = (P.xsphEps > 0.f) || P.useXSPH; break;
      case 1: frames = atoi(optarg); break;
      case 2: dump = optarg; break;
      default: break; /* -p -F -H -B: UI only */
    }
  }
  tau_sph *s;
  TAU_OR_DIE(tau_sph_create(&P, 0, NULL, &s));
  TAU_OR_DIE(tau_sph_init(s));
  const double t0 = cli_now();
  float t = 0, tau = 0;
  long long step = 0;
  for (int f = 0; f < frames; ++f) {
    TAU_OR_DIE(tau_sph_step(s, 1));
    if ((f + 1) % (100 * stride) == 0 || f + 1 == frames) {
      TAU_OR_DIE(tau_sph_clock(s, &t, &tau, &step));
      printf("step %lld  t=%.6f  tau=%.6f\n", step, t, tau);
    }
  }
  TAU_OR_DIE(tau_sph_sync(s));
  const double secs = cli_now() - t0;
  printf("%lld sub-steps of %d particles in %.3f s: %.1f Mparticle-updates/s\n", tau_sph_substeps_done(s),
         P.N, secs, (double)tau_sph_substeps_done(s) * P.N / secs / 1e6);
  if (dump) {
    float *pos = (float *)malloc((size_t)P.N * 8), *vel = (float *)malloc((size_t)P.N * 8);
    TAU_OR_DIE(tau_sph_download(s, pos, vel, NULL, NULL));
    void *planes[2] = {pos, vel};
    cli_dump(dump, 2, 4, 2, P.N, 1, step, (double)t, planes);
    free(pos);
    free(vel);
  }
  TAU_OR_DIE(tau_sph_destroy(s));
  return 0;
}
