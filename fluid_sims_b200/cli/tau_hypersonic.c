/* tau_hypersonic — C host of the reference's CPU solver tau_hypersonic.c (BASELINE config 1) over libtau_b200.so.
 * The reference binary takes no arguments: it opens a raylib window (W = H = 300, :12-13), steps
 * STEPS_PER_FRAME = 2 times per frame (:18, :710-713), renders one of four views (M cycles view_mode :707-708;
 * 2 = "speed mode", README.md:1) and prints t on screen.  Without raylib this host keeps that frame loop and makes
 * the interactive state explicit:
 *   --nx N --ny N      grid (default 300 300; config 1 uses 256 256)
 *   --frames F         frames to run (default 100), --steps-per-frame K (default 2, :18), or --steps S in total
 *   --view MODE        0 log(rho), 1 log(p), 2 speed, 3 schlieren — or the names rho | p | speed | schlieren
 *   --speed-mode       == --view 2
 *   --ppm FILE         the last frame's pixels as a binary PPM (what UpdateTexture would have shown, :788)
 *   --dump FILE        the state planes rho, mx, my, E (TAUDUMP1, cli_common.h) for parity checks
 * The update path (step_physics :500-674) and the render loops (:713-786) run on the device (fp64, 0 ulp). */
#include "cli_common.h"

static int parse_view(const char *s) {
  if (!strcmp(s, "rho")) return 0;
  if (!strcmp(s, "p")) return 1;
  if (!strcmp(s, "speed")) return 2;
  if (!strcmp(s, "schlieren")) return 3;
  return atoi(s);
}

int main(int argc, char **argv) {
  int W = 300, H = 300, frames = 100, spf = 2, steps = -1, view = 0;
  const char *ppm = NULL, *dump = NULL;
  for (int i = 1; i < argc; ++i) {
    const char *a = argv[i];
    const int has = i + 1 < argc;
    if (!strcmp(a, "--nx") && has) { W = atoi(argv[++i]); continue; }
    if (!strcmp(a, "--ny") && has) { H = atoi(argv[++i]); continue; }
    if (!strcmp(a, "--frames") && has) { frames = atoi(argv[++i]); continue; }
    if (!strcmp(a, "--steps-per-frame") && has) { spf = atoi(argv[++i]); continue; }
    if (!strcmp(a, "--steps") && has) { steps = atoi(argv[++i]); continue; }
    if (!strcmp(a, "--view") && has) { view = parse_view(argv[++i]); continue; }
    if (!strcmp(a, "--speed-mode")) { view = 2; continue; }
    if (!strcmp(a, "--ppm") && has) { ppm = argv[++i]; continue; }
    if (!strcmp(a, "--dump") && has) { dump = argv[++i]; continue; }
    if (!strcmp(a, "-h") || !strcmp(a, "--help")) {
      printf("Usage: %s [--nx N] [--ny N] [--frames F] [--steps-per-frame K] [--steps S] [--view MODE | --speed-mode]\n"
             "          [--ppm FILE] [--dump FILE]\n", argv[0]);
      return 0;
    }
    fprintf(stderr, "Unknown or incomplete argument: %s\n", a);
    return 1;
  }
  if (view < 0 || view > 3 || spf < 1 || frames < 0) {
    fprintf(stderr, "Invalid --view / --steps-per-frame / --frames\n");
    return 1;
  }
  if (steps >= 0) { frames = steps / spf; }
  tau_hypc *s;
  TAU_OR_DIE(tau_hypc_create(W, H, 0, NULL, &s));
  TAU_OR_DIE(tau_hypc_init(s)); /* init_sim :702 */
  uint32_t *rgba = (uint32_t *)malloc((size_t)W * H * sizeof(uint32_t));
  double mm[2] = {0, 0}, t = 0;
  const double w0 = cli_now();
  for (int f = 0; f < frames; ++f) {
    TAU_OR_DIE(tau_hypc_step(s, spf));                      /* :710-713 */
    if (ppm) TAU_OR_DIE(tau_hypc_render(s, view, rgba, mm)); /* :713-788 */
  }
  if (steps >= 0 && steps % spf) TAU_OR_DIE(tau_hypc_step(s, steps % spf));
  TAU_OR_DIE(tau_hypc_sync(s));
  const double secs = cli_now() - w0;
  TAU_OR_DIE(tau_hypc_clock(s, &t, NULL));
  static const char *modes[4] = {"log(rho)", "log(p)", "speed", "schlieren"};
  const long long done = tau_hypc_steps_done(s);
  printf("t = %.4f\n%s\n", t, modes[view]); /* the two text overlays of the reference's frame, :796-802 */
  printf("%lld steps of %dx%d in %.3f s -> %.1f Mcell-updates/s\n", done, W, H, secs, (double)done * W * H / secs / 1e6);
  if (ppm) {
    if (frames == 0) TAU_OR_DIE(tau_hypc_render(s, view, rgba, mm));
    FILE *fp = fopen(ppm, "wb");
    if (!fp) { fprintf(stderr, "cannot open %s for writing\n", ppm); return 1; }
    fprintf(fp, "P6\n%d %d\n255\n", W, H);
    for (size_t i = 0; i < (size_t)W * H; ++i) {
      const unsigned char px[3] = {(unsigned char)(rgba[i] & 255u), (unsigned char)((rgba[i] >> 8) & 255u),
                                   (unsigned char)((rgba[i] >> 16) & 255u)};
      fwrite(px, 1, 3, fp);
    }
    fclose(fp);
    printf("view %d: value range [%.6g, %.6g] -> %s\n", view, mm[0], mm[1], ppm);
  }
  if (dump) {
    const size_t n = (size_t)W * H;
    double *u = (double *)malloc(4 * n * sizeof(double));
    double *pl[4] = {u, u + n, u + 2 * n, u + 3 * n};
    TAU_OR_DIE(tau_hypc_download(s, pl, NULL));
    void *planes[4] = {pl[0], pl[1], pl[2], pl[3]};
    cli_dump(dump, 4, 8, W, H, 1, done, t, planes);
    free(u);
  }
  free(rgba);
  TAU_OR_DIE(tau_hypc_destroy(s));
  return 0;
}
