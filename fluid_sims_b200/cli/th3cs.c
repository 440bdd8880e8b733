/* th3cs — C host of the reference's headless `.4spl` exporter (th3cs.cu main :1062-1259) over
 * libtau_b200.so.  The reference hard-codes everything: 64^3, 60 frames, 4 steps per frame, a 256-entry
 * thermal palette, output "tau_hypersonic.4spl"; those are the defaults here, with --n N, --frames N,
 * --steps-per-frame N, --out FILE added.  Per frame the reference copies the schlieren volume to the host
 * and quantises it there; here the palette indices are computed on the device (tau_hyp3d_export_frame). */
#include "cli_common.h"

int main(int argc, char **argv) {
  int n = 64, frames = 60, steps_per_frame = 4, pSize = 256;
  const char *out = "tau_hypersonic.4spl";
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--n") && i + 1 < argc) n = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--frames") && i + 1 < argc) frames = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--steps-per-frame") && i + 1 < argc) steps_per_frame = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
    else { fprintf(stderr, "Usage: %s [--n N] [--frames N] [--steps-per-frame N] [--out FILE]\n", argv[0]); return 1; }
  }
  if (n < 8 || frames < 1 || steps_per_frame < 1) { fprintf(stderr, "th3cs: bad --n / --frames / --steps-per-frame\n"); return 1; }
  tau_hyp3d_params p;
  tau_hyp3d_default_params(&p, n, n, n);
  tau_hyp3d *sim;
  TAU_OR_DIE(tau_hyp3d_create(&p, 0, 0, n, NULL, &sim));
  TAU_OR_DIE(tau_hyp3d_init(sim));
  const size_t N = (size_t)n * n * n;
  uint8_t *indices = (uint8_t *)malloc(N * (size_t)frames);
  float *palette = (float *)malloc(sizeof(float) * 12 * (size_t)pSize);
  if (!indices || !palette) { fprintf(stderr, "th3cs: out of host memory\n"); return 1; }
  tau_4spl_thermal_palette(palette, pSize);                               /* :1136-1144 */
  printf("Running Hypersonic CFD for %d frames...\n", frames);            /* :1150 */
  const double w0 = cli_now();
  for (int f = 0; f < frames; ++f) {
    TAU_OR_DIE(tau_hyp3d_step(sim, steps_per_frame));                     /* :1155-1190 */
    TAU_OR_DIE(tau_hyp3d_export_frame(sim, indices + (size_t)f * N, NULL)); /* :1193-1222 */
    float t;
    TAU_OR_DIE(tau_hyp3d_clock(sim, &t, NULL, NULL, NULL));
    printf("Frame %d/%d processed (t=%.6f)\n", f + 1, frames, t);         /* :1223 */
  }
  const double secs = cli_now() - w0;
  printf("Writing simulation video to %s...\n", out);                     /* :1233 */
  TAU_OR_DIE(tau_4spl_write(out, n, n, n, frames, pSize, 0x0004u, palette, indices)); /* :1226-1231 */
  printf("Export Complete!\n");
  printf("%d frames (%d steps of %d^3) in %.3f s: %.1f Mcell-updates/s incl. export\n", frames,
         frames * steps_per_frame, n, secs, (double)frames * steps_per_frame * N / secs / 1e6);
  free(indices);
  free(palette);
  TAU_OR_DIE(tau_hyp3d_destroy(sim));
  return 0;
}
