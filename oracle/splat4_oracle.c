/* splat4_oracle.c — CPU restatement of the host-side frame quantisation and palette of the reference's
 * `.4spl` exporter (th3cs.cu main).  TEST INFRASTRUCTURE ONLY: only tests/ may call this.
 *
 * Follows th3cs.cu line by line: palette :1136-1144, per-frame min/max :1199-1205, palette index
 * :1207-1222.  Pinning: the exporter's 4splat.c is missing from the reference repository, so the reference
 * binary cannot be built — but its whole main() runs on the CPU: oracle/ref_drivers/ref_th3cs_host.cpp
 * compiles th3cs.cu for the host (sizes made variable by sed, <<< >>> launches rewritten mechanically by
 * tests/hostemu/hostemu_build.py, kernels executed by the fiber emulator of tests/hostemu/hostemu.h) with four stub
 * 4splat functions that capture what main() hands them.  Its output (tests/golden/th3cs_ref_host.npz:
 * 24^3, 48 frames, generator tests/golden/make_golden_host.py) is reproduced index for index, frame for
 * frame by hyp3d_oracle.c (k_step + d_tau controller + vis mode 8) followed by this file
 * (tests/test_oracle_cpu.py::test_th3cs_golden_pins_the_3d_and_4spl_oracles).
 */
#include <math.h>
#include <stdint.h>

void oracle_4spl_palette(float *palette, int pSize) {                 /* :1136-1144 */
  for (int i = 0; i < pSize; i++) {
    float t_val = (float)i / (pSize - 1.0f);
    float r = fminf(1.0f, t_val * 2.5f);
    float g = fmaxf(0.0f, fminf(1.0f, t_val * 2.5f - 0.5f));
    float b = fmaxf(0.0f, fminf(1.0f, t_val * 2.5f - 1.5f));
    float *e = palette + 12 * (long)i;                                 /* create_splat4D(0,1,0,1,0,1,0,1,r,g,b,1) */
    e[0] = 0; e[1] = 1; e[2] = 0; e[3] = 1; e[4] = 0; e[5] = 1; e[6] = 0; e[7] = 1;
    e[8] = r; e[9] = g; e[10] = b; e[11] = 1.0f;
  }
}

int oracle_4spl_index(float norm) {                                    /* :1215-1218 */
  norm = powf(norm, 0.65f);
  int pIdx = (int)(norm * 255.0f);
  if (pIdx > 255) pIdx = 255;
  if (pIdx < 0) pIdx = 0;
  return pIdx;
}

void oracle_4spl_frame_indices(const float *sch, long n, uint8_t *out, float minmax[2]) {   /* :1199-1222 */
  float min_val = 1e30f, max_val = -1e30f;
  for (long i = 0; i < n; ++i) {
    min_val = fminf(min_val, sch[i]);
    max_val = fmaxf(max_val, sch[i]);
  }
  float range = fmaxf(max_val - min_val, 1e-12f);
  for (long i = 0; i < n; ++i) {
    float norm = (sch[i] - min_val) / range;
    out[i] = (uint8_t)oracle_4spl_index(norm);
  }
  minmax[0] = min_val;
  minmax[1] = max_val;
}

/* exhaustive: is oracle_4spl_index non-decreasing over every float in [0, 1]?  (the product's device
 * quantiser rests on that; ~1.07e9 powf calls) -> number of violations */
long oracle_4spl_index_monotone_violations(uint32_t first_bits, uint32_t last_bits) {
  long bad = 0;
  int prev = -1;
  for (uint32_t b = first_bits; b <= last_bits; ++b) {
    union { uint32_t u; float f; } x = {b};
    int k = oracle_4spl_index(x.f);
    if (k < prev) ++bad;
    prev = k;
    if (b == 0xffffffffu) break;
  }
  return bad;
}
