/* cli_common.h — shared helpers of the C host programs: raw field dump + wall clock.
 * Dump format (new; the reference has no state output, SURVEY.md 5 "checkpoint/resume: none"):
 *   char magic[8] = "TAUDUMP1"; int32 nplanes, elem_bytes, d0, d1, d2; int64 step; double t;
 *   then nplanes planes of d0*d1*d2 elements, in the reference's plane order. */
#ifndef TAU_CLI_COMMON_H
#define TAU_CLI_COMMON_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "tau_b200.h"

static inline double cli_now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static inline int cli_dump(const char *path, int nplanes, int elem_bytes, int d0, int d1, int d2,
                           long long step, double t, void *const *planes) {
  FILE *f = fopen(path, "wb");
  if (!f) {
    fprintf(stderr, "cannot open %s for writing\n", path);
    return -1;
  }
  int32_t hdr[5] = {nplanes, elem_bytes, d0, d1, d2};
  int64_t st = step;
  fwrite("TAUDUMP1", 1, 8, f);
  fwrite(hdr, sizeof(hdr), 1, f);
  fwrite(&st, sizeof(st), 1, f);
  fwrite(&t, sizeof(t), 1, f);
  for (int p = 0; p < nplanes; ++p) fwrite(planes[p], (size_t)elem_bytes, (size_t)d0 * d1 * d2, f);
  fclose(f);
  return 0;
}
#endif
