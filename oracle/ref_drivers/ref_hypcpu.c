/* ref_hypcpu.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 * Compiles the reference CPU solver (tau_hypersonic.c or tau_hypersonic_simd.c) through the fake
 * raylib header with main() renamed, and exposes its file-static step_physics()/init_sim() state.
 * REF_SRC is a build-time copy (oracle/_ref/.gen, deleted after the build, never committed) whose
 * two unguarded `#define W/H` lines were rewritten by sed to the requested grid size.
 * Used (a) as the reference CPU baseline timed by bench.py (`cpu_baseline.kind = "reference"`,
 * `bench.py --impl reference`) and (b) to pin oracle/hypcpu_oracle.c. */
#define _POSIX_C_SOURCE 200809L
#include <time.h>
#define main ref_hypcpu_main
#include REF_SRC
#undef main

void ref_hypcpu_dims(int *w, int *h) { *w = W; *h = H; }
void ref_hypcpu_init(void) { init_sim(); }
double ref_hypcpu_time(void) { return sim_t; }

/* run n steps of the reference's per-step host entry point; returns wall seconds spent in them */
double ref_hypcpu_steps(int n) {
  struct timespec a, b;
  clock_gettime(CLOCK_MONOTONIC, &a);
  for (int i = 0; i < n; ++i) {
    step_physics(); /* advances sim_t itself (tau_hypersonic.c:673) */
  }
  clock_gettime(CLOCK_MONOTONIC, &b);
  return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}

/* AoS state -> SoA planes (N doubles each) + mask */
void ref_hypcpu_get(double *rho, double *mx, double *my, double *E, unsigned char *m) {
  for (int i = 0; i < W * H; ++i) {
    rho[i] = U[i].rho; mx[i] = U[i].mx; my[i] = U[i].my; E[i] = U[i].E; m[i] = mask[i];
  }
}
void ref_hypcpu_set(const double *rho, const double *mx, const double *my, const double *E,
                    const unsigned char *m) {
  for (int i = 0; i < W * H; ++i) {
    U[i].rho = rho[i]; U[i].mx = mx[i]; U[i].my = my[i]; U[i].E = E[i]; mask[i] = m[i];
  }
}
