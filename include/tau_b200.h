/* tau_b200.h — C-ABI of the B200-native per-timestep update path.
 *
 * The reference (seanwevans/fluid-sims @ 78429ea) has no FFI/plugin boundary: each solver's
 * per-step host sequence is inline code in main().  This header introduces that
 * boundary: one opaque handle per solver, plain pointers and sizes only, no C++/torch types.
 * Every entry point cites the reference host sequence it replaces (file:line into the reference).
 *
 * Conventions
 *   - all functions return 0 on success or a negative errno-style code; tau_last_error() returns
 *     the message for the calling thread.  The CLI binaries wrap calls in TAU_OR_DIE, which prints
 *     the message to stderr and exits — the reference's CK()/gpuAssert() policy
 *     (tau_hypersonic_cuda.cu:69-75, tau_gray_scott.cu:29-41).
 *   - "planes" are row-major SoA arrays in the reference's layout and order.
 *   - *_step() only enqueues work on the handle's stream; nothing inside it synchronises with the
 *     host.  *_download(), *_clock() and *_sync() synchronise.
 *   - slab mode (multi-GPU): a handle owns rows [y_begin, y_begin+ny_local) of the global grid plus
 *     `halo` ghost rows on each side, which the caller fills between steps (NCCL send/recv on the
 *     device pointers returned by *_halo_ptrs).  Single-GPU: y_begin=0, ny_local=ny.
 *   - there is no CPU fallback: creating a handle without a CUDA device fails with -ENODEV.
 */
#ifndef TAU_B200_H
#define TAU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAU_B200_ABI_VERSION 1

const char *tau_last_error(void);
int tau_abi_version(void);
/* number of visible CUDA devices (0 when none / no driver); never fails */
int tau_device_count(void);

/* ------------------------------------------------------------------------------------------ */
/* Gray-Scott reaction-diffusion (reference: tau_gray_scott.cu)                                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct tau_gs tau_gs;

/* mirrors `struct Params` tau_gray_scott.cu:43-61 (simulation fields only) */
typedef struct tau_gs_params {
  int nx, ny;
  float dx, dt, Du, Dv, feed, kill;
  unsigned seed;
} tau_gs_params;

/* defaults of tau_gray_scott.cu:43-61; nx = ny = 128 (headless default, :293-296) */
void tau_gs_default_params(tau_gs_params *p);
/* host-side initial pattern, identical to init_pattern() tau_gray_scott.cu:173-204 */
void tau_gs_init_pattern(float *u, float *v, int nx, int ny, unsigned seed);

/* replaces the cudaMalloc block tau_gray_scott.cu:301-306.  stream may be NULL (own stream). */
int tau_gs_create(const tau_gs_params *p, int device, int y_begin, int ny_local, void *stream,
                  tau_gs **out);
/* init_pattern + H2D (tau_gray_scott.cu:299-309); in slab mode uploads this rank's rows */
int tau_gs_init(tau_gs *h);
/* inject caller state: u, v are full ny_local x nx host planes */
int tau_gs_upload(tau_gs *h, const float *u, const float *v);
/* THE hot path: nsteps x { step_kernel; swap } of tau_gray_scott.cu:321-329, without the per-step
 * cudaDeviceSynchronize.  Single-GPU handles apply the periodic wrap themselves; slab handles
 * (ny_local < ny) expect the ghost rows to have been exchanged before every step. */
int tau_gs_step(tau_gs *h, int nsteps);
int tau_gs_download(tau_gs *h, float *u, float *v);
int tau_gs_sync(tau_gs *h);
/* device pointers of the CURRENT state planes, pointing at ghost row -1 (pitch nx floats,
 * ny_local+2 rows). */
int tau_gs_device_planes(tau_gs *h, float **u, float **v);
long long tau_gs_steps_done(tau_gs *h);
/* kernels launched by this handle so far (for bench.py's gpu_launches) */
long long tau_gs_launch_count(tau_gs *h);
/* device time of the most recent tau_gs_step() call in ms (CUDA events on the handle's stream) */
int tau_gs_last_step_ms(tau_gs *h, float *ms);
int tau_gs_destroy(tau_gs *h);


/* ------------------------------------------------------------------------------------------ */
/* 2-D hypersonic compressible flow (reference: tau_hypersonic_cuda.cu)                         */
/* ------------------------------------------------------------------------------------------ */
typedef struct tau_hyp2d tau_hyp2d;

/* mirrors `struct SimConfig` tau_hypersonic_cuda.cu:37-50 */
typedef struct tau_hyp2d_config {
  double gamma, cfl, visc_nu, visc_rho, visc_e, inflow_mach;
  double geom_x0, geom_cy, geom_Rb, geom_Rn, geom_theta;
  int steps_per_frame;
} tau_hyp2d_config;

#define TAU_F32 0
#define TAU_F64 1

/* default_config() :1394-1409; the reference derives the geometry from its compile-time H, here
 * from the runtime H (W, H are runtime: the reference's `#define W 8192 / H 1024` :28-29 become
 * arguments). */
void tau_hyp2d_default_config(tau_hyp2d_config *c, int W, int H);
/* the validation block of parse_args() :1545-1637 (same messages, returned not printed) */
int tau_hyp2d_validate_config(const tau_hyp2d_config *c);
/* replaces main()'s allocation block :1748-1811.  dtype: TAU_F32 (BASELINE "fp32") or TAU_F64
 * (the reference's arithmetic).  Slab: rows [y_begin, y_begin+h_local), 2 ghost rows each side. */
int tau_hyp2d_create(const tau_hyp2d_config *c, int W, int H, int dtype, int device, int y_begin,
                     int h_local, void *stream, tau_hyp2d **out);
/* k_init :740-770 (+ the first max-wavespeed scan) */
int tau_hyp2d_init(tau_hyp2d *h);
/* inject caller state: 4 host planes rho,mx,my,E of h_local x W elements in the handle's dtype and
 * (optional) the body mask */
int tau_hyp2d_upload(tau_hyp2d *h, const void *const planes[4], const uint8_t *mask);
/* THE hot path: nsteps x the loop body :1833-1889 (inflow column, max wavespeed, dt, predict,
 * x/y fluxes, update+diffusion, swap, sim_t += dt) as one fused kernel per step, dt on device. */
int tau_hyp2d_step(tau_hyp2d *h, int nsteps);
/* sim_t (:1888) and the dt of the most recent step; synchronises */
int tau_hyp2d_clock(tau_hyp2d *h, double *sim_t, double *dt_last);
int tau_hyp2d_download(tau_hyp2d *h, void *const planes[4], uint8_t *mask);
/* enqueue-only forms (host buffers must be pinned and stay valid until tau_hyp2d_sync): with two
 * handles on two streams a frame loop overlaps the upload of frame i+1 with the download of frame i */
int tau_hyp2d_upload_async(tau_hyp2d *h, const void *const planes[4], const uint8_t *mask);
int tau_hyp2d_download_async(tau_hyp2d *h, void *const planes[4], uint8_t *mask);
/* multi-GPU (peers attached): upload the next frame's owned rows with NO host exchange — the hand-over of the
 * ghost rows, the all-reduce(max) of the wavespeed and the barrier between frames run on the device, in stream
 * order (see hypersonic2d.cu).  Every rank calls it for the same frame; the body mask stays as uploaded. */
int tau_hyp2d_upload_peers_async(tau_hyp2d *h, const void *const planes[4]);
int tau_hyp2d_sync(tau_hyp2d *h);
/* slab plumbing: device pointers of the current planes (4 contiguous planes of (h_local+4) x W,
 * starting at ghost row -2), the mask (same row layout) and the max-wavespeed scalar the next
 * step will read (all-reduce it with MAX across ranks before tau_hyp2d_step). */
int tau_hyp2d_device_state(tau_hyp2d *h, void **planes, uint8_t **mask, double **maxspeed_slot);
/* Multi-GPU without per-step host work (one process per GPU, one box): export CUDA-IPC handles of
 * this handle's planes and control block (3 x 64 bytes), all-gather them across ranks, attach.
 * Afterwards the step kernel pushes its boundary rows into the neighbours' ghost rows over NVLink
 * and tau_hyp2d_step(h, n) may advance many steps per call; the wavespeed all-reduce and the step
 * barrier run on the device.  After init/upload the caller exchanges ghost rows + the wavespeed
 * once on the host side and calls tau_hyp2d_peers_ready(). */
int tau_hyp2d_ipc_export(tau_hyp2d *h, void *out, size_t out_bytes);
int tau_hyp2d_ipc_attach(tau_hyp2d *h, int rank, int world, const void *all_handles,
                         const int *h_locals);
int tau_hyp2d_peers_ready(tau_hyp2d *h);
/* unmap the peers' memory again: every rank detaches, the processes meet at a barrier, then the handles are
 * destroyed (CUDA wants importers to close before the exporter frees).  tau_hyp2d_destroy falls back to it. */
int tau_hyp2d_ipc_detach(tau_hyp2d *h);
/* Diagnostics of the device-side exchange: average microseconds per step spent waiting for the
 * peers' messages, computing, and between steps; out[3] = steps counted since peers_ready. */
int tau_hyp2d_peer_timing(tau_hyp2d *h, double out[4]);
/* Render pass of the frame loop :1892-1926 (k_render_vals, k_reduce_minmax, k_compute_inv_range,
 * k_render_pixels) on the current state.  view_mode as the reference's `view_mode` (:1195-1228):
 * 0 log rho, 1 log p, 2 speed, 3 log |grad rho|, 4 asinh(vorticity), 5 Mach, 6 log(p/rho).
 * rgba: h_local x W pixels, byte order R,G,B,A (uchar4 of :688-690), body cells (110,110,110).
 * Slabs: tau_hyp2d_render_minmax gives the slab's extrema — reduce them (min/max) across ranks and
 * hand the result to tau_hyp2d_render_pixels; modes 3/4 read the ghost rows, which the device-side
 * exchange keeps current (with the host-driven exchange, exchange once more after the last step).
 * tau_hyp2d_render = both passes on one handle. */
int tau_hyp2d_render_minmax(tau_hyp2d *h, int view_mode, double minmax[2]);
int tau_hyp2d_render_pixels(tau_hyp2d *h, int view_mode, const double minmax[2], uint32_t *rgba);
int tau_hyp2d_render(tau_hyp2d *h, int view_mode, uint32_t *rgba, double minmax_out[2]);
/* grid, dtype, slab and config of a handle (any out pointer may be NULL) */
int tau_hyp2d_describe(tau_hyp2d *h, int *W, int *H, int *dtype, int *y_begin, int *h_local,
                       tau_hyp2d_config *cfg);
/* sizes of the work-item table(s) and persistent grids the next step uses (host bookkeeping):
 * out = {items, grid CTAs, pair-kernel items, items left to the production kernel, pair grid, rest grid} */
int tau_hyp2d_work_items(tau_hyp2d *h, int out[6]);
/* restore sim_t and the step counter after tau_hyp2d_upload (checkpoint/resume) */
int tau_hyp2d_set_clock(tau_hyp2d *h, double sim_t, long long steps_done);

/* The reference's regression snapshot (`struct RegressionSnapshot`,
 * tau_hypersonic_cuda_tests.cu:20-36) and its text file (:84-125): compute_snapshot :143-176 on the
 * handle's state (host, sequential, the reference's summation order), write / read in the
 * reference's format, compare with the reference's tolerances (:527-557; returns the number of
 * failed checks, names in tau_last_error()).  A slab handle yields its partial sums. */
typedef struct tau_hyp2d_snapshot_t {
  int steps, fluid_cells;
  double sum_rho, sum_mx, sum_my, sum_E, min_rho, min_p, max_mach, checksum_rho, checksum_mx, checksum_E;
} tau_hyp2d_snapshot_t;
int tau_hyp2d_snapshot(tau_hyp2d *h, tau_hyp2d_snapshot_t *out);
int tau_hyp2d_snapshot_write(const char *path, const tau_hyp2d_snapshot_t *s);
int tau_hyp2d_snapshot_read(const char *path, tau_hyp2d_snapshot_t *s);
int tau_hyp2d_snapshot_compare(const tau_hyp2d_snapshot_t *current, const tau_hyp2d_snapshot_t *expected);
/* Raw SoA checkpoint of a handle (one file per slab) and bit-identical resume; the reference has
 * no state output.  _info reads the header so that a matching handle can be created. */
int tau_hyp2d_checkpoint_save(tau_hyp2d *h, const char *path);
int tau_hyp2d_checkpoint_info(const char *path, int *W, int *H, int *dtype, int *y_begin, int *h_local,
                              long long *steps, double *sim_t, tau_hyp2d_config *cfg);
int tau_hyp2d_checkpoint_load(tau_hyp2d *h, const char *path);
/* rows each warp marches per work item (tuning; --tile-by analogue of :1641-1685) */
int tau_hyp2d_set_seg_rows(tau_hyp2d *h, int rows);
/* the row schedule behind the work-item table as a pure host function (no device needed): layers of
 * layer_h[i] rows starting at layer_y[i]; seg_rows == 0 selects the guided (tall-to-short) schedule */
int tau_hyp2d_plan_layers(int h_local, int nstrips, int resident_warps, int seg_rows, int taper_k,
                          int min_rows, int max_rows, int *layer_y, int *layer_h, int cap);
/* the height in use (chosen by a wave model at the first step unless set explicitly) */
int tau_hyp2d_get_seg_rows(tau_hyp2d *h);
long long tau_hyp2d_steps_done(tau_hyp2d *h);
/* which step kernel the handle launches: 0 hyp2d_step (one column per lane; fp64, narrow or non-TMA grids),
 * 1 hyp2d_step_pair + hyp2d_step (two launches), 2 hyp2d_step_fused (two columns per lane, one launch) */
int tau_hyp2d_kernel_mode(tau_hyp2d *h);
long long tau_hyp2d_launch_count(tau_hyp2d *h);
int tau_hyp2d_last_step_ms(tau_hyp2d *h, float *ms);
int tau_hyp2d_destroy(tau_hyp2d *h);

/* ------------------------------------------------------------------------------------------ */
/* 3-D hypersonic flow with vibrational relaxation (reference: tau_hypersonic_3d_cuda.cu)       */
/* ------------------------------------------------------------------------------------------ */
typedef struct tau_hyp3d tau_hyp3d;

/* mirrors `struct Params` tau_hypersonic_3d_cuda.cu:21-42, plus the initial log-time clock
 * (t0, d_tau0) that main() hard-codes (:1635-1636) */
typedef struct tau_hyp3d_params {
  int nx, ny, nz;
  float dx, dy, dz;
  float cfl, u_ref, R, gamma_floor, Twall, tau_vib, theta_v;
  float sdf_cx, sdf_cy, sdf_cz, sdf_r;
  float inflow_r, inflow_p, inflow_u, inflow_v, inflow_w;
  int sponge_n;
  float sponge_strength;
  int sponge_out_n;
  float sponge_out_strength;
  float t0, d_tau0;
} tau_hyp3d_params;

/* main()'s constants :1531-1557 for an nx x ny x nz grid (dx = 1/nx ...; the reference runs 64^3) */
void tau_hyp3d_default_params(tau_hyp3d_params *p, int nx, int ny, int nz);
/* replaces the allocation block :1572-1601 incl. k_build_solid_mask.  Slab: z-planes
 * [z_begin, z_begin+nz_local) with 3 ghost planes on either side (z is periodic: a ring). */
int tau_hyp3d_create(const tau_hyp3d_params *p, int device, int z_begin, int nz_local, void *stream,
                     tau_hyp3d **out);
/* reset_sim -> k_init :1605, clock := (t0, d_tau0) */
int tau_hyp3d_init(tau_hyp3d *h);
/* inject caller state: 6 host planes xi, phix, phiy, phiz, lam, zet of nz_local*ny*nx floats,
 * index (z*ny+y)*nx+x (:152); clock2 = {t, d_tau} or NULL for (t0, d_tau0) */
int tau_hyp3d_upload(tau_hyp3d *h, const float *const planes[6], const float *clock2);
/* THE hot path: nsteps x the loop body :1679-1712 (log-time clock, k_step, d_tau controller, swap)
 * with clock and controller on the device.  Single-GPU handles only. */
int tau_hyp3d_step(tau_hyp3d *h, int nsteps);
/* slab protocol, one step: [exchange ghost planes] step_begin [all-reduce MAX of *maxs] step_end */
int tau_hyp3d_step_begin(tau_hyp3d *h);
int tau_hyp3d_step_end(tau_hyp3d *h);
/* t, d_tau for the NEXT step, and dt / max wavespeed sum of the last one (HUD :1763-1768) */
int tau_hyp3d_clock(tau_hyp3d *h, float *t, float *d_tau, float *dt_last, float *maxs_last);
int tau_hyp3d_download(tau_hyp3d *h, float *const planes[6], uint8_t *solid);
int tau_hyp3d_sync(tau_hyp3d *h);
/* k_vis :800-905 — the scalar field the volume renderer consumes: nz_local*ny*nx floats, solid
 * cells 0.  mode = VisMode :784-794 (0 |grad rho|, 1 log(1+rho), 2 log(1+p), 3 |u|, 4 Mach,
 * 5 |curl u|, 6 div u, 7 Q criterion; 8 = th3cs.cu's k_schlieren_export :641-673).  Slab handles:
 * exchange the ghost planes first. */
int tau_hyp3d_vis(tau_hyp3d *h, int mode, float *out);
/* One frame of th3cs.cu's export loop :1193-1222 — k_schlieren_export (vis mode 8: mode 0 with the
 * exporter's divisions), min/max, and per voxel (int)(powf((v-min)/range, 0.65f)*255) clamped to 0..255 —
 * all on the device; nz_local*ny*nx palette indices come back (1 B/voxel instead of the reference's
 * 4 B/voxel D2H + host loop).  Identical to the host loop by construction: the 255 steps of that function
 * are located with the host's powf (tau_4spl_index_thresholds).  minmax (optional): the frame's range. */
int tau_hyp3d_export_frame(tau_hyp3d *h, uint8_t *indices, float minmax[2]);
void tau_4spl_index_thresholds(float thr[255]);
/* `.4spl` container (th3cs.cu:17-62, 1226-1240; reader viewer.html:67-96; the writer library 4splat.c is
 * missing from the reference, see csrc/splat4.cu for what is pinned and what is our choice).
 * palette: pSize entries of 12 floats (Splat4D); indices: frames*depth*height*width bytes, x fastest. */
void tau_4spl_thermal_palette(float *palette, int pSize);
int tau_4spl_write(const char *path, int width, int height, int depth, int frames, int pSize, unsigned flags,
                   const float *palette, const uint8_t *indices);
/* dims = {width, height, depth, frames, pSize, flags}; verifies length and checksum */
int tau_4spl_info(const char *path, int dims[6]);
/* device pointers: current state (6 contiguous planes of (nz_local+6)*ny*nx floats, starting at
 * ghost plane -3) and the max-wavespeed accumulator of the running step.  Write GHOST planes only through this pointer (the
 * ring exchange): the handle keeps decoded primitives of its own planes next to the state (tau_hyp3d_upload / _init replace
 * those) */
int tau_hyp3d_device_state(tau_hyp3d *h, float **planes, float **maxs);
long long tau_hyp3d_steps_done(tau_hyp3d *h);
/* Multi-GPU from ONE process (new: the reference has none; SURVEY 8(b) create(cfg, dims, ngpus)): one z-slab handle per
 * device, z periodic = a ring.  Per step the three boundary planes of every field go to the neighbours' ghost planes by
 * cudaMemcpyPeerAsync, every device runs its step kernel, the max wavespeed sums are folded on the host (one
 * synchronisation per step; the reference's loop :1684-1700 has two) and every device's controller commits the same clock.
 * planes cover the whole grid in the reference layout.  devices = NULL: devices 0 .. ngpus-1. */
typedef struct tau_hyp3d_group tau_hyp3d_group;
int tau_hyp3d_group_create(const tau_hyp3d_params *p, int ngpus, const int *devices, tau_hyp3d_group **out);
int tau_hyp3d_group_size(tau_hyp3d_group *g);
int tau_hyp3d_group_member(tau_hyp3d_group *g, int i, tau_hyp3d **h, int *z_begin, int *nz_local);
int tau_hyp3d_group_init(tau_hyp3d_group *g);
int tau_hyp3d_group_upload(tau_hyp3d_group *g, const float *const planes[6], const float *clock2);
int tau_hyp3d_group_step(tau_hyp3d_group *g, int nsteps);
int tau_hyp3d_group_clock(tau_hyp3d_group *g, float *t, float *d_tau, float *dt_last, float *maxs_last);
int tau_hyp3d_group_download(tau_hyp3d_group *g, float *const planes[6], uint8_t *solid);
long long tau_hyp3d_group_steps_done(tau_hyp3d_group *g);
int tau_hyp3d_group_destroy(tau_hyp3d_group *g);
long long tau_hyp3d_launch_count(tau_hyp3d *h);
int tau_hyp3d_last_step_ms(tau_hyp3d *h, float *ms);
int tau_hyp3d_destroy(tau_hyp3d *h);

/* ------------------------------------------------------------------------------------------ */
/* 2-D weakly-compressible SPH (reference: tau_sph.cu)                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct tau_sph tau_sph;

/* simulation fields of `struct Params` tau_sph.cu:49-85 (same defaults) */
typedef struct tau_sph_params {
  int N;
  float boxX, boxY;
  float dTau, t0, CFL;
  float rho0, c0, gammaEOS, hMul, viscAlpha, gravity;
  int rain, useVisc, useGrav;
  int viscSub;
  int useXSPH;
  float xsphEps;
  int seed;
} tau_sph_params;

void tau_sph_default_params(tau_sph_params *p);
/* reset_particles() tau_sph.cu:493-510 on the host: pos_xy / vel_xy are N x (x, y) floats */
void tau_sph_reset_particles(const tau_sph_params *p, float *pos_xy, float *vel_xy);
/* replaces the allocation block :561-565 + ensure_cell_buffers :512-540 + derived constants
 * :573-578 */
int tau_sph_create(const tau_sph_params *p, int device, void *stream, tau_sph **out);
/* reset_particles + H2D :567-571; clock := (t0, tau = 0) */
int tau_sph_init(tau_sph *h);
int tau_sph_upload(tau_sph *h, const float *pos_xy, const float *vel_xy);
/* THE hot path: nframes x the doStep block :663-722 — per sub-step: cell keys, radix sort, cell
 * ranges, density/pressure, forces + integration, [XSPH], [rain], tau-clock.  No host sync. */
int tau_sph_step(tau_sph *h, int nframes);
/* Multi-GPU (new: the reference has none).  Particle state is replicated on every rank; the work of
 * a sub-step is sharded by sorted-slot range, i.e. by stripes of hash bins balanced by particle
 * count.  Per sub-step: shard_substep_begin (sort, densities of own + ghost rows, forces +
 * integration of the own chunk into sorted copies), all-gather the two sorted-copy arrays
 * (tau_sph_shard_buffers: N+pad x 2 floats each, `chunk` slots per rank), shard_substep_end
 * (back to original order, rain, tau-clock). */
int tau_sph_shard_config(tau_sph *h, int rank, int world);
int tau_sph_shard_buffers(tau_sph *h, float **sxy_new, float **svel_new, int *chunk);
int tau_sph_shard_substep_begin(tau_sph *h);
int tau_sph_shard_substep_end(tau_sph *h);
int tau_sph_clock(tau_sph *h, float *t, float *tau, long long *step);
/* state in ORIGINAL particle order (any pointer may be NULL) */
int tau_sph_download(tau_sph *h, float *pos_xy, float *vel_xy, float *s, float *press);
/* render pass of the frame loop :747-755 (k_clear_grid + k_rasterize :357-374 + D2H): particle
 * counts on the terminal's half-block raster, grid2[sy * W + cx], sy in [0, 2H), y flipped */
int tau_sph_rasterize(tau_sph *h, int W, int H, int *grid2);
/* (sort key, particle index) pairs of the last sub-step's radix sort, in sorted order.  key = grid row * (Gx * subx) +
 * key column, where subx = tau_sph_subx() key columns cut each grid cell (the reference's cell, edge 2h) so that a
 * neighbourhood row is a narrower slot range; key column / subx is the reference's grid_x (:141-148) exactly, i.e. the
 * order refines the reference's cell order */
int tau_sph_download_sort(tau_sph *h, unsigned *keys, unsigned *vals);
/* the sort on its own: N keys (< number of cells rounded up to a power of two) -> sorted keys and
 * the stable permutation */
int tau_sph_sort_pairs(tau_sph *h, const unsigned *keys_in, unsigned *keys_out, unsigned *vals_out);
int tau_sph_grid(tau_sph *h, int *Gx, int *Gy, float *cell, float *hh, float *mass);
int tau_sph_subx(tau_sph *h);
int tau_sph_sync(tau_sph *h);
long long tau_sph_substeps_done(tau_sph *h);
long long tau_sph_launch_count(tau_sph *h);
int tau_sph_last_step_ms(tau_sph *h, float *ms);
int tau_sph_destroy(tau_sph *h);

/* ------------------------------------------------------------------------------------------ */
/* 2-D viscous Burgers (reference: tau_burgers.cu) — SURVEY.md 8(f) rank 3                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct tau_burgers tau_burgers;

/* simulation fields of `struct Params` tau_burgers.cu:53-90 (same defaults) */
typedef struct tau_burgers_params {
  int nx, ny;
  float dx, dy;
  float nu, u0;
  float amp, bsig, swirl, rc, offx, offy, asym;
  float CFL, tau0, t0, dtau;
  int muscl, visc_substeps;
  int colehopf, ck;
  float ca;
} tau_burgers_params;

void tau_burgers_default_params(tau_burgers_params *p);
/* initialize_host :250-304 (host): phi_u, phi_v of ny*nx floats, index j*nx+i; ny = 1 with colehopf */
void tau_burgers_init_host(const tau_burgers_params *p, float *phi_u, float *phi_v);
/* replaces device_alloc :315-322 (the four flux planes and the block-max buffer are not needed) */
int tau_burgers_create(const tau_burgers_params *p, int device, void *stream, tau_burgers **out);
/* initialize_host + H2D :652-661; clock := (t0, tau0) */
int tau_burgers_init(tau_burgers *h);
/* inject caller state; clock2 = {t, tau} or NULL for (t0, tau0) */
int tau_burgers_upload(tau_burgers *h, const float *phi_u, const float *phi_v, const float *clock2);
/* THE hot path: nsteps x { do_step :677-718; tau += dtau; t *= expf(dtau) :768-769 } — one convection
 * kernel + visc_substeps viscosity kernels per step, dt and the log-time clock on the device.
 * viscosity_step's in-place update (a data race in the reference) is evaluated as the Jacobi update. */
int tau_burgers_step(tau_burgers *h, int nsteps);
int tau_burgers_clock(tau_burgers *h, float *t, float *tau, float *dt_last);
int tau_burgers_download(tau_burgers *h, float *phi_u, float *phi_v);
/* colehopf_relL2 :720-737 on the current state (handles created with colehopf = 1) */
int tau_burgers_colehopf_error(tau_burgers *h, double *rel_l2);
int tau_burgers_sync(tau_burgers *h);
long long tau_burgers_steps_done(tau_burgers *h);
long long tau_burgers_launch_count(tau_burgers *h);
int tau_burgers_last_step_ms(tau_burgers *h, float *ms);
int tau_burgers_destroy(tau_burgers *h);

/* ------------------------------------------------------------------------------------------ */
/* 2-D shallow water (reference: tau_shallow_water.cu) — SURVEY.md 8(f) rank 3                 */
/* NOT YET RUN ON HARDWARE (written after the round-1 GPU budget was spent; see NEXT.md)       */
/* ------------------------------------------------------------------------------------------ */
typedef struct tau_sw tau_sw;

/* simulation fields of `struct Params` tau_shallow_water.cu:52-89 (same defaults) */
typedef struct tau_sw_params {
  int nx, ny;
  float dx, dy;
  float g, f0, nu, H0;
  float bumpAmp, bumpSigma, CFL;
  float offx, offy, asym, swirl, swirlRc;
  float tau0, t0, dtau;
} tau_sw_params;

void tau_sw_default_params(tau_sw_params *p);
/* initialize_host :238-277 (host): sigma = log h, u, v of ny*nx floats, index j*nx+i */
void tau_sw_init_host(const tau_sw_params *p, float *sigma, float *u, float *v);
/* replaces device_alloc :280-298 (the six flux planes and the block-max buffer are not needed) */
int tau_sw_create(const tau_sw_params *p, int device, void *stream, tau_sw **out);
/* initialize_host + H2D :640-655; clock := (t0, tau0) */
int tau_sw_init(tau_sw *h);
/* inject caller state; clock2 = {t, tau} or NULL for (t0, tau0) */
int tau_sw_upload(tau_sw *h, const float *sigma, const float *u, const float *v, const float *clock2);
/* THE hot path: nsteps x { do_step :669-705; tau += dtau; t *= expf(dtau) :767-768 } — one fused
 * flux + update kernel (+ one viscosity kernel when nu > 0) per step; dt and the log-time clock on the
 * device.  viscosity_uv's in-place update (a data race in the reference) is the Jacobi update here. */
int tau_sw_step(tau_sw *h, int nsteps);
int tau_sw_clock(tau_sw *h, float *t, float *tau, float *dt_last);
int tau_sw_download(tau_sw *h, float *sigma, float *u, float *v);
int tau_sw_sync(tau_sw *h);
long long tau_sw_steps_done(tau_sw *h);
long long tau_sw_launch_count(tau_sw *h);
int tau_sw_last_step_ms(tau_sw *h, float *ms);
int tau_sw_destroy(tau_sw *h);

/* ---------------------------------------------------------------------------------------------
 * SPH sharded by hash-bin stripes with ghost-particle exchange and migration
 * Replaces the same sub-step body (tau_sph.cu:676-721) as tau_sph_step, for `world` GPUs: rank r holds
 * only the particles of its stripe of grid rows (stripes balanced by particle count) plus, per
 * sub-step, the ghost particles of one cell row on either side.  Global particle ids (= index in the
 * reference's arrays) travel with the particles.  A sub-step is four enqueue-only phases with the
 * neighbour exchanges between them (fixed-capacity messages with their counts in a header; all counts
 * the kernels need are device-resident: no host synchronisation):
 *   phase 0 -> exchange buffers 0..3 (migrants + boundary rows: ids, pos, vel) -> phase 1
 *           -> exchange buffers 4..7 (rho, p/rho^2 of the boundary rows)      -> phase 2
 *           -> [XSPH only: exchange buffers 0..3 again (integrated boundary rows)] -> phase 3
 * "exchange": send buffer b+0 to rank-1 and b+1 to rank+1, receive b+2 from rank-1 and b+3 from rank+1
 * (whole buffers; a chain — y is not periodic).  Results are bit-identical to the single-GPU handle.
 * ------------------------------------------------------------------------------------------- */
typedef struct tau_sph_stripe tau_sph_stripe;
int tau_sph_stripe_create(const tau_sph_params *p, int device, void *stream, int rank, int world, tau_sph_stripe **out);
/* the WHOLE particle set (N x (x, y), reference order); the rank keeps its stripe; clock := (t0, 0) */
int tau_sph_stripe_upload(tau_sph_stripe *h, const float *pos_xy, const float *vel_xy);
int tau_sph_stripe_xbuf(tau_sph_stripe *h, int which, void **ptr, long long *words);
int tau_sph_stripe_phase(tau_sph_stripe *h, int phase);
/* re-balancing between sub-steps: hist_begin -> all-reduce(SUM) of *rows ints at *hist (device) -> hist_apply */
int tau_sph_stripe_hist_begin(tau_sph_stripe *h, void **hist, int *rows);
int tau_sph_stripe_hist_apply(tau_sph_stripe *h);
/* n_own, n_ghost, error bits, message high-water mark, capacity, message capacity, first row, end row (synchronises) */
int tau_sph_stripe_status(tau_sph_stripe *h, int out[8]);
/* the rank's owned particles in local order: ids[n], pos / vel (n x 2), s, press (n); *n_out = n <= capacity */
int tau_sph_stripe_download(tau_sph_stripe *h, unsigned *ids, float *pos_xy, float *vel_xy, float *s, float *press, int *n_out);
int tau_sph_stripe_clock(tau_sph_stripe *h, float *t, float *tau, long long *step);
int tau_sph_stripe_capacity(tau_sph_stripe *h, int *cap, int *xcap);
int tau_sph_stripe_sync(tau_sph_stripe *h);
long long tau_sph_stripe_substeps_done(tau_sph_stripe *h);
long long tau_sph_stripe_launch_count(tau_sph_stripe *h);
int tau_sph_stripe_destroy(tau_sph_stripe *h);

/* ---------------------------------------------------------------------------------------------
 * 2-D hypersonic on several GPUs of one box from ONE process ("create(cfg, dims, ngpus)")
 * Replaces the same step loop (tau_hypersonic_cuda.cu:1833-1889) and allocation block (:1748-1819) as
 * tau_hyp2d_*, y-slab decomposed (chain, 2 ghost rows): one slab handle per device, peers
 * reached through cudaDeviceEnablePeerAccess.  Per step every device runs ONE kernel that also pushes
 * its boundary rows into the neighbours' ghost rows over NVLink and sends its max wavespeed to every
 * peer with one release store (all-reduce(max) + step barrier); the host only enqueues.  Results are
 * bit-identical to the single-GPU handle.  planes / mask / rgba cover the whole H x W grid.
 * devices == NULL: devices 0 .. ngpus-1.  ngpus == 1 is allowed (one full-domain handle).
 * ------------------------------------------------------------------------------------------- */
typedef struct tau_hyp2d_group tau_hyp2d_group;
int tau_hyp2d_group_create(const tau_hyp2d_config *cfg, int W, int H, int dtype, int ngpus, const int *devices,
                           tau_hyp2d_group **out);
int tau_hyp2d_group_size(tau_hyp2d_group *g);
/* the i-th slab handle and its rows [y_begin, y_begin + h_local) — for per-slab calls (snapshot, timing) */
int tau_hyp2d_group_member(tau_hyp2d_group *g, int i, tau_hyp2d **h, int *y_begin, int *h_local);
int tau_hyp2d_group_init(tau_hyp2d_group *g);
int tau_hyp2d_group_upload(tau_hyp2d_group *g, const void *const planes[4], const uint8_t *mask);
int tau_hyp2d_group_step(tau_hyp2d_group *g, int nsteps);
int tau_hyp2d_group_sync(tau_hyp2d_group *g);
int tau_hyp2d_group_clock(tau_hyp2d_group *g, double *sim_t, double *dt_last);
int tau_hyp2d_group_download(tau_hyp2d_group *g, void *const planes[4], uint8_t *mask);
int tau_hyp2d_group_render(tau_hyp2d_group *g, int view_mode, uint32_t *rgba, double minmax_out[2]);
long long tau_hyp2d_group_launch_count(tau_hyp2d_group *g);
int tau_hyp2d_group_destroy(tau_hyp2d_group *g);

/* ---------------------------------------------------------------------------------------------
 * tau_hypersonic (CPU reference solver, tau_hypersonic.c — BASELINE config 1: 256 x 256, "speed mode")
 * Replaces: `static void step_physics(void)` (tau_hypersonic.c:500-674) on the file-static
 * `Cons U[W*H]`, `mask[]`, `sim_t` (:38-43), `init_sim` (:450-475) and the render loops of main()
 * (:713-786).  fp64 on the device; built without FMA contraction: results equal the reference's own
 * object code (gcc -O3, x86-64) to 0 ulp.  Planes are the SoA view of the reference's AoS state:
 * rho, mx, my, E, each H x W doubles, index y*W+x (:46).  W and H are run-time here (the reference
 * fixes them with #define W/H, :12-13).
 * ------------------------------------------------------------------------------------------- */
typedef struct tau_hypc tau_hypc;
int tau_hypc_create(int W, int H, int device, void *stream, tau_hypc **out);
/* init_sim :450-475 on the host (pure arithmetic, no device needed) */
void tau_hypc_init_host(int W, int H, double *rho, double *mx, double *my, double *E, uint8_t *mask);
int tau_hypc_init(tau_hypc *h);
int tau_hypc_upload(tau_hypc *h, const double *const planes[4], const uint8_t *mask, double sim_t);
/* THE hot path: nsteps x step_physics — two kernels per step, dt (compute_dt :477-498) and sim_t on the device */
int tau_hypc_step(tau_hypc *h, int nsteps);
int tau_hypc_clock(tau_hypc *h, double *sim_t, double *dt_last);
int tau_hypc_download(tau_hypc *h, double *const planes[4], uint8_t *mask);
/* view_mode as the reference's `view_mode` (:44): 0 log rho, 1 log p, 2 speed, 3 schlieren; rgba: W*H pixels
 * (r | g<<8 | b<<16 | 255<<24, the byte order of the reference's `pixels`); minmax_out may be NULL */
int tau_hypc_render(tau_hypc *h, int view_mode, uint32_t *rgba, double minmax_out[2]);
int tau_hypc_sync(tau_hypc *h);
long long tau_hypc_steps_done(tau_hypc *h);
long long tau_hypc_launch_count(tau_hypc *h);
int tau_hypc_last_step_ms(tau_hypc *h, float *ms);
int tau_hypc_destroy(tau_hypc *h);

#ifdef __cplusplus
}
#endif

#define TAU_OR_DIE(call)                                          \
  do {                                                            \
    int _rc = (call);                                             \
    if (_rc != 0) {                                               \
      fprintf(stderr, "%s (%s)\n", tau_last_error(), #call);      \
      exit(1);                                                    \
    }                                                             \
  } while (0)

#endif /* TAU_B200_H */
