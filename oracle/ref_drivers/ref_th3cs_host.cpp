// ref_th3cs_host.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
// Runs the reference's `.4spl` exporter th3cs.cu — its whole main(): k_build_solid_mask, k_init, 4 x k_step
// per frame with the host-side d_tau controller, k_schlieren_export, the host min/max + palette-index loop,
// the palette — ON THE CPU.  REF_SRC is th3cs.cu after (1) sed of the three hard-coded sizes
// (hp.nx/ny/nz = 64, frames = 60) into REF_N / REF_FRAMES and (2) tests/hostemu/hostemu_build.py's mechanical
// rewriting of `k<<<g, b, s>>>(args)` into emulator launches; it is compiled against tests/hostemu/hostemu.h
// (fibers with real __syncthreads semantics; __expf/__logf are libm's) and deleted after compilation.
// The four functions of the missing 4splat.c are provided here and simply capture what main() hands them.
#include "hostemu.h"

#include <stdint.h>
#include <stdio.h>

static int g_n = 16, g_frames = 2;
#define REF_N g_n
#define REF_FRAMES g_frames
static FILE *ref_null_file(const char *, const char *) { return tmpfile(); }
#define fopen ref_null_file
#define main ref_th3cs_main
#include REF_SRC
#undef main
#undef fopen

static Splat4DHeader g_header;
static std::vector<Splat4D> g_palette;
static std::vector<uint8_t> g_indices;
static int g_written = 0;

extern "C" {
Splat4D create_splat4D(float mu_x, float sigma_x, float mu_y, float sigma_y, float mu_z, float sigma_z, float mu_t,
                       float sigma_t, float r, float g, float b, float alpha) {
  return Splat4D{mu_x, sigma_x, mu_y, sigma_y, mu_z, sigma_z, mu_t, sigma_t, r, g, b, alpha};
}
Splat4DHeader create_splat4DHeader(uint32_t width, uint32_t height, uint32_t depth, uint32_t frames, uint32_t pSize,
                                   uint32_t flags) {
  Splat4DHeader h;
  memset(&h, 0, sizeof(h));
  h.width = width; h.height = height; h.depth = depth; h.frames = frames; h.pSize = pSize; h.flags = flags;
  return h;
}
Splat4DVideo create_splat4DVideo(Splat4DHeader header, Splat4D *splats, uint64_t *idxs) {
  g_header = header;
  g_palette.assign(splats, splats + header.pSize);
  const size_t n = (size_t)header.width * header.height * header.depth * header.frames;
  g_indices.resize(n);
  for (size_t i = 0; i < n; ++i) g_indices[i] = (uint8_t)idxs[i];   // main() stores 0..255 in each uint64_t
  Splat4DVideo v;
  memset(&v, 0, sizeof(v));
  v.header = header;
  return v;
}
bool write_splat4DVideo(FILE *, Splat4DVideo *) { g_written++; return true; }

// runs the exporter for an n^3 grid and `frames` frames; header6 = {width, height, depth, frames, pSize, flags};
// palette: pSize*12 floats; indices: frames*n^3 bytes.  Returns main()'s exit code.
int ref_th3cs_host_run(int n, int frames, uint32_t *header6, float *palette, uint8_t *indices) {
  g_n = n;
  g_frames = frames;
  g_written = 0;
  const int rc = ref_th3cs_main();
  if (rc != 0 || g_written != 1) return rc ? rc : -1;
  header6[0] = g_header.width; header6[1] = g_header.height; header6[2] = g_header.depth;
  header6[3] = g_header.frames; header6[4] = g_header.pSize; header6[5] = g_header.flags;
  memcpy(palette, g_palette.data(), g_palette.size() * sizeof(Splat4D));
  memcpy(indices, g_indices.data(), g_indices.size());
  return 0;
}
}
