// splat4.cu — ".4spl" volume-video container written by the reference's `th3cs` exporter (SURVEY.md 8(f)
// rank 4).  Host code only.
//
// The reference links an external `4splat.c` that is NOT in its repository (th3cs.cu:17-62 declares its
// API; reference Makefile:28-30,96-97), so the byte layout is inferred from the two places that pin it:
//   * `Splat4DHeader` th3cs.cu:27-33 — 32 bytes: magic u32, version u8[4], width, height, depth, frames,
//     pSize, flags (u32 each) — and `Splat4D` :22-25, 12 floats = 48 bytes per palette entry;
//   * the reader, viewer.html:67-96 — little-endian; width/height/depth/frames/pSize at byte offsets
//     8/12/16/20/24; palette at 32 (r, g, b at +32/+36/+40 of each entry); indices at 32 + 48*pSize,
//     ONE BYTE per voxel per frame (flags 0x0004 = "Float32 precision, 8-bit index width", th3cs.cu:1226).
// Not pinned by anything in the reference, and therefore our choice (documented in INTEGRATION.md): the
// magic (bytes "4SPL"), the version {1,0,0,0}, and the footer `Splat4DFooter` :41-45 written packed as
// checksum u32 (CRC-32 of everything before it), idxoffset u64 (= 32 + 48*pSize), end u32 ("LPS4").
// viewer.html ignores all three.
#include "common.cuh"
#include "../../include/tau_b200.h"

#include <stdio.h>
#include <stdlib.h>

namespace {
struct Crc32Table {
  uint32_t t[256];
  Crc32Table() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t x = i;
      for (int k = 0; k < 8; ++k) x = (x >> 1) ^ (0xEDB88320u & (0u - (x & 1u)));
      t[i] = x;
    }
  }
};
uint32_t crc32_update(uint32_t c, const unsigned char *p, size_t n) {
  static const Crc32Table tab;  // function-local static: initialised once, thread-safe
  for (size_t i = 0; i < n; ++i) c = tab.t[(c ^ p[i]) & 0xffu] ^ (c >> 8);
  return c;
}
void put32(unsigned char *b, uint32_t v) { for (int i = 0; i < 4; ++i) b[i] = (unsigned char)(v >> (8 * i)); }
void put64(unsigned char *b, uint64_t v) { for (int i = 0; i < 8; ++i) b[i] = (unsigned char)(v >> (8 * i)); }
uint32_t get32(const unsigned char *b) { return b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24); }
constexpr uint32_t MAGIC = 0x4C505334u;  // "4SPL"
constexpr uint32_t END = 0x3453504Cu;    // "LPS4"
}  // namespace

extern "C" {

// th3cs.cu:1136-1144: thermal map black -> red -> yellow -> white; create_splat4D(0,1,0,1,0,1,0,1,r,g,b,1)
void tau_4spl_thermal_palette(float *palette, int pSize) {
  for (int i = 0; i < pSize; ++i) {
    const float t_val = (float)i / (pSize - 1.0f);
    const float r = fminf(1.0f, t_val * 2.5f);
    const float g = fmaxf(0.0f, fminf(1.0f, t_val * 2.5f - 0.5f));
    const float b = fmaxf(0.0f, fminf(1.0f, t_val * 2.5f - 1.5f));
    float *e = palette + 12 * (size_t)i;  // mu_x, sigma_x, mu_y, sigma_y, mu_z, sigma_z, mu_t, sigma_t, r, g, b, alpha
    e[0] = 0.f; e[1] = 1.f; e[2] = 0.f; e[3] = 1.f; e[4] = 0.f; e[5] = 1.f; e[6] = 0.f; e[7] = 1.f;
    e[8] = r; e[9] = g; e[10] = b; e[11] = 1.0f;
  }
}

int tau_4spl_write(const char *path, int width, int height, int depth, int frames, int pSize, unsigned flags,
                   const float *palette, const uint8_t *indices) {
  TAU_REQUIRE(path && palette && indices, "tau_4spl_write: null argument");
  TAU_REQUIRE(width > 0 && height > 0 && depth > 0 && frames > 0 && pSize > 0 && pSize <= 256,
              "tau_4spl_write: bad shape %d x %d x %d x %d frames, palette %d", width, height, depth, frames, pSize);
  FILE *fp = fopen(path, "wb");
  TAU_REQUIRE(fp, "tau_4spl_write: cannot open %s", path);
  unsigned char hdr[32];
  put32(hdr, MAGIC);
  hdr[4] = 1; hdr[5] = 0; hdr[6] = 0; hdr[7] = 0;
  put32(hdr + 8, (uint32_t)width); put32(hdr + 12, (uint32_t)height); put32(hdr + 16, (uint32_t)depth);
  put32(hdr + 20, (uint32_t)frames); put32(hdr + 24, (uint32_t)pSize); put32(hdr + 28, flags);
  uint32_t crc = 0xffffffffu;
  bool ok = fwrite(hdr, 1, 32, fp) == 32;
  crc = crc32_update(crc, hdr, 32);
  for (int i = 0; ok && i < pSize * 12; ++i) {  // little-endian floats, whatever the host is
    uint32_t bits;
    memcpy(&bits, &palette[i], 4);
    unsigned char b[4];
    put32(b, bits);
    ok = fwrite(b, 1, 4, fp) == 4;
    crc = crc32_update(crc, b, 4);
  }
  const size_t nidx = (size_t)width * height * depth * frames;
  ok = ok && fwrite(indices, 1, nidx, fp) == nidx;
  crc = crc32_update(crc, indices, nidx);
  unsigned char ftr[16];
  put32(ftr, crc ^ 0xffffffffu);
  put64(ftr + 4, 32ull + 48ull * (uint64_t)pSize);
  put32(ftr + 12, END);
  ok = ok && fwrite(ftr, 1, 16, fp) == 16;
  ok = (fclose(fp) == 0) && ok;
  TAU_REQUIRE(ok, "tau_4spl_write: short write to %s", path);
  return TAU_OK;
}

// header fields + integrity check of a file tau_4spl_write produced: dims = {width, height, depth, frames, pSize, flags}
int tau_4spl_info(const char *path, int dims[6]) {
  TAU_REQUIRE(path && dims, "tau_4spl_info: null argument");
  FILE *fp = fopen(path, "rb");
  TAU_REQUIRE(fp, "tau_4spl_info: cannot open %s", path);
  unsigned char hdr[32];
  if (fread(hdr, 1, 32, fp) != 32 || get32(hdr) != MAGIC) {
    fclose(fp);
    tau_set_error("tau_4spl_info: %s is not a .4spl file", path);
    return TAU_ERR_INVALID;
  }
  for (int i = 0; i < 6; ++i) dims[i] = (int)get32(hdr + 8 + 4 * i);
  const size_t body = 48ull * dims[4] + (size_t)dims[0] * dims[1] * dims[2] * dims[3];
  uint32_t crc = crc32_update(0xffffffffu, hdr, 32);
  unsigned char buf[65536];
  size_t left = body;
  while (left) {
    const size_t want = left < sizeof(buf) ? left : sizeof(buf), got = fread(buf, 1, want, fp);
    if (got != want) break;
    crc = crc32_update(crc, buf, got);
    left -= got;
  }
  unsigned char ftr[16];
  const bool ok = left == 0 && fread(ftr, 1, 16, fp) == 16 && get32(ftr) == (crc ^ 0xffffffffu) && get32(ftr + 12) == END;
  fclose(fp);
  TAU_REQUIRE(ok, "tau_4spl_info: %s is truncated or its checksum does not match", path);
  return TAU_OK;
}

}  // extern "C"
