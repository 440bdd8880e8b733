// sph.cu — 2-D weakly-compressible SPH update path for sm_100a.  Replaces the per-sub-step host
// sequence of the reference `tau_sph` (tau_sph.cu:676-721):
//     k_clear_heads -> k_build_cells -> k_density_pressure_cell -> k_forces_cell -> k_integrate
//     [-> k_xsph_cell -> k_apply_xsph] [-> k_rain] -> tau-clock update
//
// What is restructured:
//   * neighbour search: the reference threads per-cell LINKED LISTS with atomicExch (:165-176), so
//     every neighbour visit is a dependent pointer chase and the summation order changes from run
//     to run.  Here particles get an integer cell key, a stable LSD RADIX SORT (hand-written:
//     per-warp digit histograms -> one scan -> stable scatter using warp match/ballot ranks) orders
//     (key, particle) pairs, and cell [start, end) ranges index position/velocity copies gathered
//     into sorted order.  The sort is bit-exact (== a stable CPU sort of the same keys) and makes the
//     whole step deterministic.
//   * per-particle sums: 8 lanes cooperate on one particle, striding over the candidate slot ranges
//     of its neighbourhood (three contiguous ranges, one per cell row, coalesced loads), and fold
//     their partial sums with warp shuffles.  The sort key cuts every cell into SUBX key columns, so
//     a range is 2.25 cells wide instead of 3 (see SUBX); key column / SUBX is the reference's grid_x.
//   * rho_j = expf(s_j) and p_j / rho_j^2, which the reference recomputes for every PAIR
//     (:242-244), are computed once per particle (identical values).
//   * k_integrate is fused into the force kernel (forces read the sorted copies, so updating the
//     state in place is safe).
//   * k_rain's write race (two spawns hitting the same particle, :389-391) is resolved
//     deterministically: the highest spawn index wins (what a sequential loop would do).
// Arithmetic keeps the reference's expression trees; this TU is compiled with the reference's
// -use_fast_math (reference Makefile:93-94).  The host step control (:663-722) has no device
// dependency and stays on the host, bit-identical.
#include "common.cuh"
#include "../../include/tau_b200.h"

#include <math.h>
#include <stdlib.h>
#include <new>
#include <random>
#include <vector>

namespace {

constexpr int SORT_WARPS = 8;            // warps per CTA in the sort kernels
constexpr int SORT_SEG = 1024;           // keys per warp segment
constexpr int SORT_MAX_BITS = 10;        // digit width upper bound (1024 bins: shared-memory histograms of the sort kernels)
constexpr int SORT_DEFAULT_BITS = 7;     // digit width used: measured at 2^21 particles (20-bit keys): 3 x 7 bits 1.570 ms per sub-step, 2 x 10 bits 1.608
constexpr int GROUP = 8;                 // lanes cooperating on one particle
// Sort-key columns per grid cell.  The reference's cell (edge 2h = the support radius, :512-540) makes the 3x3 search
// visit 9 cells = 36 h^2 for a support disc of 4 pi h^2: 35 % of the candidates pass the distance test.  The key keeps the
// reference's rows but cuts every cell into SUBX columns: a row of the neighbourhood is still ONE contiguous slot range,
// now SUBX key columns either side of the particle's own — (2 SUBX + 1) / SUBX = 2.25 cells wide instead of 3, 25 % fewer
// candidates.  SUBX is a power of two: x / (cell / SUBX) == SUBX * (x / cell) exactly, so key column >> log2(SUBX) IS the
// reference's grid_x (:141-148) and the order is a refinement of the reference's cell order.
#ifndef SPH_SUBX
#define SPH_SUBX 4
#endif
constexpr int SUBX = SPH_SUBX;
static_assert((SUBX & (SUBX - 1)) == 0 && SUBX >= 1, "power of two");

struct Consts {
  int N, Gx, Gy;   // Gx = KEY columns per row (SUBX per grid cell), Gy = grid rows
  float cellx;     // width of a key column = cell / SUBX
  float cell, mass, h, rho0, c0, gammaEOS, viscAlpha, gx, gy, boxX, boxY, alpha, xsphEps;
  int useVisc, useGrav;
  int k_begin, k_end;   // slots this launch works on (single GPU: [0, N))
  int write_state;      // forces kernel updates pos/vel itself (single GPU) or only the sorted copies
};

// grid_x / grid_y tau_sph.cu:141-157
__device__ __forceinline__ int grid_c(float x, float cell, int G) {
  int g = (int)floorf(x / cell);
  if (g < 0) g = 0;
  if (g >= G) g = G - 1;
  return g;
}
// W_cubic :105-116 (alpha = 10/(7 pi h^2) is evaluated in double by the reference: M_PI)
__device__ __forceinline__ float W_cubic(float r, float h, float alpha) {
  float q = r / h;
  if (q < 1.0f) {
    float q2 = q * q, q3 = q2 * q;
    return alpha * (1.f - 1.5f * q2 + 0.75f * q3);
  } else if (q < 2.0f) {
    float t = 2.f - q;
    return alpha * 0.25f * t * t * t;
  }
  return 0.f;
}
// gradW_cubic :118-133
__device__ __forceinline__ float2 gradW_cubic(float2 rij, float r, float h, float alpha) {
  if (r <= 1e-8f || r >= 2.0f * h) return make_float2(0.f, 0.f);
  float q = r / h;
  float dWdq;
  if (q < 1.0f) dWdq = alpha * (-3.0f * q + 2.25f * q * q);
  else {
    float t = 2.0f - q;
    dWdq = alpha * (-0.75f * t * t);
  }
  float invr = 1.0f / r;
  float dWdr = dWdq / h;
  return make_float2(dWdr * rij.x * invr, dWdr * rij.y * invr);
}

// Branch-free forms used by the pair loops.  A warp sweeps 32 (particle, candidate) pairs per
// iteration and ~35 % of them are inside the support, so a data-dependent branch around the
// kernel evaluation is taken by some lane in practically every iteration: it saves nothing and
// costs the BSSY/BRA/BSYNC bookkeeping (26 % of the density kernel's instructions, ncu).  Both
// polynomial pieces are evaluated with the reference's expression trees and selected; pairs
// outside the support contribute an exact +0.  `inv_h` is rcp.approx(h): under -use_fast_math
// `r / h` is div.approx = r * rcp.approx(h), so hoisting the reciprocal out of the loop keeps
// the bits.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float W_cubic_sel(float r, float inv_h, float alpha) {  // :105-116
  const float q = r * inv_h;
  const float q2 = q * q, q3 = q2 * q;
  const float w1 = alpha * (1.f - 1.5f * q2 + 0.75f * q3);
  const float t = 2.f - q;
  const float w2 = alpha * 0.25f * t * t * t;
  return q < 1.0f ? w1 : (q < 2.0f ? w2 : 0.f);
}

// gradW_cubic :118-133 with the reciprocal of h hoisted by the caller (same bits, see above)
__device__ __forceinline__ float2 gradW_cubic_h(float2 rij, float r, float h, float inv_h, float alpha) {
  if (r <= 1e-8f || r >= 2.0f * h) return make_float2(0.f, 0.f);
  float q = r * inv_h;
  float dWdq;
  if (q < 1.0f) dWdq = alpha * (-3.0f * q + 2.25f * q * q);
  else {
    float t = 2.0f - q;
    dWdq = alpha * (-0.75f * t * t);
  }
  float invr = 1.0f / r;
  float dWdr = dWdq * inv_h;
  return make_float2(dWdr * rij.x * invr, dWdr * rij.y * invr);
}

// ---- cell keys (k_build_cells :165-176, integer part) ------------------------------------------
__global__ void sph_keys(const float2 *__restrict__ pos, unsigned *__restrict__ keys,
                         unsigned *__restrict__ vals, Consts c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.N) return;
  const float2 p = pos[i];
  keys[i] = (unsigned)(grid_c(p.y, c.cell, c.Gy) * c.Gx + grid_c(p.x, c.cellx, c.Gx));
  vals[i] = (unsigned)i;
}

// ---- stable LSD radix sort, one digit per pass ------------------------------------------------------
// Each warp owns a contiguous segment of SORT_SEG keys.  hist[digit * nwarps + warp] counts are
// scanned digit-major, so that equal digits keep segment order and, inside a segment, load order.
__global__ void __launch_bounds__(SORT_WARPS * 32)
sort_hist(const unsigned *__restrict__ keys, unsigned *__restrict__ hist, int n, int shift, int bits,
          int nwarps) {
  __shared__ unsigned cnt[SORT_WARPS][1 << SORT_MAX_BITS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int gw = blockIdx.x * SORT_WARPS + w;
  const int nb = 1 << bits;
  for (int b = lane; b < nb; b += 32) cnt[w][b] = 0;
  __syncwarp();
  if (gw < nwarps) {
    const int beg = gw * SORT_SEG, end = min(beg + SORT_SEG, n);
    for (int i = beg + lane; i < end; i += 32)
      atomicAdd(&cnt[w][(keys[i] >> shift) & (nb - 1)], 1u);
    __syncwarp();
    for (int b = lane; b < nb; b += 32) hist[(size_t)b * nwarps + gw] = cnt[w][b];
  }
}

// exclusive scan of `n` counters in place, two kernels: every CTA scans a tile of SCAN_TILE
// counters (coalesced, warp-shuffle scan) and publishes its total; the second kernel adds to each
// tile the sum of the totals before it.
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_PER_THREAD = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tiles(unsigned *__restrict__ a, unsigned *__restrict__ totals, int n) {
  __shared__ unsigned warp_sum[SCAN_THREADS / 32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int base = blockIdx.x * SCAN_TILE + t * SCAN_PER_THREAD;
  unsigned v[SCAN_PER_THREAD], sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; ++k) {
    v[k] = (base + k < n) ? a[base + k] : 0u;
    sum += v[k];
  }
  unsigned inc = sum;  // inclusive scan of the per-thread sums across the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) warp_sum[w] = inc;
  __syncthreads();
  if (w == 0) {
    unsigned ws = warp_sum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, ws, o);
      if (lane >= o) ws += u;
    }
    warp_sum[lane] = ws;  // inclusive over warps
  }
  __syncthreads();
  unsigned run = (inc - sum) + (w ? warp_sum[w - 1] : 0u);
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; ++k) {
    if (base + k < n) a[base + k] = run;
    run += v[k];
  }
  if (t == SCAN_THREADS - 1) totals[blockIdx.x] = run;
}
__global__ void __launch_bounds__(SCAN_THREADS)
scan_add(unsigned *__restrict__ a, const unsigned *__restrict__ totals, int n) {
  __shared__ unsigned off;
  if (threadIdx.x < 32) {
    unsigned s = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += 32) s += totals[b];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) off = s;
  }
  __syncthreads();
  const unsigned o = off;
  if (blockIdx.x == 0) return;
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; ++k)
    if (base + k < n) a[base + k] += o;
}

__global__ void __launch_bounds__(SORT_WARPS * 32)
sort_scatter(const unsigned *__restrict__ keys_in, const unsigned *__restrict__ vals_in,
             unsigned *__restrict__ keys_out, unsigned *__restrict__ vals_out,
             const unsigned *__restrict__ offs, int n, int shift, int bits, int nwarps) {
  __shared__ unsigned base[SORT_WARPS][1 << SORT_MAX_BITS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int gw = blockIdx.x * SORT_WARPS + w;
  if (gw >= nwarps) return;
  const int nb = 1 << bits;
  for (int b = lane; b < nb; b += 32) base[w][b] = offs[(size_t)b * nwarps + gw];
  __syncwarp();
  const int beg = gw * SORT_SEG, end = min(beg + SORT_SEG, n);
  for (int i0 = beg; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool ok = i < end;
    const unsigned k = ok ? keys_in[i] : 0u, v = ok ? vals_in[i] : 0u;
    const unsigned d = ok ? ((k >> shift) & (nb - 1)) : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, d);   // lanes with the same digit
    const unsigned rank = __popc(peers & ((1u << lane) - 1));  // earlier lanes first: stable
    unsigned dst = 0;
    if (ok) dst = base[w][d] + rank;
    __syncwarp();
    if (ok && rank == 0) base[w][d] += __popc(peers);          // group leader advances the bin
    __syncwarp();
    if (ok) {
      keys_out[dst] = k;
      vals_out[dst] = v;
    }
  }
}

// ---- cell ranges + gather into sorted order -----------------------------------------------------------
// cellStart[c] = first sorted slot whose key is >= c (c = 0..M), so [cellStart[c], cellStart[c+1])
// is cell c (empty cells included) and a run of consecutive cells of one grid row is ONE contiguous
// slot range — the 3x3 neighbourhood is three ranges, not nine.
__global__ void sph_cell_start(const unsigned *__restrict__ keys, int *__restrict__ cellStart, int n, int M) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > M) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] < (unsigned)c) lo = mid + 1;
    else hi = mid;
  }
  cellStart[c] = lo;
}
// (the force sweep reads position AND velocity of a candidate: one 16-byte record per slot, one address, one load)
__global__ void sph_gather(const unsigned *__restrict__ vals, const float2 *__restrict__ pos,
                           const float2 *__restrict__ vel, float2 *__restrict__ sxy,
                           float4 *__restrict__ spv, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const unsigned i = vals[k];
  const float2 x = pos[i], v = vel[i];
  sxy[k] = x;
  spv[k] = make_float4(x.x, x.y, v.x, v.y);
}

__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = GROUP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Neighbour sweep: the GROUP lanes of a group stride over the three contiguous slot ranges (one per
// cell row) of a particle's 3x3 neighbourhood; `test` is the distance check (~35 % pass), `heavy`
// the kernel / gradient evaluation.
// (Tried on B200 and rejected, see profiles/sph_r1_ncu_full_compaction_variant.txt: compacting the hits through a ballot + shared
// queue so that `heavy` runs on full batches — the queue bookkeeping costs more instructions than
// the predicated-off lanes it saves, both with 8-lane groups and with one particle per warp.)
template <typename Test, typename Heavy>
__device__ __forceinline__ void neighbour_sweep(const int *__restrict__ cellStart, const Consts &c,
                                                int gx, int gy, int g, Test test, Heavy heavy) {
  const int cxl = max(gx - SUBX, 0), cxr = min(gx + SUBX, c.Gx - 1);
#pragma unroll
  for (int oy = -1; oy <= 1; ++oy) {
    const int cy = gy + oy;
    if ((unsigned)cy >= (unsigned)c.Gy) continue;
    const int end = cellStart[cy * c.Gx + cxr + 1];
    for (int j = cellStart[cy * c.Gx + cxl] + g; j < end; j += GROUP)
      if (test(j)) heavy(j);
  }
}
// the same sweep with an unconditional body (see W_cubic_sel)
template <typename Body>
__device__ __forceinline__ void neighbour_sweep_all(const int *__restrict__ cellStart, const Consts &c,
                                                    int gx, int gy, int g, Body body) {
  const int cxl = max(gx - SUBX, 0), cxr = min(gx + SUBX, c.Gx - 1);
#pragma unroll
  for (int oy = -1; oy <= 1; ++oy) {
    const int cy = gy + oy;
    if ((unsigned)cy >= (unsigned)c.Gy) continue;
    const int end = cellStart[cy * c.Gx + cxr + 1];
#pragma unroll 2
    for (int j = cellStart[cy * c.Gx + cxl] + g; j < end; j += GROUP) body(j);
  }
}

// ---- density + pressure (k_density_pressure_cell :178-213) --------------------------------------------
__global__ void __launch_bounds__(256)
sph_density(const float2 *__restrict__ sxy, const unsigned *__restrict__ vals,
            const int *__restrict__ cellStart, float2 *__restrict__ srp, float *__restrict__ s_out, float *__restrict__ press_out,
            const int *__restrict__ range, Consts c) {
  const int g = threadIdx.x & (GROUP - 1);
  const int k = (range ? range[0] : 0) + (blockIdx.x * blockDim.x + threadIdx.x) / GROUP;
  const bool valid = k < (range ? range[1] : c.N);
  const float2 xi = sxy[valid ? k : 0];
  const int gx = grid_c(xi.x, c.cellx, c.Gx), gy = grid_c(xi.y, c.cell, c.Gy);
  const float inv_h = rcp_approx(c.h);
  float rho = 0.f;
  if (valid) {
    // (no pair test: W is exactly 0 outside the support, and rho + 0 == rho)
    neighbour_sweep_all(cellStart, c, gx, gy, g, [&](int j) {
      const float2 xj = sxy[j];
      const float rx = xi.x - xj.x, ry = xi.y - xj.y;
      rho += c.mass * W_cubic_sel(sqrtf(rx * rx + ry * ry), inv_h, c.alpha);
    });
  }
  rho = group_sum(rho);
  if (valid && g == 0) {
    const float si = logf(fmaxf(rho, 1e-6f));
    const float rr = expf(si);
    const float ratio = rr / c.rho0;
    float p = (c.c0 * c.c0) * c.rho0 * (powf(ratio, c.gammaEOS) - 1.0f) / c.gammaEOS;
    p = fmaxf(p, 0.0f);
    const unsigned i = vals[k];
    s_out[i] = si;
    press_out[i] = p;
    // what k_forces_cell recomputes per pair: rho_j = expf(s[j]) and p_j / (rho_j * rho_j)
    srp[k] = make_float2(rr, p / (rr * rr));
  }
}

// ---- forces (k_forces_cell :215-272) + symplectic-Euler integration (k_integrate :324-355) ------------
// gradW_cubic :118-133 for a pair the caller has already tested (1e-16 < r^2 < 4h^2): the reference's own range test on r
// (it can still fire by one rounding of the square root) becomes a select on dW/dr instead of a branch around the body;
// a pair it rejects contributes an exact +-0 to both sums, which leaves them unchanged.
__device__ __forceinline__ float2 gradW_cubic_sel(float2 rij, float r, float h, float inv_h, float alpha) {
  const float q = r * inv_h;
  float dWdq;
  if (q < 1.0f) dWdq = alpha * (-3.0f * q + 2.25f * q * q);
  else {
    const float t = 2.0f - q;
    dWdq = alpha * (-0.75f * t * t);
  }
  const float invr = 1.0f / r;
  float dWdr = dWdq * inv_h;
  if (r <= 1e-8f || r >= 2.0f * h) dWdr = 0.f;
  return make_float2(dWdr * rij.x * invr, dWdr * rij.y * invr);
}
// the sweep of neighbour_sweep with the candidate's record loaded once and handed to both halves
template <typename Load, typename Test, typename Heavy>
__device__ __forceinline__ void neighbour_sweep_rec(const int *__restrict__ cellStart, const Consts &c, int gx, int gy, int g,
                                                    Load load, Test test, Heavy heavy) {
  const int cxl = max(gx - SUBX, 0), cxr = min(gx + SUBX, c.Gx - 1);
#pragma unroll
  for (int oy = -1; oy <= 1; ++oy) {
    const int cy = gy + oy;
    if ((unsigned)cy >= (unsigned)c.Gy) continue;
    const int end = cellStart[cy * c.Gx + cxr + 1];
    for (int j = cellStart[cy * c.Gx + cxl] + g; j < end; j += GROUP) {
      const auto rec = load(j);
      if (test(rec)) heavy(j, rec);
    }
  }
}
__global__ void __launch_bounds__(256)
sph_forces_integrate(const float4 *__restrict__ spv, const float2 *__restrict__ srp, const unsigned *__restrict__ vals,
                     const int *__restrict__ cellStart, float2 *__restrict__ pos, float2 *__restrict__ vel, float2 *__restrict__ acc,
                     float2 *__restrict__ sxy_new, float2 *__restrict__ svel_new, float dt, Consts c,
                     const int *__restrict__ range) {
  const int g = threadIdx.x & (GROUP - 1);
  // slots this launch integrates: [k_begin, k_end) from the host, or a device-resident range (stripe shards)
  const int k = (range ? range[0] : c.k_begin) + (blockIdx.x * blockDim.x + threadIdx.x) / GROUP;
  const bool valid = k < (range ? range[1] : c.k_end);
  const int kk = valid ? k : 0;
  const float4 pvi = spv[kk];
  const float2 xi = make_float2(pvi.x, pvi.y), vi = make_float2(pvi.z, pvi.w), rpi = srp[kk];
  const float rhoi = rpi.x, pri = rpi.y;
  const int gx = grid_c(xi.x, c.cellx, c.Gx), gy = grid_c(xi.y, c.cell, c.Gy);
  const float twoh = 2.f * c.h, twoh2 = twoh * twoh;
  const float inv_h = rcp_approx(c.h);
  const float eta2 = 0.01f * c.h * c.h, visc_c = -c.viscAlpha * c.c0;  // loop invariants of :247-249
  float ax = 0.f, ay = 0.f;
  if (valid) {
    // (the branch-free form of the density sweep does not pay here: measured +9 % instructions —
    // the pair test skips two dependent loads and two divisions for the 65 % of pairs outside
    // the support whenever a whole warp iteration misses, which the ragged ends of the 8-lane
    // groups make common enough.  The reference's `j != i` (:232) is implied by its r^2 <= 1e-16 test.)
    neighbour_sweep_rec(
        cellStart, c, gx, gy, g, [&](int j) { return spv[j]; },
        [&](const float4 &pj) {
          const float rx = xi.x - pj.x, ry = xi.y - pj.y;
          const float r2 = rx * rx + ry * ry;
          return !(r2 >= twoh2 || r2 <= 1e-16f);
        },
        [&](int j, const float4 &pj) {
          const float2 rij = make_float2(xi.x - pj.x, xi.y - pj.y);
          const float r2 = rij.x * rij.x + rij.y * rij.y;
          const float r = sqrtf(r2);
          const float2 gW = gradW_cubic_sel(rij, r, c.h, inv_h, c.alpha);
          const float2 rpj = srp[j];
          const float common = -c.mass * (pri + rpj.y);
          ax += common * gW.x;
          ay += common * gW.y;
          if (c.useVisc) {
            const float vx = vi.x - pj.z, vy = vi.y - pj.w;
            const float dot = vx * rij.x + vy * rij.y;
            if (dot < 0.f) {
              const float mu = (c.h * dot) / (r2 + eta2);
              const float rhoBar = 0.5f * (rhoi + rpj.x);
              const float Pi_ij = (visc_c * mu) / rhoBar;
              ax += -c.mass * Pi_ij * gW.x;
              ay += -c.mass * Pi_ij * gW.y;
            }
          }
        });
  }
  ax = group_sum(ax);
  ay = group_sum(ay);
  if (valid && g == 0) {
    if (c.useGrav) {
      ax += c.gx;
      ay += c.gy;
    }
    const unsigned i = vals[k];
    acc[i] = make_float2(ax, ay);
    float2 v = vi, x = xi;   // k_integrate
    v.x += ax * dt;
    v.y += ay * dt;
    x.x += v.x * dt;
    x.y += v.y * dt;
    const float e = 0.2f;
    if (x.x < 0.f) { x.x = 0.f; v.x = -e * v.x; }
    if (x.x > c.boxX) { x.x = c.boxX; v.x = -e * v.x; }
    if (x.y < 0.f) { x.y = 0.f; v.y = -e * v.y; }
    if (x.y > c.boxY) { x.y = c.boxY; v.y = -e * v.y; }
    if (c.write_state) {
      pos[i] = x;
      vel[i] = v;
    }
    if (sxy_new) {  // XSPH needs the post-integration state in (old) sorted order
      sxy_new[k] = x;
      svel_new[k] = v;
    }
  }
}

// ---- XSPH (k_xsph_cell :274-313 + k_apply_xsph :315-322) --------------------------------------------
// Runs after integration on the updated positions/velocities but with the cell structure built
// before it, exactly like the reference (its lists are not rebuilt between the two kernels).
__global__ void __launch_bounds__(256)
sph_xsph(const float2 *__restrict__ sxy_new, const float2 *__restrict__ svel_new,
         const float2 *__restrict__ srp, const unsigned *__restrict__ vals,
         const int *__restrict__ cellStart, float2 *__restrict__ dvel, Consts c, const int *__restrict__ range) {
  const int g = threadIdx.x & (GROUP - 1);
  const int k = (range ? range[0] : 0) + (blockIdx.x * blockDim.x + threadIdx.x) / GROUP;
  const bool valid = k < (range ? range[1] : c.N);
  const int kk = valid ? k : 0;
  const float2 xi = sxy_new[kk], vi = svel_new[kk];
  const float rhoi = srp[kk].x;
  const int gx = grid_c(xi.x, c.cellx, c.Gx), gy = grid_c(xi.y, c.cell, c.Gy);
  const float twoh = 2.f * c.h, twoh2 = twoh * twoh;
  float dx = 0.f, dy = 0.f;
  if (valid) {
    neighbour_sweep(
        cellStart, c, gx, gy, g,
        [&](int j) {
          const float2 xj = sxy_new[j];
          const float rx = xi.x - xj.x, ry = xi.y - xj.y;
          return (j != k) && (rx * rx + ry * ry < twoh2);
        },
        [&](int j) {
          const float2 xj = sxy_new[j];
          const float rx = xi.x - xj.x, ry = xi.y - xj.y;
          const float w = W_cubic(sqrtf(rx * rx + ry * ry), c.h, c.alpha);
          const float rhoBar = 0.5f * (rhoi + srp[j].x);
          const float2 vj = svel_new[j];
          dx += (c.mass / rhoBar) * (vj.x - vi.x) * w;
          dy += (c.mass / rhoBar) * (vj.y - vi.y) * w;
        });
  }
  dx = group_sum(dx);
  dy = group_sum(dy);
  if (valid && g == 0) dvel[vals[k]] = make_float2(c.xsphEps * dx, c.xsphEps * dy);
}
__global__ void sph_apply_xsph(float2 *__restrict__ vel, const float2 *__restrict__ dvel, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    vel[i].x += dvel[i].x;
    vel[i].y += dvel[i].y;
  }
}

// ---- multi-GPU (replicated state, sharded work) helpers -------------------------------------------
// slots whose density a rank needs: its own slots [k0, k1) plus every slot in the cell rows just
// below / above them (the 3x3 search of an own particle reaches one cell row further)
__global__ void sph_ghost_range(const unsigned *__restrict__ keys, const int *__restrict__ cellStart,
                                int *__restrict__ range, int k0, int k1, int Gx, int Gy) {
  if (threadIdx.x || blockIdx.x) return;
  const int r0 = (int)(keys[k0] / (unsigned)Gx) - 1, r1 = (int)(keys[k1 - 1] / (unsigned)Gx) + 1;
  range[0] = cellStart[max(r0, 0) * Gx];
  range[1] = cellStart[min(r1 + 1, Gy) * Gx];
}
// sorted-order state -> original particle order (after the all-gather of every rank's slots)
__global__ void sph_scatter_state(const unsigned *__restrict__ vals, const float2 *__restrict__ sxy_new,
                                  const float2 *__restrict__ svel_new, float2 *__restrict__ pos,
                                  float2 *__restrict__ vel, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const unsigned i = vals[k];
  pos[i] = sxy_new[k];
  vel[i] = svel_new[k];
}

// ---- rain (k_rain :377-392), write collisions resolved: the highest spawn index wins ---------------
__device__ __forceinline__ void rain_draw(int k, unsigned seed, int N, float boxX, float boxY,
                                          float &x, float &y, int &i) {
  unsigned s = seed ^ (k * 1664525u + 1013904223u);
  s = s * 1664525u + 1013904223u;
  float rx = (s & 0x00FFFFFF) / 16777216.f;
  s = s * 1664525u + 1013904223u;
  x = rx * (boxX * 0.8f) + 0.1f * boxX;
  float ry = (s & 0x00FFFFFF) / 16777216.f;
  y = boxY * (0.9f + 0.08f * ry);
  i = (int)(s % (unsigned)N);
}
__global__ void sph_rain_claim(int *__restrict__ winner, int N, int nspawn, float boxX, float boxY,
                               unsigned seed) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nspawn) return;
  float x, y;
  int i;
  rain_draw(k, seed, N, boxX, boxY, x, y, i);
  atomicMax(&winner[i], k);
}
__global__ void sph_rain_write(float2 *pos, float2 *vel, int *__restrict__ winner, int N, int nspawn,
                               float boxX, float boxY, float c0, unsigned seed) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nspawn) return;
  float x, y;
  int i;
  rain_draw(k, seed, N, boxX, boxY, x, y, i);
  if (winner[i] == k) {
    pos[i] = make_float2(x, y);
    vel[i] = make_float2(0.f, -0.5f * c0);
    winner[i] = -1;  // re-arm
  }
}
// ---- render pass: k_clear_grid + k_rasterize (:357-374) -----------------------------------------
// Particle counts on the terminal's half-block raster (W x 2H).  2 M particles fall on a few 10^4
// cells, so the reference's one-atomic-per-particle contends ~100-fold per address; here the lanes
// of a warp that hit the same cell are merged first (__match_any_sync) and one of them adds the
// group's population.  Integer result, identical to the reference's.
__global__ void sph_rasterize(const float2 *__restrict__ pos, int N, int *__restrict__ grid2, int W, int H,
                              float boxX, float boxY) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int cell = -1;
  if (i < N) {
    const float2 p = pos[i];
    const int cx = (int)(p.x / boxX * (W - 1));
    const int sy = (int)((boxY - p.y) / boxY * (2 * H - 1));  // flip y
    if ((unsigned)cx < (unsigned)W && (unsigned)sy < (unsigned)(2 * H)) cell = sy * W + cx;
  }
  const unsigned peers = __match_any_sync(0xffffffffu, cell);
  if (cell >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&grid2[cell], __popc(peers));
}

__global__ void sph_fill_int(int *a, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

}  // namespace

struct tau_sph {
  tau_sph_params p;
  int device;
  cudaStream_t stream;
  bool own_stream;
  // particle state in ORIGINAL index order (the reference's arrays, tau_sph.cu:561-565)
  float2 *pos, *vel, *acc;
  float *s, *press;
  // sorted-order scratch
  unsigned *keys[2], *vals[2], *hist, *scan_totals;
  float2 *sxy, *srp, *sxy_new, *svel_new;
  float4 *spv;  // sorted (x, y, vx, vy) records
  int *cellStart, *winner;
  int sorted_buf;  // which keys/vals buffer holds the last sort result
  // multi-GPU sharding (replicated state): this rank integrates slots [rank*chunk, ...)
  int rank, world, chunk, sub_k;
  float dTau_accum, dt_sub;
  int *range;
  int *grid2;          // render raster (device), allocated on first use
  size_t grid2_cap;
  // derived constants (:573-578, ensure_cell_buffers :512-540)
  float mass, h, cell, alpha;
  int Gx, Gy, Gxk, M, key_bits, nwarps;  // Gx x Gy: the reference's grid; Gxk = Gx * SUBX key columns; M = Gxk * Gy keys
  // host step control (:663-722)
  float t, tau, rain_carry;
  long long step, substeps, launches;
  cudaEvent_t ev0, ev1;
  bool timed, have_state;
};

namespace {

Consts make_consts(const tau_sph *h) {
  Consts c;
  c.N = h->p.N;
  c.Gx = h->Gxk;
  c.cellx = h->cell / SUBX;
  c.Gy = h->Gy;
  c.cell = h->cell;
  c.mass = h->mass;
  c.h = h->h;
  c.rho0 = h->p.rho0;
  c.c0 = h->p.c0;
  c.gammaEOS = h->p.gammaEOS;
  c.viscAlpha = h->p.viscAlpha;
  c.gx = 0.f;
  c.gy = -(h->p.useGrav ? h->p.gravity : 0.f);
  c.boxX = h->p.boxX;
  c.boxY = h->p.boxY;
  c.alpha = h->alpha;
  c.xsphEps = h->p.xsphEps;
  c.useVisc = h->p.useVisc;
  c.useGrav = h->p.useGrav;
  c.k_begin = 0;
  c.k_end = h->p.N;
  c.write_state = 1;
  return c;
}

// widest digit of a pass: SORT_MAX_BITS, or TAU_SPH_SORT_BITS (1 .. SORT_MAX_BITS) for experiments — the result of the sort
// does not depend on it
int sort_digit_bits() {
  static int bits = 0;
  if (bits == 0) {
    const char *e = getenv("TAU_SPH_SORT_BITS");
    const int v = e ? atoi(e) : SORT_DEFAULT_BITS;
    bits = v < 1 ? 1 : (v > SORT_MAX_BITS ? SORT_MAX_BITS : v);
  }
  return bits;
}
// stable radix sort of (keys[0], vals[0]), n pairs with keys < 2^key_bits; returns the buffer index holding the result
int radix_sort_pairs(unsigned *const keys[2], unsigned *const vals[2], unsigned *hist, unsigned *scan_totals, int n,
                     int key_bits, cudaStream_t stream, long long *launches) {
  const int nwarps = (n + SORT_SEG - 1) / SORT_SEG;
  const int digit_max = sort_digit_bits();
  const int passes = (key_bits + digit_max - 1) / digit_max;
  const int bits = (key_bits + passes - 1) / passes;
  const int blocks = (nwarps + SORT_WARPS - 1) / SORT_WARPS;
  int src = 0;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * bits;
    sort_hist<<<blocks, SORT_WARPS * 32, 0, stream>>>(keys[src], hist, n, shift, bits, nwarps);
    {
      const int m = (1 << bits) * nwarps, tiles = (m + SCAN_TILE - 1) / SCAN_TILE;
      scan_tiles<<<tiles, SCAN_THREADS, 0, stream>>>(hist, scan_totals, m);
      scan_add<<<tiles, SCAN_THREADS, 0, stream>>>(hist, scan_totals, m);
    }
    sort_scatter<<<blocks, SORT_WARPS * 32, 0, stream>>>(keys[src], vals[src], keys[src ^ 1], vals[src ^ 1], hist, n, shift,
                                                         bits, nwarps);
    *launches += 4;
    src ^= 1;
  }
  return src;
}
int radix_sort(tau_sph *h) {
  return radix_sort_pairs(h->keys, h->vals, h->hist, h->scan_totals, h->p.N, h->key_bits, h->stream, &h->launches);
}

// first half of a sub-step: keys, sort, cell ranges, gather, density, forces + integration.
// Sharded handles only integrate their own slots (into the sorted copies) — the caller all-gathers
// sxy_new / svel_new across ranks before substep_finish().
int substep_compute(tau_sph *h, float dt_sub) {
  Consts c = make_consts(h);
  const int n = c.N, BS = 256, GS = (n + BS - 1) / BS;
  const bool sharded = h->world > 1;
  sph_keys<<<GS, BS, 0, h->stream>>>(h->pos, h->keys[0], h->vals[0], c);
  const int sb = radix_sort(h);
  h->sorted_buf = sb;
  sph_cell_start<<<(h->M + 1 + BS - 1) / BS, BS, 0, h->stream>>>(h->keys[sb], h->cellStart, n, h->M);
  sph_gather<<<GS, BS, 0, h->stream>>>(h->vals[sb], h->pos, h->vel, h->sxy, h->spv, n);
  const int GSall = (int)(((size_t)n * GROUP + BS - 1) / BS);
  if (sharded) {
    c.k_begin = h->rank * h->chunk;
    c.k_end = min(n, c.k_begin + h->chunk);
    c.write_state = 0;
    sph_ghost_range<<<1, 32, 0, h->stream>>>(h->keys[sb], h->cellStart, h->range, c.k_begin, c.k_end, h->Gxk,
                                             h->Gy);
    h->launches++;
  }
  sph_density<<<GSall, BS, 0, h->stream>>>(h->sxy, h->vals[sb], h->cellStart, h->srp, h->s, h->press,
                                           sharded ? h->range : nullptr, c);
  const bool xsph = h->p.useXSPH && h->p.xsphEps > 0.f;
  const int GSown = (int)(((size_t)(c.k_end - c.k_begin) * GROUP + BS - 1) / BS);
  sph_forces_integrate<<<GSown, BS, 0, h->stream>>>(h->spv, h->srp, h->vals[sb], h->cellStart,
                                                    h->pos, h->vel, h->acc,
                                                    (xsph || sharded) ? h->sxy_new : nullptr,
                                                    (xsph || sharded) ? h->svel_new : nullptr, dt_sub, c, nullptr);
  h->launches += 5;
  if (xsph) {
    const int GSg = GSall;
    sph_xsph<<<GSg, BS, 0, h->stream>>>(h->sxy_new, h->svel_new, h->srp, h->vals[sb], h->cellStart,
                                        h->acc, c, nullptr);
    sph_apply_xsph<<<GS, BS, 0, h->stream>>>(h->vel, h->acc, n);
    h->launches += 2;
  }
  TAU_CUDA(cudaGetLastError());
  return TAU_OK;
}

// second half: (sharded: sorted copies of ALL slots -> original order) + rain
int substep_finish(tau_sph *h, float dt_sub) {
  const int n = h->p.N, BS = 256, GS = (n + BS - 1) / BS;
  if (h->world > 1) {
    sph_scatter_state<<<GS, BS, 0, h->stream>>>(h->vals[h->sorted_buf], h->sxy_new, h->svel_new, h->pos,
                                                h->vel, n);
    h->launches++;
  }
  if (h->p.rain) {  // :706-716
    h->rain_carry += 0.02f * h->p.N * dt_sub;
    const int nspawn = (int)h->rain_carry;
    h->rain_carry -= nspawn;
    if (nspawn > 0) {
      const int BSr = 128, GSr = (nspawn + BSr - 1) / BSr;
      const unsigned seed = (unsigned)(h->p.seed + h->step);
      sph_rain_claim<<<GSr, BSr, 0, h->stream>>>(h->winner, n, nspawn, h->p.boxX, h->p.boxY, seed);
      sph_rain_write<<<GSr, BSr, 0, h->stream>>>(h->pos, h->vel, h->winner, n, nspawn, h->p.boxX,
                                                 h->p.boxY, h->p.c0, seed);
      h->launches += 2;
    }
  }
  h->substeps++;
  TAU_CUDA(cudaGetLastError());
  return TAU_OK;
}

int substep(tau_sph *h, float dt_sub) {
  int rc = substep_compute(h, dt_sub);
  if (rc) return rc;
  return substep_finish(h, dt_sub);
}

}  // namespace

extern "C" {

void tau_sph_default_params(tau_sph_params *p) {  // struct Params tau_sph.cu:49-85
  p->N = 1 << 16;
  p->boxX = 1.0f;
  p->boxY = 1.0f;
  p->dTau = 1.0f;
  p->t0 = 1.0f;
  p->CFL = 1.0f;
  p->rho0 = 1.0f;
  p->c0 = 1.0f;
  p->gammaEOS = 1.0f;
  p->hMul = 2.0f;
  p->viscAlpha = 0.25f;
  p->gravity = 9.81f;
  p->rain = 1;
  p->useVisc = 1;
  p->useGrav = 1;
  p->viscSub = 1;
  p->useXSPH = 0;
  p->xsphEps = 0.25f;
  p->seed = 69420;
}

// reset_particles :493-510 — jittered lattice from std::mt19937 + uniform_real_distribution<float>
// (libstdc++'s stream, as the reference binary produces it)
void tau_sph_reset_particles(const tau_sph_params *P, float *pos_xy, float *vel_xy) {
  std::mt19937 rng(P->seed);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  int nSide = (int)sqrtf((float)P->N);
  int nx = nSide, ny = (P->N + nSide - 1) / nSide;
  float padX = 0.05f * P->boxX, padY = 0.05f * P->boxY;
  float width = P->boxX - 2 * padX, height = 0.6f * P->boxY - padY;
  for (int i = 0; i < P->N; ++i) {
    int ix = i % nx, iy = i / nx;
    float fx = (ix + 0.5f) / nx, fy = (iy + 0.5f) / ny;
    float x = padX + fx * width, y = padY + fy * height;
    x += (U(rng) - 0.5f) * 0.2f * width / nx;
    y += (U(rng) - 0.5f) * 0.2f * height / ny;
    pos_xy[2 * i] = x;
    pos_xy[2 * i + 1] = y;
    vel_xy[2 * i] = 0.f;
    vel_xy[2 * i + 1] = 0.f;
  }
}

int tau_sph_create(const tau_sph_params *p, int device, void *stream, tau_sph **out) {
  TAU_REQUIRE(p && out, "tau_sph_create: null argument");
  TAU_REQUIRE(p->N > 0, "tau_sph_create: N must be positive");
  TAU_REQUIRE(p->boxX > 0.f && p->boxY > 0.f && p->hMul > 0.f, "tau_sph_create: bad box / hMul");
  if (tau_device_count() <= 0) {
    tau_set_error("tau_sph_create: no CUDA device (this library has no CPU fallback)");
    return TAU_ERR_NODEV;
  }
  TAU_CUDA(cudaSetDevice(device));
  tau_sph *h = new (std::nothrow) tau_sph();
  if (!h) return TAU_ERR_NOMEM;
  memset(h, 0, sizeof(*h));
  h->p = *p;
  h->device = device;
  if (stream) {
    h->stream = (cudaStream_t)stream;
    h->own_stream = false;
  } else {
    TAU_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  // derived constants :573-578
  const float area = p->boxX * p->boxY;
  h->mass = (p->rho0 * area) / p->N;
  const float spacing = sqrtf(area / p->N);
  h->h = p->hMul * spacing;
  h->alpha = (float)(10.0f / (7.0f * M_PI * h->h * h->h));  // W_cubic's alpha, double as in :107
  // ensure_cell_buffers :512-540
  h->cell = 2.0f * h->h;
  h->Gx = (int)ceilf(p->boxX / h->cell);
  h->Gy = (int)ceilf(p->boxY / h->cell);
  if (h->Gx < 1) h->Gx = 1;
  if (h->Gy < 1) h->Gy = 1;
  h->Gxk = h->Gx * SUBX;
  h->M = h->Gxk * h->Gy;
  h->key_bits = 1;
  while ((1 << h->key_bits) < h->M) h->key_bits++;
  h->nwarps = (p->N + SORT_SEG - 1) / SORT_SEG;
  const size_t n = (size_t)p->N;
  TAU_CUDA(cudaMalloc(&h->pos, n * sizeof(float2)));
  TAU_CUDA(cudaMalloc(&h->vel, n * sizeof(float2)));
  TAU_CUDA(cudaMalloc(&h->acc, n * sizeof(float2)));
  TAU_CUDA(cudaMalloc(&h->s, n * sizeof(float)));
  TAU_CUDA(cudaMalloc(&h->press, n * sizeof(float)));
  for (int b = 0; b < 2; ++b) {
    TAU_CUDA(cudaMalloc(&h->keys[b], n * sizeof(unsigned)));
    TAU_CUDA(cudaMalloc(&h->vals[b], n * sizeof(unsigned)));
  }
  TAU_CUDA(cudaMalloc(&h->hist, (size_t)(1 << SORT_MAX_BITS) * h->nwarps * sizeof(unsigned)));
  TAU_CUDA(cudaMalloc(&h->scan_totals,
                      (((size_t)(1 << SORT_MAX_BITS) * h->nwarps + SCAN_TILE - 1) / SCAN_TILE + 1) * sizeof(unsigned)));
  TAU_CUDA(cudaMalloc(&h->sxy, n * sizeof(float2)));
  TAU_CUDA(cudaMalloc(&h->spv, n * sizeof(float4)));
  TAU_CUDA(cudaMalloc(&h->srp, n * sizeof(float2)));
  TAU_CUDA(cudaMalloc(&h->sxy_new, (n + 64) * sizeof(float2)));   // + padding: all-gather chunks
  TAU_CUDA(cudaMalloc(&h->svel_new, (n + 64) * sizeof(float2)));
  TAU_CUDA(cudaMalloc(&h->range, 2 * sizeof(int)));
  h->rank = 0;
  h->world = 1;
  h->chunk = p->N;
  TAU_CUDA(cudaMalloc(&h->cellStart, (size_t)(h->M + 1) * sizeof(int)));
  TAU_CUDA(cudaMalloc(&h->winner, n * sizeof(int)));
  sph_fill_int<<<(p->N + 255) / 256, 256, 0, h->stream>>>(h->winner, p->N, -1);
  TAU_CUDA(cudaMemsetAsync(h->acc, 0, n * sizeof(float2), h->stream));
  TAU_CUDA(cudaMemsetAsync(h->s, 0, n * sizeof(float), h->stream));
  TAU_CUDA(cudaMemsetAsync(h->press, 0, n * sizeof(float), h->stream));
  TAU_CUDA(cudaEventCreate(&h->ev0));
  TAU_CUDA(cudaEventCreate(&h->ev1));
  h->tau = 0.f;
  h->t = p->t0 * expf(h->tau);  // :578
  h->rain_carry = 0.f;
  *out = h;
  return TAU_OK;
}

int tau_sph_upload(tau_sph *h, const float *pos_xy, const float *vel_xy) {
  TAU_REQUIRE(h && pos_xy && vel_xy, "tau_sph_upload: null argument");
  TAU_CUDA(cudaSetDevice(h->device));
  const size_t bytes = (size_t)h->p.N * sizeof(float2);
  TAU_CUDA(cudaMemcpyAsync(h->pos, pos_xy, bytes, cudaMemcpyHostToDevice, h->stream));
  TAU_CUDA(cudaMemcpyAsync(h->vel, vel_xy, bytes, cudaMemcpyHostToDevice, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  h->have_state = true;
  return TAU_OK;
}

int tau_sph_init(tau_sph *h) {
  TAU_REQUIRE(h, "tau_sph_init: null handle");
  std::vector<float> p(2 * (size_t)h->p.N), v(2 * (size_t)h->p.N);
  tau_sph_reset_particles(&h->p, p.data(), v.data());
  h->tau = 0.f;
  h->t = h->p.t0 * expf(h->tau);
  h->rain_carry = 0.f;
  h->step = 0;
  return tau_sph_upload(h, p.data(), v.data());
}

// nframes x the doStep block :663-722 (each frame = viscSub sub-steps)
int tau_sph_step(tau_sph *h, int nframes) {
  TAU_REQUIRE(h, "tau_sph_step: null handle");
  TAU_REQUIRE(nframes >= 0, "tau_sph_step: nframes must be >= 0");
  TAU_REQUIRE(h->have_state, "tau_sph_step: no state (call tau_sph_init or tau_sph_upload)");
  TAU_REQUIRE(h->world == 1, "tau_sph_step: sharded handles advance with tau_sph_shard_substep_begin/end");
  TAU_CUDA(cudaSetDevice(h->device));
  TAU_CUDA(cudaEventRecord(h->ev0, h->stream));
  const tau_sph_params &P = h->p;
  for (int f = 0; f < nframes; ++f) {
    float dTau_accum = 0.f;
    const int K = (P.viscSub > 0 ? P.viscSub : 1);
    const float dt_try = h->t * P.dTau;
    const float dt_cfl = P.CFL * h->h / (P.c0 * (1.0f + 2.0f * P.viscAlpha));
    const float dt_eff = fminf(dt_try, dt_cfl);
    const float dt_sub = dt_eff / K;
    for (int k = 0; k < K; ++k) {
      int rc = substep(h, dt_sub);
      if (rc) return rc;
      const float dTau_actual = dt_sub / fmaxf(h->t, 1e-9f);
      dTau_accum += dTau_actual;
      h->t = P.t0 * expf(h->tau + dTau_accum);
    }
    h->tau += dTau_accum;
    h->step++;
  }
  TAU_CUDA(cudaEventRecord(h->ev1, h->stream));
  h->timed = true;
  return TAU_OK;
}

// ---- multi-GPU: replicated particle state, work sharded by sorted-slot (= cell-stripe) range -------
// Every rank holds all N particles and performs the (cheap) sort; rank r computes densities for its
// slot chunk plus the ghost cell rows around it and integrates its own chunk into the sorted copies
// sxy_new / svel_new.  Between begin and end the caller all-gathers those two arrays (equal chunks
// of `chunk` slots; NCCL over NVLink).  XSPH is not supported in this mode.
int tau_sph_shard_config(tau_sph *h, int rank, int world) {
  TAU_REQUIRE(h, "tau_sph_shard_config: null handle");
  TAU_REQUIRE(world >= 1 && rank >= 0 && rank < world, "tau_sph_shard_config: bad rank %d of %d", rank, world);
  TAU_REQUIRE(!(world > 1 && h->p.useXSPH && h->p.xsphEps > 0.f),
              "tau_sph_shard_config: XSPH needs a second ghost exchange and is single-GPU only");
  const int chunk = (h->p.N + world - 1) / world;
  TAU_REQUIRE((size_t)chunk * world <= (size_t)h->p.N + 64, "tau_sph_shard_config: N=%d does not split into "
              "%d chunks within the padding", h->p.N, world);
  h->rank = rank;
  h->world = world;
  h->chunk = chunk;
  h->sub_k = 0;
  h->dTau_accum = 0.f;
  return TAU_OK;
}

int tau_sph_shard_buffers(tau_sph *h, float **sxy_new, float **svel_new, int *chunk) {
  TAU_REQUIRE(h, "tau_sph_shard_buffers: null handle");
  if (sxy_new) *sxy_new = reinterpret_cast<float *>(h->sxy_new);
  if (svel_new) *svel_new = reinterpret_cast<float *>(h->svel_new);
  if (chunk) *chunk = h->chunk;
  return TAU_OK;
}

int tau_sph_shard_substep_begin(tau_sph *h) {
  TAU_REQUIRE(h && h->have_state, "tau_sph_shard_substep_begin: no state");
  TAU_CUDA(cudaSetDevice(h->device));
  const tau_sph_params &P = h->p;
  const int K = (P.viscSub > 0 ? P.viscSub : 1);
  if (h->sub_k == 0) {  // frame start :663-669
    const float dt_try = h->t * P.dTau;
    const float dt_cfl = P.CFL * h->h / (P.c0 * (1.0f + 2.0f * P.viscAlpha));
    h->dt_sub = fminf(dt_try, dt_cfl) / K;
    h->dTau_accum = 0.f;
  }
  return substep_compute(h, h->dt_sub);
}

int tau_sph_shard_substep_end(tau_sph *h) {
  TAU_REQUIRE(h && h->have_state, "tau_sph_shard_substep_end: no state");
  const tau_sph_params &P = h->p;
  const int K = (P.viscSub > 0 ? P.viscSub : 1);
  int rc = substep_finish(h, h->dt_sub);
  if (rc) return rc;
  h->dTau_accum += h->dt_sub / fmaxf(h->t, 1e-9f);  // :718-720
  h->t = P.t0 * expf(h->tau + h->dTau_accum);
  if (++h->sub_k == K) {
    h->tau += h->dTau_accum;
    h->step++;
    h->sub_k = 0;
  }
  return TAU_OK;
}

int tau_sph_clock(tau_sph *h, float *t, float *tau, long long *step) {
  TAU_REQUIRE(h, "tau_sph_clock: null handle");
  if (t) *t = h->t;
  if (tau) *tau = h->tau;
  if (step) *step = h->step;
  return TAU_OK;
}

int tau_sph_download(tau_sph *h, float *pos_xy, float *vel_xy, float *s, float *press) {
  TAU_REQUIRE(h, "tau_sph_download: null handle");
  TAU_CUDA(cudaSetDevice(h->device));
  const size_t n = (size_t)h->p.N;
  if (pos_xy) TAU_CUDA(cudaMemcpyAsync(pos_xy, h->pos, n * sizeof(float2), cudaMemcpyDeviceToHost, h->stream));
  if (vel_xy) TAU_CUDA(cudaMemcpyAsync(vel_xy, h->vel, n * sizeof(float2), cudaMemcpyDeviceToHost, h->stream));
  if (s) TAU_CUDA(cudaMemcpyAsync(s, h->s, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  if (press) TAU_CUDA(cudaMemcpyAsync(press, h->press, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

// Render pass of the frame loop :747-755 (k_clear_grid, k_rasterize, D2H): particle counts on a
// W x 2H raster (two vertical samples per terminal row, `halfblocks`), row-major, y flipped.
int tau_sph_rasterize(tau_sph *h, int W, int H, int *grid2) {
  TAU_REQUIRE(h && grid2, "tau_sph_rasterize: null argument");
  TAU_REQUIRE(W >= 1 && H >= 1 && (size_t)W * 2 * H <= (size_t)1 << 28, "tau_sph_rasterize: bad raster %d x %d", W, H);
  TAU_CUDA(cudaSetDevice(h->device));
  const size_t cells = (size_t)W * 2 * H;
  if (h->grid2_cap < cells) {
    if (h->grid2) TAU_CUDA(cudaFree(h->grid2));
    TAU_CUDA(cudaMalloc(&h->grid2, cells * sizeof(int)));
    h->grid2_cap = cells;
  }
  TAU_CUDA(cudaMemsetAsync(h->grid2, 0, cells * sizeof(int), h->stream));
  sph_rasterize<<<(h->p.N + 255) / 256, 256, 0, h->stream>>>(h->pos, h->p.N, h->grid2, W, H, h->p.boxX, h->p.boxY);
  h->launches++;
  TAU_CUDA(cudaGetLastError());
  TAU_CUDA(cudaMemcpyAsync(grid2, h->grid2, cells * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

// the (cell key, particle index) pairs of the most recent sub-step's radix sort, sorted order
int tau_sph_download_sort(tau_sph *h, unsigned *keys, unsigned *vals) {
  TAU_REQUIRE(h && keys && vals, "tau_sph_download_sort: null argument");
  TAU_REQUIRE(h->substeps > 0, "tau_sph_download_sort: no sub-step has run yet");
  const size_t n = (size_t)h->p.N;
  TAU_CUDA(cudaMemcpyAsync(keys, h->keys[h->sorted_buf], n * 4, cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaMemcpyAsync(vals, h->vals[h->sorted_buf], n * 4, cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

// stand-alone access to the sort for tests: sorts n (key, value) pairs with keys < 2^key_bits
int tau_sph_sort_pairs(tau_sph *h, const unsigned *keys_in, unsigned *keys_out, unsigned *vals_out) {
  TAU_REQUIRE(h && keys_in && keys_out && vals_out, "tau_sph_sort_pairs: null argument");
  TAU_CUDA(cudaSetDevice(h->device));
  const size_t n = (size_t)h->p.N;
  std::vector<unsigned> iota(n);
  for (size_t i = 0; i < n; ++i) iota[i] = (unsigned)i;
  TAU_CUDA(cudaMemcpyAsync(h->keys[0], keys_in, n * 4, cudaMemcpyHostToDevice, h->stream));
  TAU_CUDA(cudaMemcpyAsync(h->vals[0], iota.data(), n * 4, cudaMemcpyHostToDevice, h->stream));
  const int sb = radix_sort(h);
  TAU_CUDA(cudaMemcpyAsync(keys_out, h->keys[sb], n * 4, cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaMemcpyAsync(vals_out, h->vals[sb], n * 4, cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  TAU_CUDA(cudaGetLastError());
  return TAU_OK;
}

int tau_sph_grid(tau_sph *h, int *Gx, int *Gy, float *cell, float *hh, float *mass) {
  TAU_REQUIRE(h, "tau_sph_grid: null handle");
  if (Gx) *Gx = h->Gx;
  if (Gy) *Gy = h->Gy;
  if (cell) *cell = h->cell;
  if (hh) *hh = h->h;
  if (mass) *mass = h->mass;
  return TAU_OK;
}
// key columns per grid cell (SUBX): the sort key of a particle is row * (Gx * subx) + key column, and
// key column / subx is the reference's grid_x
int tau_sph_subx(tau_sph *h) {
  (void)h;
  return SUBX;
}

int tau_sph_sync(tau_sph *h) {
  TAU_REQUIRE(h, "tau_sph_sync: null handle");
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

long long tau_sph_substeps_done(tau_sph *h) { return h ? h->substeps : -1; }
long long tau_sph_launch_count(tau_sph *h) { return h ? h->launches : -1; }

int tau_sph_last_step_ms(tau_sph *h, float *ms) {
  TAU_REQUIRE(h && ms, "tau_sph_last_step_ms: null argument");
  TAU_REQUIRE(h->timed, "tau_sph_last_step_ms: no step has been timed yet");
  TAU_CUDA(cudaEventSynchronize(h->ev1));
  TAU_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return TAU_OK;
}

int tau_sph_destroy(tau_sph *h) {
  if (!h) return TAU_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  void *ptrs[] = {h->grid2, h->range, h->winner, h->cellStart, h->svel_new, h->sxy_new, h->srp, h->spv, h->sxy,
                  h->scan_totals, h->hist, h->vals[1], h->keys[1], h->vals[0], h->keys[0], h->press, h->s, h->acc,
                  h->vel, h->pos};
  for (void *p : ptrs) cudaFree(p);
  cudaEventDestroy(h->ev1);
  cudaEventDestroy(h->ev0);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return TAU_OK;
}

}  // extern "C"

#include "sph_stripes.inc"
