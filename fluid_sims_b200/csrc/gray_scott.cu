// gray_scott.cu — Gray-Scott reaction-diffusion update path for sm_100a.
//
// Replaces the per-step host sequence of the reference `tgs` binary
// (tau_gray_scott.cu:321-329: step_kernel<<<>>> + cudaDeviceSynchronize + 2 swaps) and its kernel
// (tau_gray_scott.cu:141-171).  Arithmetic follows the reference expression tree exactly and this
// TU is compiled with the reference's own math flags (-use_fast_math: ftz, fmad, div.approx —
// Makefile:78-79 of the reference), so on identical inputs the planes are bit-identical.
//
// Design (HBM-bound stencil, 16 algorithmic bytes per cell):
//   * planes carry one ghost row above and below (pitch nx, ny_local+2 rows); a 128x32 tile plus
//     halo is staged in shared memory by two TMA box loads (u and v) completing on one mbarrier;
//   * each thread owns a float4 of 4 consecutive cells in x for 4 consecutive rows: x-neighbours
//     come from warp shuffles (edge lanes read the halo columns), y-neighbours are reused from
//     registers, results leave as 128-bit streaming stores;
//   * the periodic wrap costs nothing in the common case: x-wrap is a halo-column patch in edge
//     tiles only; y-wrap is the ghost rows, which the kernel itself refreshes in the output plane
//     (single-GPU) or the slab owner's neighbours provide (multi-GPU);
//   * no per-step host synchronisation.
#include "common.cuh"
#include "../../include/tau_b200.h"

#include <algorithm>
#include <new>
#include <vector>

namespace {

constexpr int GS_TX = 128;                 // tile width  (cells)  = 32 lanes x float4
constexpr int GS_TY = 32;                  // tile height (cells)  = 8 warps x 4 rows
constexpr int GS_HX = 4;                   // halo columns staged each side (TMA needs 16-byte rows)
constexpr int GS_SW = GS_TX + 2 * GS_HX;   // shared tile pitch (floats)
constexpr int GS_SH = GS_TY + 2;           // shared tile rows
constexpr int GS_THREADS = 256;
constexpr int GS_ROWS_PER_WARP = GS_TY / (GS_THREADS / 32);

struct GsConsts {
  float Du, Dv, dt, dx, feed, kill;
};

// One cell of tau_gray_scott.cu:153-170, same operand order, same operators.
__device__ __forceinline__ void gs_point(float u, float v, float uR, float uL, float uP, float uM,
                                         float vR, float vL, float vP, float vM,
                                         const GsConsts &c, float &un, float &vn) {
  float lap_u = (uR + uL + uP + uM - 4.0f * u) / (c.dx * c.dx);
  float lap_v = (vR + vL + vP + vM - 4.0f * v) / (c.dx * c.dx);
  float uvv = u * v * v;
  float du = c.Du * lap_u - uvv + c.feed * (1.0f - u);
  float dv = c.Dv * lap_v + uvv - (c.feed + c.kill) * v;
  un = u + c.dt * du;
  vn = v + c.dt * dv;
}

__device__ __forceinline__ float4 lds_f4(const float *p) {
  return *reinterpret_cast<const float4 *>(p);
}

// TMA path: nx % 4 == 0.  Planes point at ghost row -1.
__global__ void __launch_bounds__(GS_THREADS)
gs_step_tma(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_v,
            const float *__restrict__ u_in, const float *__restrict__ v_in,
            float *__restrict__ u_out, float *__restrict__ v_out, int nx, int ny_local,
            GsConsts c, int wrap_y) {
  __shared__ alignas(128) float su[GS_SH * GS_SW];
  __shared__ alignas(128) float sv[GS_SH * GS_SW];
  __shared__ alignas(8) uint64_t bar;

  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * GS_TX;
  const int y0 = blockIdx.y * GS_TY;  // local row of the tile's first cell; plane row = y0 + 1

  if (tid == 0) {
    tau::mbar_init(&bar, 1);
    tau::mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    tau::mbar_expect_tx(&bar, 2u * GS_SH * GS_SW * sizeof(float));
    // box origin: column x0-4, plane row y0 (= local row y0-1)
    tau::tma_load_2d(su, &tm_u, x0 - GS_HX, y0, &bar);
    tau::tma_load_2d(sv, &tm_v, x0 - GS_HX, y0, &bar);
  }
  tau::mbar_wait(&bar, 0);

  // periodic wrap in x: only tiles touching the left/right domain edge patch their halo column
  const bool edgeL = (x0 == 0);
  const bool edgeR = (x0 + GS_TX >= nx);
  if (edgeL || edgeR) {
    const int rows = min(GS_SH, ny_local + 2 - y0);
    if (tid < GS_SH) {
      if (tid < rows) {
        const size_t g = (size_t)(y0 + tid) * nx;
        if (edgeL) {
          su[tid * GS_SW + GS_HX - 1] = u_in[g + nx - 1];
          sv[tid * GS_SW + GS_HX - 1] = v_in[g + nx - 1];
        }
        if (edgeR) {
          su[tid * GS_SW + (nx - x0) + GS_HX] = u_in[g];
          sv[tid * GS_SW + (nx - x0) + GS_HX] = v_in[g];
        }
      }
    }
    __syncthreads();
  }

  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int r0 = warp * GS_ROWS_PER_WARP;  // first tile row of this warp
  const int cx = GS_HX + 4 * lane;         // shared column of this thread's float4
  const int gx = x0 + 4 * lane;

  float4 uM = lds_f4(&su[(r0)*GS_SW + cx]);  // row above the first (tile row r0-1 -> smem row r0)
  float4 vM = lds_f4(&sv[(r0)*GS_SW + cx]);
  float4 uC = lds_f4(&su[(r0 + 1) * GS_SW + cx]);
  float4 vC = lds_f4(&sv[(r0 + 1) * GS_SW + cx]);

#pragma unroll
  for (int r = 0; r < GS_ROWS_PER_WARP; ++r) {
    const int srow = r0 + r + 1;  // smem row of the centre
    float4 uP = lds_f4(&su[(srow + 1) * GS_SW + cx]);
    float4 vP = lds_f4(&sv[(srow + 1) * GS_SW + cx]);

    float uLft = __shfl_up_sync(0xffffffffu, uC.w, 1);
    float vLft = __shfl_up_sync(0xffffffffu, vC.w, 1);
    float uRgt = __shfl_down_sync(0xffffffffu, uC.x, 1);
    float vRgt = __shfl_down_sync(0xffffffffu, vC.x, 1);
    if (lane == 0) {
      uLft = su[srow * GS_SW + GS_HX - 1];
      vLft = sv[srow * GS_SW + GS_HX - 1];
    }
    if (lane == 31) {
      uRgt = su[srow * GS_SW + GS_HX + GS_TX];
      vRgt = sv[srow * GS_SW + GS_HX + GS_TX];
    }

    float4 un, vn;
    gs_point(uC.x, vC.x, uC.y, uLft, uP.x, uM.x, vC.y, vLft, vP.x, vM.x, c, un.x, vn.x);
    gs_point(uC.y, vC.y, uC.z, uC.x, uP.y, uM.y, vC.z, vC.x, vP.y, vM.y, c, un.y, vn.y);
    gs_point(uC.z, vC.z, uC.w, uC.y, uP.z, uM.z, vC.w, vC.y, vP.z, vM.z, c, un.z, vn.z);
    gs_point(uC.w, vC.w, uRgt, uC.z, uP.w, uM.w, vRgt, vC.z, vP.w, vM.w, c, un.w, vn.w);

    const int y = y0 + r0 + r;  // local row
    if (gx < nx && y < ny_local) {
      const size_t o = (size_t)(y + 1) * nx + gx;
      tau::stg_stream_f4(u_out + o, un);
      tau::stg_stream_f4(v_out + o, vn);
      if (wrap_y) {  // keep the output plane's ghost rows periodic (single-GPU handles)
        if (y == 0) {
          const size_t og = (size_t)(ny_local + 1) * nx + gx;
          tau::stg_stream_f4(u_out + og, un);
          tau::stg_stream_f4(v_out + og, vn);
        }
        if (y == ny_local - 1) {
          tau::stg_stream_f4(u_out + gx, un);
          tau::stg_stream_f4(v_out + gx, vn);
        }
      }
    }
    uM = uC;
    vM = vC;
    uC = uP;
    vC = vP;
  }
}

// Generic path for any nx (the reference sizes the grid from the terminal, tau_gray_scott.cu:287-292,
// so odd widths are legal).  One thread per cell, ghost rows for y, compare-based wrap for x.
__global__ void __launch_bounds__(256)
gs_step_generic(const float *__restrict__ u_in, const float *__restrict__ v_in,
                float *__restrict__ u_out, float *__restrict__ v_out, int nx, int ny_local,
                GsConsts c, int wrap_y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= nx || j >= ny_local) return;
  const int ip = (i + 1 == nx) ? 0 : i + 1;
  const int im = (i == 0) ? nx - 1 : i - 1;
  const size_t row = (size_t)(j + 1) * nx;
  float un, vn;
  gs_point(u_in[row + i], v_in[row + i], u_in[row + ip], u_in[row + im], u_in[row + nx + i],
           u_in[row - nx + i], v_in[row + ip], v_in[row + im], v_in[row + nx + i],
           v_in[row - nx + i], c, un, vn);
  u_out[row + i] = un;
  v_out[row + i] = vn;
  if (wrap_y) {
    if (j == 0) {
      u_out[(size_t)(ny_local + 1) * nx + i] = un;
      v_out[(size_t)(ny_local + 1) * nx + i] = vn;
    }
    if (j == ny_local - 1) {
      u_out[i] = un;
      v_out[i] = vn;
    }
  }
}

// ghost rows <- periodic images (single-GPU handles, after init/upload)
__global__ void gs_fill_ghost_rows(float *u, float *v, int nx, int ny_local) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx) return;
  u[i] = u[(size_t)ny_local * nx + i];
  v[i] = v[(size_t)ny_local * nx + i];
  u[(size_t)(ny_local + 1) * nx + i] = u[(size_t)nx + i];
  v[(size_t)(ny_local + 1) * nx + i] = v[(size_t)nx + i];
}

}  // namespace

struct tau_gs {
  tau_gs_params p;
  int device;
  int y_begin, ny_local;
  bool slab;      // ny_local < ny: ghost rows are the caller's job
  bool use_tma;
  cudaStream_t stream;
  bool own_stream;
  float *u[2], *v[2];
  CUtensorMap tm_u[2], tm_v[2];
  int cur;
  long long steps, launches;
  cudaEvent_t ev0, ev1;
  bool timed;
};

extern "C" {

void tau_gs_default_params(tau_gs_params *p) {
  p->nx = 128;
  p->ny = 128;
  p->dx = 1.0f;
  p->dt = 1.0f;
  p->Du = 0.2f;
  p->Dv = 0.1f;
  p->feed = 0.03f;
  p->kill = 0.06f;
  p->seed = 1337u;
}

// Host restatement of init_pattern (tau_gray_scott.cu:173-204): u=1, v=0; centred square of
// half-width min(nx,ny)/12 set to (0.5, 0.25); 64 xorshift32 speckles set to (0.35, 0.65).
void tau_gs_init_pattern(float *u, float *v, int nx, int ny, unsigned seed) {
  const size_t n = (size_t)nx * ny;
  for (size_t k = 0; k < n; ++k) {
    u[k] = 1.0f;
    v[k] = 0.0f;
  }
  const int cx = nx / 2, cy = ny / 2, r = std::min(nx, ny) / 12;
  for (int j = -r; j <= r; ++j)
    for (int i = -r; i <= r; ++i) {
      const int x = (cx + i + nx) % nx, y = (cy + j + ny) % ny;
      u[(size_t)y * nx + x] = 0.50f;
      v[(size_t)y * nx + x] = 0.25f;
    }
  uint32_t s = seed ? seed : 1u;
  auto rng = [&]() {
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    return s;
  };
  for (int k = 0; k < 64; ++k) {
    const int x = (int)(rng() % (uint32_t)nx);
    const int y = (int)(rng() % (uint32_t)ny);
    u[(size_t)y * nx + x] = 0.35f;
    v[(size_t)y * nx + x] = 0.65f;
  }
}

int tau_gs_create(const tau_gs_params *p, int device, int y_begin, int ny_local, void *stream,
                  tau_gs **out) {
  TAU_REQUIRE(p && out, "tau_gs_create: null argument");
  TAU_REQUIRE(p->nx > 0 && p->ny > 0, "tau_gs_create: nx, ny must be positive (got %d x %d)",
              p->nx, p->ny);
  TAU_REQUIRE(y_begin >= 0 && ny_local > 0 && y_begin + ny_local <= p->ny,
              "tau_gs_create: slab rows [%d,%d) outside [0,%d)", y_begin, y_begin + ny_local,
              p->ny);
  if (tau_device_count() <= 0) {
    tau_set_error("tau_gs_create: no CUDA device (this library has no CPU fallback)");
    return TAU_ERR_NODEV;
  }
  TAU_CUDA(cudaSetDevice(device));
  tau_gs *h = new (std::nothrow) tau_gs();
  if (!h) return TAU_ERR_NOMEM;
  h->p = *p;
  h->device = device;
  h->y_begin = y_begin;
  h->ny_local = ny_local;
  h->slab = (ny_local != p->ny);
  h->use_tma = (p->nx % 4 == 0) && (p->nx >= 4);
  h->cur = 0;
  h->steps = 0;
  h->launches = 0;
  h->timed = false;
  if (stream) {
    h->stream = (cudaStream_t)stream;
    h->own_stream = false;
  } else {
    TAU_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  const size_t bytes = (size_t)p->nx * (ny_local + 2) * sizeof(float);
  for (int b = 0; b < 2; ++b) {
    TAU_CUDA(cudaMalloc(&h->u[b], bytes));
    TAU_CUDA(cudaMalloc(&h->v[b], bytes));
    TAU_CUDA(cudaMemsetAsync(h->u[b], 0, bytes, h->stream));
    TAU_CUDA(cudaMemsetAsync(h->v[b], 0, bytes, h->stream));
  }
  if (h->use_tma) {
    const uint64_t dims[2] = {(uint64_t)p->nx, (uint64_t)(ny_local + 2)};
    const uint64_t strides[1] = {(uint64_t)p->nx * sizeof(float)};
    const uint32_t box[2] = {GS_SW, GS_SH};
    for (int b = 0; b < 2; ++b) {
      int rc = tau_make_tensor_map(&h->tm_u[b], h->u[b], 4, 2, dims, strides, box);
      if (rc) return rc;
      rc = tau_make_tensor_map(&h->tm_v[b], h->v[b], 4, 2, dims, strides, box);
      if (rc) return rc;
    }
  }
  TAU_CUDA(cudaEventCreate(&h->ev0));
  TAU_CUDA(cudaEventCreate(&h->ev1));
  *out = h;
  return TAU_OK;
}

static int gs_after_upload(tau_gs *h) {
  if (!h->slab) {
    gs_fill_ghost_rows<<<(h->p.nx + 255) / 256, 256, 0, h->stream>>>(h->u[h->cur], h->v[h->cur],
                                                                      h->p.nx, h->ny_local);
    h->launches++;
    TAU_CUDA(cudaGetLastError());
  }
  return TAU_OK;
}

int tau_gs_upload(tau_gs *h, const float *u, const float *v) {
  TAU_REQUIRE(h && u && v, "tau_gs_upload: null argument");
  TAU_CUDA(cudaSetDevice(h->device));
  const size_t bytes = (size_t)h->p.nx * h->ny_local * sizeof(float);
  TAU_CUDA(cudaMemcpyAsync(h->u[h->cur] + h->p.nx, u, bytes, cudaMemcpyHostToDevice, h->stream));
  TAU_CUDA(cudaMemcpyAsync(h->v[h->cur] + h->p.nx, v, bytes, cudaMemcpyHostToDevice, h->stream));
  int rc = gs_after_upload(h);
  if (rc) return rc;
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

int tau_gs_init(tau_gs *h) {
  TAU_REQUIRE(h, "tau_gs_init: null handle");
  const size_t n = (size_t)h->p.nx * h->p.ny;
  std::vector<float> u(n), v(n);
  tau_gs_init_pattern(u.data(), v.data(), h->p.nx, h->p.ny, h->p.seed);
  const size_t off = (size_t)h->y_begin * h->p.nx;
  h->steps = 0;
  return tau_gs_upload(h, u.data() + off, v.data() + off);
}

int tau_gs_step(tau_gs *h, int nsteps) {
  TAU_REQUIRE(h, "tau_gs_step: null handle");
  TAU_REQUIRE(nsteps >= 0, "tau_gs_step: nsteps must be >= 0");
  TAU_REQUIRE(!(h->slab && nsteps > 1),
              "tau_gs_step: a slab handle advances one step per call (ghost rows must be "
              "exchanged in between)");
  TAU_CUDA(cudaSetDevice(h->device));
  const GsConsts c{h->p.Du, h->p.Dv, h->p.dt, h->p.dx, h->p.feed, h->p.kill};
  const int wrap_y = h->slab ? 0 : 1;
  TAU_CUDA(cudaEventRecord(h->ev0, h->stream));
  for (int s = 0; s < nsteps; ++s) {
    const int a = h->cur, b = a ^ 1;
    if (h->use_tma) {
      dim3 grid((h->p.nx + GS_TX - 1) / GS_TX, (h->ny_local + GS_TY - 1) / GS_TY);
      gs_step_tma<<<grid, GS_THREADS, 0, h->stream>>>(h->tm_u[a], h->tm_v[a], h->u[a], h->v[a],
                                                       h->u[b], h->v[b], h->p.nx, h->ny_local, c,
                                                       wrap_y);
    } else {
      dim3 block(32, 8), grid((h->p.nx + 31) / 32, (h->ny_local + 7) / 8);
      gs_step_generic<<<grid, block, 0, h->stream>>>(h->u[a], h->v[a], h->u[b], h->v[b], h->p.nx,
                                                     h->ny_local, c, wrap_y);
    }
    h->launches++;
    h->cur = b;
    h->steps++;
  }
  TAU_CUDA(cudaEventRecord(h->ev1, h->stream));
  h->timed = true;
  TAU_CUDA(cudaGetLastError());
  return TAU_OK;
}

int tau_gs_download(tau_gs *h, float *u, float *v) {
  TAU_REQUIRE(h && u && v, "tau_gs_download: null argument");
  TAU_CUDA(cudaSetDevice(h->device));
  const size_t bytes = (size_t)h->p.nx * h->ny_local * sizeof(float);
  TAU_CUDA(cudaMemcpyAsync(u, h->u[h->cur] + h->p.nx, bytes, cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaMemcpyAsync(v, h->v[h->cur] + h->p.nx, bytes, cudaMemcpyDeviceToHost, h->stream));
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

int tau_gs_sync(tau_gs *h) {
  TAU_REQUIRE(h, "tau_gs_sync: null handle");
  TAU_CUDA(cudaStreamSynchronize(h->stream));
  return TAU_OK;
}

int tau_gs_device_planes(tau_gs *h, float **u, float **v) {
  TAU_REQUIRE(h && u && v, "tau_gs_device_planes: null argument");
  *u = h->u[h->cur];
  *v = h->v[h->cur];
  return TAU_OK;
}

long long tau_gs_steps_done(tau_gs *h) { return h ? h->steps : -1; }
long long tau_gs_launch_count(tau_gs *h) { return h ? h->launches : -1; }

int tau_gs_last_step_ms(tau_gs *h, float *ms) {
  TAU_REQUIRE(h && ms, "tau_gs_last_step_ms: null argument");
  TAU_REQUIRE(h->timed, "tau_gs_last_step_ms: no step has been timed yet");
  TAU_CUDA(cudaEventSynchronize(h->ev1));
  TAU_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return TAU_OK;
}

int tau_gs_destroy(tau_gs *h) {
  if (!h) return TAU_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (int b = 1; b >= 0; --b) {
    cudaFree(h->v[b]);
    cudaFree(h->u[b]);
  }
  cudaEventDestroy(h->ev1);
  cudaEventDestroy(h->ev0);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return TAU_OK;
}

}  // extern "C"
