// hypersonic2d_fused.cuh — EXPERIMENTAL: ONE step kernel for pair mode (TAU_HYP2D_PAIR=2 at handle creation).
// NOT YET RUN ON HARDWARE; the default path does not touch it.  Verified on the CPU emulator like the pair
// kernel (tests/test_hostemu_cpu.py).
//
// TAU_HYP2D_PAIR=1 launches two kernels per step — hyp2d_step_pair on the interior body-free 60-column items,
// then hyp2d_step on the rest — and the second cannot overlap the first (both are ordered behind the previous
// step), so its ramp and tail are paid per step: a few microseconds that matter at 512-row slabs (N = 8).
// This kernel claims both kinds of item from ONE table (bit 30 of the descriptor marks a pair item) and is
// nothing but the included text of the two marches around the production kernel's prologue and epilogue, so
// it owns the step's bookkeeping and the multi-GPU message like hyp2d_step does.  The shared-memory ring of a
// warp is sized for the larger (pair) slots; the production march uses the front of each.
#pragma once

namespace {

constexpr unsigned HF_PAIR_BIT = 0x40000000u;

__global__ void __launch_bounds__(H2_WARPS * 32, HP_MIN_CTAS)
hyp2d_step_fused(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmPair,
                 const Params<float> P, const float *__restrict__ Uin, float *__restrict__ Uout,
                 const uint8_t *__restrict__ mask, const uint2 *__restrict__ items, Ctrl *__restrict__ ctrl,
                 int step_slot, const PeerPush peer) {
  using R = float;
  constexpr bool USE_TMA = true;
#define H2_WARP_RING_ELEMS (H2_NS * HP_SLOT)
#include "hypersonic2d_prologue.inc"
#undef H2_WARP_RING_ELEMS
  while (item < (unsigned)P.nitems) {
    if (desc.x & HF_PAIR_BIT) {
      const Ring2 ring2{ring_base};
#define HP_TM tmPair
#define HP_NITEMS P.nitems
#define HP_CLAIM_CTR ctrl->next_item
#define HP_HALF_DT f2(half_dt)
#define HP_RING ring2
#define HP_RING_BASE ring_base
#include "hypersonic2d_pair_item.inc"
#undef HP_TM
#undef HP_NITEMS
#undef HP_CLAIM_CTR
#undef HP_HALF_DT
#undef HP_RING
#undef HP_RING_BASE
    } else {
#include "hypersonic2d_item.inc"
    }
  }  // work-item loop
#include "hypersonic2d_epilogue.inc"
}

}  // namespace
