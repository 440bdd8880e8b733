"""TEST INFRASTRUCTURE ONLY — the stripe-sharded SPH solver (fluid_sims_b200/csrc/sph_stripes.inc) run on the CPU emulator:
`world` rank handles live in ONE process and step in lock-step; the neighbour exchanges are plain memmoves between the
handles' message buffers ("device" memory is host memory in the emulator), the histogram all-reduce a numpy sum.  The result
must equal the single-GPU handle (same emulated library) BIT FOR BIT.  Run with TAU_B200_LIB=build/hostemu/libsph_hostemu.so
(tests/test_hostemu_cpu.py does, in a subprocess); prints one JSON line."""
import ctypes as C
import json
import sys

import numpy as np

from fluid_sims_b200 import sph as S


def lockstep(handles, nsub, rebalance_every):
    w = len(handles)
    xsph = handles[0].params.useXSPH and handles[0].params.xsphEps > 0

    def exchange(base):
        for r, h in enumerate(handles):          # r's send_hi -> (r+1)'s recv_lo ; r's send_lo -> (r-1)'s recv_hi
            if r + 1 < w:
                src, dst = h._bufs[base + 1], handles[r + 1]._bufs[base + 2]
                C.memmove(dst[0], src[0], 4 * src[1])
            if r > 0:
                src, dst = h._bufs[base], handles[r - 1]._bufs[base + 3]
                C.memmove(dst[0], src[0], 4 * src[1])

    for k in range(nsub):
        for phase, base in ((0, 0), (1, 4), (2, 0 if xsph else None), (3, None)):
            for h in handles:
                S.check(S._st_phase(h._handle, phase))
            if base is not None and w > 1:
                exchange(base)
        if w > 1 and rebalance_every and (k + 1) % rebalance_every == 0:
            ptrs = []
            for h in handles:
                p, n = C.c_void_p(), C.c_int()
                S.check(S._st_hist_begin(h._handle, C.byref(p), C.byref(n)))
                ptrs.append((p.value, n.value))
            hists = [np.ctypeslib.as_array((C.c_int * n).from_address(p)) for p, n in ptrs]
            total = np.sum(hists, axis=0).astype(np.int32)
            assert int(total.sum()) == handles[0].params.N, "owned sets are not a partition of the particles"
            for hs in hists:
                hs[:] = total
            for h in handles:
                S.check(S._st_hist_apply(h._handle))


def main():
    N, world, frames = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    over = json.loads(sys.argv[4]) if len(sys.argv) > 4 else {}
    rebalance_every = over.pop("rebalance_every", 4)
    P = S.Params(N=N, **over)
    pos0, vel0 = S.reset_particles(P)
    one = S.SPH(P).upload(pos0, vel0)
    one.step(frames)
    ref = one.download()
    handles = [S.SPHStripes(P, r, world).upload(pos0, vel0) for r in range(world)]
    K = P.viscSub if P.viscSub > 0 else 1
    lockstep(handles, frames * K, rebalance_every)
    st = [h.status() for h in handles]
    got = S.assemble([h.download_local() for h in handles], N)
    moved = float(np.abs(ref[0] - pos0).max())
    print(json.dumps({
        "pos_equal": bool(np.array_equal(got[0], ref[0])), "vel_equal": bool(np.array_equal(got[1], ref[1])),
        "s_max_diff": float(np.abs(got[2] - ref[2]).max()), "s_mismatches": int((got[2] != ref[2]).sum()),
        "press_mismatches": int((got[3] != ref[3]).sum()), "clock_equal": handles[0].clock() == one.clock(),
        "status": st, "moved": moved,
        "max_pos_diff": float(np.abs(got[0] - ref[0]).max())}))


if __name__ == "__main__":
    main()
