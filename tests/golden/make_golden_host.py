"""Golden vectors that need NO GPU: the reference's own kernel bodies compiled for the host (g++ against
oracle/shims/hostcuda/cuda_runtime.h) and emulated thread by thread by oracle/ref_drivers/ref_*_host.cpp.
Run where /root/reference exists:   make -C oracle ref && python tests/golden/make_golden_host.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

SW_CASES = (
    ("a", dict(nx=96, ny=64, nu=0.0, dtau=1e-3), 40),                                   # the default (violent) field
    ("b", dict(nx=100, ny=70, nu=0.0, dtau=1.0, offx=10, offy=-5, swirlRc=12, bumpSigma=6, bumpAmp=50), 40),
    ("c", dict(nx=48, ny=40, nu=0.0, dtau=0.05, H0=1.0, bumpAmp=0.5, bumpSigma=4, asym=0.3, swirl=0.2, swirlRc=8,
               offx=0, offy=0, dx=2.0, dy=1.5), 60),
    ("d", dict(nx=37, ny=19, nu=0.0, dtau=0.02, H0=5.0, bumpAmp=2.0, bumpSigma=3, asym=0.0, swirl=0.0, offx=3,
               offy=-2), 50),
)


def sw():
    out = {}
    for tag, kw, steps in SW_CASES:
        prm = oracle.sw_params(**kw)
        s0, u0, v0 = oracle.ref_sw_host_init(prm)
        s, u, v, ck, dts = oracle.ref_sw_host_run(prm, s0, u0, v0, steps)
        out.update({f"s0_{tag}": s0, f"u0_{tag}": u0, f"v0_{tag}": v0, f"s_{tag}": s, f"u_{tag}": u, f"v_{tag}": v,
                    f"clock_{tag}": np.array(ck), f"dts_{tag}": dts, f"p19_{tag}": prm.as19(),
                    f"steps_{tag}": np.array(steps)})
    np.savez_compressed(os.path.join(OUT, "sw_ref_host.npz"), **out)
    print("shallow-water host golden written")


TH3CS_N, TH3CS_FRAMES = 24, 48


def th3cs():
    # the reference's whole .4spl exporter (th3cs.cu main(): k_init, 4 x k_step per frame with the host d_tau
    # controller, k_schlieren_export, host min/max + palette index loop) on the CPU emulator, 24^3 x 48 frames
    hdr, pal, idx = oracle.ref_th3cs_host_run(TH3CS_N, TH3CS_FRAMES)
    np.savez_compressed(os.path.join(OUT, "th3cs_ref_host.npz"), header=np.array(list(hdr.values()), np.uint32),
                        palette=pal, indices=idx)
    print("th3cs host golden written", hdr)


def hypcpu():
    # BASELINE config 1: the reference's CPU solver tau_hypersonic.c compiled at 256 x 256 with its own flags
    # (oracle/_ref/libref_hypcpu_256x256.so), init_sim + 10 warm-up + 200 steps of step_physics
    r = oracle.RefHypCpu(256, 256)
    r.init()
    p0, mask = r.get()
    r.steps(10)
    t10 = r.sim_t
    r.steps(200)
    p210, _ = r.get()
    np.savez_compressed(os.path.join(OUT, "hypcpu_ref_256x256.npz"), mask=mask, rho0=p0[0], E0=p0[3],
                        rho=p210[0], mx=p210[1], my=p210[2], E=p210[3], sim_t=np.array([t10, r.sim_t]))
    print("tau_hypersonic.c 256x256 golden written", t10, r.sim_t)


if __name__ == "__main__":
    for w in (sys.argv[1:] or ["sw", "th3cs", "hypcpu"]):
        globals()[w]()
