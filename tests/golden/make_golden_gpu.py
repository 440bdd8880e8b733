"""Generate golden fixtures from the REFERENCE's own kernels (oracle/_ref) on a GPU box.

    gpurun -- python tests/golden/make_golden_gpu.py        # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/                 # then commit

The reference holds no golden vectors for these solvers; these files pin the
CPU oracle (oracle/*.c) to what the reference code itself computes on a B200.  Inputs are seeded;
each file records the generating arguments.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)


def gs():
    # (a) reference init pattern + 200 steps at defaults, 96x64
    u0, v0 = oracle.ref_gs_init_pattern(96, 64, 1337)
    u1, v1 = oracle.ref_gs_run(u0, v0, 200)
    # (b) random field, non-default coefficients, odd sizes, dx = 0.5 (power of two)
    rng = np.random.default_rng(7)
    ur = rng.random((45, 70), dtype=np.float32)
    vr = (rng.random((45, 70), dtype=np.float32) * 0.5).astype(np.float32)
    kw = dict(Du=0.16, Dv=0.08, dt=0.25, dx=0.5, feed=0.0367, kill=0.0649)
    ur1, vr1 = oracle.ref_gs_run(ur, vr, 33, **kw)
    # (c) long run where v decays into the flush-to-zero range
    ul, vl = oracle.ref_gs_run(u0, v0 * np.float32(1e-30), 400)
    np.savez_compressed(os.path.join(OUT, "gs_ref.npz"), u0=u0, v0=v0, u200=u1, v200=v1, ur=ur,
                        vr=vr, ur33=ur1, vr33=vr1, kw=np.array(list(kw.values()), np.float32),
                        ul400=ul, vl400=vl)
    print("gs golden written")


def hyp2d():
    for (W, H, steps) in ((256, 128, 60), (200, 120, 25)):
        r = oracle.ref(f"ref_hyp2d_{W}x{H}")
        cfg11 = np.zeros(11)
        r.ref_hyp2d_default_config.argtypes = [oracle.f64p]
        r.ref_hyp2d_default_config(cfg11)
        p0, mask, _, _, _ = oracle.ref_hyp2d_run(W, H, cfg11, 0)
        out = {"cfg11": cfg11, "mask": mask, "steps": np.array(steps)}
        for k, p in zip(("rho0", "mx0", "my0", "E0"), p0):
            out[k] = p
        p1, _, t, dts, _ = oracle.ref_hyp2d_run(W, H, cfg11, steps)
        for k, p in zip(("rho", "mx", "my", "E"), p1):
            out[k] = p
        out["sim_t"] = np.array(t)
        out["dts"] = dts
        # a second configuration that exercises walls harder: lower Mach, bigger body
        cfg_b = cfg11.copy()
        cfg_b[5] = 3.0           # inflow_mach
        cfg_b[6] = 40.0          # geom_x0
        p2, mask_b, t2, dts2, _ = oracle.ref_hyp2d_run(W, H, cfg_b, steps)
        out["cfg11_b"] = cfg_b
        out["mask_b"] = mask_b
        for k, p in zip(("rho_b", "mx_b", "my_b", "E_b"), p2):
            out[k] = p
        out["sim_t_b"] = np.array(t2)
        np.savez_compressed(os.path.join(OUT, f"hyp2d_ref_{W}x{H}.npz"), **out)
    # device-helper vectors: HLLC on random state pairs, limited reconstruction on random triples
    rng = np.random.default_rng(11)
    n = 4096
    cfg11 = np.zeros(11)
    r = oracle.ref("ref_hyp2d_256x128")
    r.ref_hyp2d_default_config(cfg11)
    prim = np.empty((n, 2, 4))
    prim[..., 0] = rng.uniform(0.05, 8.0, (n, 2))
    prim[..., 1] = rng.normal(0, 12.0, (n, 2))
    prim[..., 2] = rng.normal(0, 12.0, (n, 2))
    prim[..., 3] = rng.uniform(0.02, 60.0, (n, 2))
    prim[: n // 8, 1, :] = prim[: n // 8, 0, :]          # equal states
    prim[n // 8: n // 4, :, 1] += 40.0                   # supersonic to the right
    g = cfg11[0]
    cons = np.empty((n, 8))
    for s in range(2):
        rho, u, v, p = (prim[:, s, k] for k in range(4))
        cons[:, 4 * s + 0] = rho
        cons[:, 4 * s + 1] = rho * u
        cons[:, 4 * s + 2] = rho * v
        cons[:, 4 * s + 3] = p / (g - 1) + 0.5 * rho * (u * u + v * v)
    hllc = oracle.ref_hyp2d_eval(256, 128, cfg11, 0, cons)
    tri = np.empty((n, 12))
    base = rng.uniform(0.1, 5.0, (n, 4))
    for k in range(3):
        tri[:, 4 * k:4 * k + 4] = base + rng.normal(0, 0.8, (n, 4)) * (rng.random((n, 1)) < 0.8)
    tri[: n // 16, 0] = -0.5   # force the positivity repair loop
    recon = oracle.ref_hyp2d_eval(256, 128, cfg11, 1, tri)
    np.savez_compressed(os.path.join(OUT, "hyp2d_ref_helpers.npz"), cfg11=cfg11, hllc_in=cons,
                        hllc_out=hllc, recon_in=tri, recon_out=recon)
    print("hyp2d golden written")


def hyp3d():
    # the reference's own k_step from k_init, small grids (block 8x8x4 does not divide 20 / 28)
    for (nx, ny, nz, steps) in ((32, 28, 20, 30),):
        prm = oracle.hyp3d_params(nx, ny, nz)
        p0, solid, _, _, _, _ = oracle.ref_hyp3d_run(prm, 0)
        p1, _, clock, dts, mx, _ = oracle.ref_hyp3d_run(prm, steps)
        out = {"steps": np.array(steps), "solid": solid, "clock": np.array(clock), "dts": dts, "maxs": mx}
        for k, a, b in zip(("xi", "phix", "phiy", "phiz", "lam", "zet"), p0, p1):
            out[k + "0"] = a
            out[k] = b
        # a developed state: start late in the inflow ramp so that shocks and the sponge matter
        p2, _, clock2, dts2, mx2, _ = oracle.ref_hyp3d_run(prm, steps, planes=p1, clock=(0.015, 2e-3))
        for k, b in zip(("xi", "phix", "phiy", "phiz", "lam", "zet"), p2):
            out[k + "_b"] = b
        out["clock_b"] = np.array(clock2)
        out["dts_b"] = dts2
        np.savez_compressed(os.path.join(OUT, f"hyp3d_ref_{nx}x{ny}x{nz}.npz"), **out)
    print("hyp3d golden written")


def sph():
    # the reference's own kernels (linked-list neighbour search) on its own initial lattice
    out = {}
    for tag, N, frames, kw in (("a", 4096, 10, {}), ("b", 6000, 8, dict(useXSPH=1, rain=0, viscSub=2))):
        prm = oracle.sph_params(N, **kw)
        pos0, vel0 = oracle.ref_sph_reset_particles(prm)
        pos, vel, acc, s, pr, ck, _ = oracle.ref_sph_run(prm, pos0, vel0, frames)
        out.update({f"pos0_{tag}": pos0, f"vel0_{tag}": vel0, f"pos_{tag}": pos, f"vel_{tag}": vel,
                    f"s_{tag}": s, f"press_{tag}": pr, f"clock_{tag}": ck, f"p19_{tag}": prm.as19(),
                    f"frames_{tag}": np.array(frames)})
    np.savez_compressed(os.path.join(OUT, "sph_ref.npz"), **out)
    print("sph golden written")


def burgers():
    # the reference's own kernels where they are deterministic (nu = 0: viscosity_step does not mix cells)
    # plus one viscous 1-D Cole-Hopf case (ny = 1: a single block row; the in-place race is confined to
    # block seams) and the render-free init of each
    out = {}
    cases = (("a", dict(nx=96, ny=64, dtau=1e-3, nu=0.0, swirl=0.2, amp=0.3), 40),
             ("b", dict(nx=96, ny=64, dtau=1e-3, nu=0.0, swirl=0.2, amp=0.3, muscl=1), 40),
             ("c", dict(nx=96, ny=64, dtau=1e-3, nu=0.0), 10),
             ("d", dict(nx=300, colehopf=1, dtau=5e-3, t0=1e-3, nu=0.0, ck=2), 200))
    for tag, kw, steps in cases:
        prm = oracle.burgers_params(**kw)
        u0, v0 = oracle.ref_burgers_init(prm)
        u, v, ck, dts, _ = oracle.ref_burgers_run(prm, u0, v0, steps)
        out.update({f"u0_{tag}": u0, f"v0_{tag}": v0, f"u_{tag}": u, f"v_{tag}": v, f"clock_{tag}": np.array(ck),
                    f"dts_{tag}": dts, f"p22_{tag}": prm.as22(), f"steps_{tag}": np.array(steps)})
    np.savez_compressed(os.path.join(OUT, "burgers_ref.npz"), **out)
    print("burgers golden written")


def sw():
    # the reference's own shallow-water kernels (nu = 0: no racy viscosity_uv); complements the host-emulated
    # fixture sw_ref_host.npz (make_golden_host.py) with the -use_fast_math device intrinsics
    out = {}
    gentle = dict(H0=2.0, bumpAmp=0.4, bumpSigma=5, asym=0.3, swirl=0.05, swirlRc=10, offx=3, offy=-2)
    cases = (("a", dict(nx=96, ny=64, dtau=0.02, nu=0.0, **gentle), 40),
             ("b", dict(nx=70, ny=37, dtau=0.05, nu=0.0, dx=2.0, dy=1.5, **gentle), 25),
             ("c", dict(nx=96, ny=64, dtau=1e-3, nu=0.0), 10))
    for tag, kw, steps in cases:
        prm = oracle.sw_params(**kw)
        s0, u0, v0 = oracle.ref_sw_init(prm)
        s, u, v, ck, dts, _ = oracle.ref_sw_run(prm, s0, u0, v0, steps)
        out.update({f"s0_{tag}": s0, f"u0_{tag}": u0, f"v0_{tag}": v0, f"s_{tag}": s, f"u_{tag}": u, f"v_{tag}": v,
                    f"clock_{tag}": np.array(ck), f"dts_{tag}": dts, f"p19_{tag}": prm.as19(),
                    f"steps_{tag}": np.array(steps)})
    np.savez_compressed(os.path.join(OUT, "sw_ref.npz"), **out)
    print("shallow-water golden written")


if __name__ == "__main__":
    which = sys.argv[1:] or ["gs", "hyp2d", "hyp3d", "sph", "burgers"]
    for w in which:
        globals()[w]()
