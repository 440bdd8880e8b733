#!/usr/bin/env python
"""bench_all.py — secondary benchmarks of the other BASELINE configs (the headline is bench.py):

  config 3  tau_gray_scott 8192x8192                  -> Mcell-updates/s, HBM roofline (16 B/cell)
  config 4  tau_hypersonic_3d_cuda n^3 (default 256, --n3 512; z-slabs under torchrun)
                                                        -> Mcell-updates/s, roofline at 49 B/cell
  config 5  tau_sph 2M particles                       -> Mparticle-updates/s (per sub-step)
  (next)    tau_burgers 4096x4096 (only when named: `bench_all.py burgers`)

Each line also carries `reference_gpu`: the reference's own kernels (oracle/_ref, recompiled for
sm_100a) timed on the same GPU on the same problem — test infrastructure used as a yardstick, never
as the product path.  One JSON object per line on stdout.

  python bench_all.py [gs] [hyp3d] [sph] [burgers] [--steps K]
  torchrun --nproc-per-node 8 bench_all.py hyp3d --n3 512        # config 4 proper
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


P31 = (1 << 31) - 1


def crc_hex(vals):
    import zlib
    return f"{zlib.crc32(repr([int(v) for v in vals]).encode()) & 0xFFFFFFFF:08x}"


def particles_crc(ids, pos, vel):
    """order-independent checksum of a particle set keyed by global id: partial sums add across ranks"""
    import numpy as np
    w = (ids.astype(np.int64) % 65521) + 1
    out = []
    for arr in (pos, vel):
        bits = np.ascontiguousarray(arr, np.float32).view(np.uint32).astype(np.int64)
        out += [int(((bits[:, 0] * w) % P31).sum()), int(((bits[:, 1] * (w + 7)) % P31).sum())]
    return out


def bench_gs(a):
    import oracle
    from fluid_sims_b200.gray_scott import GrayScott, Params
    n = a.gs_n
    g = GrayScott(Params(nx=n, ny=n)).init()
    g.step(20)
    g.sync()
    g.step(a.steps)
    ms = g.last_step_ms() / a.steps
    cells = n * n
    ach = 16 * cells / (ms * 1e-3) / 1e9
    ref_ms = None
    if oracle.has_ref("ref_gs"):
        import ctypes as C
        r = oracle.ref("ref_gs")
        r.ref_gs_time.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
        t = C.c_float()
        r.ref_gs_time(n, n, min(a.steps, 200), C.byref(t))
        ref_ms = t.value / min(a.steps, 200)
    return ({"bench": "gray_scott", "grid": [n, n], "steps": a.steps, "ms_per_step": ms,
                      "value": cells / (ms * 1e-3) / 1e6, "unit": "Mcell-updates/s",
                      "roofline": {"bound": "hbm", "achieved": ach, "peak": peak(), "unit": "GB/s",
                                   "frac": ach / peak(), "algorithmic_bytes_per_cell": 16},
                      "reference_gpu": {"ms_per_step": ref_ms,
                                        "value": cells / (ref_ms * 1e-3) / 1e6 if ref_ms else None,
                                        "what": "tau_gray_scott.cu step_kernel recompiled for sm_100a"},
                      "gpu_launches": g.launch_count})


def bench_hyp3d(a):
    import numpy as np
    import torch

    import oracle
    from fluid_sims_b200 import slab
    from fluid_sims_b200.hypersonic3d import HALO, Hypersonic3D, Params
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.n3
    prm = Params.default(n, n, n)
    z0, nl = slab.partition_rows(n, world)[rank]
    ts = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(ts)
    sim = Hypersonic3D(prm, device=local, z_begin=z0, nz_local=nl, stream=ts.cuda_stream).init()
    # start late in the inflow ramp so that the bow shock forms within the warm-up
    p0, _ = sim.download()
    sim.upload(p0, (5e-3, 2e-3))
    del p0
    views = {}

    def advance(k):
        if world == 1:
            sim.step(k)
            return
        for _ in range(k):
            pp, mp = sim.device_state()
            if pp not in views:
                views[pp] = slab.wrap_plane(pp, (6, nl + 2 * HALO, n, n), torch.float32, local)
            if mp not in views:
                views[mp] = slab.wrap_plane(mp, (1,), torch.float32, local)
            slab.exchange_halos([views[pp]], HALO, periodic=True, dim=1)
            sim.step_begin()
            dist.all_reduce(views[mp], op=dist.ReduceOp.MAX)
            sim.step_end()

    advance(a.warm3)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    advance(a.steps3)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps3
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # checksum of the owned planes (weights by global cell index): equal across N <=> the z-slab ring is bit-identical
    pp, _ = sim.device_state()
    v = slab.wrap_plane(pp, (6, nl + 2 * HALO, n, n), torch.float32, local)[:, HALO:HALO + nl]
    bits = v.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    gz = torch.arange(z0, z0 + nl, device=bits.device, dtype=torch.int64).view(1, nl, 1, 1)
    gy = torch.arange(n, device=bits.device, dtype=torch.int64).view(1, 1, n, 1)
    gx = torch.arange(n, device=bits.device, dtype=torch.int64).view(1, 1, 1, n)
    wgt = ((gz * n + gy) * n + gx) % 65521 + 1
    acc = torch.stack([bits.sum(dim=(1, 2, 3)), ((bits * wgt).sum(dim=3) % P31).sum(dim=(1, 2))])
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    acc[1] %= P31
    crc = crc_hex(acc.flatten().tolist())
    del v, bits, wgt
    cells = n ** 3
    ach = 49 * (cells / world) / (ms * 1e-3) / 1e9
    ref = None
    if world == 1 and oracle.has_ref("ref_hyp3d") and n <= 512:
        op = oracle.hyp3d_params(n, n, n)
        *_, rms = oracle.ref_hyp3d_run(op, a.steps3, clock=(5e-3, 2e-3))
        ref = rms / a.steps3
    if rank == 0:
        return ({"bench": "hypersonic3d", "grid": [n, n, n], "n_gpus": world,
                          "steps": a.steps3, "ms_per_step": ms,
                          "value": cells / (ms * 1e-3) / 1e6, "unit": "Mcell-updates/s",
                          "roofline": {"bound": "hbm", "achieved": ach, "peak": peak(), "unit": "GB/s",
                                       "frac": ach / peak(), "algorithmic_bytes_per_cell": 49,
                                       "note": "compute/MUFU bound by >10x (WENO5+HLLC), see DESIGN.md"},
                          "reference_gpu": {"ms_per_step": ref,
                                            "value": cells / (ref * 1e-3) / 1e6 if ref else None,
                                            "what": "tau_hypersonic_3d_cuda.cu k_step recompiled for "
                                                    "sm_100a incl. its 2 blocking 4-byte copies per step"},
                          "clock": sim.clock(), "state_crc": crc, "steps_from_init": a.warm3 + a.steps3,
                          "parallelism": f"z-slab ring x{world}"})
    return None


def bench_sph(a):
    import torch

    import oracle
    from fluid_sims_b200 import slab
    from fluid_sims_b200.sph import SPH, Params, reset_particles
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = a.sph_n
    P = Params(N=N)
    ts = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(ts)
    if world == 1:
        s = SPH(P, device=local, stream=ts.cuda_stream).init()
        advance = s.step
    else:   # hash-bin stripes: every rank holds its stripe only; ghost rows + migrants by NCCL send / recv
        from fluid_sims_b200.sph import SPHStripes, nccl_plumbing
        exchange, allreduce_sum = nccl_plumbing(local)
        s = SPHStripes(P, rank, world, device=local, stream=ts.cuda_stream, exchange=exchange,
                       allreduce_sum=allreduce_sum).init()

        def advance(k):
            for _ in range(k):
                s.substep()

    advance(5)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    import time
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    w0 = time.perf_counter()
    advance(a.steps_sph)
    host_ms = (time.perf_counter() - w0) * 1e3 / a.steps_sph     # time the host needs to ENQUEUE a sub-step
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps_sph
    stripes = None
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        stripes = [None] * world
        dist.all_gather_object(stripes, s.status())
    import numpy as np
    if world == 1:
        pos, vel, _, _ = s.download()
        part = particles_crc(np.arange(N, dtype=np.uint32), pos, vel)
    else:
        ids, pos, vel, _, _ = s.download_local()
        part = particles_crc(ids, pos, vel)
        t = torch.tensor(part, device=f"cuda:{local}", dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        part = [int(x) for x in t.tolist()]
    crc = crc_hex([x % P31 for x in part])
    ref = None
    if world == 1 and oracle.has_ref("ref_sph"):
        pos0, vel0 = reset_particles(P)
        r = oracle.ref_sph_run(oracle.sph_params(N), pos0, vel0, a.steps_sph)
        ref = r[6] / a.steps_sph
    ach = 72 * N / (ms * 1e-3) / 1e9
    if rank == 0:
        return ({"bench": "sph", "particles": N, "n_gpus": world, "substeps": a.steps_sph,
                          "ms_per_substep": ms, "value": N / (ms * 1e-3) / 1e6,
                          "unit": "Mparticle-updates/s",
                          "roofline": {"bound": "hbm", "achieved": ach, "peak": peak(), "unit": "GB/s",
                                       "frac": ach / peak() / world, "algorithmic_bytes_per_particle": 72,
                                       "note": "L2-resident working set, instruction bound; HBM "
                                               "fraction is informational (SURVEY 8(d))"},
                          "reference_gpu": {"ms_per_substep": ref,
                                            "value": N / (ref * 1e-3) / 1e6 if ref else None,
                                            "what": "tau_sph.cu kernels recompiled for sm_100a"},
                          "parallelism": ("single GPU" if world == 1 else
                                          "hash-bin stripes x%d, ghost-row + migrant exchange by NCCL send/recv" % world),
                          **({"stripes": [{k: x[k] for k in ("n_own", "n_ghost", "err", "max_send", "row_begin", "row_end")}
                                          for x in stripes]} if stripes else {}),
                          "state_crc": crc, "substeps_from_init": 5 + a.steps_sph, "host_enqueue_ms_per_substep": host_ms,
                          "gpu_launches": s.launch_count})
    return None


def bench_burgers(a):
    """tau_burgers n x n (periodic), reference defaults (first-order Rusanov, K = 1)."""
    import oracle
    from fluid_sims_b200.burgers import Burgers, Params, initialize_host
    n = a.burgers_n
    P = Params(nx=n, ny=n, dtau=1e-3, rc=40.0 * n / 512, bsig=16.0 * n / 512)
    u0, v0 = initialize_host(P)
    s = Burgers(P).upload(u0, v0)
    s.step(20)
    s.sync()
    s.step(a.steps)
    ms = s.last_step_ms() / a.steps
    cells = n * n
    bytes_per_cell = 16 * (1 + max(P.visc_substeps, 1))   # each kernel reads phi_u, phi_v and writes them
    ach = bytes_per_cell * cells / (ms * 1e-3) / 1e9
    ref_ms = None
    if oracle.has_ref("ref_burgers") and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        k = min(a.steps, 100)
        *_, t = oracle.ref_burgers_run(oracle.burgers_params(nx=n, ny=n, dtau=1e-3, rc=P.rc, bsig=P.bsig), u0, v0, k)
        ref_ms = t / k
    return ({"bench": "burgers", "grid": [n, n], "steps": a.steps, "ms_per_step": ms,
                      "value": cells / (ms * 1e-3) / 1e6, "unit": "Mcell-updates/s",
                      "roofline": {"bound": "hbm", "achieved": ach, "peak": peak(), "unit": "GB/s",
                                   "frac": ach / peak(), "algorithmic_bytes_per_cell": bytes_per_cell,
                                   "note": "2 kernels per step (convection, viscosity), 16 B/cell each"},
                      "reference_gpu": {"ms_per_step": ref_ms,
                                        "value": cells / (ref_ms * 1e-3) / 1e6 if ref_ms else None,
                                        "what": "tau_burgers.cu kernels recompiled for sm_100a incl. the per-step "
                                                "D2H of the block maxima"},
                      "gpu_launches": s.launch_count})


def bench_sw(a):
    """tau_sw n x n (periodic).  Only on request (`bench_all.py sw`): the kernels have not
    run on hardware yet (NEXT.md).  A gentle field scaled with the grid; nu > 0 (2 kernels per step) as the
    reference's default."""
    import oracle
    from fluid_sims_b200.shallow_water import Params, ShallowWater, initialize_host
    n = a.sw_n
    kw = dict(nx=n, ny=n, dtau=1e-3, offx=0.1 * n, offy=0.1 * n, swirlRc=100.0 * n / 512, bumpSigma=8.0 * n / 512)
    P = Params(**kw)
    f = initialize_host(P)
    s = ShallowWater(P).upload(*f)
    s.step(20)
    s.sync()
    s.step(a.steps)
    ms = s.last_step_ms() / a.steps
    cells = n * n
    bytes_per_cell = 24 + (20 if P.nu > 0 else 0)   # update: 3 planes in + 3 out; viscosity: sigma, u, v in, u, v out
    ach = bytes_per_cell * cells / (ms * 1e-3) / 1e9
    ref_ms = None
    if oracle.has_ref("ref_sw") and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        k = min(a.steps, 100)
        *_, t = oracle.ref_sw_run(oracle.sw_params(**kw), *f, k)
        ref_ms = t / k
    return ({"bench": "shallow_water", "grid": [n, n], "steps": a.steps, "ms_per_step": ms,
                      "value": cells / (ms * 1e-3) / 1e6, "unit": "Mcell-updates/s",
                      "roofline": {"bound": "hbm", "achieved": ach, "peak": peak(), "unit": "GB/s",
                                   "frac": ach / peak(), "algorithmic_bytes_per_cell": bytes_per_cell,
                                   "note": "fused flux+update kernel 24 B/cell, viscosity kernel 20 B/cell"},
                      "reference_gpu": {"ms_per_step": ref_ms,
                                        "value": cells / (ref_ms * 1e-3) / 1e6 if ref_ms else None,
                                        "what": "tau_shallow_water.cu kernels recompiled for sm_100a incl. the "
                                                "per-step D2H of the block maxima"},
                      "gpu_launches": s.launch_count})


BENCHES = {"gs": bench_gs, "hyp3d": bench_hyp3d, "sph": bench_sph, "burgers": bench_burgers, "sw": bench_sw}


def default_args(**over):
    """the argument namespace of main() with its defaults (bench.py's `other_configs` leg calls the benches directly)"""
    a = make_parser().parse_args([])
    for k, v in over.items():
        setattr(a, k, v)
    return a


def make_parser():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=[])
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--gs-n", type=int, default=8192)
    ap.add_argument("--n3", type=int, default=256)
    ap.add_argument("--steps3", type=int, default=40)
    ap.add_argument("--warm3", type=int, default=60)
    ap.add_argument("--sph-n", type=int, default=1 << 21)
    ap.add_argument("--steps-sph", type=int, default=50)
    ap.add_argument("--burgers-n", type=int, default=4096)
    ap.add_argument("--sw-n", type=int, default=4096)
    return ap


def main():
    a = make_parser().parse_args()
    which = a.which or ["gs", "hyp3d", "sph"]
    for w in which:
        rec = BENCHES[w](a)
        if rec is not None:
            print(json.dumps(rec), flush=True)
    # one process group for the whole run (re-initialising NCCL between benches is not reliable)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
