#!/usr/bin/env python
"""bench.py — headline benchmark: Mcell-updates/s of the 2-D hypersonic update path on the
4096 x 4096 fp32 grid (BASELINE.json configs[1]), with the HBM roofline of the fused step kernel
and the reference CPU solver timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # product arm
  python bench.py --impl reference [--steps K] [--warmup W]      # reference CPU arm
  torchrun ... bench.py --gpus N ...                              # one rank per GPU (N > 1)

Definitions
  step     one solver time step of the whole grid (tau_hypersonic_cuda.cu:1833-1889), i.e. W*H
           cell-updates.  `value` = W*H*K / t_device, inputs resident in HBM, t from CUDA events on
           the launching stream, max over ranks.
  e2e      the same metric through the C-ABI with HOST buffers: every e2e step is one *frame* of
           the reference's frame loop — upload the 4 state planes from pinned host memory, run
           `steps_per_frame` solver steps (reference default 2, tau_hypersonic_cuda.cu:1407),
           download the 4 planes — all inside the timed region (wall clock, drained at the end).
           Three frames are in flight (one handle per frame in flight, each on its own stream, tau_hyp2d_upload_async /
           tau_hyp2d_download_async), so the H2D copy of one frame overlaps the D2H copy of the
           other; at N>1 every rank moves its own slab over its own PCIe link and the ghost-row hand-over after
           an upload happens on the device (tau_hyp2d_upload_peers_async; `--e2e-handover host` = the
           host-driven NCCL exchange of round 1, one frame at a time).
  roofline achieved = 33 algorithmic bytes/cell (read 4 fp32 fields + 1 mask byte, write 4 fields;
           SURVEY.md §8(d)) x W*H / average duration of one hyp2d_step launch (CUDA events over the
           timed region, in which it is the only kernel) vs the measured HBM copy bandwidth in
           MEASURED_PEAKS.json.
  cpu_baseline  the reference's own tau_hypersonic_simd.c (compiled from the reference sources into
           oracle/_ref with the reference's flags) stepping a 4096 x 256 band of the grid on one
           host core; `--impl reference` runs one independent replica of it per host core.
  state_crc  checksum of the four state planes after the timed region (develop + warmup + steps solver steps
           from k_init), computed from per-row weighted sums of the fp32 bit patterns so that ranks can add
           their parts: identical across N = 1/2/4/8 <=> the slab-decomposed run is bit-identical.
  dtype_f64  the same solver with the fp64 handle (the instantiation that holds the north-star's 1e-5 bound),
           short run, N = 1 only.
  reference_gpu  the reference's own kernels (oracle/_ref, recompiled for sm_100a) timed on the same GPU on the
           same grid — a yardstick beside the CPU baseline, N = 1 only.
  other_configs  BASELINE configs 3-5 (Gray-Scott 8192^2, 3-D hypersonic 512^3 — z-slabs at N > 1 —,
           SPH 2^21 particles), short runs of bench_all.py's benches, each with its roofline, reference_gpu (N = 1) and a
           state_crc that is equal across N when the multi-GPU run is bit-identical to the single-GPU one.
  slab_balance  N > 1 (strong scaling, peer exchange): the y-slabs are cut by measured cost, not by equal rows —
           slab.hyp2d_balanced_partition runs two short trials BEFORE the benchmark's own handles exist, compares how long
           each rank's step kernel is busy (the slabs with the bow shock, the body and the wake are 10-20 % slower per row)
           and moves the cuts; `rows` = the partition used.  The state does not depend on it (state_crc); --no-balance.
The working set (2 x 268 MB of state) is larger than the 126 MB L2, so no L2 flush is needed
between timed steps.

Timed region at N > 1: barrier + synchronize, then the W warm-up steps and the K timed steps are enqueued back
to back; ev0 is recorded IN-STREAM after the last warm-up step (the device-side inbox barrier of every step has
aligned the ranks by then), ev1 after the K-th step, then barrier + synchronize; max over ranks.  A host barrier
between warm-up and ev0 would put the ranks' host skew (~1 ms) into a region that lasts 2 ms at N = 8.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID_W = 4096
GRID_H = 4096
BYTES_PER_CELL = 33           # 4 fields R + 4 fields W (fp32) + 1 mask byte
CPU_BAND = (4096, 256)        # bounded CPU sample: a 4096 x 256 band (1/16 of the rows)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def profile_constants():
    """per-launch DRAM bytes and executed warp-instructions of the step kernel at 4096 rows, from the committed
    ncu capture (profiles/hyp2d_step_traffic.json); both scale with the rows a launch updates."""
    path = os.path.join(ROOT, "profiles", "hyp2d_step_traffic.json")
    try:
        with open(path) as f:
            j = json.load(f)
        return j
    except Exception:
        return {}


def state_crc(torch, dist, planes_view, y0, hl, W, halo, world, dev):
    """order-independent-across-ranks checksum of the owned rows of the 4 planes (see module docstring)"""
    import zlib
    P31 = (1 << 31) - 1
    bits = planes_view[:, halo:halo + hl, :].contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    rows = torch.arange(y0, y0 + hl, device=bits.device, dtype=torch.int64).view(1, hl, 1)
    W2 = bits.shape[2]                                      # 2 W words per row for an fp64 handle
    cols = torch.arange(W2, device=bits.device, dtype=torch.int64).view(1, 1, W2)
    wgt = (rows * W2 + cols) % 65521 + 1
    s1 = bits.sum(dim=(1, 2))                               # < 2^56
    s2 = ((bits * wgt).sum(dim=2) % P31).sum(dim=1)         # row sums < 2^60, then < 2^31 each
    acc = torch.stack([s1, s2]).to(torch.int64)
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    acc[1] %= P31
    vals = [int(v) for v in acc.flatten().tolist()]
    return f"{zlib.crc32(repr(vals).encode()) & 0xFFFFFFFF:08x}", vals


def bind_to_gpu_numa_node(torch, dev):
    """Pin this process (and therefore its first-touch pinned host buffers) to the CPU cores of the NUMA node the GPU hangs off:
    at N > 1 every rank moves its slab over its own PCIe link, and a buffer on the other socket makes that traffic cross the
    socket interconnect.  Returns a short description for the e2e record; does nothing where sysfs does not tell."""
    try:
        pr = torch.cuda.get_device_properties(dev)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
        if node < 0 or nodes < 2:
            return f"single NUMA node (gpu {bus}: node {node} of {nodes})"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"gpu {bus} on node {node} of {nodes}: no allowed core there"
        os.sched_setaffinity(0, cpus)
        return f"bound to NUMA node {node} of {nodes} ({len(cpus)} cores) for gpu {bus}"
    except Exception as e:      # noqa: BLE001
        return f"not bound ({type(e).__name__}: {e})"[:120]


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                 "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        smax = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) > 8:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU solver from oracle/_ref
# --------------------------------------------------------------------------------------------
def _cpu_replica(args):
    nsteps, warm = args
    import oracle
    r = oracle.RefHypCpu(CPU_BAND[0], CPU_BAND[1], simd=True)
    r.init()
    if warm:
        r.steps(warm)
    return r.steps(nsteps)


def cpu_reference(nsteps, warm, procs):
    """Aggregate Mcell-updates/s of `procs` independent replicas (the solver is single-threaded:
    tau_hypersonic_simd.c has no OpenMP/pthreads)."""
    import multiprocessing as mp
    cells = CPU_BAND[0] * CPU_BAND[1]
    if procs == 1:
        secs = [_cpu_replica((nsteps, warm))]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            secs = pool.map(_cpu_replica, [(nsteps, warm)] * procs)
    t = max(secs)
    return procs * cells * nsteps / t / 1e6, t / nsteps * 1e3


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    val, ms = cpu_reference(a.steps, a.warmup, cores)
    sample = (f"{a.steps} x step_physics() of tau_hypersonic_simd.c (gcc -O3 -mavx2 -mfma) on a "
              f"{CPU_BAND[0]}x{CPU_BAND[1]} band, one independent replica per host core")
    print(json.dumps({
        "impl": "reference", "metric": "Mcell-updates/s", "value": val, "unit": "Mcell-updates/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "tau_hypersonic 4096x4096 (CPU reference steps a 4096x256 band per "
                               "replica; solver is fp64 and single-threaded)"},
        "cpu_baseline": {"value": val, "unit": "Mcell-updates/s", "cores": cores,
                         "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------
# product arm
# --------------------------------------------------------------------------------------------
def run_product(a):
    import faulthandler

    import numpy as np
    import torch

    faulthandler.dump_traceback_later(a.trace_after, repeat=False, file=sys.stderr)   # where is it, if it hangs?

    from fluid_sims_b200 import slab
    from fluid_sims_b200.hypersonic2d import HALO, Hypersonic2D, SimConfig

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    dev = local if world > 1 else 0

    W = a.grid_w or GRID_W
    H = (a.grid_h or GRID_H) * (world if a.scaling == "weak" else 1)
    cfg = SimConfig.default(W, H)
    # an explicit side stream: torch's legacy default stream has handle 0, which the C-ABI reads as
    # "create your own stream" — events recorded on it would not see the kernels
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    parts = slab.partition_rows(H, world, slab.hyp2d_row_costs(cfg) if a.cost_weighted_slabs else None)
    balance = None
    if world > 1 and a.exchange == "peer" and a.scaling == "strong" and not a.no_balance and not a.cost_weighted_slabs:
        # measured load balancing (slab.hyp2d_balanced_partition): the slabs that hold the bow shock, the body and the wake
        # are 10-20 % slower per row; the state does not depend on where the cuts are (state_crc below)
        def trial(yb, hloc):
            t = Hypersonic2D(cfg, dtype=a.dtype, device=dev, y_begin=yb, h_local=hloc, stream=stream)
            if a.seg_rows:
                t.set_seg_rows(a.seg_rows)
            return t.init()
        try:
            parts, busy = slab.hyp2d_balanced_partition(trial, H, steps=min(a.develop, 600) or 100)
            balance = {"rows": [c for _, c in parts], "busy_us_last_trial": [round(b, 1) for b in busy]}
        except Exception as e:                      # never lose the measurement to the balancer
            parts = slab.partition_rows(H, world)
            balance = {"error": repr(e)[:200]}
        agreed = [None] * world                     # one partition for all ranks, or equal rows for all
        dist.all_gather_object(agreed, parts)
        if any(p != agreed[0] for p in agreed):
            parts = slab.partition_rows(H, world)
            balance = {"error": "ranks disagreed on the balanced partition; equal rows used"}
    y0, hl = parts[rank]
    sim = Hypersonic2D(cfg, dtype=a.dtype, device=dev, y_begin=y0, h_local=hl, stream=stream)
    if a.seg_rows:
        sim.set_seg_rows(a.seg_rows)
    sim.init()
    tdt = torch.float32 if a.dtype == "f32" else torch.float64

    views = {}

    def state_views():
        pp, mp, sp = sim.device_state()
        if pp not in views:
            views[pp] = slab.wrap_plane(pp, (4, hl + 2 * HALO, W), tdt, dev)
        if sp not in views:
            views[sp] = slab.wrap_plane(sp, (1,), torch.float64, dev)
        return views[pp], views[sp], mp

    peer = world > 1 and a.exchange == "peer"

    def resync():
        """after init/upload on a slab: ghost rows + wavespeed once over NCCL, arm device barrier"""
        if world > 1:
            slab.hyp2d_sync_state(sim)
            if peer:
                sim.peers_ready()
                dist.barrier()

    def advance(n):
        if world == 1 or peer:
            sim.step(n)      # peer mode: halo push + max all-reduce + barrier are device-side
            return
        for _ in range(n):   # host-driven exchange (NCCL P2P + all-reduce every step)
            planes, speed, _ = state_views()
            slab.exchange_halos([planes], HALO, periodic=False, dim=1)
            dist.all_reduce(speed, op=dist.ReduceOp.MAX)
            sim.step(1)

    if peer:
        slab.hyp2d_attach_peers(sim)
    resync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # develop the flow first (uniform inflow under-exercises the limiter/HLLC branches)
    advance(a.develop)
    barrier()

    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_driven = world > 1 and not peer     # NCCL exchange per step: nothing aligns the ranks on the device
    advance(a.warmup)
    if host_driven:
        barrier()
    launches0 = sim.launch_count
    ev0.record()                             # in-stream, right behind the last warm-up step (see module docstring)
    advance(a.steps)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = sim.launch_count - launches0
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    peer_timing = None
    if peer:
        mine = sim.peer_timing()
        peer_timing = [None] * world
        dist.all_gather_object(peer_timing, {k: round(v, 2) for k, v in mine.items()})
    crc, crc_parts = state_crc(torch, dist, state_views()[0], y0, hl, W, HALO, world, dev)
    clock_t, clock_dt = sim.clock()

    cells = W * H
    value = cells * a.steps / (ms * 1e-3) / 1e6
    kernel_ms = ms / a.steps
    peak, peak_src = measured_peak()
    # roofline of the dominant kernel (hyp2d_step): per launch it updates this rank's slab
    rows_mean = H / world                    # (balanced slabs differ in height; the step time is the max over ranks)
    ach = BYTES_PER_CELL * (W * rows_mean) / (kernel_ms * 1e-3) / 1e9
    bpc = BYTES_PER_CELL if a.dtype == "f32" else 65
    if a.dtype != "f32":
        ach = bpc * (W * rows_mean) / (kernel_ms * 1e-3) / 1e9

    prof = profile_constants()
    frac_rows = rows_mean / 4096.0 * (W / 4096.0)
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    inst = prof.get("warp_instructions_per_launch")
    issue_frac = (inst * frac_rows / (kernel_ms * 1e-3)) / (sm_count * 4 * sm_mhz * 1e6) if inst else None
    line = None
    if rank == 0:
        line = {
            "metric": "Mcell-updates/s", "value": value, "unit": "Mcell-updates/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
            "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": f"tau_hypersonic_cuda {W}x{H} {a.dtype}: k_init state developed "
                                   f"for {a.develop} steps, then timed",
                       "grid": [W, H], "slab_rows_per_gpu": hl, "parallelism": f"y-slab x{world}" + (f" ({a.exchange} halo exchange)" if world > 1 else ""),
                       "cache": "state (2 x 268 MB) larger than L2 (126 MB): no flush needed",
                       "seg_rows": sim.seg_rows, "step_kernel": sim.kernel_name},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak,
                         "traffic": prof["dram_bytes_per_launch"] * frac_rows if "dram_bytes_per_launch" in prof else None,
                         "traffic_source": prof.get("source", None) and (prof["source"] + "; scaled by the rows one launch updates"),
                         "peak_source": f"{peak_src} (MEASURED_PEAKS.json hbm_gbs)",
                         "algorithmic_bytes_per_cell": bpc, "kernel": sim.kernel_name,
                         "kernel_ms": kernel_ms,
                         "limiter": "fp32-issue", "issue_frac": issue_frac,
                         "issue_peak": f"{sm_count} SMs x 4 schedulers x {sm_mhz:.0f} MHz warp-instructions/s",
                         "note": "the kernel is FP32-issue bound, not HBM bound: `frac` is the HBM fraction at the algorithmic 33 B/cell, "
                                 "`issue_frac` the executed warp-instructions (committed ncu capture, scaled by rows) over the issue peak; DESIGN.md 4.1"},
            "cpu_baseline": None, "e2e": None, "gpu_launches": int(launches), "clocks": clocks,
            "state_crc": crc, "sim_t": clock_t,
            **({"peer_timing": peer_timing} if peer_timing else {}),
            **({"slab_balance": balance} if balance else {}),
        }
    # ---- from here on nothing may cost the headline: a watchdog prints what has been measured and exits -------
    def emit():
        if rank == 0:
            print(json.dumps(line), flush=True)

    done = threading.Event()

    def watchdog():
        if not done.wait(a.total_timeout):
            if rank == 0:
                line["watchdog"] = f"sections after the timed region did not finish within {a.total_timeout} s"
            emit()
            os._exit(0)

    threading.Thread(target=watchdog, daemon=True).start()
    # ---- e2e: frames through the C-ABI with host buffers -------------------------------------
    e2e = None
    if not a.no_e2e:
        numa = bind_to_gpu_numa_node(torch, dev) if (world > 1 and not a.no_numa_bind) else None
        np_dt = np.float32 if a.dtype == "f32" else np.float64
        host_in = [torch.empty((hl, W), dtype=tdt).pin_memory() for _ in range(4)]
        planes, _ = sim.download()
        for t_, p_ in zip(host_in, planes):
            t_.numpy()[...] = p_
        del planes
        import ctypes as C
        from fluid_sims_b200 import hypersonic2d as h2
        host_out = [torch.empty((hl, W), dtype=tdt).pin_memory() for _ in range(4)]
        in_ptrs = (C.c_void_p * 4)(*[t_.data_ptr() for t_ in host_in])
        out_ptrs = (C.c_void_p * 4)(*[t_.data_ptr() for t_ in host_out])

        pipelined = world == 1
        if pipelined:
            # several frames in flight: extra handles on their own streams, so that the upload of
            # frame i+1 (H2D copy engine) overlaps the download of frame i (D2H copy engine)
            lanes = [(sim, out_ptrs, host_out)]
            lane_streams = []
            for _ in range(a.e2e_lanes - 1):
                st = torch.cuda.Stream(device=dev)
                sb = Hypersonic2D(cfg, dtype=a.dtype, device=dev, y_begin=y0, h_local=hl, stream=st.cuda_stream)
                if a.seg_rows:
                    sb.set_seg_rows(a.seg_rows)
                sb.init()
                ho = [torch.empty((hl, W), dtype=tdt).pin_memory() for _ in range(4)]
                lanes.append((sb, (C.c_void_p * 4)(*[t_.data_ptr() for t_ in ho]), ho))
                lane_streams.append(st)  # keep the stream alive

            def frame(i):
                h, optr, _ = lanes[i % len(lanes)]
                h2.check(h2._sync(h._handle))          # this lane's previous frame has landed
                h2.check(h2._upload_async(h._handle, in_ptrs, C.c_void_p(0)))
                h2.check(h2._step(h._handle, a.steps_per_frame))
                h2.check(h2._download_async(h._handle, optr, C.c_void_p(0)))

            def drain():
                for h, _, _ in lanes:
                    h2.check(h2._sync(h._handle))
        elif a.e2e_peers_async and a.exchange == "peer":
            # the frame hand-over between ranks on the device (validated at N = 2 in round 2: 4.06 ms per frame
            # with three lanes against 6.19 ms host-driven) — no NCCL exchange, host synchronisation or barrier
            # inside the frame loop; with --e2e-lanes > 1 every
            # rank runs that many slab handles (each with its own peer attachments) on their own streams, so the
            # upload of frame i+1 overlaps the download of frame i
            lanes = [(sim, out_ptrs, host_out)]
            lane_streams = []
            for _ in range(a.e2e_lanes - 1):
                st = torch.cuda.Stream(device=dev)
                sb = Hypersonic2D(cfg, dtype=a.dtype, device=dev, y_begin=y0, h_local=hl, stream=st.cuda_stream)
                if a.seg_rows:
                    sb.set_seg_rows(a.seg_rows)
                sb.init()
                slab.hyp2d_attach_peers(sb)
                slab.hyp2d_sync_state(sb)
                sb.peers_ready()
                dist.barrier()
                ho = [torch.empty((hl, W), dtype=tdt).pin_memory() for _ in range(4)]
                lanes.append((sb, (C.c_void_p * 4)(*[t_.data_ptr() for t_ in ho]), ho))
                lane_streams.append(st)

            def frame(i):
                h, optr, _ = lanes[i % len(lanes)]
                h2.check(h2._upload_peers_async(h._handle, in_ptrs))
                h2.check(h2._step(h._handle, a.steps_per_frame))
                h2.check(h2._download_async(h._handle, optr, C.c_void_p(0)))

            def drain():
                for h, _, _ in lanes:
                    h2.check(h2._sync(h._handle))
        else:
            def frame(i):
                h2.check(h2._upload(sim._handle, in_ptrs, C.c_void_p(0)))
                resync()
                advance(a.steps_per_frame)
                h2.check(h2._download(sim._handle, out_ptrs, C.c_void_p(0)))

            def drain():
                pass

        for i in range(2 * a.e2e_lanes):
            frame(i)
        drain()
        barrier()
        t0 = time.perf_counter()
        for i in range(a.e2e_frames):
            frame(i)
        drain()
        barrier()
        dt_wall = time.perf_counter() - t0
        if pipelined:
            # both lanes computed the same frame from the same input: identical results
            for _, _, ho in lanes[1:]:
                assert all(torch.equal(x, y) for x, y in zip(host_out, ho)), "pipelined lanes differ"
        if world > 1:
            t = torch.tensor([dt_wall], device=f"cuda:{dev}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_wall = float(t.item())
        nbytes = 4 * hl * W * np.dtype(np_dt).itemsize
        e2e = {"value": cells * a.steps_per_frame * a.e2e_frames / dt_wall / 1e6,
               "unit": "Mcell-updates/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "steps_per_frame": a.steps_per_frame, "frames": a.e2e_frames,
               "frames_in_flight": a.e2e_lanes if (pipelined or (a.e2e_peers_async and a.exchange == "peer")) else 1,
               "ms_per_frame": dt_wall / a.e2e_frames * 1e3}
        if world > 1:
            e2e["host_buffers"] = numa
            e2e["handover"] = "device" if (a.e2e_peers_async and a.exchange == "peer") else "host"
            if a.e2e_peers_async and a.exchange == "peer":
                for sb, _, _ in lanes[1:]:      # the extra lanes' peer mappings go before their planes do
                    slab.hyp2d_detach_peers(sb)
                    sb.close()

    if rank == 0:
        line["e2e"] = e2e
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        v, _ = cpu_reference(a.cpu_steps, 1, 1)
        cpu = {"value": v, "unit": "Mcell-updates/s", "cores": 1, "kind": "reference",
               "sample": (f"{a.cpu_steps} x step_physics() of tau_hypersonic_simd.c (gcc -O3 -mavx2 "
                          f"-mfma, oracle/_ref) on a {CPU_BAND[0]}x{CPU_BAND[1]} band of the grid, "
                          f"1 thread (the solver is single-threaded); host has {os.cpu_count()} cores")}
        if a.cpu_full_steps > 0:
            try:
                import oracle
                r = oracle.RefHypCpu(GRID_W, GRID_H, simd=True)
                r.init()
                secs = r.steps(a.cpu_full_steps)
                cpu["full_grid"] = {"value": GRID_W * GRID_H * a.cpu_full_steps / secs / 1e6, "unit": "Mcell-updates/s",
                                    "sample": f"{a.cpu_full_steps} x step_physics() on the full {GRID_W}x{GRID_H} grid, 1 thread"}
            except Exception as e:      # noqa: BLE001 (library not built for this size)
                cpu["full_grid"] = {"error": str(e)[:200]}

    # ---- the fp64 handle (holds the north-star's 1e-5 bound) and the reference's own kernels, N = 1 only ----
    if rank == 0:
        line["cpu_baseline"] = cpu
    f64 = ref_gpu = None
    if world == 1 and not a.no_extras:
        try:
            s64 = Hypersonic2D(cfg, dtype="f64", device=dev, stream=stream).init()
            s64.step(a.develop_f64)
            s64.sync()
            s64.step(a.steps_f64)
            ms64 = s64.last_step_ms() / a.steps_f64
            f64 = {"dtype": "f64", "ms_per_step": ms64, "value": cells / (ms64 * 1e-3) / 1e6, "unit": "Mcell-updates/s",
                   "steps": a.steps_f64, "developed_for": a.develop_f64,
                   "roofline": {"bound": "hbm", "achieved": 65 * cells / (ms64 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": 65 * cells / (ms64 * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_cell": 65},
                   "parity": "L_inf <= 5.5e-13 vs the reference kernels after 1000 steps (tests/test_hyp2d_gpu.py)"}
            s64.close()
        except Exception as e:          # noqa: BLE001
            f64 = {"error": str(e)[:200]}
        try:
            import oracle
            if oracle.has_ref(f"ref_hyp2d_{W}x{H}"):
                *_, rms = oracle.ref_hyp2d_run(W, H, oracle.hyp2d_cfg(W, H).as11(), a.steps_ref_gpu)
                ref_gpu = {"ms_per_step": rms / a.steps_ref_gpu, "value": cells / (rms / a.steps_ref_gpu * 1e-3) / 1e6,
                           "unit": "Mcell-updates/s", "dtype": "f64", "steps": a.steps_ref_gpu,
                           "what": "tau_hypersonic_cuda.cu's own step loop (:1833-1889: 7 kernels + an 8-byte D2H and host dt "
                                   "per step) recompiled for sm_100a, from k_init, same GPU"}
        except Exception as e:          # noqa: BLE001
            ref_gpu = {"error": str(e)[:200]}

    if rank == 0 and f64:
        line["dtype_f64"] = f64
    if rank == 0 and ref_gpu:
        line["reference_gpu"] = ref_gpu
    if world > 1 and peer:   # unmap the peers' planes on every rank before any rank frees them
        try:
            slab.hyp2d_detach_peers(sim)
        except Exception as e:      # teardown must not turn a finished measurement into a failure
            print(f"[bench] peer detach: {e}", file=sys.stderr)
    sim.close()

    # ---- BASELINE configs 3-5, short runs (never allowed to cost the headline line: watchdog) -------------
    if not a.no_other:
        other = {}
        import bench_all
        # config 4 is 512^3 at every N (6.4 GB on one GPU), so that the lines' state_crc can be compared across N
        ba = bench_all.default_args(steps=200, n3=512, steps3=12, warm3=20, steps_sph=30)
        for name in (("gs",) if world == 1 else ()) + ("hyp3d", "sph"):
            try:
                rec = bench_all.BENCHES[name](ba)
            except Exception as e:      # noqa: BLE001
                rec = {"error": f"{type(e).__name__}: {e}"[:300]}
            other[name] = rec
            torch.cuda.synchronize()
        if rank == 0:
            line["other_configs"] = other
    done.set()
    emit()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--develop", type=int, default=1500,
                    help="untimed steps run first so that the bow shock exists")
    ap.add_argument("--steps-per-frame", type=int, default=2)
    ap.add_argument("--e2e-frames", type=int, default=24)
    ap.add_argument("--e2e-lanes", type=int, default=3, help="frames in flight in the e2e loop (N=1)")
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--seg-rows", type=int, default=0)
    ap.add_argument("--grid-w", type=int, default=0, help="experiments only (default 4096)")
    ap.add_argument("--grid-h", type=int, default=0, help="experiments only (default 4096)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU halo exchange: device-side peer pushes over NVLink (default) or "
                         "host-driven NCCL send/recv + all-reduce every step")
    ap.add_argument("--e2e-handover", default="device", choices=["device", "host"],
                    help="N>1, --exchange peer: hand the uploaded frame over to the neighbours on the device "
                         "(tau_hyp2d_upload_peers_async: no NCCL, host synchronisation or barrier in the frame loop; default) "
                         "or with the host-driven NCCL exchange of round 1")
    ap.add_argument("--e2e-peers-async", action="store_true", help="(round-1 spelling of --e2e-handover device)")
    ap.add_argument("--cost-weighted-slabs", action="store_true",
                    help="N>1: rows partitioned by estimated cost instead of equally (results are identical either way; measured: "
                         "no gain — 0.08676 vs 0.08679 ms/step at N = 8 — the per-rank busy times even out but the step does not shorten)")
    ap.add_argument("--no-numa-bind", action="store_true", help="N>1 e2e: do not pin the rank to its GPU's NUMA node")
    ap.add_argument("--no-balance", action="store_true", help="N>1: equal rows per slab instead of the measured balance")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the fp64-handle and reference-kernel sub-records (N=1)")
    ap.add_argument("--no-other", action="store_true", help="skip BASELINE configs 3-5 (other_configs)")
    ap.add_argument("--total-timeout", type=float, default=300.0,
                    help="seconds the sections after the timed region (e2e, baselines, other configs) may take before the "
                         "watchdog prints the line with what has been measured")
    ap.add_argument("--trace-after", type=float, default=150.0, help="dump every thread's Python stack to stderr after this many seconds")
    ap.add_argument("--cpu-full-steps", type=int, default=3, help="reference CPU solver on the full 4096^2 grid (0 = skip)")
    ap.add_argument("--develop-f64", type=int, default=100)
    ap.add_argument("--steps-f64", type=int, default=40)
    ap.add_argument("--steps-ref-gpu", type=int, default=20)
    a = ap.parse_args()
    a.e2e_peers_async = a.e2e_peers_async or a.e2e_handover == "device"
    if a.warmup < 3 and a.impl == "b200":
        a.warmup = 3
    if a.impl == "reference":
        run_reference(a)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if a.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun, one rank per GPU
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                   f"--nproc-per-node={a.gpus}", "--master-addr", "127.0.0.1", "--master-port",
                   os.environ.get("MASTER_PORT", "29517"), os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        run_product(a)


if __name__ == "__main__":
    main()
