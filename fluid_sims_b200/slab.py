"""Slab (row-block) domain decomposition across the GPUs of one box.

One process per GPU (`torch.distributed`, backend nccl on GPUs / gloo on CPU for the host-logic
tests).  The reference has no multi-GPU path ("absent"), so parity is defined against
the single-GPU run: a slab-decomposed run must reproduce it bit-for-bit.

Each rank owns rows [y_begin, y_begin + ny_local) of every plane plus `halo` ghost rows above and
below.  Per step the ranks exchange `halo` boundary rows with their two neighbours:

  * periodic direction (Gray-Scott y, 3-D z)  -> the ranks form a ring,
  * clamped direction (2-D hypersonic y)       -> a chain; the edge ranks' outer ghosts are the
                                                  solver's own boundary condition.

The exchange is written against `torch.distributed` P2P ops on plain tensors, so the very same code
runs over NCCL/NVLink on device planes and over gloo on CPU tensors in the tests.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def partition_rows(n: int, parts: int, weights=None) -> List[Tuple[int, int]]:
    """Balanced contiguous partition of n rows: [(begin, count)] * parts (earlier ranks get the
    remainder).  `weights` (one relative cost per row): equal cost per part instead of equal rows — results do
    not depend on the partition (a slab run is bit-identical to the single-GPU run for every decomposition)."""
    if parts <= 0 or n < parts:
        raise ValueError(f"cannot split {n} rows over {parts} ranks")
    if weights is not None and parts > 1:
        import itertools
        w = [float(x) for x in weights]
        if len(w) != n or min(w) <= 0:
            raise ValueError("weights: one positive cost per row")
        cum = list(itertools.accumulate(w))
        cuts, r = [0], 0
        for k in range(1, parts):
            target = cum[-1] * k / parts
            while r < n and cum[r] <= target:
                r += 1
            r = max(r, cuts[-1] + 1)
            r = min(r, n - (parts - k))
            cuts.append(r)
        cuts.append(n)
        return [(cuts[k], cuts[k + 1] - cuts[k]) for k in range(parts)]
    base, rem = divmod(n, parts)
    out, b = [], 0
    for r in range(parts):
        c = base + (1 if r < rem else 0)
        out.append((b, c))
        b += c
    return out


class _DevArray:
    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3, "strides": None}


_TYPESTR = {torch.float32: "<f4", torch.float64: "<f8", torch.uint8: "|u1", torch.int32: "<i4"}


def wrap_plane(ptr: int, shape, dtype: torch.dtype, device: int | None = None) -> torch.Tensor:
    """Zero-copy torch view of a device plane owned by libtau_b200 (for NCCL halo traffic)."""
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    return torch.as_tensor(_DevArray(ptr, shape, _TYPESTR[dtype]), device=dev)


def exchange_halos(planes: Sequence[torch.Tensor], halo: int, periodic: bool, group=None,
                   dim: int = 0) -> None:
    """Fill the ghost rows of every plane from the neighbouring ranks.

    planes: tensors whose dimension `dim` is the decomposed one, of extent ny_local + 2*halo.
    Rows [halo, 2*halo) go up (to rank-1's bottom ghosts), rows [-2*halo, -halo) go down (to
    rank+1's top ghosts).  Non-periodic: rank 0 has no upper neighbour, the last rank no lower one
    (their outer ghosts stay whatever the solver's boundary condition put there).
    """
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        if periodic:
            for p in planes:
                n = p.shape[dim]
                p.narrow(dim, 0, halo).copy_(p.narrow(dim, n - 2 * halo, halo))
                p.narrow(dim, n - halo, halo).copy_(p.narrow(dim, halo, halo))
        return
    up = rank - 1 if rank > 0 else (world - 1 if periodic else None)
    dn = rank + 1 if rank < world - 1 else (0 if periodic else None)
    # Per plane the ops are issued as [send_up, send_dn, recv_dn, recv_up].  With distinct
    # neighbours the order is irrelevant; in a 2-rank ring both neighbours are the same peer and
    # messages between one pair of ranks match in issue order: the peer's first message (its top
    # rows, sent "up") is my bottom ghost, its second (bottom rows) my top ghost.
    ops, keep = [], []
    for p in planes:
        n = p.shape[dim]
        sends, recvs = [], []
        if up is not None:
            s_up = p.narrow(dim, halo, halo).contiguous()
            r_up = torch.empty_like(s_up)
            sends.append(dist.P2POp(dist.isend, s_up, up, group))
            recvs.append(dist.P2POp(dist.irecv, r_up, up, group))
            keep.append((p.narrow(dim, 0, halo), r_up))
        if dn is not None:
            s_dn = p.narrow(dim, n - 2 * halo, halo).contiguous()
            r_dn = torch.empty_like(s_dn)
            sends.append(dist.P2POp(dist.isend, s_dn, dn, group))
            recvs.insert(0, dist.P2POp(dist.irecv, r_dn, dn, group))
            keep.append((p.narrow(dim, n - halo, halo), r_dn))
        ops += sends + recvs
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for dst, r in keep:
        dst.copy_(r)


def allreduce_max_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


def hyp2d_row_costs(cfg, extra: float = 0.07):
    """Relative cost per grid row of the 2-D hypersonic step: rows that cross the body (masked march, subsonic shock layer
    and wake: no warp-uniform supersonic short cut) cost about 7 % more (measured at N = 8: 80 vs 76.5 us for the slabs that
    hold the body).  For partition_rows(..., weights=...)."""
    lo, hi = cfg.geom_cy - cfg.geom_Rb, cfg.geom_cy + cfg.geom_Rb
    return [1.0 + (extra if lo <= y + 0.5 <= hi else 0.0) for y in range(cfg.H)]


def rebalance_rows(parts: List[Tuple[int, int]], busy_us: Sequence[float], fixed_us: float = 10.0,
                   min_rows: int = 8) -> List[Tuple[int, int]]:
    """New contiguous partition of the same rows from the time each part's step kernel was busy (`busy_us[r]`, e.g.
    Hypersonic2D.peer_timing()["busy_us"]): a row of part r is taken to cost (busy_us[r] - fixed_us) / rows_r, and the
    cuts move so that every part gets the same modelled cost.  Pure host arithmetic; the state does not depend on the
    partition (every decomposition is bit-identical to one GPU), so this is load balancing only."""
    n = sum(c for _, c in parts)
    w: List[float] = []
    for (b, c), t in zip(parts, busy_us):
        w += [max(float(t) - fixed_us, 1e-3) / c] * c
    out = partition_rows(n, len(parts), w)
    if min(c for _, c in out) < min_rows:
        return list(parts)
    return out


def hyp2d_balanced_partition(make_sim, H: int, steps: int, iters: int = 2, group=None):
    """Measured load balancing of the y-slabs of the 2-D hypersonic solver (peer mode): `make_sim(y_begin, h_local)` builds
    this rank's initialised handle; every rank runs `steps` steps on the current partition, the ranks compare how long their
    step kernels were busy (the rows that hold the bow shock, the body and its wake cost 10-20 % more than free-stream
    rows), and the cuts move (rebalance_rows).  Returns (partition, [busy_us per rank of the last measurement]).  The
    trial handles are destroyed; results do not depend on the partition."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = partition_rows(H, world)
    busy: List[float] = []
    for _ in range(iters):
        y0, hl = parts[rank]
        sim = make_sim(y0, hl)
        hyp2d_attach_peers(sim, group)
        hyp2d_sync_state(sim, group)
        sim.peers_ready()
        dist.barrier(group=group)
        sim.step(steps)
        sim.sync()
        mine = float(sim.peer_timing()["busy_us"])
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=group)
        busy = [float(x) for x in gathered]
        hyp2d_detach_peers(sim, group)
        sim.close()
        dist.barrier(group=group)
        parts = rebalance_rows(parts, busy)
    return parts, busy


def hyp2d_attach_peers(sim, group=None) -> None:
    """All-gather the CUDA-IPC handles of every rank's planes/control block and attach them, so that
    the 2-D hypersonic step kernel pushes its boundary rows straight into the neighbours' ghost rows
    (NVLink peer stores) and the wavespeed all-reduce + step barrier run on the device."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = (sim.ipc_export(), sim.h_local)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    sim.ipc_attach(rank, world, [g[0] for g in gathered], [g[1] for g in gathered])


def hyp2d_detach_peers(sim, group=None) -> None:
    """Teardown counterpart of hyp2d_attach_peers: every rank unmaps its peers' memory, then all meet at a
    barrier, so that no rank frees planes another rank still has mapped.  Call before closing the handles."""
    sim.ipc_detach()
    if dist.is_initialized():
        dist.barrier(group=group)


def hyp2d_sync_state(sim, group=None) -> None:
    """Host-driven (NCCL) exchange of the CURRENT state's ghost rows, the static mask's ghost rows
    and the max-wavespeed scalar — needed once after init()/upload(); arms the device barrier when
    peers are attached."""
    from .hypersonic2d import HALO
    pp, mp, sp = sim.device_state()
    tdt = torch.float32 if sim.dtype == "f32" else torch.float64
    W, hl, dev = sim.cfg.W, sim.h_local, sim.device
    planes = wrap_plane(pp, (4, hl + 2 * HALO, W), tdt, dev)
    mask = wrap_plane(mp, (hl + 2 * HALO, W), torch.uint8, dev)
    speed = wrap_plane(sp, (1,), torch.float64, dev)
    exchange_halos([mask], HALO, periodic=False, group=group, dim=0)
    exchange_halos([planes], HALO, periodic=False, group=group, dim=1)
    dist.all_reduce(speed, op=dist.ReduceOp.MAX, group=group)
    torch.cuda.synchronize()
    dist.barrier(group=group)
