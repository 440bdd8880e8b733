"""Host-side slab logic on CPU with torch.distributed/gloo (world sizes 2 and 3): the same
`fluid_sims_b200.slab` code that moves ghost rows over NCCL on the GPUs.  The per-slab stepper is
the CPU oracle, so a decomposed run must reproduce the single-domain oracle bit-for-bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from fluid_sims_b200 import slab


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gs_worker(rank, world, port, nx, ny, steps, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    f32p = oracle.f32p
    oracle.lib.oracle_gs_step_rows.argtypes = [f32p] * 4 + [C.c_int, C.c_int] + [C.c_float] * 6 + [C.c_int, C.c_int]
    u0, v0 = oracle.gs_init_pattern(nx, ny)
    y0, nl = slab.partition_rows(ny, world)[rank]
    u = torch.zeros(nl + 2, nx)
    v = torch.zeros(nl + 2, nx)
    u[1:-1] = torch.from_numpy(u0[y0:y0 + nl])
    v[1:-1] = torch.from_numpy(v0[y0:y0 + nl])
    for _ in range(steps):
        slab.exchange_halos([u, v], 1, periodic=True)
        un, vn = np.zeros((nl + 2, nx), np.float32), np.zeros((nl + 2, nx), np.float32)
        oracle.lib.oracle_gs_step_rows(u.numpy(), v.numpy(), un, vn, nx, nl + 2, 0.2, 0.1, 1.0, 1.0,
                                       0.03, 0.06, 1, nl + 1)
        u[1:-1] = torch.from_numpy(un[1:-1])
        v[1:-1] = torch.from_numpy(vn[1:-1])
    np.save(os.path.join(out, f"u{rank}.npy"), u[1:-1].numpy())
    np.save(os.path.join(out, f"v{rank}.npy"), v[1:-1].numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_exchange_reproduces_single_domain(world, tmp_path):
    nx, ny, steps = 48, 37, 12
    mp.spawn(_gs_worker, args=(world, _free_port(), nx, ny, steps, str(tmp_path)), nprocs=world, join=True)
    u = np.vstack([np.load(tmp_path / f"u{r}.npy") for r in range(world)])
    v = np.vstack([np.load(tmp_path / f"v{r}.npy") for r in range(world)])
    u0, v0 = oracle.gs_init_pattern(nx, ny)
    eu, ev = oracle.gs_run(u0, v0, steps)
    assert np.array_equal(u, eu) and np.array_equal(v, ev)


def _chain_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    halo, nl, W = 2, 5, 7
    # planes laid out (field, rows, W) like the 2-D hypersonic state: decomposed dimension = 1
    p = torch.full((4, nl + 2 * halo, W), -1.0, dtype=torch.float64)
    for r in range(nl):
        p[:, halo + r, :] = 100 * rank + r
    slab.exchange_halos([p], halo, periodic=False, dim=1)
    m = torch.tensor([float(rank + 1)], dtype=torch.float64)
    slab.allreduce_max_(m)
    np.save(os.path.join(out, f"p{rank}.npy"), p.numpy())
    np.save(os.path.join(out, f"m{rank}.npy"), m.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_chain_exchange_and_max(world, tmp_path):
    mp.spawn(_chain_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    halo, nl = 2, 5
    for r in range(world):
        p = np.load(tmp_path / f"p{r}.npy")
        assert np.load(tmp_path / f"m{r}.npy")[0] == world
        top, bot = p[:, :halo, :], p[:, -halo:, :]
        if r == 0:
            assert (top == -1).all()                         # outer ghosts untouched (solver BC)
        else:
            assert (top[:, 0] == 100 * (r - 1) + nl - 2).all() and (top[:, 1] == 100 * (r - 1) + nl - 1).all()
        if r == world - 1:
            assert (bot == -1).all()
        else:
            assert (bot[:, 0] == 100 * (r + 1)).all() and (bot[:, 1] == 100 * (r + 1) + 1).all()


def _hyp3d_worker(rank, world, port, n, steps, out):
    """z-slab ring of the 3-D solver: 3 ghost planes per side, max-wavespeed all-reduce feeding the
    d_tau controller (tau_hypersonic_3d_cuda.cu:1680-1704) on every rank — the protocol of
    tau_hyp3d_step_begin / _end, with the CPU oracle as the per-slab stepper."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    nx, ny, nz = n
    prm = oracle.hyp3d_params(nx, ny, nz)
    planes, solid = oracle.hyp3d_init(prm)
    z0, nl = slab.partition_rows(nz, world)[rank]
    G = 3
    loc = oracle.hyp3d_params(nx, ny, nl + 2 * G)
    loc.dz = prm.dz                      # the slab is a window of the global grid, not a smaller grid
    loc.sdf_cz = prm.sdf_cz - (z0 - G) * prm.dz   # sphere centre in slab-local coordinates
    step = oracle.lib.oracle_hyp3d_step_planes
    step.argtypes = [C.POINTER(oracle.Hyp3dParams), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), oracle.u8p,
                     C.c_float, C.c_float, C.c_int, C.c_int]
    step.restype = C.c_float
    full = [np.asarray(p_, np.float32).reshape(nz, ny, nx) for p_ in planes]
    sol_full = np.asarray(solid, np.uint8).reshape(nz, ny, nx)
    st = torch.zeros(6, nl + 2 * G, ny, nx)
    for f in range(6):
        st[f, G:G + nl] = torch.from_numpy(full[f][z0:z0 + nl].copy())
    sol = torch.zeros(nl + 2 * G, ny, nx, dtype=torch.uint8)
    sol[G:G + nl] = torch.from_numpy(sol_full[z0:z0 + nl].copy())
    slab.exchange_halos([sol], G, periodic=True, dim=0)       # static: once
    libm = C.CDLL("libm.so.6")           # the oracle's clock uses libm's expf: same function, same bits
    libm.expf.argtypes, libm.expf.restype = [C.c_float], C.c_float
    t, d_tau = np.float32(0.012), np.float32(2e-3)
    for _ in range(steps):
        slab.exchange_halos([st], G, periodic=True, dim=1)
        t = np.float32(t * np.float32(libm.expf(float(d_tau))))
        dt = np.float32(t * d_tau)
        gain = np.float32(min(max(t / np.float32(0.02), 0.0), 1.0))
        a = [np.ascontiguousarray(st[f].numpy()) for f in range(6)]
        b = [np.zeros_like(x) for x in a]
        pa = (C.c_void_p * 6)(*[x.ctypes.data for x in a])
        pb = (C.c_void_p * 6)(*[x.ctypes.data for x in b])
        maxs = step(C.byref(loc), pa, pb, np.ascontiguousarray(sol.numpy()).ravel(), dt, gain, G, G + nl)
        m = torch.tensor([maxs], dtype=torch.float32)
        slab.allreduce_max_(m)
        dt_cfl = np.float32(prm.cfl / max(np.float32(m.item()), np.float32(1e-9)))
        if dt > np.float32(1.10) * dt_cfl:
            d_tau = np.float32(d_tau * np.float32(0.80))
        elif dt < np.float32(0.85) * dt_cfl:
            d_tau = np.float32(d_tau * np.float32(1.10))
        d_tau = np.float32(min(max(d_tau, np.float32(1e-7)), np.float32(5e-2)))
        for f in range(6):
            st[f, G:G + nl] = torch.from_numpy(b[f][G:G + nl])
    np.save(os.path.join(out, f"s{rank}.npy"), st[:, G:G + nl].numpy())
    np.save(os.path.join(out, f"c{rank}.npy"), np.array([t, d_tau], np.float32))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_hyp3d_z_slab_ring_reproduces_single_domain(world, tmp_path):
    n, steps = (20, 12, 18), 5
    mp.spawn(_hyp3d_worker, args=(world, _free_port(), n, steps, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"s{r}.npy") for r in range(world)], axis=1)
    prm = oracle.hyp3d_params(*n)
    planes, solid = oracle.hyp3d_init(prm)
    ref, ck, _, _ = oracle.hyp3d_run(prm, planes, solid, steps, (0.012, 2e-3))
    for f in range(6):
        assert np.array_equal(got[f].ravel(), ref[f]), f
    for r in range(world):
        assert tuple(np.load(tmp_path / f"c{r}.npy")) == tuple(np.float32(x) for x in ck)


def test_partition_rows_cost_weighted():
    """equal cost per slab instead of equal rows (bench.py at N > 1): contiguous, complete, >= 1 row each, and the slabs that
    hold the costlier rows are shorter"""
    from fluid_sims_b200.hypersonic2d import SimConfig
    cfg = SimConfig.default(4096, 4096)
    w = slab.hyp2d_row_costs(cfg)
    assert len(w) == 4096 and min(w) == 1.0 and max(w) > 1.0
    for parts in (2, 3, 4, 8):
        p = slab.partition_rows(4096, parts, w)
        assert p[0][0] == 0 and all(p[i][0] + p[i][1] == p[i + 1][0] for i in range(parts - 1)) and p[-1][0] + p[-1][1] == 4096
        cost = [sum(w[b:b + c]) for b, c in p]
        assert max(cost) - min(cost) <= 2.2 and min(c for _, c in p) >= 1
    p8 = slab.partition_rows(4096, 8, w)
    assert p8[3][1] < p8[0][1] and p8[4][1] < p8[7][1]
    assert slab.partition_rows(10, 3, [1.0] * 10) == [(0, 3), (3, 3), (6, 4)]
    assert slab.partition_rows(5, 5, [9, 1, 1, 1, 1]) == [(i, 1) for i in range(5)]
    assert slab.partition_rows(4096, 1, w) == [(0, 4096)]


def test_rebalance_rows_moves_the_cuts_towards_equal_cost():
    """measured load balancing of the 2-D slabs (bench.py at N > 1): rows of a slab cost (busy - fixed) / rows; the new
    partition is contiguous, complete, gives the slow slabs fewer rows, and equalises the modelled cost"""
    parts = slab.partition_rows(4096, 8)
    busy = [65.8, 66.8, 72.0, 74.1, 74.5, 72.5, 67.1, 62.6]           # measured at N = 8 (profiles/r2d_bench_n8.json)
    new = slab.rebalance_rows(parts, busy)
    assert new[0][0] == 0 and all(new[i][0] + new[i][1] == new[i + 1][0] for i in range(7)) and new[-1][0] + new[-1][1] == 4096
    assert new[4][1] < 512 < new[7][1] and new[3][1] < 512 < new[0][1]
    per_row = [(b - 10.0) / 512 for b in busy]
    cost = [sum(per_row[min(y // 512, 7)] for y in range(b, b + c)) for b, c in new]
    assert max(cost) - min(cost) < 0.5 and max(cost) < max(b - 10.0 for b in busy) - 3.0
    assert slab.rebalance_rows(parts, [70.0] * 8) == parts
    assert slab.rebalance_rows([(0, 8), (8, 8)], [1000.0, 20.0]) == [(0, 8), (8, 8)]     # would leave < 8 rows: keep


def test_partition_rows():
    assert slab.partition_rows(4096, 8) == [(512 * i, 512) for i in range(8)]
    parts = slab.partition_rows(37, 3)
    assert parts == [(0, 13), (13, 12), (25, 12)] and sum(c for _, c in parts) == 37
    with pytest.raises(ValueError):
        slab.partition_rows(2, 3)
