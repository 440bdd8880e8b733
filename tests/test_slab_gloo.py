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


def test_partition_rows():
    assert slab.partition_rows(4096, 8) == [(512 * i, 512) for i in range(8)]
    parts = slab.partition_rows(37, 3)
    assert parts == [(0, 13), (13, 12), (25, 12)] and sum(c for _, c in parts) == 37
    with pytest.raises(ValueError):
        slab.partition_rows(2, 3)
