"""Host-side mirror of the reference `tau_sph` solver (tau_sph.cu) over the C-ABI: `Params`
(:49-85, simulation fields), `reset_particles` (:493-510) and the per-frame step block (:663-722),
here `SPH.step()`."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from ._lib import check, declare

_FIELDS = [("N", C.c_int), ("boxX", C.c_float), ("boxY", C.c_float), ("dTau", C.c_float),
           ("t0", C.c_float), ("CFL", C.c_float), ("rho0", C.c_float), ("c0", C.c_float),
           ("gammaEOS", C.c_float), ("hMul", C.c_float), ("viscAlpha", C.c_float),
           ("gravity", C.c_float), ("rain", C.c_int), ("useVisc", C.c_int), ("useGrav", C.c_int),
           ("viscSub", C.c_int), ("useXSPH", C.c_int), ("xsphEps", C.c_float), ("seed", C.c_int)]


class _CParams(C.Structure):
    _fields_ = _FIELDS


_h = C.c_void_p
_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_reset = declare("tau_sph_reset_particles", [C.POINTER(_CParams), _f32, _f32], None)
_create = declare("tau_sph_create", [C.POINTER(_CParams), C.c_int, C.c_void_p, C.POINTER(_h)])
_init = declare("tau_sph_init", [_h])
_upload = declare("tau_sph_upload", [_h, _f32, _f32])
_step = declare("tau_sph_step", [_h, C.c_int])
_shard_cfg = declare("tau_sph_shard_config", [_h, C.c_int, C.c_int])
_shard_buf = declare("tau_sph_shard_buffers", [_h, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int)])
_shard_begin = declare("tau_sph_shard_substep_begin", [_h])
_shard_end = declare("tau_sph_shard_substep_end", [_h])
_clock = declare("tau_sph_clock", [_h, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_longlong)])
_download = declare("tau_sph_download", [_h, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p])
_dl_sort = declare("tau_sph_download_sort", [_h, _u32, _u32])
_rasterize = declare("tau_sph_rasterize", [_h, C.c_int, C.c_int, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")])
_sort_pairs = declare("tau_sph_sort_pairs", [_h, _u32, _u32, _u32])
_grid = declare("tau_sph_grid", [_h, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float),
                                 C.POINTER(C.c_float), C.POINTER(C.c_float)])
_subx = declare("tau_sph_subx", [_h])
_sync = declare("tau_sph_sync", [_h])
_substeps = declare("tau_sph_substeps_done", [_h], C.c_longlong)
_launches = declare("tau_sph_launch_count", [_h], C.c_longlong)
_last_ms = declare("tau_sph_last_step_ms", [_h, C.POINTER(C.c_float)])
_destroy = declare("tau_sph_destroy", [_h])


@dataclass
class Params:
    N: int = 1 << 16
    boxX: float = 1.0
    boxY: float = 1.0
    dTau: float = 1.0
    t0: float = 1.0
    CFL: float = 1.0
    rho0: float = 1.0
    c0: float = 1.0
    gammaEOS: float = 1.0
    hMul: float = 2.0
    viscAlpha: float = 0.25
    gravity: float = 9.81
    rain: int = 1
    useVisc: int = 1
    useGrav: int = 1
    viscSub: int = 1
    useXSPH: int = 0
    xsphEps: float = 0.25
    seed: int = 69420

    def _c(self) -> _CParams:
        return _CParams(*[getattr(self, f[0]) for f in _FIELDS])

    def as19(self):
        return np.array([float(getattr(self, f[0])) for f in _FIELDS], np.float32)


def reset_particles(p: Params):
    """(pos, vel) as (N, 2) float32 — reset_particles(), tau_sph.cu:493-510."""
    pos = np.empty((p.N, 2), np.float32)
    vel = np.empty((p.N, 2), np.float32)
    cp = p._c()
    _reset(C.byref(cp), pos.reshape(-1), vel.reshape(-1))
    return pos, vel


class SPH:
    def __init__(self, params: Params | None = None, device: int = 0, stream: int | None = None):
        self.params = params or Params()
        self.device = device
        self._handle = _h()
        cp = self.params._c()
        check(_create(C.byref(cp), device, C.c_void_p(stream or 0), C.byref(self._handle)))

    def init(self):
        check(_init(self._handle))
        return self

    def upload(self, pos, vel):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1)
        vel = np.ascontiguousarray(vel, np.float32).reshape(-1)
        assert pos.size == 2 * self.params.N and vel.size == pos.size
        check(_upload(self._handle, pos, vel))
        return self

    def step(self, nframes: int = 1):
        check(_step(self._handle, nframes))
        return self

    # ---- multi-GPU: replicated state, sharded work ----------------------------------------------
    def shard_config(self, rank: int, world: int):
        check(_shard_cfg(self._handle, rank, world))
        return self

    def shard_buffers(self):
        """(sxy_new_ptr, svel_new_ptr, chunk) — the sorted-copy arrays to all-gather per sub-step."""
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_int()
        check(_shard_buf(self._handle, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def shard_substep(self, all_gather):
        """One sub-step of a sharded handle; `all_gather(sxy_new_ptr, svel_new_ptr, chunk)` must
        exchange the chunks between the ranks (see bench_all.py)."""
        check(_shard_begin(self._handle))
        all_gather(*self.shard_buffers())
        check(_shard_end(self._handle))

    def clock(self):
        t, tau, st = C.c_float(), C.c_float(), C.c_longlong()
        check(_clock(self._handle, C.byref(t), C.byref(tau), C.byref(st)))
        return float(t.value), float(tau.value), int(st.value)

    def download(self):
        """(pos, vel, s, press) in original particle order."""
        n = self.params.N
        pos, vel = np.empty((n, 2), np.float32), np.empty((n, 2), np.float32)
        s, pr = np.empty(n, np.float32), np.empty(n, np.float32)
        check(_download(self._handle, pos.ctypes.data, vel.ctypes.data, s.ctypes.data, pr.ctypes.data))
        return pos, vel, s, pr

    def rasterize(self, W: int, H: int):
        """Particle counts on the W x 2H half-block raster (k_rasterize, tau_sph.cu:363-374)."""
        g = np.empty((2 * H, W), np.int32)
        check(_rasterize(self._handle, W, H, g))
        return g

    def download_sort(self):
        n = self.params.N
        k, v = np.empty(n, np.uint32), np.empty(n, np.uint32)
        check(_dl_sort(self._handle, k, v))
        return k, v

    def sort_pairs(self, keys):
        keys = np.ascontiguousarray(keys, np.uint32)
        assert keys.size == self.params.N
        ko, vo = np.empty_like(keys), np.empty_like(keys)
        check(_sort_pairs(self._handle, keys, ko, vo))
        return ko, vo

    def grid(self):
        gx, gy = C.c_int(), C.c_int()
        cell, h, m = C.c_float(), C.c_float(), C.c_float()
        check(_grid(self._handle, C.byref(gx), C.byref(gy), C.byref(cell), C.byref(h), C.byref(m)))
        return dict(Gx=gx.value, Gy=gy.value, cell=cell.value, h=h.value, mass=m.value, subx=int(_subx(self._handle)))

    def sync(self):
        check(_sync(self._handle))

    @property
    def substeps_done(self) -> int:
        return int(_substeps(self._handle))

    @property
    def launch_count(self) -> int:
        return int(_launches(self._handle))

    def last_step_ms(self) -> float:
        ms = C.c_float()
        check(_last_ms(self._handle, C.byref(ms)))
        return float(ms.value)

    def close(self):
        if self._handle:
            _destroy(self._handle)
            self._handle = _h()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- the particle set sharded by hash-bin stripes (tau_sph_stripe_*, include/tau_b200.h) -------------------------
_s = C.c_void_p
_st_create = declare("tau_sph_stripe_create", [C.POINTER(_CParams), C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(_s)])
_st_upload = declare("tau_sph_stripe_upload", [_s, _f32, _f32])
_st_xbuf = declare("tau_sph_stripe_xbuf", [_s, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)])
_st_phase = declare("tau_sph_stripe_phase", [_s, C.c_int])
_st_hist_begin = declare("tau_sph_stripe_hist_begin", [_s, C.POINTER(C.c_void_p), C.POINTER(C.c_int)])
_st_hist_apply = declare("tau_sph_stripe_hist_apply", [_s])
_st_status = declare("tau_sph_stripe_status", [_s, C.POINTER(C.c_int)])
_st_download = declare("tau_sph_stripe_download", [_s, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.POINTER(C.c_int)])
_st_clock = declare("tau_sph_stripe_clock", [_s, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_longlong)])
_st_capacity = declare("tau_sph_stripe_capacity", [_s, C.POINTER(C.c_int), C.POINTER(C.c_int)])
_st_sync = declare("tau_sph_stripe_sync", [_s])
_st_substeps = declare("tau_sph_stripe_substeps_done", [_s], C.c_longlong)
_st_launches = declare("tau_sph_stripe_launch_count", [_s], C.c_longlong)
_st_destroy = declare("tau_sph_stripe_destroy", [_s])


class SPHStripes:
    """One rank of the stripe-sharded SPH solver: holds the particles of its stripe of hash-bin rows only.

    `exchange(send_lo, send_hi, recv_lo, recv_hi)` moves whole message buffers between neighbouring ranks (send_lo
    to rank-1, send_hi to rank+1, recv_lo from rank-1, recv_hi from rank+1; arguments are (device pointer, 32-bit
    words), None where there is no neighbour) and `allreduce_sum(ptr, n_ints)` sums the row histogram in place;
    `nccl_plumbing()` returns the pair for torch.distributed (NCCL send/recv over NVLink)."""

    def __init__(self, params: Params, rank: int, world: int, device: int = 0, stream: int | None = None,
                 exchange=None, allreduce_sum=None, rebalance_every: int = 16):
        self.params, self.rank, self.world, self.device = params, rank, world, device
        self.exchange, self.allreduce_sum, self.rebalance_every = exchange, allreduce_sum, rebalance_every
        self._handle = _s()
        cp = params._c()
        check(_st_create(C.byref(cp), device, C.c_void_p(stream or 0), rank, world, C.byref(self._handle)))
        self._bufs = []
        for which in range(8):
            p, w = C.c_void_p(), C.c_longlong()
            check(_st_xbuf(self._handle, which, C.byref(p), C.byref(w)))
            self._bufs.append((p.value, int(w.value)))

    def upload(self, pos, vel):
        """the WHOLE particle set on every rank (reference order); each keeps its stripe"""
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1)
        vel = np.ascontiguousarray(vel, np.float32).reshape(-1)
        assert pos.size == 2 * self.params.N and vel.size == pos.size
        check(_st_upload(self._handle, pos, vel))
        return self

    def init(self):
        return self.upload(*reset_particles(self.params))

    def _xchg(self, base):
        if self.world == 1:
            return
        lo, hi = self.rank > 0, self.rank < self.world - 1
        b = self._bufs
        self.exchange(b[base] if lo else None, b[base + 1] if hi else None, b[base + 2] if lo else None,
                      b[base + 3] if hi else None)

    def substep(self):
        """one sub-step (tau_sph.cu:676-721) of this rank's stripe; every rank must call it"""
        xsph = self.params.useXSPH and self.params.xsphEps > 0
        check(_st_phase(self._handle, 0))
        self._xchg(0)
        check(_st_phase(self._handle, 1))
        self._xchg(4)
        check(_st_phase(self._handle, 2))
        if xsph:
            self._xchg(0)
        check(_st_phase(self._handle, 3))
        if self.world > 1 and self.rebalance_every and self.substeps_done % self.rebalance_every == 0:
            self.rebalance()

    def rebalance(self):
        p, n = C.c_void_p(), C.c_int()
        check(_st_hist_begin(self._handle, C.byref(p), C.byref(n)))
        self.allreduce_sum(p.value, n.value)
        check(_st_hist_apply(self._handle))

    def step(self, nframes: int = 1):
        k = self.params.viscSub if self.params.viscSub > 0 else 1
        for _ in range(nframes * k):
            self.substep()
        return self

    def status(self):
        out = (C.c_int * 8)()
        check(_st_status(self._handle, out))
        keys = ("n_own", "n_ghost", "err", "max_send", "cap", "xcap", "row_begin", "row_end")
        return dict(zip(keys, [int(v) for v in out]))

    def download_local(self):
        """(ids, pos, vel, s, press) of the particles this rank owns, local order"""
        cap, xcap = C.c_int(), C.c_int()
        check(_st_capacity(self._handle, C.byref(cap), C.byref(xcap)))
        ids = np.empty(cap.value, np.uint32)
        pos, vel = np.empty((cap.value, 2), np.float32), np.empty((cap.value, 2), np.float32)
        s, pr = np.empty(cap.value, np.float32), np.empty(cap.value, np.float32)
        n = C.c_int()
        check(_st_download(self._handle, ids.ctypes.data, pos.ctypes.data, vel.ctypes.data, s.ctypes.data, pr.ctypes.data,
                           C.byref(n)))
        keep = ids[:n.value] != 0xFFFFFFFF       # a particle the rain re-homed to another rank leaves a hole until the next sub-step
        return ids[:n.value][keep], pos[:n.value][keep], vel[:n.value][keep], s[:n.value][keep], pr[:n.value][keep]

    def clock(self):
        t, tau, st = C.c_float(), C.c_float(), C.c_longlong()
        check(_st_clock(self._handle, C.byref(t), C.byref(tau), C.byref(st)))
        return float(t.value), float(tau.value), int(st.value)

    def sync(self):
        check(_st_sync(self._handle))

    @property
    def substeps_done(self) -> int:
        return int(_st_substeps(self._handle))

    @property
    def launch_count(self) -> int:
        return int(_st_launches(self._handle))

    def close(self):
        if self._handle:
            check(_st_destroy(self._handle))
            self._handle = _s()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def assemble(parts, N):
    """[(ids, pos, vel, s, press)] of every rank -> (pos, vel, s, press) in the reference's particle order; every id
    must appear exactly once"""
    ids = np.concatenate([p[0] for p in parts])
    assert ids.size == N and np.array_equal(np.sort(ids), np.arange(N, dtype=ids.dtype)), "stripes lost or duplicated particles"
    out = []
    for k, shape in ((1, (N, 2)), (2, (N, 2)), (3, (N,)), (4, (N,))):
        a = np.empty(shape, np.float32)
        a[ids] = np.concatenate([p[k] for p in parts])
        out.append(a)
    return tuple(out)


def nccl_plumbing(device: int):
    """(exchange, allreduce_sum) for SPHStripes over torch.distributed: whole message buffers by NCCL send / recv
    between neighbouring ranks (a chain in y), the row histogram by all_reduce — on torch's current stream, which the
    caller makes the handle's stream."""
    import torch
    import torch.distributed as dist

    from . import slab
    rank = dist.get_rank()
    views = {}

    def view(buf):
        if buf not in views:
            views[buf] = slab.wrap_plane(buf[0], (buf[1],), torch.int32, device)
        return views[buf]

    def exchange(send_lo, send_hi, recv_lo, recv_hi):
        if os.environ.get("TAU_SPH_STRIPE_NOXCHG") == "1":     # timing experiments only: results are wrong without the exchange
            return
        ops = []
        if send_lo is not None:
            ops += [dist.P2POp(dist.isend, view(send_lo), rank - 1), dist.P2POp(dist.irecv, view(recv_lo), rank - 1)]
        if send_hi is not None:
            ops += [dist.P2POp(dist.isend, view(send_hi), rank + 1), dist.P2POp(dist.irecv, view(recv_hi), rank + 1)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def allreduce_sum(ptr, n):
        dist.all_reduce(slab.wrap_plane(ptr, (n,), torch.int32, device), op=dist.ReduceOp.SUM)

    return exchange, allreduce_sum
