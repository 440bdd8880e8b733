"""Host-side mirror of the reference `tau_sph` solver (tau_sph.cu) over the C-ABI: `Params`
(:49-85, simulation fields), `reset_particles` (:493-510) and the per-frame step block (:663-722),
here `SPH.step()`."""
from __future__ import annotations

import ctypes as C
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
        return dict(Gx=gx.value, Gy=gy.value, cell=cell.value, h=h.value, mass=m.value)

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
