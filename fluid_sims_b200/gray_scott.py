"""Host-side mirror of the reference `tgs` solver (tau_gray_scott.cu) over the C-ABI.

Names follow the reference: `Params` (tau_gray_scott.cu:43-61), `init_pattern` (:173-204) and the
per-step sequence `step_kernel + swap` (:321-329), here `GrayScott.step()`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import check, declare


class _CParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("dx", C.c_float), ("dt", C.c_float),
                ("Du", C.c_float), ("Dv", C.c_float), ("feed", C.c_float), ("kill", C.c_float),
                ("seed", C.c_uint)]


_h = C.c_void_p
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_create = declare("tau_gs_create", [C.POINTER(_CParams), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.POINTER(_h)])
_init = declare("tau_gs_init", [_h])
_upload = declare("tau_gs_upload", [_h, _fp, _fp])
_step = declare("tau_gs_step", [_h, C.c_int])
_download = declare("tau_gs_download", [_h, _fp, _fp])
_sync = declare("tau_gs_sync", [_h])
_planes = declare("tau_gs_device_planes", [_h, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)])
_steps_done = declare("tau_gs_steps_done", [_h], C.c_longlong)
_launches = declare("tau_gs_launch_count", [_h], C.c_longlong)
_last_ms = declare("tau_gs_last_step_ms", [_h, C.POINTER(C.c_float)])
_destroy = declare("tau_gs_destroy", [_h])
_init_pattern = declare("tau_gs_init_pattern", [_fp, _fp, C.c_int, C.c_int, C.c_uint], None)


@dataclass
class Params:
    """Simulation fields of `struct Params` (tau_gray_scott.cu:43-61), same defaults."""
    nx: int = 128           # headless default, tau_gray_scott.cu:293-296
    ny: int = 128
    dx: float = 1.0
    dt: float = 1.0
    Du: float = 0.2
    Dv: float = 0.1
    feed: float = 0.03
    kill: float = 0.06
    seed: int = 1337

    def _c(self) -> _CParams:
        return _CParams(self.nx, self.ny, self.dx, self.dt, self.Du, self.Dv, self.feed,
                        self.kill, self.seed)


def init_pattern(nx: int, ny: int, seed: int = 1337):
    """Initial (u, v) planes — init_pattern(), tau_gray_scott.cu:173-204."""
    u = np.empty((ny, nx), np.float32)
    v = np.empty((ny, nx), np.float32)
    _init_pattern(u, v, nx, ny, seed)
    return u, v


class GrayScott:
    """One solver handle on one GPU.  `y_begin/ny_local` select a slab of rows (multi-GPU)."""

    def __init__(self, params: Params | None = None, device: int = 0, y_begin: int = 0,
                 ny_local: int | None = None, stream: int | None = None):
        self.params = params or Params()
        self.ny_local = self.params.ny if ny_local is None else ny_local
        self.y_begin = y_begin
        self.device = device
        self._handle = _h()
        cp = self.params._c()
        check(_create(C.byref(cp), device, y_begin, self.ny_local, C.c_void_p(stream or 0),
                      C.byref(self._handle)))

    @property
    def halo(self) -> int:
        return 1

    def init(self):
        check(_init(self._handle))
        return self

    def upload(self, u: np.ndarray, v: np.ndarray):
        u = np.ascontiguousarray(u, np.float32)
        v = np.ascontiguousarray(v, np.float32)
        if u.shape != (self.ny_local, self.params.nx) or v.shape != u.shape:
            raise ValueError(f"expected planes of shape {(self.ny_local, self.params.nx)}")
        check(_upload(self._handle, u, v))
        return self

    def step(self, nsteps: int = 1):
        check(_step(self._handle, nsteps))
        return self

    def download(self):
        u = np.empty((self.ny_local, self.params.nx), np.float32)
        v = np.empty_like(u)
        check(_download(self._handle, u, v))
        return u, v

    def sync(self):
        check(_sync(self._handle))

    def device_planes(self):
        """Raw device addresses (u, v) of the current planes incl. the two ghost rows."""
        pu, pv = C.c_void_p(), C.c_void_p()
        check(_planes(self._handle, C.byref(pu), C.byref(pv)))
        return pu.value, pv.value

    @property
    def steps_done(self) -> int:
        return int(_steps_done(self._handle))

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
