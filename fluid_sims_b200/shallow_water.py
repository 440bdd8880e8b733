"""Host-side mirror of the reference `tau_sw` solver (tau_shallow_water.cu) over the C-ABI: `Params`
(:52-89, simulation fields and defaults), `initialize_host` (:238-277) and the per-step sequence
`do_step` + clock (:669-705, :767-768), here `ShallowWater.step()`.

Not yet run on hardware (written after the round-1 GPU budget was spent) — see NEXT.md."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, fields

import numpy as np

from ._lib import check, declare

_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


class _CParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int)] + [(k, C.c_float) for k in (
        "dx", "dy", "g", "f0", "nu", "H0", "bumpAmp", "bumpSigma", "CFL", "offx", "offy", "asym", "swirl",
        "swirlRc", "tau0", "t0", "dtau")]


@dataclass
class Params:
    """`struct Params` tau_shallow_water.cu:52-89 (simulation fields, the struct's defaults)."""
    nx: int = 512
    ny: int = 512
    dx: float = 1.0
    dy: float = 1.0
    g: float = 9.81
    f0: float = 1.0   # parsed by the reference, used by none of its kernels
    nu: float = 0.001
    H0: float = 1000.0
    bumpAmp: float = 1.0
    bumpSigma: float = 1.0
    CFL: float = 0.5
    offx: float = 100.0
    offy: float = 100.0
    asym: float = 10.0
    swirl: float = 1.0
    swirlRc: float = 100.0
    tau0: float = 0.0
    t0: float = 1.0
    dtau: float = 1.0

    def _c(self) -> _CParams:
        c = _CParams()
        for f in fields(self):
            setattr(c, f.name, getattr(self, f.name))
        return c

    @property
    def shape(self):
        return (self.ny, self.nx)


_h = C.c_void_p
_init_host = declare("tau_sw_init_host", [C.POINTER(_CParams), _f32, _f32, _f32], None)
_create = declare("tau_sw_create", [C.POINTER(_CParams), C.c_int, C.c_void_p, C.POINTER(_h)])
_init = declare("tau_sw_init", [_h])
_upload = declare("tau_sw_upload", [_h, _f32, _f32, _f32, C.c_void_p])
_step = declare("tau_sw_step", [_h, C.c_int])
_clock = declare("tau_sw_clock", [_h] + [C.POINTER(C.c_float)] * 3)
_download = declare("tau_sw_download", [_h, _f32, _f32, _f32])
_sync = declare("tau_sw_sync", [_h])
_steps_done = declare("tau_sw_steps_done", [_h], C.c_longlong)
_launches = declare("tau_sw_launch_count", [_h], C.c_longlong)
_last_ms = declare("tau_sw_last_step_ms", [_h, C.POINTER(C.c_float)])
_destroy = declare("tau_sw_destroy", [_h])


def initialize_host(p: Params):
    """(sigma = log h, u, v), each (ny, nx) float32 — host code in the reference too."""
    s, u, v = (np.zeros(p.shape, np.float32) for _ in range(3))
    _init_host(C.byref(p._c()), s.ravel(), u.ravel(), v.ravel())
    return s, u, v


class ShallowWater:
    def __init__(self, params: Params, device: int = 0, stream: int | None = None):
        self.params = params
        self._handle = _h()
        check(_create(C.byref(params._c()), device, C.c_void_p(stream or 0), C.byref(self._handle)))

    def init(self):
        check(_init(self._handle))
        return self

    def upload(self, sigma, u, v, clock=None):
        a = [np.ascontiguousarray(x, np.float32).reshape(self.params.shape).ravel() for x in (sigma, u, v)]
        ck = None if clock is None else np.array(clock, np.float32)
        check(_upload(self._handle, *a, C.c_void_p(ck.ctypes.data if ck is not None else 0)))
        return self

    def step(self, nsteps: int = 1):
        check(_step(self._handle, nsteps))
        return self

    def clock(self):
        """(t, tau, dt_eff of the last step)."""
        t, tau, dt = C.c_float(), C.c_float(), C.c_float()
        check(_clock(self._handle, C.byref(t), C.byref(tau), C.byref(dt)))
        return float(t.value), float(tau.value), float(dt.value)

    def download(self):
        s, u, v = (np.empty(self.params.shape, np.float32) for _ in range(3))
        check(_download(self._handle, s.ravel(), u.ravel(), v.ravel()))
        return s, u, v

    def sync(self):
        check(_sync(self._handle))

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
