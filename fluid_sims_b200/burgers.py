"""Host-side mirror of the reference `tau_burgers` solver (tau_burgers.cu) over the C-ABI: `Params`
(:53-90, simulation fields and defaults), `initialize_host` (:250-304) and the per-step sequence
`do_step` + clock (:677-718, :768-769), here `Burgers.step()`."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, fields

import numpy as np

from ._lib import check, declare

_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


class _CParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int)] + [(k, C.c_float) for k in (
        "dx", "dy", "nu", "u0", "amp", "bsig", "swirl", "rc", "offx", "offy", "asym", "CFL", "tau0", "t0",
        "dtau")] + [("muscl", C.c_int), ("visc_substeps", C.c_int), ("colehopf", C.c_int), ("ck", C.c_int),
                    ("ca", C.c_float)]


@dataclass
class Params:
    """`struct Params` tau_burgers.cu:53-90 (simulation fields, the reference's defaults)."""
    nx: int = 512
    ny: int = 512
    dx: float = 1.0
    dy: float = 1.0
    nu: float = 0.1
    u0: float = 1.0
    amp: float = 1.0
    bsig: float = 16.0
    swirl: float = 10.0
    rc: float = 40.0
    offx: float = 0.0
    offy: float = 0.0
    asym: float = 0.0
    CFL: float = 0.45
    tau0: float = 0.0
    t0: float = 1.0
    dtau: float = 1.0
    muscl: int = 0
    visc_substeps: int = 1
    colehopf: int = 0
    ck: int = 4
    ca: float = 0.5

    def _c(self) -> _CParams:
        c = _CParams()
        for f in fields(self):
            setattr(c, f.name, getattr(self, f.name))
        return c

    @property
    def shape(self):
        return (1 if self.colehopf else self.ny, self.nx)  # :649-650


_h = C.c_void_p
_init_host = declare("tau_burgers_init_host", [C.POINTER(_CParams), _f32, _f32], None)
_create = declare("tau_burgers_create", [C.POINTER(_CParams), C.c_int, C.c_void_p, C.POINTER(_h)])
_init = declare("tau_burgers_init", [_h])
_upload = declare("tau_burgers_upload", [_h, _f32, _f32, C.c_void_p])
_step = declare("tau_burgers_step", [_h, C.c_int])
_clock = declare("tau_burgers_clock", [_h] + [C.POINTER(C.c_float)] * 3)
_download = declare("tau_burgers_download", [_h, _f32, _f32])
_ch_err = declare("tau_burgers_colehopf_error", [_h, C.POINTER(C.c_double)])
_sync = declare("tau_burgers_sync", [_h])
_steps_done = declare("tau_burgers_steps_done", [_h], C.c_longlong)
_launches = declare("tau_burgers_launch_count", [_h], C.c_longlong)
_last_ms = declare("tau_burgers_last_step_ms", [_h, C.POINTER(C.c_float)])
_destroy = declare("tau_burgers_destroy", [_h])


def initialize_host(p: Params):
    u, v = np.zeros(p.shape, np.float32), np.zeros(p.shape, np.float32)
    _init_host(C.byref(p._c()), u.ravel(), v.ravel())
    return u, v


class Burgers:
    def __init__(self, params: Params, device: int = 0, stream: int | None = None):
        self.params = params
        self._handle = _h()
        check(_create(C.byref(params._c()), device, C.c_void_p(stream or 0), C.byref(self._handle)))

    def init(self):
        check(_init(self._handle))
        return self

    def upload(self, phi_u, phi_v, clock=None):
        u = np.ascontiguousarray(phi_u, np.float32).reshape(self.params.shape)
        v = np.ascontiguousarray(phi_v, np.float32).reshape(self.params.shape)
        ck = None if clock is None else np.array(clock, np.float32)
        check(_upload(self._handle, u.ravel(), v.ravel(), C.c_void_p(ck.ctypes.data if ck is not None else 0)))
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
        u, v = np.empty(self.params.shape, np.float32), np.empty(self.params.shape, np.float32)
        check(_download(self._handle, u.ravel(), v.ravel()))
        return u, v

    def colehopf_error(self) -> float:
        e = C.c_double()
        check(_ch_err(self._handle, C.byref(e)))
        return float(e.value)

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
