"""Host-side mirror of the reference `tau3d` solver (tau_hypersonic_3d_cuda.cu) over the C-ABI.
`Params` mirrors `struct Params` (:21-42) with main()'s hard-coded values (:1531-1557) as defaults;
`Hypersonic3D.step()` is the loop body :1679-1712."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import check, declare

_FIELDS = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int)] + \
          [(n, C.c_float) for n in ("dx", "dy", "dz", "cfl", "u_ref", "R", "gamma_floor", "Twall",
                                    "tau_vib", "theta_v", "sdf_cx", "sdf_cy", "sdf_cz", "sdf_r",
                                    "inflow_r", "inflow_p", "inflow_u", "inflow_v", "inflow_w")] + \
          [("sponge_n", C.c_int), ("sponge_strength", C.c_float), ("sponge_out_n", C.c_int),
           ("sponge_out_strength", C.c_float), ("t0", C.c_float), ("d_tau0", C.c_float)]


class _CParams(C.Structure):
    _fields_ = _FIELDS


_h = C.c_void_p
_default = declare("tau_hyp3d_default_params", [C.POINTER(_CParams), C.c_int, C.c_int, C.c_int], None)
_create = declare("tau_hyp3d_create", [C.POINTER(_CParams), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                       C.POINTER(_h)])
_init = declare("tau_hyp3d_init", [_h])
_upload = declare("tau_hyp3d_upload", [_h, C.POINTER(C.c_void_p), C.c_void_p])
_step = declare("tau_hyp3d_step", [_h, C.c_int])
_step_begin = declare("tau_hyp3d_step_begin", [_h])
_step_end = declare("tau_hyp3d_step_end", [_h])
_clock = declare("tau_hyp3d_clock", [_h] + [C.POINTER(C.c_float)] * 4)
_download = declare("tau_hyp3d_download", [_h, C.POINTER(C.c_void_p), C.c_void_p])
_sync = declare("tau_hyp3d_sync", [_h])
_vis = declare("tau_hyp3d_vis", [_h, C.c_int, np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")])
VIS_MODES = ("schlieren_rho", "log_rho", "log_p", "speed", "mach", "vort_mag", "div", "q_criterion",
             "schlieren_export")   # 8: th3cs.cu k_schlieren_export :641-673
_export_frame = declare("tau_hyp3d_export_frame", [_h, np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS"),
                                                   C.POINTER(C.c_float)])
_devstate = declare("tau_hyp3d_device_state", [_h, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)])
_steps_done = declare("tau_hyp3d_steps_done", [_h], C.c_longlong)
_launches = declare("tau_hyp3d_launch_count", [_h], C.c_longlong)
_last_ms = declare("tau_hyp3d_last_step_ms", [_h, C.POINTER(C.c_float)])
_destroy = declare("tau_hyp3d_destroy", [_h])
_g = C.c_void_p
_g_create = declare("tau_hyp3d_group_create", [C.POINTER(_CParams), C.c_int, C.c_void_p, C.POINTER(_g)])
_g_init = declare("tau_hyp3d_group_init", [_g])
_g_upload = declare("tau_hyp3d_group_upload", [_g, C.POINTER(C.c_void_p), C.c_void_p])
_g_step = declare("tau_hyp3d_group_step", [_g, C.c_int])
_g_clock = declare("tau_hyp3d_group_clock", [_g] + [C.POINTER(C.c_float)] * 4)
_g_download = declare("tau_hyp3d_group_download", [_g, C.POINTER(C.c_void_p), C.c_void_p])
_g_destroy = declare("tau_hyp3d_group_destroy", [_g])

HALO = 3
PLANES = ("xi", "phix", "phiy", "phiz", "lam", "zet")


@dataclass
class Params:
    nx: int = 64
    ny: int = 64
    nz: int = 64
    dx: float = 0.0
    dy: float = 0.0
    dz: float = 0.0
    cfl: float = 0.3333
    u_ref: float = 10.0
    R: float = 10.0
    gamma_floor: float = 1.1
    Twall: float = 0.02
    tau_vib: float = 2e-4
    theta_v: float = 0.2
    sdf_cx: float = 0.5
    sdf_cy: float = 0.5
    sdf_cz: float = 0.5
    sdf_r: float = 0.25
    inflow_r: float = 0.02
    inflow_p: float = 0.02
    inflow_u: float = 100.0
    inflow_v: float = 0.0
    inflow_w: float = 0.0
    sponge_n: int = 24
    sponge_strength: float = 0.05
    sponge_out_n: int = 24
    sponge_out_strength: float = 0.05
    t0: float = 1e-5
    d_tau0: float = 1e-3

    @classmethod
    def default(cls, nx: int = 64, ny: int = 64, nz: int = 64, **over) -> "Params":
        c = _CParams()
        _default(C.byref(c), nx, ny, nz)
        kw = {f[0]: getattr(c, f[0]) for f in _FIELDS}
        kw.update(over)
        return cls(**kw)

    def _c(self) -> _CParams:
        return _CParams(*[getattr(self, f[0]) for f in _FIELDS])


class Hypersonic3D:
    def __init__(self, params: Params | None = None, device: int = 0, z_begin: int = 0,
                 nz_local: int | None = None, stream: int | None = None):
        self.params = params or Params.default()
        self.nz_local = self.params.nz if nz_local is None else nz_local
        self.z_begin = z_begin
        self.device = device
        self._handle = _h()
        cp = self.params._c()
        check(_create(C.byref(cp), device, z_begin, self.nz_local, C.c_void_p(stream or 0),
                      C.byref(self._handle)))

    @property
    def shape(self):
        return (self.nz_local, self.params.ny, self.params.nx)

    def init(self):
        check(_init(self._handle))
        return self

    def upload(self, planes, clock=None):
        arrs = [np.ascontiguousarray(p, np.float32).reshape(self.shape) for p in planes]
        ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in arrs])
        ck = None
        if clock is not None:
            ck = np.array(clock, np.float32)
        check(_upload(self._handle, ptrs, C.c_void_p(ck.ctypes.data if ck is not None else 0)))
        return self

    def step(self, nsteps: int = 1):
        check(_step(self._handle, nsteps))
        return self

    def step_begin(self):
        check(_step_begin(self._handle))

    def step_end(self):
        check(_step_end(self._handle))

    def clock(self):
        """(t, d_tau) for the next step, (dt, max wavespeed sum) of the last one."""
        v = [C.c_float() for _ in range(4)]
        check(_clock(self._handle, *[C.byref(x) for x in v]))
        return tuple(float(x.value) for x in v)

    def download(self):
        arrs = [np.empty(self.shape, np.float32) for _ in range(6)]
        solid = np.empty(self.shape, np.uint8)
        ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in arrs])
        check(_download(self._handle, ptrs, C.c_void_p(solid.ctypes.data)))
        return arrs, solid

    def vis(self, mode: int):
        """k_vis (tau_hypersonic_3d_cuda.cu:800-905): the diagnostic scalar field, shape (nz_local, ny, nx)."""
        out = np.empty(self.shape, np.float32)
        check(_vis(self._handle, mode, out))
        return out

    def export_frame(self):
        """One frame of th3cs.cu's `.4spl` export loop (:1193-1222) on the device: ((nz_local, ny, nx) uint8
        palette indices, (min, max) of the schlieren field)."""
        idx = np.empty(self.shape, np.uint8)
        mm = (C.c_float * 2)()
        check(_export_frame(self._handle, idx.ravel(), mm))
        return idx, (float(mm[0]), float(mm[1]))

    def sync(self):
        check(_sync(self._handle))

    def device_state(self):
        p, m = C.c_void_p(), C.c_void_p()
        check(_devstate(self._handle, C.byref(p), C.byref(m)))
        return p.value, m.value

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


class Hypersonic3DGroup:
    """The 3-D solver on `ngpus` devices from ONE process (tau_hyp3d_group_*): one z-slab handle per device, the periodic
    ring's ghost planes by cudaMemcpyPeerAsync, the max wavespeed folded on the host once per step.  Planes cover the whole
    grid in the reference layout; results are bit-identical to Hypersonic3D on one device."""

    def __init__(self, params: Params | None = None, ngpus: int = 1):
        self.params = params or Params.default()
        self.ngpus = ngpus
        self._handle = _g()
        cp = self.params._c()
        check(_g_create(C.byref(cp), ngpus, None, C.byref(self._handle)))

    @property
    def shape(self):
        return (self.params.nz, self.params.ny, self.params.nx)

    def init(self):
        check(_g_init(self._handle))
        return self

    def upload(self, planes, clock=None):
        arrs = [np.ascontiguousarray(p, np.float32).reshape(self.shape) for p in planes]
        ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in arrs])
        ck = np.array(clock, np.float32) if clock is not None else None
        check(_g_upload(self._handle, ptrs, C.c_void_p(ck.ctypes.data if ck is not None else 0)))
        return self

    def step(self, nsteps: int = 1):
        check(_g_step(self._handle, nsteps))
        return self

    def clock(self):
        v = [C.c_float() for _ in range(4)]
        check(_g_clock(self._handle, *[C.byref(x) for x in v]))
        return tuple(float(x.value) for x in v)

    def download(self):
        out = [np.empty(self.shape, np.float32) for _ in range(6)]
        solid = np.empty(self.shape, np.uint8)
        ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in out])
        check(_g_download(self._handle, ptrs, C.c_void_p(solid.ctypes.data)))
        return out, solid

    def close(self):
        if self._handle:
            check(_g_destroy(self._handle))
            self._handle = _g()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
