"""Host-side mirror of the reference's CPU solver `tau_hypersonic` (tau_hypersonic.c, BASELINE config 1) over the
C-ABI (tau_hypc_*).  Names follow the reference: `init_sim` (:450), `step_physics` (:500-674) -> `step()`,
`view_mode` (:44; 2 = "speed mode") -> `render()`.  The arithmetic runs on the device in fp64 without FMA
contraction and equals the reference's own object code to 0 ulp.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, declare

_h = C.c_void_p
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_create = declare("tau_hypc_create", [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(_h)])
_init_host = declare("tau_hypc_init_host", [C.c_int, C.c_int, _f64p, _f64p, _f64p, _f64p, _u8p], None)
_init = declare("tau_hypc_init", [_h])
_upload = declare("tau_hypc_upload", [_h, C.POINTER(C.c_void_p), C.c_void_p, C.c_double])
_step = declare("tau_hypc_step", [_h, C.c_int])
_clock = declare("tau_hypc_clock", [_h, C.POINTER(C.c_double), C.POINTER(C.c_double)])
_download = declare("tau_hypc_download", [_h, C.POINTER(C.c_void_p), C.c_void_p])
_render = declare("tau_hypc_render", [_h, C.c_int, C.c_void_p, C.POINTER(C.c_double)])
_sync = declare("tau_hypc_sync", [_h])
_steps_done = declare("tau_hypc_steps_done", [_h], C.c_longlong)
_launches = declare("tau_hypc_launch_count", [_h], C.c_longlong)
_last_ms = declare("tau_hypc_last_step_ms", [_h, C.POINTER(C.c_float)])
_destroy = declare("tau_hypc_destroy", [_h])

VIEW_MODES = {"rho": 0, "p": 1, "speed": 2, "schlieren": 3}   # tau_hypersonic.c:44


def init_sim(W: int, H: int):
    """init_sim (:450-475) on the host: ([rho, mx, my, E] as (H, W) float64, mask (H, W) uint8)."""
    planes = [np.empty((H, W), np.float64) for _ in range(4)]
    mask = np.empty((H, W), np.uint8)
    _init_host(W, H, *planes, mask)
    return planes, mask


class HypersonicC:
    """One tau_hypersonic.c simulation on one GPU."""

    def __init__(self, W: int = 300, H: int = 300, device: int = 0, stream: int = 0):   # W, H: :12-13
        self.W, self.H = W, H
        self._handle = _h()
        check(_create(W, H, device, C.c_void_p(stream), C.byref(self._handle)))

    def init(self) -> "HypersonicC":
        check(_init(self._handle))
        return self

    def upload(self, planes, mask, sim_t: float = 0.0) -> "HypersonicC":
        arrs = [np.ascontiguousarray(p, np.float64).reshape(self.H, self.W) for p in planes]
        m = np.ascontiguousarray(mask, np.uint8).reshape(self.H, self.W)
        ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in arrs])
        check(_upload(self._handle, ptrs, C.c_void_p(m.ctypes.data), sim_t))
        return self

    def step(self, n: int = 1) -> "HypersonicC":
        """n x step_physics."""
        check(_step(self._handle, n))
        return self

    def clock(self):
        """(sim_t, dt of the last step)."""
        t, dt = C.c_double(), C.c_double()
        check(_clock(self._handle, C.byref(t), C.byref(dt)))
        return float(t.value), float(dt.value)

    def download(self):
        arrs = [np.empty((self.H, self.W), np.float64) for _ in range(4)]
        mask = np.empty((self.H, self.W), np.uint8)
        ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in arrs])
        check(_download(self._handle, ptrs, C.c_void_p(mask.ctypes.data)))
        return arrs, mask

    def render(self, view_mode=2):
        """main()'s render loops (:713-786): (rgba (H, W, 4) uint8, (min, max))."""
        mode = VIEW_MODES.get(view_mode, view_mode)
        px = np.empty((self.H, self.W), np.uint32)
        mm = (C.c_double * 2)()
        check(_render(self._handle, int(mode), C.c_void_p(px.ctypes.data), mm))
        return px.view(np.uint8).reshape(self.H, self.W, 4), (float(mm[0]), float(mm[1]))

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
            check(_destroy(self._handle))
            self._handle = _h()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
