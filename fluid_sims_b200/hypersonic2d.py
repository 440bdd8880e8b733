"""Host-side mirror of the reference `tau_2d_hypersonic_cuda` solver (tau_hypersonic_cuda.cu) over
the C-ABI.  Names follow the reference: `SimConfig` (:37-50), `default_config` (:1394-1409), the
flag validation of `parse_args` (:1545-1637) and the per-step sequence (:1833-1889), here
`Hypersonic2D.step()`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, fields

import numpy as np

from ._lib import check, declare, lib


class _CConfig(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("gamma", "cfl", "visc_nu", "visc_rho", "visc_e",
                                          "inflow_mach", "geom_x0", "geom_cy", "geom_Rb", "geom_Rn",
                                          "geom_theta")] + [("steps_per_frame", C.c_int)]


_h = C.c_void_p
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_default = declare("tau_hyp2d_default_config", [C.POINTER(_CConfig), C.c_int, C.c_int], None)
_validate = declare("tau_hyp2d_validate_config", [C.POINTER(_CConfig)])
_create = declare("tau_hyp2d_create", [C.POINTER(_CConfig), C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_void_p, C.POINTER(_h)])
_init = declare("tau_hyp2d_init", [_h])
_upload = declare("tau_hyp2d_upload", [_h, C.POINTER(C.c_void_p), C.c_void_p])
_step = declare("tau_hyp2d_step", [_h, C.c_int])
_clock = declare("tau_hyp2d_clock", [_h, C.POINTER(C.c_double), C.POINTER(C.c_double)])
_download = declare("tau_hyp2d_download", [_h, C.POINTER(C.c_void_p), C.c_void_p])
_sync = declare("tau_hyp2d_sync", [_h])
_upload_async = declare("tau_hyp2d_upload_async", [_h, C.POINTER(C.c_void_p), C.c_void_p])
_download_async = declare("tau_hyp2d_download_async", [_h, C.POINTER(C.c_void_p), C.c_void_p])
_upload_peers_async = declare("tau_hyp2d_upload_peers_async", [_h, C.POINTER(C.c_void_p)])
_devstate = declare("tau_hyp2d_device_state", [_h, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                               C.POINTER(C.c_void_p)])
_kernel_mode = declare("tau_hyp2d_kernel_mode", [_h])
_set_seg = declare("tau_hyp2d_set_seg_rows", [_h, C.c_int])
_get_seg = declare("tau_hyp2d_get_seg_rows", [_h])
_ipc_export = declare("tau_hyp2d_ipc_export", [_h, C.c_void_p, C.c_size_t])
_ipc_attach = declare("tau_hyp2d_ipc_attach", [_h, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)])
_peers_ready = declare("tau_hyp2d_peers_ready", [_h])
_ipc_detach = declare("tau_hyp2d_ipc_detach", [_h])
_peer_timing = declare("tau_hyp2d_peer_timing", [_h, C.POINTER(C.c_double)])
_render_mm = declare("tau_hyp2d_render_minmax", [_h, C.c_int, C.POINTER(C.c_double)])
_render_px = declare("tau_hyp2d_render_pixels", [_h, C.c_int, C.POINTER(C.c_double), C.c_void_p])
_render = declare("tau_hyp2d_render", [_h, C.c_int, C.c_void_p, C.POINTER(C.c_double)])


class Snapshot(C.Structure):
    """`struct RegressionSnapshot` (tau_hypersonic_cuda_tests.cu:20-36)."""
    _fields_ = [("steps", C.c_int), ("fluid_cells", C.c_int)] + [
        (k, C.c_double) for k in ("sum_rho", "sum_mx", "sum_my", "sum_E", "min_rho", "min_p", "max_mach",
                                  "checksum_rho", "checksum_mx", "checksum_E")]

    def as_tuple(self):
        return tuple(getattr(self, f[0]) for f in self._fields_)

    def write(self, path: str):
        check(_snap_write(path.encode(), C.byref(self)))

    @classmethod
    def read(cls, path: str) -> "Snapshot":
        s = cls()
        check(_snap_read(path.encode(), C.byref(s)))
        return s

    def failures_against(self, expected: "Snapshot"):
        """The reference's verification (:527-557): list of failed check names (empty = pass)."""
        n = _snap_cmp(C.byref(self), C.byref(expected))
        if n < 0:
            check(n)
        return [] if n == 0 else lib.tau_last_error().decode().split("; ")


_snapshot = declare("tau_hyp2d_snapshot", [_h, C.POINTER(Snapshot)])
_snap_write = declare("tau_hyp2d_snapshot_write", [C.c_char_p, C.POINTER(Snapshot)])
_snap_read = declare("tau_hyp2d_snapshot_read", [C.c_char_p, C.POINTER(Snapshot)])
_snap_cmp = declare("tau_hyp2d_snapshot_compare", [C.POINTER(Snapshot), C.POINTER(Snapshot)])
_ckpt_save = declare("tau_hyp2d_checkpoint_save", [_h, C.c_char_p])
_ckpt_load = declare("tau_hyp2d_checkpoint_load", [_h, C.c_char_p])
_ckpt_info = declare("tau_hyp2d_checkpoint_info", [C.c_char_p] + [C.POINTER(C.c_int)] * 5 +
                     [C.POINTER(C.c_longlong), C.POINTER(C.c_double), C.c_void_p])
_set_clock = declare("tau_hyp2d_set_clock", [_h, C.c_double, C.c_longlong])
VIEW_MODES = ("log_rho", "log_p", "speed", "log_grad_rho", "asinh_vorticity", "mach", "log_p_over_rho")
IPC_BYTES = 3 * 64
_steps_done = declare("tau_hyp2d_steps_done", [_h], C.c_longlong)
_launches = declare("tau_hyp2d_launch_count", [_h], C.c_longlong)
_last_ms = declare("tau_hyp2d_last_step_ms", [_h, C.POINTER(C.c_float)])
_destroy = declare("tau_hyp2d_destroy", [_h])

HALO = 2


@dataclass
class SimConfig:
    """`struct SimConfig` (tau_hypersonic_cuda.cu:37-50) plus the grid size, which the reference
    fixes at compile time (`#define W 8192 / H 1024`, :28-29)."""
    W: int
    H: int
    gamma: float = 1.1
    cfl: float = 0.25
    visc_nu: float = 5e-2
    visc_rho: float = 5e-2
    visc_e: float = 2e-2
    inflow_mach: float = 25.0
    geom_x0: float = 125.0
    geom_cy: float = 0.0
    geom_Rb: float = 0.0
    geom_Rn: float = 0.0
    geom_theta: float = 0.0
    steps_per_frame: int = 2

    @classmethod
    def default(cls, W: int = 8192, H: int = 1024, **over) -> "SimConfig":
        """default_config() (:1394-1409) — geometry derived from H."""
        c = _CConfig()
        _default(C.byref(c), W, H)
        kw = {f[0]: getattr(c, f[0]) for f in _CConfig._fields_}
        kw.update(over)
        return cls(W=W, H=H, **kw)

    def _c(self) -> _CConfig:
        return _CConfig(*[getattr(self, f[0]) for f in _CConfig._fields_])

    def validate(self) -> None:
        c = self._c()
        check(_validate(C.byref(c)))


_NP = {"f32": np.float32, "f64": np.float64}


class Hypersonic2D:
    """One solver handle on one GPU (`y_begin/h_local` select a slab of rows for multi-GPU)."""

    def __init__(self, cfg: SimConfig, dtype: str = "f32", device: int = 0, y_begin: int = 0,
                 h_local: int | None = None, stream: int | None = None):
        if dtype not in _NP:
            raise ValueError("dtype must be 'f32' or 'f64'")
        self.cfg = cfg
        self.dtype = dtype
        self.np_dtype = _NP[dtype]
        self.h_local = cfg.H if h_local is None else h_local
        self.y_begin = y_begin
        self.device = device
        self._handle = _h()
        cc = cfg._c()
        check(_create(C.byref(cc), cfg.W, cfg.H, 0 if dtype == "f32" else 1, device, y_begin,
                      self.h_local, C.c_void_p(stream or 0), C.byref(self._handle)))

    def init(self):
        check(_init(self._handle))
        return self

    def upload(self, planes, mask=None):
        arrs = [np.ascontiguousarray(p, self.np_dtype).reshape(self.h_local, self.cfg.W)
                for p in planes]
        ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in arrs])
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask, np.uint8).reshape(self.h_local, self.cfg.W)
        check(_upload(self._handle, ptrs, C.c_void_p(m.ctypes.data if m is not None else 0)))
        return self

    def step(self, nsteps: int = 1):
        check(_step(self._handle, nsteps))
        return self

    @property
    def kernel_name(self) -> str:
        """the step kernel(s) this handle launches per step"""
        return ("hyp2d_step", "hyp2d_step_pair+hyp2d_step", "hyp2d_step_fused")[_kernel_mode(self._handle)]

    def clock(self):
        """(sim_t, dt of the last step)."""
        t, dt = C.c_double(), C.c_double()
        check(_clock(self._handle, C.byref(t), C.byref(dt)))
        return float(t.value), float(dt.value)

    def download(self):
        """([rho, mx, my, E], mask) as (h_local, W) arrays in the handle's dtype."""
        arrs = [np.empty((self.h_local, self.cfg.W), self.np_dtype) for _ in range(4)]
        mask = np.empty((self.h_local, self.cfg.W), np.uint8)
        ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in arrs])
        check(_download(self._handle, ptrs, C.c_void_p(mask.ctypes.data)))
        return arrs, mask

    def sync(self):
        check(_sync(self._handle))

    # ---- regression snapshot + checkpoint/resume (tau_hypersonic_cuda_tests.cu:84-176) -------
    def snapshot(self) -> Snapshot:
        out = Snapshot()
        check(_snapshot(self._handle, C.byref(out)))
        return out

    def checkpoint_save(self, path: str):
        check(_ckpt_save(self._handle, path.encode()))
        return self

    def checkpoint_load(self, path: str):
        check(_ckpt_load(self._handle, path.encode()))
        return self

    @staticmethod
    def checkpoint_info(path: str):
        v = [C.c_int() for _ in range(5)]
        steps, t = C.c_longlong(), C.c_double()
        check(_ckpt_info(path.encode(), *[C.byref(x) for x in v], C.byref(steps), C.byref(t), None))
        return {"W": v[0].value, "H": v[1].value, "dtype": "f64" if v[2].value else "f32",
                "y_begin": v[3].value, "h_local": v[4].value, "steps": steps.value, "sim_t": t.value}

    # ---- render pass (reference frame loop, tau_hypersonic_cuda.cu:1892-1926) ---------------
    def render_minmax(self, view_mode: int):
        """(min, max) of the view value over this handle's fluid cells."""
        mm = (C.c_double * 2)()
        check(_render_mm(self._handle, view_mode, mm))
        return float(mm[0]), float(mm[1])

    def render_pixels(self, view_mode: int, vmin: float, vmax: float):
        """(h_local, W, 4) uint8 RGBA for the given (global) value range."""
        out = np.empty((self.h_local, self.cfg.W, 4), np.uint8)
        mm = (C.c_double * 2)(vmin, vmax)
        check(_render_px(self._handle, view_mode, mm, C.c_void_p(out.ctypes.data)))
        return out

    def render(self, view_mode: int = 0):
        """Both passes: ((h_local, W, 4) uint8 RGBA, (min, max))."""
        out = np.empty((self.h_local, self.cfg.W, 4), np.uint8)
        mm = (C.c_double * 2)()
        check(_render(self._handle, view_mode, C.c_void_p(out.ctypes.data), mm))
        return out, (float(mm[0]), float(mm[1]))

    def device_state(self):
        """(planes_ptr, mask_ptr, maxspeed_ptr) raw device addresses for the slab exchange."""
        p, m, s = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(_devstate(self._handle, C.byref(p), C.byref(m), C.byref(s)))
        return p.value, m.value, s.value

    # ---- multi-GPU peer plumbing (CUDA IPC) -------------------------------------------------
    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(IPC_BYTES)
        check(_ipc_export(self._handle, buf, IPC_BYTES))
        return buf.raw

    def ipc_attach(self, rank: int, world: int, handles, h_locals):
        blob = b"".join(handles)
        assert len(blob) == world * IPC_BYTES
        hl = (C.c_int * world)(*h_locals)
        check(_ipc_attach(self._handle, rank, world, blob, hl))
        return self

    def ipc_detach(self):
        check(_ipc_detach(self._handle))
        return self

    def peers_ready(self):
        check(_peers_ready(self._handle))
        return self

    def peer_timing(self):
        """Average microseconds per step: (waiting for peers, computing, launch gap, steps counted)."""
        out = (C.c_double * 4)()
        check(_peer_timing(self._handle, out))
        return {"wait_us": out[0], "busy_us": out[1], "gap_us": out[2], "steps": int(out[3])}

    def upload_peers_async(self, planes):
        """Multi-GPU, peers attached: enqueue the upload of the next frame's owned rows (4 pinned host arrays or
        raw addresses) with the ghost-row hand-over, the wavespeed all-reduce and the inter-frame barrier done
        on the device.  Every rank calls it for the same frame; buffers stay valid until sync()."""
        addrs = [p if isinstance(p, int) else (p.data_ptr() if hasattr(p, "data_ptr") else p.ctypes.data) for p in planes]
        check(_upload_peers_async(self._handle, (C.c_void_p * 4)(*addrs)))
        return self

    def set_seg_rows(self, rows: int):
        check(_set_seg(self._handle, rows))
        return self

    @property
    def seg_rows(self) -> int:
        return int(_get_seg(self._handle))

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


# ---- several GPUs of one box from ONE process (tau_hyp2d_group_*, include/tau_b200.h) ---------------------------
_g = C.c_void_p
_g_create = declare("tau_hyp2d_group_create", [C.POINTER(_CConfig), C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.POINTER(C.c_int), C.POINTER(_g)])
_g_size = declare("tau_hyp2d_group_size", [_g])
_g_member = declare("tau_hyp2d_group_member", [_g, C.c_int, C.POINTER(_h), C.POINTER(C.c_int), C.POINTER(C.c_int)])
_g_init = declare("tau_hyp2d_group_init", [_g])
_g_upload = declare("tau_hyp2d_group_upload", [_g, C.POINTER(C.c_void_p), C.c_void_p])
_g_step = declare("tau_hyp2d_group_step", [_g, C.c_int])
_g_sync = declare("tau_hyp2d_group_sync", [_g])
_g_clock = declare("tau_hyp2d_group_clock", [_g, C.POINTER(C.c_double), C.POINTER(C.c_double)])
_g_download = declare("tau_hyp2d_group_download", [_g, C.POINTER(C.c_void_p), C.c_void_p])
_g_render = declare("tau_hyp2d_group_render", [_g, C.c_int, C.c_void_p, C.POINTER(C.c_double)])
_g_launches = declare("tau_hyp2d_group_launch_count", [_g], C.c_longlong)
_g_destroy = declare("tau_hyp2d_group_destroy", [_g])


class Hypersonic2DGroup:
    """The 2-D solver y-slab decomposed over `ngpus` devices of one box, driven from this one process
    (no torchrun, no NCCL): same calls as Hypersonic2D, planes cover the whole grid."""

    def __init__(self, cfg: SimConfig, ngpus: int, dtype: str = "f32", devices=None):
        if dtype not in _NP:
            raise ValueError("dtype must be 'f32' or 'f64'")
        self.cfg, self.dtype, self.np_dtype, self.ngpus = cfg, dtype, _NP[dtype], ngpus
        self._handle = _g()
        cc = cfg._c()
        dv = (C.c_int * ngpus)(*devices) if devices is not None else None
        check(_g_create(C.byref(cc), cfg.W, cfg.H, 0 if dtype == "f32" else 1, ngpus, dv, C.byref(self._handle)))

    def slabs(self):
        """[(y_begin, h_local)] of the member handles."""
        out = []
        for i in range(int(_g_size(self._handle))):
            y0, hl = C.c_int(), C.c_int()
            check(_g_member(self._handle, i, None, C.byref(y0), C.byref(hl)))
            out.append((y0.value, hl.value))
        return out

    def init(self):
        check(_g_init(self._handle))
        return self

    def upload(self, planes, mask=None):
        arrs = [np.ascontiguousarray(p, self.np_dtype).reshape(self.cfg.H, self.cfg.W) for p in planes]
        ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in arrs])
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8).reshape(self.cfg.H, self.cfg.W)
        check(_g_upload(self._handle, ptrs, C.c_void_p(m.ctypes.data if m is not None else 0)))
        return self

    def step(self, nsteps: int = 1):
        check(_g_step(self._handle, nsteps))
        return self

    def sync(self):
        check(_g_sync(self._handle))

    def clock(self):
        t, dt = C.c_double(), C.c_double()
        check(_g_clock(self._handle, C.byref(t), C.byref(dt)))
        return float(t.value), float(dt.value)

    def download(self):
        arrs = [np.empty((self.cfg.H, self.cfg.W), self.np_dtype) for _ in range(4)]
        mask = np.empty((self.cfg.H, self.cfg.W), np.uint8)
        ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in arrs])
        check(_g_download(self._handle, ptrs, C.c_void_p(mask.ctypes.data)))
        return arrs, mask

    def render(self, view_mode=0):
        mode = VIEW_MODES.index(view_mode) if isinstance(view_mode, str) else int(view_mode)
        px = np.empty((self.cfg.H, self.cfg.W, 4), np.uint8)
        mm = (C.c_double * 2)()
        check(_g_render(self._handle, mode, C.c_void_p(px.ctypes.data), mm))
        return px, (float(mm[0]), float(mm[1]))

    @property
    def launch_count(self) -> int:
        return int(_g_launches(self._handle))

    def close(self):
        if self._handle:
            check(_g_destroy(self._handle))
            self._handle = _g()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
