"""TEST INFRASTRUCTURE ONLY — ctypes harness over build/hostemu/libhypersonic2d_hostemu.so (the product's
hypersonic2d.cu run by the CPU fiber emulator; see hostemu.h)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import build as hostemu_build  # noqa: E402

from fluid_sims_b200.hypersonic2d import SimConfig, _CConfig  # noqa: E402  (config struct layout only)

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(hostemu_build.build("hypersonic2d"))
        h = C.c_void_p
        L.tau_hyp2d_default_config.argtypes = [C.POINTER(_CConfig), C.c_int, C.c_int]
        L.tau_hyp2d_default_config.restype = None
        L.tau_hyp2d_create.argtypes = [C.POINTER(_CConfig)] + [C.c_int] * 6 + [C.c_void_p, C.POINTER(h)]
        L.tau_hyp2d_init.argtypes = [h]
        L.tau_hyp2d_upload.argtypes = [h, C.POINTER(C.c_void_p), C.c_void_p]
        L.tau_hyp2d_step.argtypes = [h, C.c_int]
        L.tau_hyp2d_clock.argtypes = [h, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.tau_hyp2d_download.argtypes = [h, C.POINTER(C.c_void_p), C.c_void_p]
        L.tau_hyp2d_set_seg_rows.argtypes = [h, C.c_int]
        L.tau_hyp2d_launch_count.argtypes = [h]
        L.tau_hyp2d_launch_count.restype = C.c_longlong
        L.tau_hyp2d_destroy.argtypes = [h]
        L.tau_hyp2d_work_items.argtypes = [h, C.POINTER(C.c_int)]
        L.tau_hostemu_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(f"rc={rc}: {lib().tau_hostemu_last_error().decode()}")


def default_cfg(W, H, **over):
    c = _CConfig()
    lib().tau_hyp2d_default_config(C.byref(c), W, H)
    for k, v in over.items():
        setattr(c, k, v)
    return c


def run(W, H, steps, dtype, planes=None, mask=None, seg_rows=None, pair=False, chunks=None, **over):
    """-> (planes, mask, sim_t, dts, launches) from the emulated product library"""
    L = lib()
    npdt = np.float32 if dtype == "f32" else np.float64
    cc = default_cfg(W, H, **over)
    os.environ.pop("TAU_HYP2D_PAIR", None)
    if pair:
        os.environ["TAU_HYP2D_PAIR"] = "1"
    h = C.c_void_p()
    try:
        check(L.tau_hyp2d_create(C.byref(cc), W, H, 0 if dtype == "f32" else 1, 0, 0, H, None, C.byref(h)))
    finally:
        os.environ.pop("TAU_HYP2D_PAIR", None)
    if seg_rows:
        check(L.tau_hyp2d_set_seg_rows(h, seg_rows))
    if planes is None:
        check(L.tau_hyp2d_init(h))
    else:
        arrs = [np.ascontiguousarray(p, npdt).reshape(H, W) for p in planes]
        ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in arrs])
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8).reshape(H, W)
        check(L.tau_hyp2d_upload(h, ptrs, C.c_void_p(m.ctypes.data if m is not None else 0)))
    wi = (C.c_int * 6)()
    check(L.tau_hyp2d_work_items(h, wi))
    run.last_work_items = list(wi)
    dts, t, dt = [], C.c_double(), C.c_double()
    for n in (chunks or [1] * steps):
        check(L.tau_hyp2d_step(h, n))
        check(L.tau_hyp2d_clock(h, C.byref(t), C.byref(dt)))
        dts.append(dt.value)
    out = [np.empty((H, W), npdt) for _ in range(4)]
    m = np.empty((H, W), np.uint8)
    ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in out])
    check(L.tau_hyp2d_download(h, ptrs, C.c_void_p(m.ctypes.data)))
    check(L.tau_hyp2d_clock(h, C.byref(t), C.byref(dt)))
    n = L.tau_hyp2d_launch_count(h)
    L.tau_hyp2d_destroy(h)
    return out, m, t.value, np.array(dts), n
