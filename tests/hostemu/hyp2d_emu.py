"""TEST INFRASTRUCTURE ONLY — ctypes harness over build/hostemu/libhypersonic2d_hostemu.so (the product's
hypersonic2d.cu run by the CPU fiber emulator; see hostemu.h)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import hostemu_build  # noqa: E402

from fluid_sims_b200.hypersonic2d import SimConfig, _CConfig  # noqa: E402  (config struct layout only)

_lib = None


def lib():
    global _lib
    if _lib is None:
        san = os.environ.get("TAU_HC_SANITIZE") == "1"   # UBSan alignment + bounds build (run it in a subprocess)
        L = C.CDLL(hostemu_build.build("hypersonic2d", tag="_san" if san else "", sanitize=san))
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


def run(W, H, steps, dtype, planes=None, mask=None, seg_rows=None, pair=False, chunks=None, reinit_after=0, **over):
    """-> (planes, mask, sim_t, dts, launches) from the emulated product library"""
    L = lib()
    npdt = np.float32 if dtype == "f32" else np.float64
    cc = default_cfg(W, H, **over)
    os.environ.pop("TAU_HYP2D_PAIR", None)
    if pair:
        os.environ["TAU_HYP2D_PAIR"] = str(int(pair))      # True / 1: pair + production kernels; 2: the fused kernel
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
    if reinit_after:      # a few steps, then tau_hyp2d_init again: every rotating counter must start over
        check(L.tau_hyp2d_step(h, reinit_after))
        check(L.tau_hyp2d_init(h))
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


def run_slabs(W, H, steps, dtype, world, pair=False, frames=(), reverse_ranks=False, cuts=None, **over):
    """`world` y-slab handles in ONE process, stepped in turn: the emulated counterpart of one process per
    GPU with CUDA-IPC peers (fluid_sims_b200/slab.py: hyp2d_sync_state + hyp2d_attach_peers).  Each rank's
    step kernel pushes its boundary rows into the neighbours' ghost rows and sends its max wavespeed to
    every peer's inbox; the next step's kernel polls the inbox.  -> (planes, mask, sim_t per rank, open IPC
    mappings after destroy)"""
    L = lib()
    L.tau_hyp2d_device_state.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 3
    L.tau_hyp2d_ipc_export.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.tau_hyp2d_ipc_attach.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    L.tau_hyp2d_peers_ready.argtypes = [C.c_void_p]
    L.tau_hostemu_ipc_open_mappings.restype = C.c_longlong
    npdt = np.float32 if dtype == "f32" else np.float64
    cc = default_cfg(W, H, **over)
    base, rem = divmod(H, world)
    parts, y = [], 0
    for r in range(world):
        hl = cuts[r] if cuts is not None else base + (r < rem)   # cuts: rows per rank (unequal slabs of a balanced run)
        parts.append((y, hl))
        y += hl
    assert y == H
    os.environ.pop("TAU_HYP2D_PAIR", None)
    if pair:
        os.environ["TAU_HYP2D_PAIR"] = str(int(pair))
    hs = []
    try:
        for y0, hl in parts:
            h = C.c_void_p()
            check(L.tau_hyp2d_create(C.byref(cc), W, H, 0 if dtype == "f32" else 1, 0, y0, hl, None, C.byref(h)))
            check(L.tau_hyp2d_init(h))
            hs.append(h)
    finally:
        os.environ.pop("TAU_HYP2D_PAIR", None)
    # hyp2d_sync_state: ghost rows of the mask and of the current planes, max of the wavespeed slot
    views = []
    for h, (y0, hl) in zip(hs, parts):
        pp, mp, sp = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(L.tau_hyp2d_device_state(h, C.byref(pp), C.byref(mp), C.byref(sp)))
        planes = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_float if dtype == "f32" else C.c_double)),
                                       shape=(4, hl + 4, W))
        mask = np.ctypeslib.as_array(C.cast(mp, C.POINTER(C.c_uint8)), shape=(hl + 4, W))
        speed = np.ctypeslib.as_array(C.cast(sp, C.POINTER(C.c_double)), shape=(1,))
        views.append((planes, mask, speed))
    for r in range(world - 1):
        (pu, mu, _), (pd, md, _) = views[r], views[r + 1]
        hu = parts[r][1]
        pd[:, 0:2] = pu[:, hu:hu + 2]
        md[0:2] = mu[hu:hu + 2]
        pu[:, hu + 2:hu + 4] = pd[:, 2:4]
        mu[hu + 2:hu + 4] = md[2:4]
    smax = max(float(v[2][0]) for v in views)
    for v in views:
        v[2][0] = smax
    # hyp2d_attach_peers
    blob = b""
    for h in hs:
        buf = C.create_string_buffer(192)
        check(L.tau_hyp2d_ipc_export(h, buf, 192))
        blob += buf.raw
    hl_arr = (C.c_int * world)(*[p[1] for p in parts])
    for r, h in enumerate(hs):
        check(L.tau_hyp2d_ipc_attach(h, r, world, blob, hl_arr))
        check(L.tau_hyp2d_peers_ready(h))
    for _ in range(steps):
        for h in hs:
            check(L.tau_hyp2d_step(h, 1))
    # further frames: (full-domain planes, steps) — handed over on the "device" (tau_hyp2d_upload_peers_async):
    # every rank enqueues its upload first (a rank's next step waits for all peers' announcements)
    L.tau_hyp2d_upload_peers_async.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    order = list(zip(hs, parts))
    if reverse_ranks:       # the ranks are independent processes: no call order between them may matter
        order.reverse()
        hs_step = hs[::-1]
    else:
        hs_step = hs
    for fplanes, fsteps in frames:
        keep = []
        for h, (y0, hl) in order:
            arrs = [np.ascontiguousarray(np.asarray(p, npdt).reshape(H, W)[y0:y0 + hl]) for p in fplanes]
            keep.append(arrs)
            check(L.tau_hyp2d_upload_peers_async(h, (C.c_void_p * 4)(*[a.ctypes.data for a in arrs])))
        for _ in range(fsteps):
            for h in hs_step:
                check(L.tau_hyp2d_step(h, 1))
    outs, masks, ts = [], [], []
    for h, (y0, hl) in zip(hs, parts):
        out = [np.empty((hl, W), npdt) for _ in range(4)]
        m = np.empty((hl, W), np.uint8)
        ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in out])
        check(L.tau_hyp2d_download(h, ptrs, C.c_void_p(m.ctypes.data)))
        t, dt = C.c_double(), C.c_double()
        check(L.tau_hyp2d_clock(h, C.byref(t), C.byref(dt)))
        outs.append(out)
        masks.append(m)
        ts.append(t.value)
    for h in hs:
        L.tau_hyp2d_destroy(h)
    planes = [np.concatenate([o[k] for o in outs], axis=0) for k in range(4)]
    return planes, np.concatenate(masks, axis=0), ts, int(L.tau_hostemu_ipc_open_mappings())
