"""TEST INFRASTRUCTURE ONLY — ctypes harness over the product's hypersonic3d.cu run by the CPU fiber
emulator (see hostemu.h).  `packed=True` is the default build (packed WENO5 pair), `packed=False` builds it with -DT3_SCALAR_WENO."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import hostemu_build  # noqa: E402

from fluid_sims_b200.hypersonic3d import _CParams  # noqa: E402  (tau_hyp3d_params layout only)

_libs = {}


def cparams(prm, t0=1e-5, d_tau0=1e-3):
    """oracle.Hyp3dParams lacks tau_hyp3d_params' trailing (t0, d_tau0): convert instead of casting"""
    c = _CParams()
    for name, _ in prm._fields_:
        setattr(c, name, getattr(prm, name))
    c.t0, c.d_tau0 = t0, d_tau0
    return c


def lib(packed=False):
    if packed not in _libs:
        L = C.CDLL(hostemu_build.build("hypersonic3d") if packed
                   else hostemu_build.build("hypersonic3d", defines=("T3_SCALAR_WENO",), tag="_scalar"))
        h = C.c_void_p
        L.tau_hyp3d_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(h)]
        L.tau_hyp3d_init.argtypes = [h]
        L.tau_hyp3d_upload.argtypes = [h, C.POINTER(C.c_void_p), C.c_void_p]
        L.tau_hyp3d_step.argtypes = [h, C.c_int]
        L.tau_hyp3d_clock.argtypes = [h] + [C.POINTER(C.c_float)] * 4
        L.tau_hyp3d_download.argtypes = [h, C.POINTER(C.c_void_p), C.c_void_p]
        L.tau_hyp3d_destroy.argtypes = [h]
        L.tau_hyp3d_step_begin.argtypes = [h]
        L.tau_hyp3d_step_end.argtypes = [h]
        L.tau_hyp3d_device_state.argtypes = [h, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.tau_hyp3d_vis.argtypes = [h, C.c_int, np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")]
        L.tau_hyp3d_export_frame.argtypes = [h, np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS"),
                                             C.POINTER(C.c_float)]
        L.tau_hostemu_last_error.restype = C.c_char_p
        _libs[packed] = L
    return _libs[packed]


def run(prm, planes, steps, clock, packed=False):
    """prm: oracle.Hyp3dParams (same layout as tau_hyp3d_params).  -> (planes, solid, (t, d_tau, dt, maxs))"""
    L = lib(packed)

    def check(rc):
        if rc != 0:
            raise RuntimeError(f"rc={rc}: {L.tau_hostemu_last_error().decode()}")
    h = C.c_void_p()
    cp = cparams(prm)
    check(L.tau_hyp3d_create(C.byref(cp), 0, 0, prm.nz, None, C.byref(h)))
    shape = (prm.nz, prm.ny, prm.nx)
    if planes is None:
        check(L.tau_hyp3d_init(h))
    else:
        arrs = [np.ascontiguousarray(p, np.float32).reshape(shape) for p in planes]
        ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in arrs])
        ck = np.array(clock, np.float32)
        check(L.tau_hyp3d_upload(h, ptrs, C.c_void_p(ck.ctypes.data)))
    check(L.tau_hyp3d_step(h, steps))
    out = [np.empty(shape, np.float32) for _ in range(6)]
    solid = np.empty(shape, np.uint8)
    ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in out])
    check(L.tau_hyp3d_download(h, ptrs, C.c_void_p(solid.ctypes.data)))
    v = [C.c_float() for _ in range(4)]
    check(L.tau_hyp3d_clock(h, *[C.byref(x) for x in v]))
    L.tau_hyp3d_destroy(h)
    return out, solid, tuple(float(x.value) for x in v)


def vis_and_export(prm, planes, clock=(0.012, 2e-3), modes=(0, 8)):
    """-> ({mode: vis field}, palette indices, (min, max)) of the uploaded state"""
    L = lib(False)
    h = C.c_void_p()
    cp = cparams(prm)
    assert L.tau_hyp3d_create(C.byref(cp), 0, 0, prm.nz, None, C.byref(h)) == 0
    shape = (prm.nz, prm.ny, prm.nx)
    arrs = [np.ascontiguousarray(p, np.float32).reshape(shape) for p in planes]
    ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in arrs])
    ck = np.array(clock, np.float32)
    assert L.tau_hyp3d_upload(h, ptrs, C.c_void_p(ck.ctypes.data)) == 0
    vis = {}
    for m in modes:
        out = np.empty(shape, np.float32)
        assert L.tau_hyp3d_vis(h, m, out.ravel()) == 0, L.tau_hostemu_last_error()
        vis[m] = out
    idx = np.empty(shape, np.uint8)
    mm = (C.c_float * 2)()
    assert L.tau_hyp3d_export_frame(h, idx.ravel(), mm) == 0, L.tau_hostemu_last_error()
    L.tau_hyp3d_destroy(h)
    return vis, idx, (float(mm[0]), float(mm[1]))


def export_video(prm, frames, steps_per_frame=4):
    """th3cs.cu's main loop through the product: init, then per frame `steps_per_frame` steps + export_frame.
    -> (frames, nz, ny, nx) uint8"""
    L = lib(False)
    h = C.c_void_p()
    cp = cparams(prm)
    assert L.tau_hyp3d_create(C.byref(cp), 0, 0, prm.nz, None, C.byref(h)) == 0
    assert L.tau_hyp3d_init(h) == 0
    out = np.empty((frames, prm.nz, prm.ny, prm.nx), np.uint8)
    for f in range(frames):
        assert L.tau_hyp3d_step(h, steps_per_frame) == 0
        assert L.tau_hyp3d_export_frame(h, out[f].ravel(), None) == 0
    L.tau_hyp3d_destroy(h)
    return out


def run_slabs(prm, planes, steps, clock, world, packed=False):
    """The z-slab ring of tests/_mgpu_worker.py in one process: `world` slab handles, ghost planes copied between their
    state buffers through tau_hyp3d_device_state (in the emulator a device pointer is a host pointer), the max wavespeed
    reduced by hand between step_begin and step_end.  -> (planes of the whole domain, [clock per handle])"""
    L = lib(packed)
    H = 3
    nz, ny, nx = prm.nz, prm.ny, prm.nx
    cp = cparams(prm)
    cuts = [(r * nz // world, (r + 1) * nz // world - r * nz // world) for r in range(world)]
    full = [np.ascontiguousarray(p, np.float32).reshape(nz, ny, nx) for p in planes]
    hs = []
    for z0, nl in cuts:
        h = C.c_void_p()
        assert L.tau_hyp3d_create(C.byref(cp), 0, z0, nl, None, C.byref(h)) == 0, L.tau_hostemu_last_error()
        arrs = [np.ascontiguousarray(f[z0:z0 + nl]) for f in full]
        ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in arrs])
        ck = np.array(clock, np.float32)
        assert L.tau_hyp3d_upload(h, ptrs, C.c_void_p(ck.ctypes.data)) == 0
        hs.append(h)

    def state(h, nl):
        pp, mp = C.c_void_p(), C.c_void_p()
        assert L.tau_hyp3d_device_state(h, C.byref(pp), C.byref(mp)) == 0
        st = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_float)), shape=(6, nl + 2 * H, ny, nx))
        return st, np.ctypeslib.as_array(C.cast(mp, C.POINTER(C.c_float)), shape=(1,))
    for _ in range(steps):
        sts = [state(h, nl) for h, (_, nl) in zip(hs, cuts)]
        for r in range(world):   # z is periodic: a ring
            lo, hi = sts[(r - 1) % world][0], sts[(r + 1) % world][0]
            me, nl = sts[r][0], cuts[r][1]
            me[:, :H] = lo[:, cuts[(r - 1) % world][1]:cuts[(r - 1) % world][1] + H]   # the lower neighbour's last H planes
            me[:, nl + H:] = hi[:, H:2 * H]                                              # the upper neighbour's first H
        for h in hs:
            assert L.tau_hyp3d_step_begin(h) == 0, L.tau_hostemu_last_error()
        m = max(float(s[1][0]) for s in sts)
        for s_ in sts:
            s_[1][0] = m
        for h in hs:
            assert L.tau_hyp3d_step_end(h) == 0
    out = [np.empty((nz, ny, nx), np.float32) for _ in range(6)]
    clocks = []
    for h, (z0, nl) in zip(hs, cuts):
        part = [np.empty((nl, ny, nx), np.float32) for _ in range(6)]
        ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in part])
        assert L.tau_hyp3d_download(h, ptrs, None) == 0
        for f in range(6):
            out[f][z0:z0 + nl] = part[f]
        v = [C.c_float() for _ in range(4)]
        assert L.tau_hyp3d_clock(h, *[C.byref(x) for x in v]) == 0
        clocks.append(tuple(float(x.value) for x in v))
        L.tau_hyp3d_destroy(h)
    return out, clocks


def run_group(prm, planes, steps, clock, ngpus, packed=True, chunks=(None,)):
    """tau_hyp3d_group_*: ONE process, one z-slab handle per (pretend) device, ghost planes by cudaMemcpyPeerAsync, host max.
    `planes` None: group_init.  `chunks`: the steps are issued in these pieces (None = all at once).
    -> (planes of the whole grid, solid, (t, d_tau, dt, maxs))"""
    L = lib(packed)
    gp = C.c_void_p
    L.tau_hyp3d_group_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(gp)]
    L.tau_hyp3d_group_init.argtypes = [gp]
    L.tau_hyp3d_group_upload.argtypes = [gp, C.POINTER(C.c_void_p), C.c_void_p]
    L.tau_hyp3d_group_step.argtypes = [gp, C.c_int]
    L.tau_hyp3d_group_clock.argtypes = [gp] + [C.POINTER(C.c_float)] * 4
    L.tau_hyp3d_group_download.argtypes = [gp, C.POINTER(C.c_void_p), C.c_void_p]
    L.tau_hyp3d_group_destroy.argtypes = [gp]
    L.tau_hyp3d_group_size.argtypes = [gp]
    g = gp()
    cp = cparams(prm)
    assert L.tau_hyp3d_group_create(C.byref(cp), ngpus, None, C.byref(g)) == 0, L.tau_hostemu_last_error()
    assert L.tau_hyp3d_group_size(g) == ngpus
    shape = (prm.nz, prm.ny, prm.nx)
    if planes is None:
        assert L.tau_hyp3d_group_init(g) == 0
    else:
        arrs = [np.ascontiguousarray(p, np.float32).reshape(shape) for p in planes]
        ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in arrs])
        ck = np.array(clock, np.float32)
        assert L.tau_hyp3d_group_upload(g, ptrs, C.c_void_p(ck.ctypes.data)) == 0, L.tau_hostemu_last_error()
    left = steps
    for c in chunks:
        k = left if c is None else min(c, left)
        assert L.tau_hyp3d_group_step(g, k) == 0, L.tau_hostemu_last_error()
        left -= k
    assert left == 0
    out = [np.empty(shape, np.float32) for _ in range(6)]
    solid = np.empty(shape, np.uint8)
    ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in out])
    assert L.tau_hyp3d_group_download(g, ptrs, C.c_void_p(solid.ctypes.data)) == 0
    v = [C.c_float() for _ in range(4)]
    assert L.tau_hyp3d_group_clock(g, *[C.byref(x) for x in v]) == 0
    L.tau_hyp3d_group_destroy(g)
    return out, solid, tuple(float(x.value) for x in v)
