"""The PRODUCT's Burgers and shallow-water translation units (fluid_sims_b200/csrc/{burgers,shallow_water}.cu
— kernels AND their host-side step logic) executed on the CPU by the fiber emulator in tests/hostemu/, and
compared with the CPU oracle.  This checks code, it is not a code path: the emulated library is built under
build/hostemu/ by this test and loaded by nothing else (tests/hostemu/hostemu.h explains the model).

Because the emulated kernels use the host's libm with -ffp-contract=off — exactly what oracle/ uses — a
kernel that keeps the reference's expression trees must match the oracle BIT FOR BIT; tile/halo indexing,
face bookkeeping, buffer rotation, the device-side dt / clock slots and barrier placement (the emulator
aborts on barrier divergence, and fresh "device" memory is filled with garbage) are all exercised.
What it cannot show: anything about -use_fast_math intrinsics, memory ordering between blocks, or speed.
burgers.cu has been validated on a B200 (tests/test_burgers_gpu.py); running it here as well validates
the emulator.  shallow_water.cu has not run on hardware yet — this is its strongest check so far."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import oracle

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostemu"))
import build as hostemu_build  # noqa: E402

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def swlib():
    lib = C.CDLL(hostemu_build.build("shallow_water"))
    P = C.POINTER(oracle.SwParams)     # same field order as tau_sw_params (checked below)
    lib.tau_sw_create.argtypes = [P, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.tau_sw_upload.argtypes = [C.c_void_p, f32p, f32p, f32p, C.c_void_p]
    lib.tau_sw_init.argtypes = [C.c_void_p]
    lib.tau_sw_step.argtypes = [C.c_void_p, C.c_int]
    lib.tau_sw_clock.argtypes = [C.c_void_p] + [C.POINTER(C.c_float)] * 3
    lib.tau_sw_download.argtypes = [C.c_void_p, f32p, f32p, f32p]
    lib.tau_sw_destroy.argtypes = [C.c_void_p]
    lib.tau_sw_launch_count.argtypes = [C.c_void_p]
    lib.tau_sw_launch_count.restype = C.c_longlong
    lib.tau_hostemu_last_error.restype = C.c_char_p
    return lib


def sw_emulated(lib, prm, s0, u0, v0, steps, chunks=1, use_init=False):
    h = C.c_void_p()
    assert lib.tau_sw_create(C.byref(prm), 0, None, C.byref(h)) == 0, lib.tau_hostemu_last_error()
    if use_init:
        assert lib.tau_sw_init(h) == 0
    else:
        a = [np.ascontiguousarray(x, np.float32).ravel() for x in (s0, u0, v0)]
        assert lib.tau_sw_upload(h, *a, None) == 0
    for _ in range(chunks):
        assert lib.tau_sw_step(h, steps // chunks) == 0
    out = [np.empty(prm.nx * prm.ny, np.float32) for _ in range(3)]
    assert lib.tau_sw_download(h, *out) == 0
    t, tau, dt = C.c_float(), C.c_float(), C.c_float()
    assert lib.tau_sw_clock(h, C.byref(t), C.byref(tau), C.byref(dt)) == 0
    n = lib.tau_sw_launch_count(h)
    lib.tau_sw_destroy(h)
    return [o.reshape(prm.shape) for o in out], (t.value, tau.value, dt.value), n


GENTLE = dict(H0=2.0, bumpAmp=0.4, bumpSigma=5, asym=0.3, swirl=0.05, swirlRc=10, offx=3, offy=-2)


def test_param_struct_layouts_agree():
    from fluid_sims_b200.shallow_water import _CParams
    assert [f[0] for f in _CParams._fields_] == [f[0] for f in oracle.SwParams._fields_]
    assert C.sizeof(_CParams) == C.sizeof(oracle.SwParams) == 19 * 4


@pytest.mark.parametrize("kw,steps", [
    (dict(nx=96, ny=64, dtau=0.02, nu=0.0, **GENTLE), 30),                 # whole 32x16 tiles, 1 kernel per step
    (dict(nx=96, ny=64, dtau=0.02, nu=0.05, **GENTLE), 30),                # + Jacobi viscosity, sigma pointer swap
    (dict(nx=70, ny=37, dtau=0.05, nu=0.02, dx=2.0, dy=1.5, **GENTLE), 25),    # ragged tiles in x and y
    (dict(nx=33, ny=5, dtau=0.02, nu=0.0, **GENTLE), 20),                  # one cell past a tile; ny < tile
    (dict(nx=7, ny=3, dtau=0.02, nu=0.01, **GENTLE), 12),                  # grid smaller than the halo'd tile
    (dict(nx=64, ny=48, dtau=1e-3, nu=0.0), 12),                           # default (violent, H0 = 1000) field
    (dict(nx=64, ny=48, dtau=1.0, nu=0.001, offx=5, offy=5), 30),          # reference defaults: t overflows the CFL cap
])
def test_shallow_water_product_code_equals_oracle_bit_for_bit(swlib, kw, steps):
    prm = oracle.sw_params(**kw)
    s0, u0, v0 = oracle.sw_init(prm)
    (s, u, v), ck, n = sw_emulated(swlib, prm, s0, u0, v0, steps)
    es, eu, ev, eck, dts = oracle.sw_run(prm, s0, u0, v0, steps)
    assert np.array_equal(s, es) and np.array_equal(u, eu) and np.array_equal(v, ev)
    assert ck[0] == eck[0] and ck[1] == eck[1] and ck[2] == dts[-1]
    assert n == 1 + steps * (2 if prm.nu > 0 else 1)
    assert np.abs(s - s0).max() > 1e-4


def test_shallow_water_step_chunking_init_and_reupload(swlib):
    prm = oracle.sw_params(nx=70, ny=37, dtau=0.05, nu=0.02, **GENTLE)
    s0, u0, v0 = oracle.sw_init(prm)
    a, cka, _ = sw_emulated(swlib, prm, s0, u0, v0, 24)
    b, ckb, _ = sw_emulated(swlib, prm, s0, u0, v0, 24, chunks=24)     # odd/even step parity of every slot
    c, ckc, _ = sw_emulated(swlib, prm, None, None, None, 24, use_init=True)
    assert all(np.array_equal(x, y) and np.array_equal(x, z) for x, y, z in zip(a, b, c)) and cka == ckb == ckc
    # a state uploaded in the middle of a run (steps_done odd) picks the right clock / wavespeed slots
    h = C.c_void_p()
    assert swlib.tau_sw_create(C.byref(prm), 0, None, C.byref(h)) == 0
    assert swlib.tau_sw_init(h) == 0 and swlib.tau_sw_step(h, 7) == 0
    ck = np.array([1.5, 0.25], np.float32)
    assert swlib.tau_sw_upload(h, s0.ravel(), u0.ravel(), v0.ravel(), ck.ctypes.data) == 0
    assert swlib.tau_sw_step(h, 9) == 0
    out = [np.empty(prm.nx * prm.ny, np.float32) for _ in range(3)]
    swlib.tau_sw_download(h, *out)
    swlib.tau_sw_destroy(h)
    es, eu, ev, _, _ = oracle.sw_run(prm, s0, u0, v0, 9, clock=(1.5, 0.25))
    assert np.array_equal(out[0].reshape(prm.shape), es) and np.array_equal(out[1].reshape(prm.shape), eu)
    # errors are loud
    h = C.c_void_p()
    assert swlib.tau_sw_create(C.byref(prm), 0, None, C.byref(h)) == 0
    assert swlib.tau_sw_step(h, 1) == -22 and b"no state" in swlib.tau_hostemu_last_error()
    swlib.tau_sw_destroy(h)
    bad = oracle.sw_params(nx=0)
    assert swlib.tau_sw_create(C.byref(bad), 0, None, C.byref(h)) == -22


# ---- Burgers: GPU-validated code through the same emulator (validates the emulator) ------------------------
@pytest.fixture(scope="module")
def bglib():
    lib = C.CDLL(hostemu_build.build("burgers"))
    P = C.POINTER(oracle.BurgersParams)
    lib.tau_burgers_create.argtypes = [P, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.tau_burgers_upload.argtypes = [C.c_void_p, f32p, f32p, C.c_void_p]
    lib.tau_burgers_step.argtypes = [C.c_void_p, C.c_int]
    lib.tau_burgers_clock.argtypes = [C.c_void_p] + [C.POINTER(C.c_float)] * 3
    lib.tau_burgers_download.argtypes = [C.c_void_p, f32p, f32p]
    lib.tau_burgers_destroy.argtypes = [C.c_void_p]
    lib.tau_hostemu_last_error.restype = C.c_char_p
    return lib


@pytest.mark.parametrize("kw,steps", [
    (dict(nx=96, ny=64, dtau=1e-3, swirl=0.2, amp=0.3), 20),
    (dict(nx=70, ny=37, dtau=2e-3, visc_substeps=3, swirl=0.2, amp=0.3, muscl=1), 15),
    (dict(nx=33, ny=5, dtau=1e-3, muscl=1, rc=4.0, bsig=3.0), 10),
    (dict(nx=300, colehopf=1, dtau=5e-3, t0=1e-3, nu=0.5, ck=2), 100),
])
def test_burgers_product_code_equals_oracle(bglib, kw, steps):
    prm = oracle.burgers_params(**kw)
    u0, v0 = oracle.burgers_init(prm)
    h = C.c_void_p()
    assert bglib.tau_burgers_create(C.byref(prm), 0, None, C.byref(h)) == 0, bglib.tau_hostemu_last_error()
    assert bglib.tau_burgers_upload(h, u0.ravel(), v0.ravel(), None) == 0
    assert bglib.tau_burgers_step(h, steps) == 0
    n = u0.size
    u, v = np.empty(n, np.float32), np.empty(n, np.float32)
    assert bglib.tau_burgers_download(h, u, v) == 0
    t, tau, dt = C.c_float(), C.c_float(), C.c_float()
    bglib.tau_burgers_clock(h, C.byref(t), C.byref(tau), C.byref(dt))
    bglib.tau_burgers_destroy(h)
    eu, ev, eck, dts = oracle.burgers_run(prm, u0, v0, steps)
    err = max(float(np.abs(u.reshape(eu.shape) - eu).max()), float(np.abs(v.reshape(ev.shape) - ev).max()))
    print(f"\nburgers emulated vs oracle {kw}: {err:.3e}")
    # burgers.cu evaluates u = u0 sinh(phi) once per tile cell and reuses it, like the oracle: identical
    assert err == 0.0
    assert dt.value == dts[-1] and abs(t.value - eck[0]) <= 1e-6 * eck[0]
