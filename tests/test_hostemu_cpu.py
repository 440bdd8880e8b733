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
import hostemu_build  # noqa: E402

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
    (dict(nx=64, ny=48, dtau=1.0, nu=0.001, offx=5, offy=5), 100),         # reference defaults: t = e^step reaches inf at step 89
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


# ---- 2-D hypersonic solver: the HEADLINE kernel and its experimental two-columns-per-lane variant -----------
# hypersonic2d.cu runs whole in the emulator: tensor maps + TMA box loads (zero fill), the mbarrier ring,
# the persistent grid with its device work queue (the pretend device has TAU_HC_SMS x TAU_HC_CTAS_PER_SM
# resident CTAs), warp-shuffle column marching, the last-CTA-out reduction and the device-side clock.
# Its 11 inline-PTX statements are replaced by hostemu_build.py (rcp/sqrt.approx -> exact, %tid -> threadIdx, ...).
import hyp2d_emu  # noqa: E402  (tests/hostemu)

NAMES = ("rho", "mx", "my", "E")


def rel_linf(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


def _random_state_with_walls(W, H, holes=0.0005):
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:H, 0:W]
    rho = 1.0 + 0.3 * np.sin(xx / 9.0) * np.cos(yy / 7.0)
    u = 3.0 + 0.5 * np.cos(xx / 11.0)
    v = 0.7 * np.sin(yy / 5.0)
    p = 1.0 + 0.2 * np.cos((xx + yy) / 13.0)
    g = 1.1
    planes = [rho, rho * u, rho * v, p / (g - 1) + 0.5 * rho * (u * u + v * v)]
    mask = np.zeros((H, W), np.uint8)
    mask[40:56, 50:70] = 1
    mask[0:3, 100:110] = 1          # touches y = 0
    mask[H - 2:, 20:30] = 1         # touches y = H-1
    mask[60:64, 0:2] = 1            # touches x = 0
    mask[10:14, W - 3:] = 1         # touches x = W-1
    mask[rng.random((H, W)) < holes] = 1
    return planes, mask


@pytest.fixture
def pretend_device(monkeypatch):
    def set_(sms, ctas):
        monkeypatch.setenv("TAU_HC_SMS", str(sms))
        monkeypatch.setenv("TAU_HC_CTAS_PER_SM", str(ctas))
    return set_


@pytest.mark.parametrize("W,H,steps,seg,sms", [(64, 33, 12, 8, 3), (36, 7, 8, 4, 1), (124, 64, 10, 64, 2),
                                               (203, 57, 10, None, 3)])   # 203: W*8 % 16 != 0, the non-TMA loader
def test_hyp2d_production_kernel_f64_equals_oracle(pretend_device, W, H, steps, seg, sms):
    pretend_device(sms, 2)
    x0 = min(125.0, W / 3.0)
    cfg = oracle.hyp2d_cfg(W, H, geom_x0=x0)
    planes, mask = oracle.hyp2d_init(cfg)
    ref, t_ref, dts_ref = oracle.hyp2d_run(cfg, planes, mask, steps)
    out, m, t, dts, _ = hyp2d_emu.run(W, H, steps, "f64", seg_rows=seg, geom_x0=x0)
    assert np.array_equal(m.ravel(), mask)
    for k, a, b in zip(NAMES, out, ref):
        assert rel_linf(a, b) < 1e-12, k          # measured 1e-15 (contraction differs, nothing else)
    assert abs(t - t_ref) < 1e-13 and np.abs(dts - dts_ref).max() < 1e-15


def test_hyp2d_production_kernel_uploaded_state_with_walls(pretend_device):
    pretend_device(3, 2)
    W, H, steps = 128, 96, 10
    planes, mask = _random_state_with_walls(W, H, holes=0.002)
    ref, t_ref, _ = oracle.hyp2d_run(oracle.hyp2d_cfg(W, H), planes, mask.ravel(), steps)
    out, _, t, _, _ = hyp2d_emu.run(W, H, steps, "f64", planes=planes, mask=mask)
    assert max(rel_linf(a, b) for a, b in zip(out, ref)) < 1e-12 and abs(t - t_ref) < 1e-13
    out, _, t, _, _ = hyp2d_emu.run(W, H, steps, "f32", planes=planes, mask=mask, chunks=[3, 7])
    assert max(rel_linf(a, b) for a, b in zip(out, ref)) < 5e-6     # measured 1e-6


@pytest.mark.parametrize("W,H,steps,seg,sms,ctas", [
    (308, 96, 25, None, 3, 2),     # guided schedule; 30 pair items + 72 production items
    (308, 96, 12, 24, 1, 1),       # one resident CTA: every item but the first four comes off the device queue
    (1000, 40, 8, 8, 16, 2),       # more CTAs than items in the tail: late claims find the queue empty
    (136, 20, 10, None, 3, 2),     # the narrowest grid pair mode accepts
])
def test_hyp2d_pair_kernel_equals_production_kernel(pretend_device, W, H, steps, seg, sms, ctas):
    """hypersonic2d_pair.cuh (TAU_HYP2D_PAIR=1: two columns per lane, packed fp32x2 arithmetic) has not run
    on hardware yet.  Here it computes the interior body-free items of a smooth random field with walls, the
    production kernel the rest, and the result is compared with the production kernel alone (fp32 both:
    they differ by FMA contraction only) and with the fp64 oracle."""
    pretend_device(sms, ctas)
    planes, mask = _random_state_with_walls(W, H) if H >= 64 else (None, None)
    kw = {} if planes is not None else dict(geom_x0=min(125.0, W / 3.0))
    if planes is not None:
        ref, t_ref, _ = oracle.hyp2d_run(oracle.hyp2d_cfg(W, H), planes, mask.ravel(), steps)
    else:
        cfg = oracle.hyp2d_cfg(W, H, **kw)
        p0, m0 = oracle.hyp2d_init(cfg)
        ref, t_ref, _ = oracle.hyp2d_run(cfg, p0, m0, steps)
    a, _, ta, dtsa, na = hyp2d_emu.run(W, H, steps, "f32", planes=planes, mask=mask, seg_rows=seg, **kw)
    b, _, tb, dtsb, nb = hyp2d_emu.run(W, H, steps, "f32", planes=planes, mask=mask, seg_rows=seg, pair=True, **kw)
    items = hyp2d_emu.run.last_work_items
    assert items[2] > 0 and items[3] > 0 and nb == na + steps      # the pair kernel did launch, once per step
    assert max(rel_linf(x, y) for x, y in zip(b, a)) < 2e-6         # measured <= 2.3e-7
    assert max(rel_linf(x, y) for x, y in zip(b, ref)) < 5e-6       # measured 1.2e-6, same as the production kernel
    assert np.abs(dtsa - dtsb).max() <= 1e-9 * dtsa.max() and abs(ta - tb) <= 1e-9 * ta


def test_hyp2d_pair_mode_declines_what_it_cannot_do(pretend_device):
    pretend_device(3, 2)
    for W, H, dtype in ((137, 21, "f32"), (200, 40, "f64"), (120, 40, "f32")):   # no TMA / fp64 / too narrow
        a, *_ = hyp2d_emu.run(W, H, 4, dtype, geom_x0=W / 3.0)
        b, *_ = hyp2d_emu.run(W, H, 4, dtype, pair=True, geom_x0=W / 3.0)
        assert hyp2d_emu.run.last_work_items[2] == 0
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("pair", [False, 1, 2])
@pytest.mark.parametrize("after", [1, 2, 4])
def test_hyp2d_reinit_after_a_few_steps_starts_every_rotating_counter_over(pretend_device, pair, after):
    """tau_hyp2d_init cleared Ctrl and the step count but not the pair kernel's separately allocated claim
    counters; re-initialising after 1, 4, 7 ... steps then made the first step claim from a stale counter and skip items.
    A run that is re-initialised after `after` steps must equal a fresh run bit for bit (production, pair and fused modes)."""
    pretend_device(3, 2)
    W, H, steps = 308, 96, 5
    a, _, ta, dtsa, _ = hyp2d_emu.run(W, H, steps, "f32", pair=pair, geom_x0=W / 3.0)
    b, _, tb, dtsb, _ = hyp2d_emu.run(W, H, steps, "f32", pair=pair, reinit_after=after, geom_x0=W / 3.0)
    assert ta == tb and np.array_equal(dtsa, dtsb) and all(np.array_equal(x, y) for x, y in zip(a, b))


# ---- 3-D hypersonic solver, default build (packed WENO5 pair) and the scalar form (-DT3_SCALAR_WENO) ----------
import hyp3d_emu  # noqa: E402  (tests/hostemu)


@pytest.fixture(scope="module")
def developed_3d_flow():
    prm = oracle.hyp3d_params(32, 28, 20)
    planes, solid = oracle.hyp3d_init(prm)
    dev, _, _, _ = oracle.hyp3d_run(prm, planes, solid, 150, (5e-3, 2e-3))   # bow shock formed, |phi_y,z| ~ 1
    assert min(float(np.abs(a).max()) for a in dev) > 0.5
    return prm, dev, solid


@pytest.mark.parametrize("steps,tol", [(1, dict(f=2e-6, lam=3e-3, zet=5e-4)), (10, dict(f=2e-5, lam=6e-3, zet=5e-3))])
def test_hyp3d_default_and_packed_weno_builds(developed_3d_flow, steps, tol):
    """lam / zet (log-type thermodynamic variables) amplify fp32 rounding differences to ~1e-3 within one
    step at a few cells of the shock layer — the same size on the GPU (tests/test_hyp3d_gpu.py TOL_DEV) —
    while xi / phi stay at the 1e-7 level; an indexing or pairing mistake would show there at O(0.1).
    The packed build (the default since its first hardware run in round 2) must sit inside the same envelope, against the oracle AND
    against the default build."""
    prm, dev, solid = developed_3d_flow
    clock = (0.012, 2e-3)
    ref, ck_ref, _, _ = oracle.hyp3d_run(prm, dev, solid, steps, clock)
    d, sol, ckd = hyp3d_emu.run(prm, dev, steps, clock)
    p, _, ckp = hyp3d_emu.run(prm, dev, steps, clock, packed=True)
    assert np.array_equal(sol.ravel(), solid)

    def inside(x, y):
        e = [float(np.abs(np.asarray(a).ravel() - np.asarray(b).ravel()).max()) for a, b in zip(x, y)]
        assert max(e[:4]) < tol["f"] and e[4] < tol["lam"] and e[5] < tol["zet"], e
    inside(d, ref)
    inside(p, ref)
    inside(p, d)
    for ck in (ckd, ckp):
        assert abs(ck[0] - ck_ref[0]) <= 1e-6 * ck_ref[0] and abs(ck[1] - ck_ref[1]) <= 1e-5 * ck_ref[1]


# ---- Gray-Scott and SPH (GPU-validated code; CPU regression checks of the same sources) ----------------------
class _GsParams(C.Structure):       # tau_gs_params
    _fields_ = [("nx", C.c_int), ("ny", C.c_int)] + [(k, C.c_float) for k in ("dx", "dt", "Du", "Dv", "feed", "kill")] + \
               [("seed", C.c_uint)]


@pytest.mark.parametrize("nx,ny,steps", [(128, 64, 20), (100, 37, 15), (256, 130, 10), (33, 5, 8)])
def test_gray_scott_product_code_equals_oracle(nx, ny, steps):
    """gs_step_tma (TMA 2-D box loads of a tile + halo, mbarrier) where nx*4 % 16 == 0, gs_step_generic
    otherwise.  The oracle spells the GPU's FMA contraction out (it is bit-exact against the reference kernel
    on a B200); this build has contraction off, hence a 2-ulp bound instead of equality."""
    L = C.CDLL(hostemu_build.build("gray_scott"))
    L.tau_gs_default_params.argtypes = [C.POINTER(_GsParams)]
    L.tau_gs_create.argtypes = [C.POINTER(_GsParams), C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    L.tau_gs_upload.argtypes = [C.c_void_p, f32p, f32p]
    L.tau_gs_step.argtypes = [C.c_void_p, C.c_int]
    L.tau_gs_download.argtypes = [C.c_void_p, f32p, f32p]
    L.tau_gs_destroy.argtypes = [C.c_void_p]
    p = _GsParams()
    L.tau_gs_default_params(C.byref(p))
    p.nx, p.ny = nx, ny
    u0, v0 = oracle.gs_init_pattern(nx, ny)
    h = C.c_void_p()
    assert L.tau_gs_create(C.byref(p), 0, 0, ny, None, C.byref(h)) == 0
    assert L.tau_gs_upload(h, u0.ravel(), v0.ravel()) == 0 and L.tau_gs_step(h, steps) == 0
    u, v = np.empty(nx * ny, np.float32), np.empty(nx * ny, np.float32)
    assert L.tau_gs_download(h, u, v) == 0
    L.tau_gs_destroy(h)
    eu, ev = oracle.gs_run(u0, v0, steps, Du=p.Du, Dv=p.Dv, dt=p.dt, dx=p.dx, feed=p.feed, kill=p.kill)
    assert np.abs(u.reshape(ny, nx) - eu).max() < 1e-6 and np.abs(v.reshape(ny, nx) - ev).max() < 1e-6
    assert np.abs(eu - u0).max() > 1e-3


@pytest.mark.parametrize("N,frames,over", [(2048, 3, {}), (3000, 2, dict(useXSPH=1)),
                                           (4096, 2, dict(rain=0, viscSub=2))])
def test_sph_product_code_equals_oracle(N, frames, over):
    """the whole sub-step: cell keys, the hand-written LSD radix sort (__match_any_sync digit ranking, tile
    scans), cell ranges, density, forces + integration, XSPH, rain, tau-clock"""
    L = C.CDLL(hostemu_build.build("sph"))
    P = C.POINTER(oracle.SphParams)
    u32p = oracle.u32p
    L.tau_sph_reset_particles.argtypes = [P, f32p, f32p]
    L.tau_sph_create.argtypes = [P, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    L.tau_sph_upload.argtypes = [C.c_void_p, f32p, f32p]
    L.tau_sph_step.argtypes = [C.c_void_p, C.c_int]
    L.tau_sph_download.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p]
    L.tau_sph_download_sort.argtypes = [C.c_void_p, u32p, u32p]
    L.tau_sph_clock.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_longlong)]
    L.tau_sph_destroy.argtypes = [C.c_void_p]
    p = oracle.sph_params(N, **over)
    pos, vel = np.zeros(2 * N, np.float32), np.zeros(2 * N, np.float32)
    L.tau_sph_reset_particles(C.byref(p), pos, vel)
    h = C.c_void_p()
    assert L.tau_sph_create(C.byref(p), 0, None, C.byref(h)) == 0
    assert L.tau_sph_upload(h, pos, vel) == 0 and L.tau_sph_step(h, frames) == 0
    po, ve, s, pr = np.zeros(2 * N, np.float32), np.zeros(2 * N, np.float32), np.zeros(N, np.float32), np.zeros(N, np.float32)
    assert L.tau_sph_download(h, po, ve, s, pr) == 0
    k, v = np.zeros(N, np.uint32), np.zeros(N, np.uint32)
    assert L.tau_sph_download_sort(h, k, v) == 0
    t, tau, st = C.c_float(), C.c_float(), C.c_longlong()
    L.tau_sph_clock(h, C.byref(t), C.byref(tau), C.byref(st))
    L.tau_sph_destroy(h)
    epos, evel, _, es, epr, ck = oracle.sph_run(p, pos, vel, frames)
    assert np.abs(po.reshape(-1, 2) - epos).max() < 1e-6 and np.abs(ve.reshape(-1, 2) - evel).max() < 1e-5
    assert np.abs(s - es).max() < 1e-5 and np.abs(pr - epr).max() < 1e-5      # measured 1-2e-6
    assert t.value == ck.t and st.value == ck.step
    assert np.all(np.diff(k.astype(np.int64)) >= 0) and np.array_equal(np.sort(v), np.arange(N, dtype=np.uint32))
    key_of = np.empty(N, np.int64)   # ... and the permutation is the STABLE sort of the keys
    key_of[v] = k
    assert np.array_equal(v, np.argsort(key_of, kind="stable").astype(np.uint32))


@pytest.mark.parametrize("dtype,world", [("f64", 2), ("f64", 3), ("f32", 3)])
def test_hyp2d_device_side_slab_exchange_is_bit_identical_to_one_domain(pretend_device, dtype, world):
    """The multi-GPU protocol of the 2-D solver with `world` handles in one process standing in for one
    process per GPU (CUDA IPC emulated as plain pointers): every step kernel pushes its boundary rows into
    the neighbours' ghost rows, sends its max wavespeed to every peer's inbox with one store, and the next
    step polls the inbox — no host exchange after the initial hyp2d_sync_state.  Bit-identical to the
    single-domain run (as measured on 2/4/8 B200s), and tau_hyp2d_destroy unmaps what ipc_attach opened."""
    pretend_device(3, 2)
    W, H, steps = 200, 120, 12
    a, m, t, _, _ = hyp2d_emu.run(W, H, steps, dtype, geom_x0=W / 3.0)
    b, mb, ts, open_mappings = hyp2d_emu.run_slabs(W, H, steps, dtype, world, geom_x0=W / 3.0)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and np.array_equal(m, mb)
    assert all(tt == t for tt in ts) and open_mappings == 0


@pytest.mark.parametrize("dtype,cuts", [("f32", (37, 20, 41, 22)), ("f64", (8, 70, 42)), ("f32", (60, 60))])
def test_hyp2d_unequal_slabs_are_bit_identical_to_one_domain(pretend_device, dtype, cuts):
    """What the measured load balancing (slab.hyp2d_balanced_partition) produces: slabs of different heights, cuts through
    the body.  The state must not depend on where the cuts are."""
    pretend_device(3, 2)
    W, H, steps = 200, 120, 12
    a, m, t, _, _ = hyp2d_emu.run(W, H, steps, dtype, geom_x0=W / 3.0)
    b, mb, ts, open_mappings = hyp2d_emu.run_slabs(W, H, steps, dtype, len(cuts), cuts=cuts, geom_x0=W / 3.0)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and np.array_equal(m, mb)
    assert all(tt == t for tt in ts) and open_mappings == 0


def test_hyp2d_pair_kernel_in_slab_mode(pretend_device):
    """pair kernel first, production kernel second (it owns the step's bookkeeping and the peer message);
    both push boundary rows.  Different items go to the pair kernel than in the single-domain run, so the
    comparison is to rounding (FMA contraction), not bit for bit."""
    pretend_device(3, 2)
    W, H, steps = 200, 120, 12
    a, _, t, _, _ = hyp2d_emu.run(W, H, steps, "f32", geom_x0=W / 3.0)
    b, _, ts, open_mappings = hyp2d_emu.run_slabs(W, H, steps, "f32", 2, pair=True, geom_x0=W / 3.0)
    assert max(rel_linf(x, y) for x, y in zip(b, a)) < 2e-6 and all(abs(tt - t) <= 1e-9 * t for tt in ts)
    assert open_mappings == 0


def test_hyp2d_vector_accesses_are_aligned_and_in_bounds():
    """UBSan (alignment, bounds) build of the emulated 2-D solver — production, pair and slab modes — in a
    subprocess (a report aborts it).  float2 / uint2 carry CUDA's alignment in the emulator's headers."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "hostemu", "sanitized_run.py")],
                       capture_output=True, text=True, timeout=600)
    if r.returncode != 0 and "cannot find -lubsan" in r.stderr:
        pytest.skip("libubsan not available")
    assert r.returncode == 0 and "sanitized run clean" in r.stdout, r.stderr[-2000:]


def test_hyp3d_prim_side_buffer_is_bit_identical_to_decoding_every_tile(developed_3d_flow, monkeypatch):
    """Since round 2 the step kernel writes decode(new state) next to the state and the next step's tile builds load it
    instead of decoding their 7.7x-amplified halo themselves (TAU_HYP3D_PRIMS=0: the old kernel).  Same function of the
    same numbers: the two must agree BIT FOR BIT — single domain, and z-slab rings whose ghost planes arrive encoded and
    are decoded in the tile build."""
    prm, dev, solid = developed_3d_flow
    steps, clock = 4, (0.012, 2e-3)
    monkeypatch.setenv("TAU_HYP3D_PRIMS", "0")
    a, _, cka = hyp3d_emu.run(prm, dev, steps, clock, packed=True)
    monkeypatch.setenv("TAU_HYP3D_PRIMS", "1")
    b, _, ckb = hyp3d_emu.run(prm, dev, steps, clock, packed=True)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and cka == ckb
    assert max(np.abs(x - np.asarray(y).reshape(x.shape)).max() for x, y in zip(a, dev)) > 1e-3   # the flow moved
    for world in (2, 3):
        c, cks = hyp3d_emu.run_slabs(prm, dev, steps, clock, world, packed=True)
        assert all(np.array_equal(x, y) for x, y in zip(a, c)), world
        assert all(ck[:2] == cka[:2] for ck in cks), world


@pytest.mark.parametrize("ngpus", [1, 2, 3, 4])
def test_hyp3d_group_equals_single_domain(developed_3d_flow, monkeypatch, ngpus):
    """tau_hyp3d_group_* (ONE process, one z-slab handle per pretend device, ghost planes by cudaMemcpyPeerAsync around the
    periodic ring, max wavespeed folded on the host, every controller committing the same clock) reproduces the single
    handle bit for bit: a developed flow uploaded with its clock, stepped in uneven pieces, and the k_init state."""
    monkeypatch.setenv("TAU_HC_DEVICES", str(ngpus))
    prm, dev, solid = developed_3d_flow
    steps, clock = 5, (0.012, 2e-3)
    a, sol_a, cka = hyp3d_emu.run(prm, dev, steps, clock, packed=True)
    b, sol_b, ckb = hyp3d_emu.run_group(prm, dev, steps, clock, ngpus, chunks=(2, 1, None))
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and np.array_equal(sol_a, sol_b) and cka == ckb
    c, _, ckc = hyp3d_emu.run(prm, None, 3, None, packed=True)
    d, _, ckd = hyp3d_emu.run_group(prm, None, 3, None, ngpus)
    assert all(np.array_equal(x, y) for x, y in zip(c, d)) and ckc == ckd


def test_hyp3d_4spl_frame_export_is_bit_identical_to_the_reference_host_loop(developed_3d_flow):
    """tau_hyp3d_export_frame (never run on hardware): schlieren field (vis mode 8 = th3cs.cu's
    k_schlieren_export), min/max by ordered-integer atomics, palette index by binary search over the 255
    steps of (int)(powf(norm, 0.65f)*255).  Integer work: identical to the host loop th3cs.cu:1199-1222
    (oracle/splat4_oracle.c) applied to the same field."""
    prm, dev, solid = developed_3d_flow
    vis, idx, mm = hyp3d_emu.vis_and_export(prm, dev, modes=tuple(range(9)))
    want_idx, want_mm = oracle.splat4_frame_indices(vis[8])
    assert np.array_equal(idx, want_idx) and mm == want_mm
    assert len(np.unique(idx)) > 100 and idx.max() == 255 and idx.min() == 0      # a real picture
    # the field itself against the oracle's k_vis / k_schlieren_export restatements
    # (every k_vis mode: the tile-staged kernel and the oracle share libm here, so the fields are identical)
    for mode in range(9):
        assert np.array_equal(vis[mode], oracle.hyp3d_vis(prm, dev, solid, mode)), mode
    # 64^3 is the exporter's grid: dx = 1/64 is a power of two, modes 0 and 8 then agree to the last bit
    assert np.abs(vis[8] - vis[0]).max() <= 1e-6 * np.abs(vis[0]).max()


def test_th3cs_main_loop_through_the_product_matches_the_reference_exporter():
    """init -> (4 steps + tau_hyp3d_export_frame) x 48 through the emulated product against the reference
    exporter's own output (tests/golden/th3cs_ref_host.npz, see tests/test_oracle_cpu.py).  The product's step
    kernel differs from the reference's in rounding (face-once fluxes, reciprocal-based division), so a few
    voxels sit on the other side of a palette step late in the run: measured 3 of 13 824, off by one."""
    import os as _os
    g = np.load(_os.path.join(_os.path.dirname(__file__), "golden", "th3cs_ref_host.npz"))
    idx = g["indices"]
    frames, n = idx.shape[0], idx.shape[1]
    out = hyp3d_emu.export_video(oracle.hyp3d_params(n, n, n), frames)
    d = np.abs(out.astype(int) - idx.astype(int)).reshape(frames, -1)
    assert (d[:24] == 0).all()                                     # half the run: identical
    assert d.max() <= 2 and (d != 0).sum(axis=1).max() <= 0.005 * d.shape[1]


def test_hyp2d_headline_width_and_height(pretend_device):
    """the benchmarked grid's extents, one axis at a time (4096 x 64 and 256 x 4096): 137 strips with the
    16-column remainder strip at the right edge, TMA boxes hanging over x = W, 69 pair items per layer, the
    guided layer schedule over 4096 rows — production and pair kernels against the fp64 oracle"""
    pretend_device(8, 5)
    for W, H, steps in ((4096, 64, 3), (256, 4096, 2)):
        yy, xx = np.mgrid[0:H, 0:W]
        rho = 1.0 + 0.3 * np.sin(xx / 9.0) * np.cos(yy / 7.0)
        u, v = 3.0 + 0.5 * np.cos(xx / 11.0), 0.7 * np.sin(yy / 5.0)
        p = 1.0 + 0.2 * np.cos((xx + yy) / 13.0)
        planes = [rho, rho * u, rho * v, p / 0.1 + 0.5 * rho * (u * u + v * v)]
        mask = np.zeros((H, W), np.uint8)
        mask[H // 2:H // 2 + 12, W // 8:W // 8 + 40] = 1
        mask[0:3, W - 90:W - 80] = 1
        mask[10:14, W - 3:] = 1
        mask[H - 4:, 0:2] = 1
        ref, t_ref, _ = oracle.hyp2d_run(oracle.hyp2d_cfg(W, H), planes, mask.ravel(), steps)
        a, _, ta, _, _ = hyp2d_emu.run(W, H, steps, "f32", planes=planes, mask=mask)
        b, _, tb, _, _ = hyp2d_emu.run(W, H, steps, "f32", planes=planes, mask=mask, pair=True)
        items = hyp2d_emu.run.last_work_items
        assert items[2] > 0.3 * items[0] and 2 * items[2] + items[3] >= items[0]   # a pair item = two strips; the rest is cut into 8-row pieces
        assert max(rel_linf(x, y) for x, y in zip(a, ref)) < 2e-6
        assert max(rel_linf(x, y) for x, y in zip(b, ref)) < 2e-6 and max(rel_linf(x, y) for x, y in zip(b, a)) < 1e-6
        assert abs(ta - t_ref) <= 1e-6 * t_ref and ta == tb


@pytest.mark.parametrize("order", ["reverse", "random"])
def test_results_do_not_depend_on_block_order(swlib, pretend_device, monkeypatch, order):
    """a GPU promises no block order: reversed and shuffled launches (TAU_HC_BLOCK_ORDER) must give the same
    bits — the clear-two-steps-ahead control slots, the device work queue (items go to different CTAs), the
    last-CTA-out reduction, the peer messages"""
    pretend_device(3, 2)
    prm = oracle.sw_params(nx=70, ny=37, dtau=0.05, nu=0.02, dx=2.0, dy=1.5, **GENTLE)
    s0, u0, v0 = oracle.sw_init(prm)
    W, H, steps = 200, 120, 8
    base_sw, ck_sw, _ = sw_emulated(swlib, prm, s0, u0, v0, 15)
    base64, _, t64, _, _ = hyp2d_emu.run(W, H, steps, "f64", geom_x0=W / 3.0)
    base_pair, _, tp, _, _ = hyp2d_emu.run(W, H, steps, "f32", pair=True, geom_x0=W / 3.0)
    monkeypatch.setenv("TAU_HC_BLOCK_ORDER", order)
    got_sw, ck2, _ = sw_emulated(swlib, prm, s0, u0, v0, 15)
    assert all(np.array_equal(a, b) for a, b in zip(base_sw, got_sw)) and ck_sw == ck2
    got64, _, t2, _, _ = hyp2d_emu.run(W, H, steps, "f64", geom_x0=W / 3.0)
    assert all(np.array_equal(a, b) for a, b in zip(base64, got64)) and t2 == t64
    got_pair, _, tp2, _, _ = hyp2d_emu.run(W, H, steps, "f32", pair=True, geom_x0=W / 3.0)
    assert all(np.array_equal(a, b) for a, b in zip(base_pair, got_pair)) and tp2 == tp
    slabs, _, ts, _ = hyp2d_emu.run_slabs(W, H, steps, "f64", 3, geom_x0=W / 3.0)
    assert all(np.array_equal(a, b) for a, b in zip(base64, slabs)) and all(t == t64 for t in ts)


def test_hyp2d_frame_handover_between_ranks_on_the_device(pretend_device):
    """tau_hyp2d_upload_peers_async (never run on hardware): a frame upload as a pseudo-step of the control-slot
    rotation — wait for the peers' last-step messages, H2D, wavespeed scan, push of the boundary rows into the
    neighbours' ghost rows, one message per peer — replaces the host-driven exchange (NCCL + barrier) between
    frames.  Three frames with 4 / 1 / 5 steps (every residue of the three-slot rotation), 2 and 3 ranks:
    bit-identical to one domain that gets the same frames through tau_hyp2d_upload."""
    pretend_device(3, 2)
    W, H = 200, 120
    yy, xx = np.mgrid[0:H, 0:W]

    def state(k):
        rho = 1.0 + 0.3 * np.sin(xx / (9.0 + k)) * np.cos(yy / 7.0)
        u, v = 3.0 + 0.5 * np.cos(xx / 11.0), 0.7 * np.sin(yy / (5.0 + k))
        p = 1.0 + 0.2 * np.cos((xx + yy) / 13.0)
        return [rho, rho * u, rho * v, p / 0.1 + 0.5 * rho * (u * u + v * v)]
    frames = [(state(0), 4), (state(1), 1), (state(2), 5)]
    L = hyp2d_emu.lib()
    for dtype, npdt in (("f64", np.float64), ("f32", np.float32)):
        cc = hyp2d_emu.default_cfg(W, H, geom_x0=W / 3.0)
        h = C.c_void_p()
        hyp2d_emu.check(L.tau_hyp2d_create(C.byref(cc), W, H, 0 if dtype == "f32" else 1, 0, 0, H, None, C.byref(h)))
        hyp2d_emu.check(L.tau_hyp2d_init(h))
        hyp2d_emu.check(L.tau_hyp2d_step(h, 3))
        for fp, fs in frames:
            arrs = [np.ascontiguousarray(p, npdt) for p in fp]
            hyp2d_emu.check(L.tau_hyp2d_upload(h, (C.c_void_p * 4)(*[a.ctypes.data for a in arrs]), None))
            hyp2d_emu.check(L.tau_hyp2d_step(h, fs))
        one = [np.empty((H, W), npdt) for _ in range(4)]
        m = np.empty((H, W), np.uint8)
        hyp2d_emu.check(L.tau_hyp2d_download(h, (C.c_void_p * 4)(*[a.ctypes.data for a in one]), C.c_void_p(m.ctypes.data)))
        L.tau_hyp2d_destroy(h)
        for world in (2, 3):
            got, _, _, open_mappings = hyp2d_emu.run_slabs(W, H, 3, dtype, world, frames=frames,
                                                           reverse_ranks=(world == 3), geom_x0=W / 3.0)
            assert all(np.array_equal(a, b) for a, b in zip(one, got)) and open_mappings == 0, (dtype, world)
    got, _, _, _ = hyp2d_emu.run_slabs(W, H, 3, "f32", 2, pair=True, frames=frames, geom_x0=W / 3.0)
    assert max(rel_linf(a, b) for a, b in zip(got, one)) < 2e-6      # pair kernel: FMA rounding only


def test_c_host_programs_run_end_to_end_on_the_emulator(tmp_path):
    """The C hosts of fluid_sims_b200/cli/ (the reference's binary names and flags) with the emulated library
    preloaded in front of libtau_b200.so: argument parsing, frame loops, reports and file output of every
    program, incl. the two that have not met hardware (tau_sw, th3cs — whose .4spl file must contain the very
    frames the reference exporter produces, tests/golden/th3cs_ref_host.npz)."""
    import subprocess
    from fluid_sims_b200 import splat4
    cli = os.path.join(os.path.dirname(__file__), "..", "fluid_sims_b200", "cli")
    subprocess.run(["make", "-C", cli], check=True, capture_output=True)
    env = dict(os.environ, LD_PRELOAD=hostemu_build.build_all(), TAU_HC_SMS="3", TAU_HC_CTAS_PER_SM="2")
    out4 = str(tmp_path / "v.4spl")
    cases = {"tgs": ["--nx", "64", "--ny", "32", "--steps", "3", "--headless"],
             "tau_2d_hypersonic_cuda": ["--nx", "128", "--ny", "64", "--frames", "2"],
             "tau3d": ["--n", "16", "--frames", "1"],
             "tau_sph": ["--n", "2048", "--frames", "2"],
             "tau_burgers": ["--nx", "96", "--ny", "64", "--steps", "4", "--headless", "--dtau", "1e-3"],
             "tau_sw": ["--nx", "96", "--ny", "64", "--steps", "20", "--headless", "--dtau", "1e-3"],
             "th3cs": ["--n", "24", "--frames", "6", "--out", out4]}
    for exe, args in cases.items():
        r = subprocess.run([os.path.join(cli, exe)] + args, capture_output=True, text=True, timeout=300, env=env,
                           cwd=str(tmp_path))
        assert r.returncode == 0 and "updates/s" in r.stdout, (exe, r.stdout[-300:], r.stderr[-300:])
    # tau3d --gpus N (tau_hyp3d_group_*, one process): the dump is the --gpus 1 dump, byte for byte
    dumps = []
    for gpus in (1, 3):
        d = str(tmp_path / f"t3_{gpus}.bin")
        r = subprocess.run([os.path.join(cli, "tau3d"), "--n", "18", "--frames", "3", "--gpus", str(gpus), "--dump", d],
                           capture_output=True, text=True, timeout=300, env=dict(env, TAU_HC_DEVICES="3"), cwd=str(tmp_path))
        assert r.returncode == 0 and f"on {gpus} GPU" in r.stdout, (gpus, r.stdout[-300:], r.stderr[-300:])
        dumps.append(open(d, "rb").read())
    assert dumps[0] == dumps[1] and len(dumps[0]) > 6 * 4 * 18 ** 3
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "th3cs_ref_host.npz"))
    assert splat4.info(out4) == dict(width=24, height=24, depth=24, frames=6, pSize=256, flags=4)
    assert np.array_equal(splat4.parse(open(out4, "rb").read())["indices"], g["indices"][:6])


def test_hyp2d_fuzz_random_grids_masks_schedules(monkeypatch):
    """seeded random sweep of what a user can vary: grid 8..330 x 5..150, speckled and blocky body masks touching
    any boundary, segment heights, pretend devices (1..8 SMs x 1..5 CTAs), block orders, fp64 / fp32 / pair
    mode, 1..6 steps, and 2..4 in-process ranks — every case against the fp64 oracle (~1100 further cases of the
    same generator were run once when this test was written: no failure)."""
    rng = np.random.default_rng(11)
    for case in range(24):
        W, H, steps = int(rng.integers(8, 330)), int(rng.integers(5, 150)), int(rng.integers(1, 7))
        seg = [None, 4, 8, 13, 32, 64][int(rng.integers(0, 6))]
        monkeypatch.setenv("TAU_HC_SMS", str(int(rng.integers(1, 9))))
        monkeypatch.setenv("TAU_HC_CTAS_PER_SM", str(int(rng.integers(1, 6))))
        monkeypatch.setenv("TAU_HC_BLOCK_ORDER", ["", "reverse", "random"][int(rng.integers(0, 3))])
        yy, xx = np.mgrid[0:H, 0:W]
        rho = 1.0 + 0.3 * np.sin(xx / 9.0 + rng.random()) * np.cos(yy / 7.0)
        u, v = 3.0 * rng.random() + 0.5 * np.cos(xx / 11.0), 0.7 * np.sin(yy / 5.0)
        p = 1.0 + 0.2 * np.cos((xx + yy) / 13.0)
        planes = [rho, rho * u, rho * v, p / 0.1 + 0.5 * rho * (u * u + v * v)]
        mask = (rng.random((H, W)) < rng.choice([0.0, 0.002, 0.02])).astype(np.uint8)
        if rng.random() < 0.5:
            y0, x0 = int(rng.integers(0, H)), int(rng.integers(0, W))
            mask[y0:y0 + int(rng.integers(1, 20)), x0:x0 + int(rng.integers(1, 40))] = 1
        mask[0, 0] = 0
        ref, t_ref, _ = oracle.hyp2d_run(oracle.hyp2d_cfg(W, H), planes, mask.ravel(), steps)
        dtype = "f64" if rng.random() < 0.6 else "f32"
        pair = dtype == "f32" and rng.random() < 0.6
        out, m, t, _, _ = hyp2d_emu.run(W, H, steps, dtype, planes=planes, mask=mask, seg_rows=seg, pair=pair)
        err = max(rel_linf(a, b) for a, b in zip(out, ref))
        assert err < (1e-11 if dtype == "f64" else 2e-5) and np.array_equal(m, mask), (case, W, H, steps, seg, dtype, pair, err)
        assert abs(t - t_ref) <= 1e-6 * t_ref
    for case in range(6):      # k_init states cut into 2..4 slabs, fp64: bit-identical to one domain
        W, H, steps = int(rng.integers(40, 260)), int(rng.integers(24, 130)), int(rng.integers(2, 7))
        world = int(rng.integers(2, 5))
        monkeypatch.setenv("TAU_HC_BLOCK_ORDER", ["", "reverse", "random"][int(rng.integers(0, 3))])
        a, _, t, _, _ = hyp2d_emu.run(W, H, steps, "f64", geom_x0=W / 3.0)
        b, _, ts, open_mappings = hyp2d_emu.run_slabs(W, H, steps, "f64", world, geom_x0=W / 3.0)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)) and all(tt == t for tt in ts) and open_mappings == 0, \
            (case, W, H, steps, world)


def test_shallow_water_fuzz_bit_identical(swlib, monkeypatch):
    """40 seeded random problems (grids from 1 x 1 to 140 x 90, every combination of cell sizes, viscosity,
    clock rates, depths, bump and swirl strengths), chunked stepping, any block order: the emulated product
    equals the oracle bit for bit (150 further cases were run once when this test was written)."""
    rng = np.random.default_rng(5)
    for case in range(40):
        nx, ny, steps = int(rng.integers(1, 140)), int(rng.integers(1, 90)), int(rng.integers(1, 12))
        monkeypatch.setenv("TAU_HC_BLOCK_ORDER", ["", "reverse", "random"][int(rng.integers(0, 3))])
        kw = dict(nx=nx, ny=ny, dx=float(rng.choice([1.0, 2.0, 0.5])), dy=float(rng.choice([1.0, 1.5])),
                  nu=float(rng.choice([0.0, 0.02, 0.3])), dtau=float(rng.choice([1e-3, 0.05, 1.0])),
                  H0=float(rng.choice([1.0, 10.0, 1000.0])), bumpAmp=float(rng.choice([0.0, 0.3, 1.0])),
                  bumpSigma=float(rng.choice([1.0, 4.0])), asym=float(rng.choice([0.0, 0.3])),
                  swirl=float(rng.choice([0.0, 0.05, 1.0])), swirlRc=float(rng.choice([5.0, 100.0])),
                  offx=float(rng.integers(-5, 6)), offy=float(rng.integers(-5, 6)))
        prm = oracle.sw_params(**kw)
        s0, u0, v0 = oracle.sw_init(prm)
        (s, u, v), ck, _ = sw_emulated(swlib, prm, s0, u0, v0, steps, chunks=1 if steps % 2 else 2)
        es, eu, ev, eck, _ = oracle.sw_run(prm, s0, u0, v0, steps)
        assert all(np.array_equal(a, b, equal_nan=True) for a, b in ((s, es), (u, eu), (v, ev))), (case, kw, steps)
        assert ck[0] == eck[0] and ck[1] == eck[1]


def test_hyp2d_frame_handover_fuzz(monkeypatch):
    """random frame sequences through tau_hyp2d_upload_peers_async: 0..3 steps before the first hand-over, 1..4
    frames of 0..4 steps each (0 = two uploads in a row), 2..4 ranks, either rank call order, any block order and
    pretend device — bit-identical to one domain (30 cases of this generator were run when it was written)."""
    L = hyp2d_emu.lib()
    rng = np.random.default_rng(21)
    for case in range(6):
        W, H, world, pre = int(rng.integers(10, 50)) * 4, int(rng.integers(24, 120)), int(rng.integers(2, 5)), int(rng.integers(0, 4))
        dtype = "f64" if rng.random() < 0.5 else "f32"
        npdt = np.float64 if dtype == "f64" else np.float32
        monkeypatch.setenv("TAU_HC_BLOCK_ORDER", ["", "reverse", "random"][int(rng.integers(0, 3))])
        monkeypatch.setenv("TAU_HC_SMS", str(int(rng.integers(1, 5))))
        monkeypatch.setenv("TAU_HC_CTAS_PER_SM", str(int(rng.integers(1, 4))))
        yy, xx = np.mgrid[0:H, 0:W]

        def state(k):
            rho = 1.0 + 0.3 * np.sin(xx / (9.0 + k)) * np.cos(yy / 7.0)
            u, v = 3.0 + 0.5 * np.cos(xx / 11.0), 0.7 * np.sin(yy / (5.0 + k))
            p = 1.0 + 0.2 * np.cos((xx + yy) / 13.0)
            return [rho, rho * u, rho * v, p / 0.1 + 0.5 * rho * (u * u + v * v)]
        frames = [(state(k), int(rng.integers(0, 5))) for k in range(int(rng.integers(1, 5)))]
        cc = hyp2d_emu.default_cfg(W, H, geom_x0=W / 3.0)
        h = C.c_void_p()
        hyp2d_emu.check(L.tau_hyp2d_create(C.byref(cc), W, H, 0 if dtype == "f32" else 1, 0, 0, H, None, C.byref(h)))
        hyp2d_emu.check(L.tau_hyp2d_init(h))
        hyp2d_emu.check(L.tau_hyp2d_step(h, pre))
        for fp, fs in frames:
            arrs = [np.ascontiguousarray(p, npdt) for p in fp]
            hyp2d_emu.check(L.tau_hyp2d_upload(h, (C.c_void_p * 4)(*[a.ctypes.data for a in arrs]), None))
            hyp2d_emu.check(L.tau_hyp2d_step(h, fs))
        one = [np.empty((H, W), npdt) for _ in range(4)]
        m = np.empty((H, W), np.uint8)
        hyp2d_emu.check(L.tau_hyp2d_download(h, (C.c_void_p * 4)(*[a.ctypes.data for a in one]), C.c_void_p(m.ctypes.data)))
        L.tau_hyp2d_destroy(h)
        got, _, _, open_mappings = hyp2d_emu.run_slabs(W, H, pre, dtype, world, frames=frames,
                                                       reverse_ranks=bool(rng.integers(0, 2)), geom_x0=W / 3.0)
        assert all(np.array_equal(a, b) for a, b in zip(one, got)) and open_mappings == 0, \
            (case, W, H, world, pre, dtype, [f[1] for f in frames])


def test_hyp2d_fused_kernel_equals_the_two_kernel_pair_mode(pretend_device, monkeypatch):
    """hypersonic2d_fused.cuh (TAU_HYP2D_PAIR=2, never run on hardware): pair and production items claimed from
    ONE table by ONE kernel per step — the included text of both marches around the production kernel's prologue
    and epilogue.  Same arithmetic per cell as TAU_HYP2D_PAIR=1, so the results must be bit-identical to it;
    launches per step drop from 2 to 1; slab mode, frame hand-over and any block order included."""
    W, H, steps = 308, 96, 10
    planes, mask = _random_state_with_walls(W, H)
    for sms, ctas, order in ((3, 2, ""), (1, 1, "reverse"), (5, 3, "random")):
        pretend_device(sms, ctas)
        monkeypatch.setenv("TAU_HC_BLOCK_ORDER", order)
        a, _, ta, _, na = hyp2d_emu.run(W, H, steps, "f32", planes=planes, mask=mask)
        b, _, tb, dtsb, nb = hyp2d_emu.run(W, H, steps, "f32", planes=planes, mask=mask, pair=1)
        have_pairs = hyp2d_emu.run.last_work_items[2] > 0         # (tall layers on a tiny device can all touch a hole)
        c, _, tc, dtsc, nc = hyp2d_emu.run(W, H, steps, "f32", planes=planes, mask=mask, pair=2)
        assert all(np.array_equal(x, y) for x, y in zip(b, c)) and tb == tc and np.array_equal(dtsb, dtsc)
        assert nc == na and nb == na + (steps if have_pairs else 0)   # one kernel per step again
        assert have_pairs or sms == 1
        assert max(rel_linf(x, y) for x, y in zip(c, a)) < 2e-6
    monkeypatch.setenv("TAU_HC_BLOCK_ORDER", "")
    pretend_device(3, 2)
    one, _, t, _, _ = hyp2d_emu.run(200, 120, 8, "f32", geom_x0=66.0)
    for world in (2, 3):
        got, _, ts, open_mappings = hyp2d_emu.run_slabs(200, 120, 8, "f32", world, pair=2, geom_x0=66.0)
        assert max(rel_linf(x, y) for x, y in zip(got, one)) < 2e-6 and open_mappings == 0
        assert all(abs(tt - t) <= 1e-9 * t for tt in ts)
    # the device-side frame hand-over with the fused kernel
    yy, xx = np.mgrid[0:120, 0:200]
    rho = 1.0 + 0.3 * np.sin(xx / 9.0) * np.cos(yy / 7.0)
    u, v = 3.0 + 0.5 * np.cos(xx / 11.0), 0.7 * np.sin(yy / 5.0)
    frame = [rho, rho * u, rho * v, (1.0 + 0.2 * np.cos((xx + yy) / 13.0)) / 0.1 + 0.5 * rho * (u * u + v * v)]
    f1, _, _, _ = hyp2d_emu.run_slabs(200, 120, 3, "f32", 2, pair=1, frames=[(frame, 4)], geom_x0=66.0)
    f2, _, _, _ = hyp2d_emu.run_slabs(200, 120, 3, "f32", 2, pair=2, frames=[(frame, 4)], geom_x0=66.0)
    assert all(np.array_equal(x, y) for x, y in zip(f1, f2))


# ---- tau_hypersonic.c's update path (hypersonic_c.cu, BASELINE config 1) -----------------------------------------
@pytest.mark.parametrize("W,H,steps", [(64, 48, 40), (130, 33, 25)])
def test_hypc_product_code_equals_oracle_bit_for_bit(W, H, steps):
    """hypersonic_c.cu on the emulator (same -ffp-contract=off arithmetic as its --fmad=false device build): fields,
    mask, sim_t and dt equal the oracle — which is pinned 0 ulp on the reference's own object code — and the speed-mode
    render (sqrt only) equals the oracle's pixel for pixel."""
    lib = C.CDLL(hostemu_build.build("hypersonic_c"))
    f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    h = C.c_void_p()
    lib.tau_hypc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.tau_hypc_upload.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_double]
    lib.tau_hypc_step.argtypes = [C.c_void_p, C.c_int]
    lib.tau_hypc_download.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]
    lib.tau_hypc_clock.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.tau_hypc_render.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
    lib.tau_hypc_destroy.argtypes = [C.c_void_p]
    lib.tau_hostemu_last_error.restype = C.c_char_p
    assert lib.tau_hypc_create(W, H, 0, None, C.byref(h)) == 0, lib.tau_hostemu_last_error()
    planes, mask = oracle.hypcpu_init(W, H)
    rng = np.random.default_rng(W)
    mask = mask.reshape(H, W).copy()
    mask[H // 4:H // 4 + 3, W // 2:W // 2 + 5] = 1           # a second body: slip-wall ghosts on flat faces
    mask[H // 2, 0] = 1                                       # a body cell in the inflow column
    mask = mask.ravel()
    planes = [p * (1.0 + 0.05 * rng.random(p.size)) for p in planes]
    for p in planes[1:3]:
        p[mask == 1] = 0.0
    ptrs = (C.c_void_p * 4)(*[p.ctypes.data for p in planes])
    assert lib.tau_hypc_upload(h, ptrs, mask.ctypes.data, 0.25) == 0
    assert lib.tau_hypc_step(h, steps) == 0
    out = [np.empty(W * H, np.float64) for _ in range(4)]
    m = np.empty(W * H, np.uint8)
    assert lib.tau_hypc_download(h, (C.c_void_p * 4)(*[p.ctypes.data for p in out]), m.ctypes.data) == 0
    t, dt = C.c_double(), C.c_double()
    assert lib.tau_hypc_clock(h, C.byref(t), C.byref(dt)) == 0
    exp, et, dts = oracle.hypcpu_run(W, H, planes, mask, steps, sim_t=0.25)
    assert np.array_equal(m, mask) and t.value == et and dt.value == dts[-1]
    for a, b in zip(out, exp):
        assert np.array_equal(a, b)
    px = np.empty(W * H, np.uint32)
    mm = (C.c_double * 2)()
    assert lib.tau_hypc_render(h, 2, px.ctypes.data, mm) == 0
    ergba, emm, _ = oracle.hypcpu_render(W, H, exp, mask, 2)
    assert (mm[0], mm[1]) == emm and np.array_equal(px.view(np.uint8).reshape(H, W, 4), ergba)
    lib.tau_hypc_destroy(h)


# ---- tau_hyp2d_group: the slab handles of one process (multi-GPU behind the C boundary) --------------------------
@pytest.mark.parametrize("ngpus,dtype", [(2, "f64"), (3, "f32"), (4, "f32")])
def test_hyp2d_group_equals_single_domain(pretend_device, monkeypatch, ngpus, dtype):
    """tau_hyp2d_group_* (ONE process, one slab handle per pretend device, peers as plain pointers, hand-over by
    cudaMemcpyPeer + host max) reproduces the single-domain handle bit for bit — init and uploaded states with a
    body crossing slab boundaries, chunked stepping, download and the render pass over the whole grid."""
    from hyp2d_emu import default_cfg, lib as h2lib
    pretend_device(3, 2)
    monkeypatch.setenv("TAU_HC_DEVICES", str(ngpus))
    monkeypatch.setenv("TAU_HYP2D_GROUP_CHUNK", "1")     # emulated kernels run at launch: ranks must alternate
    monkeypatch.delenv("TAU_HYP2D_PAIR", raising=False)
    L = h2lib()
    g, h = C.c_void_p(), C.c_void_p()
    L.tau_hyp2d_group_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    for fn in ("init", "sync", "destroy"):
        getattr(L, f"tau_hyp2d_group_{fn}").argtypes = [C.c_void_p]
    L.tau_hyp2d_group_step.argtypes = [C.c_void_p, C.c_int]
    L.tau_hyp2d_group_upload.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]
    L.tau_hyp2d_group_download.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]
    L.tau_hyp2d_group_clock.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.tau_hyp2d_group_render.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
    L.tau_hyp2d_render.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
    W, H, steps = 200, 120, 9
    npdt = np.float32 if dtype == "f32" else np.float64
    cc = default_cfg(W, H, geom_x0=W / 3.0)
    dt = 0 if dtype == "f32" else 1

    def state(handle_download, handle, n):
        out = [np.empty((H, W), npdt) for _ in range(4)]
        m = np.empty((H, W), np.uint8)
        assert handle_download(handle, (C.c_void_p * 4)(*[a.ctypes.data for a in out]), m.ctypes.data) == 0
        return out, m

    assert L.tau_hyp2d_group_create(C.byref(cc), W, H, dt, ngpus, None, C.byref(g)) == 0, L.tau_hostemu_last_error()
    assert L.tau_hyp2d_create(C.byref(cc), W, H, dt, 0, 0, H, None, C.byref(h)) == 0
    assert L.tau_hyp2d_group_init(g) == 0 and L.tau_hyp2d_init(h) == 0
    assert L.tau_hyp2d_group_step(g, steps) == 0 and L.tau_hyp2d_step(h, steps) == 0
    a, ma = state(L.tau_hyp2d_group_download, g, ngpus)
    b, mb = state(L.tau_hyp2d_download, h, 1)
    assert np.array_equal(ma, mb) and all(np.array_equal(x, y) for x, y in zip(a, b))
    # an uploaded state (a second wall that crosses a slab boundary), more steps, clock, render
    rng = np.random.default_rng(ngpus)
    mb = mb.copy()
    mb[H // ngpus - 3:H // ngpus + 4, 150:158] = 1
    up = [np.ascontiguousarray(x * (1 + 0.02 * rng.random(x.shape)), npdt) for x in b]
    for x in up[1:3]:
        x[mb == 1] = 0
    ptrs = (C.c_void_p * 4)(*[x.ctypes.data for x in up])
    assert L.tau_hyp2d_group_upload(g, ptrs, mb.ctypes.data) == 0 and L.tau_hyp2d_upload(h, ptrs, mb.ctypes.data) == 0
    assert L.tau_hyp2d_group_step(g, 7) == 0 and L.tau_hyp2d_step(h, 7) == 0
    a, ma = state(L.tau_hyp2d_group_download, g, ngpus)
    b, mb2 = state(L.tau_hyp2d_download, h, 1)
    assert np.array_equal(ma, mb2) and all(np.array_equal(x, y) for x, y in zip(a, b))
    t1, d1, t2, d2 = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    L.tau_hyp2d_group_clock(g, C.byref(t1), C.byref(d1))
    L.tau_hyp2d_clock(h, C.byref(t2), C.byref(d2))
    assert (t1.value, d1.value) == (t2.value, d2.value) and t1.value > 0
    pa, pb = np.empty((H, W), np.uint32), np.empty((H, W), np.uint32)
    m1, m2 = (C.c_double * 2)(), (C.c_double * 2)()
    assert L.tau_hyp2d_group_render(g, 5, pa.ctypes.data, m1) == 0 and L.tau_hyp2d_render(h, 5, pb.ctypes.data, m2) == 0
    assert tuple(m1) == tuple(m2) and np.array_equal(pa, pb)
    assert L.tau_hyp2d_group_destroy(g) == 0 and L.tau_hyp2d_destroy(h) == 0


# ---- SPH sharded by hash-bin stripes (sph_stripes.inc): ghost exchange, migration, re-balancing, rain ------------
@pytest.mark.parametrize("N,world,frames,over", [
    (3000, 1, 4, {}), (3000, 2, 4, {}), (5000, 3, 5, dict(viscSub=3, rebalance_every=2)),
    (6000, 4, 7, dict(useXSPH=1)), (4000, 2, 6, dict(rain=0, gammaEOS=2.0, c0=2.0, useGrav=0)), (12000, 8, 5, {})])
def test_sph_stripes_equal_single_handle_bit_for_bit(N, world, frames, over):
    """`world` rank handles of the stripe-sharded solver in one process, stepping in lock-step with memmove exchanges
    (tests/hostemu/sph_stripes_emu.py), against the single-GPU handle of the same emulated library: positions, velocities
    and the clock bit-identical — through migration, ghost rows, the in-cell order by global id, stripe re-balancing, XSPH's
    third exchange and the rain's re-homing; no error bit; the owned sets partition the particles.  s / press may differ
    only on the handful of particles the rain re-homed to another rank in the last sub-step (their s stays behind)."""
    import json
    import subprocess
    so = hostemu_build.build("sph")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TAU_B200_LIB=so, PYTHONPATH=root)
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "hostemu", "sph_stripes_emu.py"), str(N), str(world),
                        str(frames), json.dumps(over)], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads(r.stdout.strip().split("\n")[-1])
    assert j["pos_equal"] and j["vel_equal"] and j["clock_equal"] and j["moved"] > 0.05, j
    assert all(s["err"] == 0 for s in j["status"]) and sum(s["n_own"] for s in j["status"]) >= N
    assert j["s_mismatches"] <= (0 if not over.get("rain", 1) else 12) and j["press_mismatches"] <= 12
    if world > 1:
        assert all(s["n_ghost"] > 0 for s in j["status"]) and max(s["max_send"] for s in j["status"]) > 0
