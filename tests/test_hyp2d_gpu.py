"""GPU parity of the fused 2-D hypersonic step (C-ABI) against the fp64 CPU oracle, the committed
golden fixtures (made by the reference kernels on a B200) and the reference kernels themselves
(oracle/_ref) on the same device.

Tolerances.  fp64 handle: the fused kernel evaluates the reference's expression trees in the same
precision; only FMA contraction/ordering differs, so we demand 1e-11 relative to the field's max
(observed ~1e-15).  fp32 handle (BASELINE config 2): the reference is fp64, so the bound is a
float-rounding bound; per-field L-inf normalised by the field's max |value|, growing with step
count because limiter/HLLC branches flip on last-bit differences near the shock.
"""
import os

import numpy as np
import pytest

import oracle
from fluid_sims_b200.hypersonic2d import Hypersonic2D, SimConfig

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ("rho", "mx", "my", "E")


def rel_linf(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


def product_run(W, H, steps, dtype, planes=None, mask=None, seg_rows=None, **over):
    cfg = SimConfig.default(W, H, **over)
    s = Hypersonic2D(cfg, dtype=dtype)
    if seg_rows:
        s.set_seg_rows(seg_rows)
    if planes is None:
        s.init()
    else:
        s.upload(planes, mask)
    dts = []
    for _ in range(steps):
        s.step(1)
        dts.append(s.clock()[1])
    out, m = s.download()
    t, _ = s.clock()
    s.close()
    return out, m, t, np.array(dts)


@pytest.mark.parametrize("name", ["256x128", "200x120"])
def test_f64_matches_reference_golden(name):
    g = np.load(os.path.join(GOLDEN, f"hyp2d_ref_{name}.npz"))
    W, H = map(int, name.split("x"))
    steps = int(g["steps"])
    out, m, t, dts = product_run(W, H, steps, "f64")
    assert np.array_equal(m.ravel(), g["mask"])
    for k, a in zip(NAMES, out):
        assert rel_linf(a, g[k]) < 1e-11, k
    assert abs(t - float(g["sim_t"])) < 1e-12
    assert np.abs(dts - g["dts"]).max() < 1e-14
    # second configuration: Mach 3, body moved upstream (walls matter more)
    out, m, t, _ = product_run(W, H, steps, "f64", inflow_mach=3.0, geom_x0=40.0)
    assert np.array_equal(m.ravel(), g["mask_b"])
    for k, a in zip(NAMES, out):
        assert rel_linf(a, g[k + "_b"]) < 1e-8, k


@pytest.mark.parametrize("W,H,steps,seg", [(256, 128, 40, None), (200, 120, 30, 16), (64, 33, 25, 8),
                                           (36, 7, 12, 4), (124, 64, 20, 64)])
def test_f64_matches_oracle(W, H, steps, seg):
    cfg = oracle.hyp2d_cfg(W, H, geom_x0=min(125.0, W / 3.0))
    planes, mask = oracle.hyp2d_init(cfg)
    ref, t_ref, dts_ref = oracle.hyp2d_run(cfg, planes, mask, steps)
    out, m, t, dts = product_run(W, H, steps, "f64", seg_rows=seg, geom_x0=min(125.0, W / 3.0))
    assert np.array_equal(m.ravel(), mask)
    for k, a, b in zip(NAMES, out, ref):
        assert rel_linf(a, b) < 1e-11, k
    assert abs(t - t_ref) < 1e-12 and np.abs(dts - dts_ref).max() < 1e-14


def test_generic_loader_odd_width_matches_oracle():
    """W*sizeof(real) not a multiple of 16: the non-TMA loader path."""
    W, H, steps = 203, 57, 20
    cfg = oracle.hyp2d_cfg(W, H, geom_x0=60.0)
    planes, mask = oracle.hyp2d_init(cfg)
    ref, _, _ = oracle.hyp2d_run(cfg, planes, mask, steps)
    for dtype, tol in (("f64", 1e-11), ("f32", 5e-4)):
        out, m, _, _ = product_run(W, H, steps, dtype, geom_x0=60.0)
        assert np.array_equal(m.ravel(), mask)
        for k, a, b in zip(NAMES, out, ref):
            assert rel_linf(a, b) < tol, (dtype, k)


def test_upload_random_state_with_walls():
    """Caller-injected state: smooth random field + a hand-made mask touching every boundary."""
    W, H, steps = 128, 96, 15
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
    mask[rng.random((H, W)) < 0.002] = 1
    cfg = oracle.hyp2d_cfg(W, H)
    ref, t_ref, _ = oracle.hyp2d_run(cfg, planes, mask.ravel(), steps)
    out, m, t, _ = product_run(W, H, steps, "f64", planes=planes, mask=mask)
    for k, a, b in zip(NAMES, out, ref):
        assert rel_linf(a, b) < 1e-11, k
    out, m, t, _ = product_run(W, H, steps, "f32", planes=planes, mask=mask)
    for k, a, b in zip(NAMES, out, ref):
        assert rel_linf(a, b) < 5e-5, k


@pytest.mark.parametrize("steps,tol", [(1, 2e-6), (10, 2e-5), (60, 5e-4)])
def test_f32_error_growth_vs_reference_golden(steps, tol):
    g = np.load(os.path.join(GOLDEN, "hyp2d_ref_256x128.npz"))
    cfg = oracle.hyp2d_cfg(256, 128)
    p0 = [g[k] for k in ("rho0", "mx0", "my0", "E0")]
    ref, _, _ = oracle.hyp2d_run(cfg, p0, g["mask"], steps)
    out, m, _, _ = product_run(256, 128, steps, "f32")
    assert np.array_equal(m.ravel(), g["mask"])
    for k, a, b in zip(NAMES, out, ref):
        assert rel_linf(a, b) < tol, k


@pytest.mark.skipif(not oracle.has_ref("ref_hyp2d_1024x512"), reason="oracle/_ref not built")
def test_f64_vs_reference_kernels_1024x512():
    W, H, steps = 1024, 512, 100
    cfg11 = oracle.hyp2d_cfg(W, H).as11()
    ref, rmask, t_ref, dts_ref, _ = oracle.ref_hyp2d_run(W, H, cfg11, steps)
    out, m, t, dts = product_run(W, H, steps, "f64")
    assert np.array_equal(m.ravel(), rmask)
    for k, a, b in zip(NAMES, out, ref):
        assert rel_linf(a, b) < 1e-10, k
    assert abs(t - t_ref) < 1e-11
    # the reference's own regression snapshot (tests:143-176) at its own tolerances (:534-557)
    cfg = oracle.hyp2d_cfg(W, H)
    sa = oracle.hyp2d_snapshot(cfg, steps, [a.astype(np.float64) for a in out], m)
    sb = oracle.hyp2d_snapshot(cfg, steps, ref, rmask)
    assert sa[1] == sb[1]
    for i in (2, 3, 4, 5, 8, 9, 10, 11):
        assert abs(sa[i] - sb[i]) <= 5e-8 * abs(sb[i]) + 1e-8
    assert abs(sa[6] - sb[6]) <= 1e-9 and abs(sa[7] - sb[7]) <= 1e-9


@pytest.mark.skipif(not oracle.has_ref("ref_hyp2d_4096x4096"), reason="oracle/_ref not built")
def test_full_size_4096_vs_reference_kernels():
    """BASELINE config 2 at full size: 20 steps from k_init, fp64 handle vs reference kernels on
    every cell, and the fp32 handle against the same fields."""
    W = H = 4096
    steps = 20
    cfg11 = oracle.hyp2d_cfg(W, H).as11()
    ref, rmask, t_ref, _, _ = oracle.ref_hyp2d_run(W, H, cfg11, steps)
    out, m, t, _ = product_run(W, H, steps, "f64")
    assert np.array_equal(m.ravel(), rmask)
    for k, a, b in zip(NAMES, out, ref):
        assert rel_linf(a, b) < 1e-10, k
    assert abs(t - t_ref) < 1e-11
    out32, m32, t32, _ = product_run(W, H, steps, "f32")
    assert np.array_equal(m32.ravel(), rmask)
    for k, a, b in zip(NAMES, out32, ref):
        assert rel_linf(a, b) < 1e-4, k
    # size-independent properties: mass only enters/leaves through the x boundaries; positivity
    assert float(out32[0].min()) > 0 and np.isfinite(out32[3]).all()


@pytest.mark.skipif(not (oracle.has_ref("ref_hyp2d_4096x4096") and oracle.has_ref("ref_hyp2d_f32_4096x4096")),
                    reason="oracle/_ref not built")
def test_full_size_4096_developed_flow_vs_reference_kernels():
    """BASELINE config 2 on a DEVELOPED flow (the bow shock exists: the state bench.py times): the fp64 handle develops
    1500 steps from k_init; from that state the reference kernels (fp64), the float-typed reference and both product
    handles advance another 100 steps.  fp64 handle: the north-star's 1e-5; fp32 handle: at most twice as far from the
    fp64 reference as the float-typed reference is (per field, L-inf and L1 relative to the field's max)."""
    W = H = 4096
    dev, steps = 1500, 100
    cfg = SimConfig.default(W, H)
    cfg11 = oracle.hyp2d_cfg(W, H).as11()
    s = Hypersonic2D(cfg, dtype="f64").init()
    s.step(dev)
    start, mask = s.download()
    assert float(np.ptp(start[0])) > 5.0                      # a shock layer: rho jumps by more than a factor 6
    ref, rmask, t_ref, _, _ = oracle.ref_hyp2d_run(W, H, cfg11, steps, planes=start, mask=mask)
    s.step(steps)
    out, _ = s.download()
    s.close()
    e64 = {k: rel_linf(a, b) for k, a, b in zip(NAMES, out, ref)}
    r32, _, _, _, _ = oracle.ref_hyp2d_run(W, H, cfg11, steps, planes=start, mask=mask, f32=True)
    s32 = Hypersonic2D(cfg, dtype="f32").upload(start, mask)
    s32.step(steps)
    out32, _ = s32.download()
    s32.close()
    e32 = {k: rel_linf(a, b) for k, a, b in zip(NAMES, out32, ref)}
    eref = {k: rel_linf(a, b) for k, a, b in zip(NAMES, r32, ref)}
    l1 = lambda a, b: float(np.abs(np.asarray(a, np.float64).ravel() - b).mean() / max(1.0, np.abs(b).max()))   # noqa: E731
    l1_32 = {k: l1(a, b) for k, a, b in zip(NAMES, out32, ref)}
    l1ref = {k: l1(a, b) for k, a, b in zip(NAMES, r32, ref)}
    print(f"\n4096^2 developed flow +{steps} steps: f64 handle rel L-inf {e64}")
    print(f"  f32 handle rel L-inf {e32} rel L1 {l1_32}\n  float-typed reference rel L-inf {eref} rel L1 {l1ref}")
    for k in NAMES:
        assert e64[k] < 1e-5, ("f64", k, e64[k])
        assert e32[k] <= 2.0 * eref[k] + 1e-6, ("f32 L-inf vs float reference", k, e32[k], eref[k])
        assert l1_32[k] <= 2.0 * l1ref[k] + 1e-8, ("f32 L1 vs float reference", k, l1_32[k], l1ref[k])


@pytest.mark.skipif(not oracle.has_ref("ref_hyp2d_1024x512"), reason="oracle/_ref not built")
def test_1000_steps_vs_reference_kernels():
    """BASELINE.json north_star: 'per-field L-inf error < 1e-5 vs the reference after 1000 steps'.
    The reference is fp64; the fp64 handle is held to that bound (L-inf relative to the field's max).
    The fp32 handle (the benchmarked configuration) cannot meet it by construction — fp32 rounding
    alone is 6e-8 per operation and limiter/HLLC branches flip near the shock — so its 1000-step
    error is bounded separately and the measured value is printed (quoted in DESIGN.md)."""
    W, H, steps = 1024, 512, 1000
    cfg11 = oracle.hyp2d_cfg(W, H).as11()
    ref, rmask, t_ref, _, _ = oracle.ref_hyp2d_run(W, H, cfg11, steps)
    s = Hypersonic2D(SimConfig.default(W, H), dtype="f64").init()
    s.step(steps)
    out, m = s.download()
    t, _ = s.clock()
    s.close()
    assert np.array_equal(m.ravel(), rmask)
    e64 = {k: rel_linf(a, b) for k, a, b in zip(NAMES, out, ref)}
    s = Hypersonic2D(SimConfig.default(W, H), dtype="f32").init()
    s.step(steps)
    out32, _ = s.download()
    t32, _ = s.clock()
    s.close()
    e32 = {k: rel_linf(a, b) for k, a, b in zip(NAMES, out32, ref)}
    l1_32 = {k: float(np.abs(np.asarray(a, np.float64).ravel() - b).mean() / max(1.0, np.abs(b).max()))
             for k, a, b in zip(NAMES, out32, ref)}
    print(f"\n1000 steps {W}x{H}: f64 rel L-inf {e64}  |t-t_ref|={abs(t - t_ref):.3e}")
    print(f"1000 steps {W}x{H}: f32 rel L-inf {e32}  rel L1 {l1_32}  |t-t_ref|={abs(t32 - t_ref):.3e}")
    for k in NAMES:
        assert e64[k] < 1e-5, ("f64", k, e64[k])
    assert abs(t - t_ref) < 1e-9
    for k in NAMES:
        assert l1_32[k] < 1e-4, ("f32 L1", k, l1_32[k])
        assert e32[k] < 2e-3, ("f32 L-inf", k, e32[k])          # measured 4-5e-4 (a few cells at the bow shock)
    assert abs(t32 - t_ref) < 1e-3 * t_ref
    # The yardstick for fp32: the REFERENCE's own algorithm evaluated in fp32 (float-typed copy of the reference
    # translation unit, oracle/gen_f32_src.py) against the same fp64 reference run.  It separates "fp32 rounding on a
    # Mach-25 shock" from "what the product's fp32 path changes" (rcp.approx, closed-form limiter, merged HLLC guards,
    # fused passes): the product may not be further from the fp64 reference than twice the float reference itself.
    if oracle.has_ref(f"ref_hyp2d_f32_{W}x{H}"):
        r32, _, tr32, _, _ = oracle.ref_hyp2d_run(W, H, cfg11, steps, f32=True)
        eref = {k: rel_linf(a, b) for k, a, b in zip(NAMES, r32, ref)}
        l1ref = {k: float(np.abs(a - b).mean() / max(1.0, np.abs(b).max())) for k, a, b in zip(NAMES, r32, ref)}
        print(f"1000 steps {W}x{H}: float-typed REFERENCE rel L-inf {eref}  rel L1 {l1ref}  |t-t_ref|={abs(tr32 - t_ref):.3e}")
        for k in NAMES:
            assert e32[k] <= 2.0 * eref[k] + 1e-6, ("f32 L-inf vs float reference", k, e32[k], eref[k])
            assert l1_32[k] <= 2.0 * l1ref[k] + 1e-8, ("f32 L1 vs float reference", k, l1_32[k], l1ref[k])


def test_multi_step_call_equals_single_steps():
    a, _, ta, _ = product_run(256, 128, 16, "f32")
    cfg = SimConfig.default(256, 128)
    s = Hypersonic2D(cfg, dtype="f32").init()
    s.step(16)
    b, _ = s.download()
    tb, _ = s.clock()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert ta == tb


def test_errors_are_loud():
    from fluid_sims_b200 import TauError
    with pytest.raises(TauError, match="gamma"):
        Hypersonic2D(SimConfig.default(64, 64, gamma=0.9))
    with pytest.raises(TauError):
        Hypersonic2D(SimConfig.default(64, 64), y_begin=60, h_local=10)
    s = Hypersonic2D(SimConfig.default(64, 64))
    with pytest.raises(TauError, match="no state"):
        s.step(1)


@pytest.mark.skipif(os.environ.get("TAU_TEST_PAIR") != "1",
                    reason="experimental packed two-column kernel (hypersonic2d_pair.cuh): written after the "
                           "round-1 GPU budget was spent, not yet run on hardware; enable with TAU_TEST_PAIR=1")
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("W,H,steps", [(256, 128, 40), (1024, 512, 100)])
def test_pair_mode_matches_production_path(W, H, steps, mode, monkeypatch):
    """TAU_HYP2D_PAIR=1: interior body-free items go through hyp2d_step_pair, the rest through hyp2d_step;
    =2: both kinds of item through the one fused kernel (bit-identical to =1).
    Same expression trees with the FMA contractions spelled out, so the two paths agree to fp32
    round-off growth (bounds of test_f32_error_growth), and both against the fp64 oracle."""
    base, m, t0, _ = product_run(W, H, steps, "f32")
    monkeypatch.setenv("TAU_HYP2D_PAIR", str(mode))
    pair, m2, t1, _ = product_run(W, H, steps, "f32")
    assert np.array_equal(m, m2)
    if mode == 2:
        monkeypatch.setenv("TAU_HYP2D_PAIR", "1")
        two, _, t2, _ = product_run(W, H, steps, "f32")
        assert all(np.array_equal(a, b) for a, b in zip(pair, two)) and t1 == t2
    for k, a, b in zip(NAMES, pair, base):
        assert rel_linf(a, b) < 5e-4, k
    assert abs(t1 - t0) < 1e-5 * t0
    cfg = oracle.hyp2d_cfg(W, H)
    planes, mask = oracle.hyp2d_init(cfg)
    ref, _, _ = oracle.hyp2d_run(cfg, planes, mask, min(steps, 40))
    if steps <= 40:
        for k, a, b in zip(NAMES, pair, ref):
            assert rel_linf(a, b) < 5e-4, k
