"""GPU parity of the 2-D hypersonic render pass (tau_hyp2d_render*, C-ABI) against the CPU oracle
and the reference's own render kernels (oracle/_ref: k_render_vals, k_reduce_minmax,
k_compute_inv_range, k_render_pixels — tau_hypersonic_cuda.cu:1178-1326) on identical states.

Bar: min/max within 1e-12 relative (the value expressions are the reference's, in fp64; only FMA
contraction may differ); pixels identical except where `255 * f(t)` lands within that round-off of an
integer boundary of the (uint8_t) truncation — at most 1 LSB on at most 1e-4 of the pixels.
"""
import numpy as np
import pytest

import oracle
from fluid_sims_b200.hypersonic2d import Hypersonic2D, SimConfig

pytestmark = pytest.mark.gpu


def developed_state(W, H, steps, **over):
    cfg = oracle.hyp2d_cfg(W, H, **over)
    planes, mask = oracle.hyp2d_init(cfg)
    planes, _, _ = oracle.hyp2d_run(cfg, planes, mask, steps)
    return cfg, planes, mask


def check_pixels(got, want, what):
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert d.max() <= 1, (what, int(d.max()))
    assert (d > 0).mean() <= 1e-4, (what, float((d > 0).mean()))


@pytest.mark.parametrize("mode", range(7))
def test_render_matches_oracle_f64(mode):
    W, H = 200, 120
    cfg, planes, mask = developed_state(W, H, 40, geom_x0=60.0)
    want, _, (mn, mx) = oracle.hyp2d_render(cfg, planes, mask, mode)
    s = Hypersonic2D(SimConfig.default(W, H, geom_x0=60.0), dtype="f64").upload(planes, mask)
    got, (gmn, gmx) = s.render(mode)
    assert abs(gmn - mn) <= 1e-12 * max(1.0, abs(mn)) and abs(gmx - mx) <= 1e-12 * max(1.0, abs(mx))
    check_pixels(got, want, f"mode {mode}")
    assert np.array_equal(got[mask.reshape(H, W) != 0], np.full((int(mask.sum()), 4), (110, 110, 110, 255), np.uint8))
    # two-pass form with a caller-chosen range == the one-call form
    again = s.render_pixels(mode, *s.render_minmax(mode))
    assert np.array_equal(again, got)
    s.close()


def test_render_walls_on_every_boundary():
    """hand-made mask touching all four edges: every branch of sample_prim_bc (modes 3, 4)."""
    W, H = 128, 96
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:H, 0:W]
    rho = 1.0 + 0.3 * np.sin(xx / 9.0) * np.cos(yy / 7.0)
    u = 3.0 + 0.5 * np.cos(xx / 11.0)
    v = 0.7 * np.sin(yy / 5.0)
    p = 1.0 + 0.2 * np.cos((xx + yy) / 13.0)
    planes = [rho, rho * u, rho * v, p / 0.1 + 0.5 * rho * (u * u + v * v)]
    mask = np.zeros((H, W), np.uint8)
    mask[40:56, 50:70] = 1
    mask[0:3, 100:110] = 1
    mask[H - 2:, 20:30] = 1
    mask[60:64, 0:2] = 1
    mask[10:14, W - 3:] = 1
    mask[rng.random((H, W)) < 0.003] = 1
    cfg = oracle.hyp2d_cfg(W, H)
    s = Hypersonic2D(SimConfig.default(W, H), dtype="f64").upload(planes, mask)
    for mode in (3, 4, 0):
        want, _, (mn, mx) = oracle.hyp2d_render(cfg, planes, mask.ravel(), mode)
        got, (gmn, gmx) = s.render(mode)
        assert abs(gmn - mn) <= 1e-12 * max(1.0, abs(mn)) and abs(gmx - mx) <= 1e-12 * max(1.0, abs(mx))
        check_pixels(got, want, f"mode {mode}")
    s.close()


@pytest.mark.skipif(not oracle.has_ref("ref_hyp2d_1024x512"), reason="oracle/_ref not built")
@pytest.mark.parametrize("mode", [0, 3, 4, 5])
def test_render_vs_reference_kernels_1024x512(mode):
    W, H = 1024, 512
    s = Hypersonic2D(SimConfig.default(W, H), dtype="f64").init()
    s.step(300)
    planes, mask = s.download()
    got, (gmn, gmx) = s.render(mode)
    cfg11 = oracle.hyp2d_cfg(W, H).as11()
    want, _, (mn, mx) = oracle.ref_hyp2d_render(W, H, cfg11, planes, mask, mode)
    assert abs(gmn - mn) <= 1e-12 * max(1.0, abs(mn)) and abs(gmx - mx) <= 1e-12 * max(1.0, abs(mx))
    check_pixels(got, want, f"mode {mode}")
    # fp32 handle on the same state: the picture is the same up to float rounding of the state.
    # (Not for the derivative views 3/4: in the uniform free stream |grad rho| is exactly 0 in fp64
    # and ~1e-7 of rounding noise in fp32, which log(1e-12 + .) turns into a different colour.)
    if mode in (0, 5):
        s32 = Hypersonic2D(SimConfig.default(W, H), dtype="f32").upload(planes, mask)
        got32, _ = s32.render(mode)
        d = np.abs(got32.astype(np.int16) - want.astype(np.int16))
        assert np.percentile(d, 99.9) <= 2
        s32.close()
    s.close()


def test_render_slabs_equal_single_domain():
    """two slab handles on one GPU with hand-exchanged ghost rows + reduced extrema == one handle"""
    W, H = 160, 96
    cfg, planes, mask = developed_state(W, H, 25, geom_x0=50.0)
    full = Hypersonic2D(SimConfig.default(W, H, geom_x0=50.0), dtype="f64").upload(planes, mask)
    want, (mn, mx) = full.render(3)
    import torch
    from fluid_sims_b200 import slab
    from fluid_sims_b200.hypersonic2d import HALO
    P = [np.asarray(p).reshape(H, W) for p in planes]
    M = np.asarray(mask).reshape(H, W)
    halves, mms = [], []
    for y0, hl in ((0, 40), (40, 56)):
        s = Hypersonic2D(SimConfig.default(W, H, geom_x0=50.0), dtype="f64", y_begin=y0, h_local=hl)
        s.upload([p[y0:y0 + hl] for p in P], M[y0:y0 + hl])
        halves.append(s)
    # ghost rows: copy the neighbour's boundary rows (what the slab exchange does)
    views = []
    for s in halves:
        pp, mp, _ = s.device_state()
        views.append((slab.wrap_plane(pp, (4, s.h_local + 2 * HALO, W), torch.float64, 0),
                      slab.wrap_plane(mp, (s.h_local + 2 * HALO, W), torch.uint8, 0)))
    (pa, ma), (pb, mb) = views
    pa[:, -HALO:, :] = pb[:, HALO:2 * HALO, :]
    ma[-HALO:, :] = mb[HALO:2 * HALO, :]
    pb[:, :HALO, :] = pa[:, -2 * HALO:-HALO, :]
    mb[:HALO, :] = ma[-2 * HALO:-HALO, :]
    torch.cuda.synchronize()
    for s in halves:
        mms.append(s.render_minmax(3))
    lo, hi = min(m[0] for m in mms), max(m[1] for m in mms)
    assert (lo, hi) == (mn, mx)
    got = np.concatenate([s.render_pixels(3, lo, hi) for s in halves], axis=0)
    assert np.array_equal(got, want)
    for s in halves + [full]:
        s.close()
