"""GPU parity of the SPH update path (C-ABI) against the reference's own kernels
(oracle/_ref/libref_sph.so, same device), the CPU oracle and numpy's stable sort.

Integer work is bit-exact: cell keys and the radix sort (== stable argsort).  Floating point: the
reference sums neighbours in linked-list order, which is an atomicExch race (tau_sph.cu:175) and
differs run to run; the product sums in sorted order across 8 cooperating lanes, so only tolerance
parity is meaningful (hard part 5).  Bounds are ~5x what fp32 reordering produces."""
import numpy as np
import pytest

import oracle
from fluid_sims_b200.sph import SPH, Params, reset_particles

pytestmark = pytest.mark.gpu


def product(P, pos0, vel0, frames):
    s = SPH(P).upload(pos0, vel0)
    s.step(frames)
    out = s.download()
    ck = s.clock()
    srt = s.download_sort()
    s.close()
    return out, ck, srt


@pytest.mark.parametrize("N", [1000, 4096, 65536, 300000])
def test_radix_sort_is_bit_exact(N):
    s = SPH(Params(N=N))
    g = s.grid()
    M = g["Gx"] * g["Gy"]
    rng = np.random.default_rng(N)
    for keys in (rng.integers(0, M, N), np.zeros(N), np.full(N, M - 1), np.arange(N)[::-1] % M,
                 rng.integers(0, 3, N)):
        keys = keys.astype(np.uint32)
        ko, vo = s.sort_pairs(keys)
        order = np.argsort(keys, kind="stable").astype(np.uint32)
        assert np.array_equal(vo, order) and np.array_equal(ko, keys[order])


def test_cell_keys_and_sort_match_oracle():
    """The sort key is (grid row, key column) with `subx` key columns per grid cell: key column // subx and the row must be
    the reference's grid_x / grid_y (:141-157, the oracle's cell key) for every particle, and the permutation the stable
    sort of the keys (ascending particle index inside a key)."""
    N = 50000
    P = Params(N=N, rain=0)
    pos0, vel0 = reset_particles(P)
    s = SPH(P).upload(pos0, vel0)
    g = s.grid()
    s.step(1)
    k, v = s.download_sort()
    s.close()
    ek, ev, _ = oracle.sph_cell_sort(oracle.sph_params(N, rain=0), pos0)
    Gx, subx = g["Gx"], g["subx"]
    cell_of = np.empty(N, np.int64)
    cell_of[ev] = ek                                   # the reference's cell of every particle
    k64 = k.astype(np.int64)
    coarse = (k64 // (Gx * subx)) * Gx + (k64 % (Gx * subx)) // subx
    assert np.array_equal(coarse, cell_of[v])
    key_of = np.empty(N, np.int64)
    key_of[v] = k64
    assert np.array_equal(v, np.argsort(key_of, kind="stable").astype(np.uint32)) and np.all(np.diff(k64) >= 0)
    if subx == 1:
        assert np.array_equal(k, ek) and np.array_equal(v, ev)


def _outliers(a, b, tol):
    d = np.abs(a - b)
    d = d.max(axis=1) if d.ndim == 2 else d
    return float((d > tol).mean()), float(d.max())


# 2^21 (BASELINE config 5) first: a failure further down must not hide it.  65536 is the reference's default N (:51).
_REF_CASES = [(1 << 21, 6, {}), (1 << 21, 4, dict(rain=0)), (65536, 30, {}), (65536, 30, dict(rain=0)),
              (65536, 12, dict(useXSPH=1)), (20000, 25, dict(viscSub=3, rain=0)),
              (65536, 12, dict(useVisc=0)), (65536, 12, dict(gammaEOS=2.0, c0=2.0, useGrav=0))]


@pytest.mark.skipif(not oracle.has_ref("ref_sph"), reason="oracle/_ref not built")
@pytest.mark.parametrize("N,frames,kw", _REF_CASES)
def test_vs_reference_kernels(N, frames, kw):
    """Against the reference's own kernels on the same device.  They are racy (atomicExch list order :175 -> the
    summation order differs run to run; k_rain collisions :377-392) and a particle that crosses a wall by one ulp in
    one run and not in the other gets v -> -0.2 v in one of them (k_integrate :338-353): an O(1) velocity difference
    from an O(1e-7) cause, which then spreads to its neighbours.  Measured on a B200 (profiles/r2_sph_parity.md):
    the product is bit-identical run to run and always within 3e-5 / 0.02 % of the sorted-order CPU oracle, while the
    reference lands between 0.015 % and 1 % (!) of the particles away from the oracle — and from itself — for the
    same input.  So the reference runs TWICE and the outlier budget is its own measured self-scatter plus a fixed
    0.2 %; the tight, deterministic gate is test_vs_cpu_oracle below."""
    P = Params(N=N, **kw)
    op = oracle.sph_params(N, **kw)
    pos0, vel0 = reset_particles(P)
    rp, rv = oracle.ref_sph_reset_particles(op)
    assert np.array_equal(pos0, rp) and np.array_equal(vel0, rv)
    (pos, vel, s, pr), ck, _ = product(P, pos0, vel0, frames)
    r = oracle.ref_sph_run(op, pos0, vel0, frames)
    r2 = oracle.ref_sph_run(op, pos0, vel0, frames)
    vtol, ptol = 5e-4 * max(1.0, np.abs(r[1]).max()), 1e-3 * max(1.0, np.abs(r[4]).max())
    for mine, k, tol, cap in ((pos, 0, 2e-5, 5e-3), (vel, 1, vtol, None), (s, 3, 5e-4, None), (pr, 4, ptol, None)):
        self_frac, self_max = _outliers(r[k], r2[k], tol)
        frac, dmax = _outliers(mine, r[k], tol)
        frac2, dmax2 = _outliers(mine, r2[k], tol)
        assert min(frac, frac2) <= 2e-3 + 2 * self_frac, (k, frac, frac2, self_frac)
        if cap is not None:
            assert min(dmax, dmax2) <= cap + 2 * self_max, (k, dmax, dmax2, self_max)
    assert ck[0] == pytest.approx(float(r[5][0]), rel=1e-6) and ck[1] == pytest.approx(float(r[5][1]), rel=1e-6)
    assert ck[2] == int(r[5][3])


@pytest.mark.parametrize("N,frames,kw,fmax,dcap", [
    (8192, 15, {}, 1e-3, 5e-5), (20000, 25, dict(viscSub=3, rain=0), 1.5e-3, 5e-4),
    (20000, 25, {}, 6e-3, 2e-3),          # dt three times as long per sub-step: measured 0.16 % / 2.2e-4
    (65536, 30, {}, 6e-3, 2e-3), (30000, 12, dict(useXSPH=1), 1.5e-3, 5e-4),
    (30000, 12, dict(gammaEOS=2.0, c0=2.0, useGrav=0), 1.5e-3, 5e-4)])
def test_vs_cpu_oracle(N, frames, kw, fmax, dcap):
    """The deterministic gate: the CPU oracle sums neighbours in the same sorted-slot order as the product (and
    resolves k_rain's write race the same way), so only the fast intrinsics (GPU) vs libm (CPU) and the 8-lane
    partial sums separate them.  Measured: max |dx| 3.0e-5 after 75 sub-steps with rain and wall bounces."""
    P = Params(N=N, **kw)
    pos0, vel0 = reset_particles(P)
    (pos, vel, s, pr), ck, _ = product(P, pos0, vel0, frames)
    o = oracle.sph_run(oracle.sph_params(N, **kw), pos0, vel0, frames)
    frac, dmax = _outliers(pos, o[0], 2e-5)
    vf, vmax = _outliers(vel, o[1], 5e-4 * max(1.0, np.abs(o[1]).max()))
    sf, smax = _outliers(s, o[3], 5e-4)
    print(f"\nsph vs CPU oracle N={N} x{frames} {kw}: pos outliers {frac:.2e} max {dmax:.2e}; vel {vf:.2e} / {vmax:.2e}; s {sf:.2e} / {smax:.2e}")
    assert frac <= fmax and dmax <= dcap, (frac, dmax)
    assert vf <= 2 * fmax and sf <= 2 * fmax, (vf, sf)
    assert ck[0] == pytest.approx(o[5].t, rel=1e-6) and ck[2] == o[5].step

# ---- frame by frame from a common state -----------------------------------------------------------------------------------
# 20 000 particles, viscSub = 3, rain on, 25 frames is the configuration that bifurcates: somewhere in its 75 sub-steps one
# particle crosses a wall (v -> -0.2 v, k_integrate :338-353) or not depending on the last bit of its acceleration, and ~1 % of
# the particles end up 3.5e-3 away.  The reference's own kernels do that to themselves from run to run (profiles/r2_sph_parity.md:
# 0.995 % / 3.5e-3 in one run of three) and the product did it against the CPU oracle when the summation order changed with the
# sort key (SUBX key columns: 1.005 % / 3.49e-3, every other configuration unchanged).  A free-running comparison of that
# configuration measures the bifurcation, not the arithmetic.  Here every frame starts from the SAME state on both sides (the
# checker's trajectory) and is compared on its own: an error in the arithmetic shows in every frame, a wall flip as a
# handful of particles in one frame.
def _frame_by_frame(P, op, frames, advance):
    """advance(pos, vel, clock) -> (pos, vel, acc, s, press, clock): one frame of the checker"""
    pos, vel = reset_particles(P)
    prod = SPH(P)
    clock = None
    worst = [0.0, 0.0, 0.0]   # outlier fraction (pos), max |dx|, outlier fraction (s)
    flips = 0
    for f in range(frames):
        prod.upload(pos, vel)                    # the product's host clock advances on its own, identically
        prod.step(1)
        mp, mv, ms, _ = prod.download()
        o = advance(pos, vel, clock)
        clock = o[5]
        frac, dmax = _outliers(mp, o[0], 2e-6)
        sf, _ = _outliers(ms, o[3], 2e-4)
        flips += int(round(frac * P.N))
        worst = [max(worst[0], frac), max(worst[1], dmax), max(worst[2], sf)]
        pos, vel = o[0], o[1]
    prod.close()
    return worst, flips


@pytest.mark.parametrize("N,frames,kw", [(20000, 25, dict(viscSub=3)), (20000, 25, dict(viscSub=3, rain=0))])
def test_vs_cpu_oracle_frame_by_frame(N, frames, kw):
    P, op = Params(N=N, **kw), oracle.sph_params(N, **kw)
    (frac, dmax, sf), flips = _frame_by_frame(P, op, frames, lambda x, v, ck: oracle.sph_run(op, x, v, 1, clock=ck))
    print(f"\nsph frame by frame vs CPU oracle N={N} x{frames} {kw}: worst frame {frac:.2e} of the particles beyond 2e-6, "
          f"max |dx| {dmax:.2e}, s beyond 2e-4: {sf:.2e}; {flips} particle-frames beyond 2e-6 in all")
    assert frac <= 5e-4 and sf <= 5e-4 and flips <= 2e-4 * N * frames, (frac, dmax, sf, flips)


@pytest.mark.skipif(not oracle.has_ref("ref_sph"), reason="oracle/_ref not built")
def test_vs_reference_kernels_frame_by_frame():
    N, frames, kw = 20000, 25, dict(viscSub=3)
    P, op = Params(N=N, **kw), oracle.sph_params(N, **kw)

    def advance(x, v, ck):
        r = oracle.ref_sph_run(op, x, v, 1, clock=ck)
        return r[0], r[1], r[2], r[3], r[4], r[5]
    (frac, dmax, sf), flips = _frame_by_frame(P, op, frames, advance)
    print(f"\nsph frame by frame vs reference kernels N={N} x{frames} {kw}: worst frame {frac:.2e} beyond 2e-6, max |dx| {dmax:.2e}, "
          f"s beyond 2e-4: {sf:.2e}; {flips} particle-frames in all")
    # (the reference's rain kernel loses colliding writes at random, :389-391: a few particles per frame)
    assert frac <= 2e-3 and sf <= 2e-3 and flips <= 1e-3 * N * frames, (frac, dmax, sf, flips)


def test_deterministic_run_to_run():
    """The reference is not (list order, rain collisions); the product is."""
    P = Params(N=100000)
    pos0, vel0 = reset_particles(P)
    a, _, _ = product(P, pos0, vel0, 12)
    b, _, _ = product(P, pos0, vel0, 12)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_errors_are_loud():
    from fluid_sims_b200 import TauError
    with pytest.raises(TauError):
        SPH(Params(N=0))
    s = SPH(Params(N=1024))
    with pytest.raises(TauError, match="no state"):
        s.step(1)


@pytest.mark.skipif(not oracle.has_ref("ref_sph"), reason="oracle/_ref not built")
@pytest.mark.parametrize("N,W,H", [(65536, 120, 40), (1 << 21, 237, 63), (1000, 7, 3)])
def test_rasterize_equals_reference_kernel(N, W, H):
    """Render pass (k_clear_grid + k_rasterize, tau_sph.cu:357-374): integer counts, bit-exact against
    the reference kernel; the plain-C oracle (no fast-math division) may move a particle that sits
    within an ulp of a raster line."""
    P = Params(N=N, rain=0)
    pos0, vel0 = reset_particles(P)
    s = SPH(P).upload(pos0, vel0)
    s.step(3)
    pos, _, _, _ = s.download()
    g = s.rasterize(W, H)
    assert g.shape == (2 * H, W) and int(g.sum()) == N and g.min() >= 0
    assert np.array_equal(g, oracle.ref_sph_rasterize(pos, W, H))
    assert np.abs(g - oracle.sph_rasterize(pos, W, H)).sum() <= max(2, N // 5000)
    s.close()
