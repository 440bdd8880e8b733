"""GPU parity of the tau_hypersonic.c update path (BASELINE config 1, C-ABI tau_hypc_*) against the reference's own
object code (oracle/_ref/libref_hypcpu_256x256.so: tau_hypersonic.c compiled with its `gcc -O3`), the committed
golden fixture made from it and the plain-C oracle.  Everything is IEEE fp64 add / mul / div / sqrt / min / max in
the reference's order and the device code is built with --fmad=false, so the bar is 0 ulp: np.array_equal."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from fluid_sims_b200.hypersonic_c import HypersonicC, init_sim

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_config1_golden_10_plus_200_steps_bit_exact():
    """W = H = 256, init_sim, 10 warm-up + 200 steps; all four fields, mask and sim_t."""
    g = np.load(os.path.join(GOLDEN, "hypcpu_ref_256x256.npz"))
    s = HypersonicC(256, 256).init()
    p0, m = s.download()
    assert np.array_equal(m.ravel(), g["mask"]) and np.array_equal(p0[0].ravel(), g["rho0"]) and np.array_equal(p0[3].ravel(), g["E0"])
    s.step(10)
    assert s.clock()[0] == g["sim_t"][0]
    s.step(200)
    out, _ = s.download()
    assert s.clock()[0] == g["sim_t"][1]
    for a, k in zip(out, ("rho", "mx", "my", "E")):
        assert np.array_equal(a.ravel(), g[k]), k
    assert s.launch_count == 1 + 2 * 210          # one wavespeed scan, then two kernels per step
    s.close()


@pytest.mark.skipif(not oracle.has_ref("ref_hypcpu_256x256"), reason="oracle/_ref not built")
def test_perturbed_state_vs_compiled_reference_bit_exact():
    """a second body, a slow pocket and a near-vacuum pocket: slip-wall ghosts on every side, subsonic HLLC
    branches, the positivity fix and the pressure repair — against the compiled reference itself"""
    W = H = 256
    r = oracle.RefHypCpu(W, H)
    r.init()
    _, mask = r.get()
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:H, 0:W]
    mask = mask.reshape(H, W).copy()
    mask[(xx - 170) ** 2 + (yy - 60) ** 2 < 15 ** 2] = 1
    mask[200:230, 120:124] = 1
    mask[100:103, 0:2] = 1                                  # a body cell in column 0 (no inflow overwrite there)
    rho = 1.0 + 0.3 * rng.random((H, W))
    u = np.where((xx > 100) & (xx < 140) & (yy > 150), 0.3, 17.0) + rng.normal(0, 0.2, (H, W))
    v = rng.normal(0, 0.5, (H, W))
    pr = np.where((xx - 60) ** 2 + (yy - 200) ** 2 < 100, 1e-9, 1.0 + 0.2 * rng.random((H, W)))
    u[mask == 1] = 0
    v[mask == 1] = 0
    planes = [rho, rho * u, rho * v, pr / 0.4 + 0.5 * rho * (u * u + v * v)]
    r.lib.ref_hypcpu_set(*[np.ascontiguousarray(p).ravel() for p in planes], mask.ravel())
    t0 = r.sim_t
    s = HypersonicC(W, H).upload(planes, mask, sim_t=t0)
    for n in (1, 11):
        r.steps(n)
        s.step(n)
        rp, _ = r.get()
        out, m = s.download()
        assert np.array_equal(m, mask) and s.clock()[0] == r.sim_t
        for a, b in zip(out, rp):
            assert np.array_equal(a.ravel(), b)
    s.close()


@pytest.mark.parametrize("W,H,steps", [(64, 48, 60), (301, 77, 40), (33, 130, 25)])
def test_other_extents_vs_oracle_bit_exact(W, H, steps):
    """the reference fixes W, H at compile time; the restatement (pinned 0 ulp on it at 256^2) checks ragged extents"""
    planes, mask = init_sim(W, H)
    op, om = oracle.hypcpu_init(W, H)
    assert np.array_equal(mask.ravel(), om) and all(np.array_equal(a.ravel(), b) for a, b in zip(planes, op))
    s = HypersonicC(W, H).init()
    s.step(steps)
    out, _ = s.download()
    exp, t, dts = oracle.hypcpu_run(W, H, op, om, steps)
    assert s.clock() == (t, dts[-1])
    for a, b in zip(out, exp):
        assert np.array_equal(a.ravel(), b)
    s.close()


def test_speed_mode_render_equals_oracle():
    """view_mode 2 ("speed mode") is sqrt and IEEE arithmetic only: pixels and min/max bit-exact; the log views may
    differ by the device's log (<= 1 ulp) -> at most 1 LSB on a handful of pixels"""
    W, H = 256, 256
    s = HypersonicC(W, H).init()
    s.step(120)
    planes, mask = s.download()
    rgba, mm = s.render("speed")
    ergba, emm, _ = oracle.hypcpu_render(W, H, planes, mask, 2)
    assert mm == emm and np.array_equal(rgba, ergba)
    for mode in (0, 1, 3):
        rgba, mm = s.render(mode)
        ergba, emm, _ = oracle.hypcpu_render(W, H, planes, mask, mode)
        assert np.allclose(mm, emm, rtol=1e-14, atol=1e-14)
        d = np.abs(rgba.astype(int) - ergba.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3
    s.close()


def test_cli_tau_hypersonic_config1_dump(tmp_path):
    """the host binary keeps the reference's name; 256 x 256, speed mode, 210 steps -> dump == golden"""
    exe = os.path.join(ROOT, "fluid_sims_b200", "cli", "tau_hypersonic")
    dump, ppm = tmp_path / "c1.dump", tmp_path / "c1.ppm"
    r = subprocess.run([exe, "--nx", "256", "--ny", "256", "--steps", "210", "--speed-mode", "--ppm", str(ppm),
                        "--dump", str(dump)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "speed" in r.stdout and "updates/s" in r.stdout
    raw = dump.read_bytes()
    assert raw[:8] == b"TAUDUMP1"
    hdr = np.frombuffer(raw, np.int32, 5, 8)
    assert list(hdr) == [4, 8, 256, 256, 1]
    step = int(np.frombuffer(raw, np.int64, 1, 28)[0])
    t = float(np.frombuffer(raw, np.float64, 1, 36)[0])
    planes = np.frombuffer(raw, np.float64, 4 * 256 * 256, 44).reshape(4, -1)
    g = np.load(os.path.join(GOLDEN, "hypcpu_ref_256x256.npz"))
    assert step == 210 and t == g["sim_t"][1]
    for a, k in zip(planes, ("rho", "mx", "my", "E")):
        assert np.array_equal(a, g[k]), k
    assert ppm.read_bytes().startswith(b"P6\n256 256\n255\n")


def test_errors_are_loud():
    from fluid_sims_b200 import TauError
    with pytest.raises(TauError, match="bad grid"):
        HypersonicC(1, 5)
    with pytest.raises(TauError, match="no state"):
        HypersonicC(32, 32).step(1)
