"""The C host program of the 2-D hypersonic solver (reference binary name tau_2d_hypersonic_cuda)
on a GPU: its dump equals the Python-API run bit for bit, its PPM frame equals the library's render
pass, the regression baseline it writes verifies in a second process, and --checkpoint/--resume
continues a run bit-identically."""
import os
import struct
import subprocess

import numpy as np
import pytest

from fluid_sims_b200.hypersonic2d import Hypersonic2D, SimConfig

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "fluid_sims_b200", "cli", "tau_2d_hypersonic_cuda")


def read_dump(path):
    """TAUDUMP1 (fluid_sims_b200/cli/cli_common.h)."""
    with open(path, "rb") as f:
        assert f.read(8) == b"TAUDUMP1"
        nplanes, es, d0, d1, d2 = struct.unpack("<5i", f.read(20))
        step, t = struct.unpack("<qd", f.read(16))
        dt = np.float64 if es == 8 else np.float32
        planes = [np.frombuffer(f.read(es * d0 * d1 * d2), dt).reshape(d1, d0) for _ in range(nplanes)]
    return planes, step, t


def run(*args):
    r = subprocess.run([BIN, *map(str, args)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


@pytest.mark.skipif(not os.path.exists(BIN), reason="CLI not built")
def test_cli_dump_render_baseline_resume(tmp_path):
    W, H, spf = 256, 128, 2
    common = ["--nx", W, "--ny", H, "--dtype", "f64", "--mach", 12, "--geom-x0", 60]
    d10, ppm, base, ck = (str(tmp_path / n) for n in ("d10.bin", "f.ppm", "base.txt", "run.ckpt"))
    out = run(*common, "--frames", 5, "--dump", d10, "--view", 3, "--ppm", ppm, "--write-baseline", base,
              "--checkpoint", ck)
    assert "SimConfig:" in out and "inflow_mach=12" in out
    s = Hypersonic2D(SimConfig.default(W, H, inflow_mach=12.0, geom_x0=60.0), dtype="f64").init()
    s.step(5 * spf)
    want, _ = s.download()
    got, step, t = read_dump(d10)
    assert step == 10 and t == s.clock()[0]
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    # PPM == render pass
    rgba, _ = s.render(3)
    raw = open(ppm, "rb").read()
    hdr = f"P6\n{W} {H}\n255\n".encode()
    assert raw.startswith(hdr)
    assert np.array_equal(np.frombuffer(raw[len(hdr):], np.uint8).reshape(H, W, 3), rgba[..., :3])
    # baseline written by one process verifies in another; a different run fails the check loudly
    assert "verified" in run(*common, "--frames", 5, "--verify-baseline", base)
    r = subprocess.run([BIN, *map(str, common), "--frames", "4", "--verify-baseline", base],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "FAIL: steps match baseline" in r.stderr
    # resume: 5 frames + checkpoint + 3 frames == 8 frames
    d_res, d_full = str(tmp_path / "res.bin"), str(tmp_path / "full.bin")
    run(*common, "--resume", ck, "--frames", 3, "--dump", d_res)
    run(*common, "--frames", 8, "--dump", d_full)
    a, sa, ta = read_dump(d_res)
    b, sb, tb = read_dump(d_full)
    assert (sa, ta) == (sb, tb) == (16, tb)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    s.close()


@pytest.mark.skipif(not os.path.exists(BIN), reason="CLI not built")
def test_cli_rejects_bad_flags_like_the_reference():
    r = subprocess.run([BIN, "--gamma", "0.9"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Invalid --gamma" in r.stderr  # parse_args :1545
    r = subprocess.run([BIN, "--bogus"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Unknown or incomplete argument" in r.stderr


def test_tau_burgers_cli_matches_api_and_cole_hopf(tmp_path):
    exe = os.path.join(ROOT, "fluid_sims_b200", "cli", "tau_burgers")
    if not os.path.exists(exe):
        pytest.skip("CLI not built")
    from fluid_sims_b200.burgers import Burgers, Params
    dump = str(tmp_path / "b.bin")
    r = subprocess.run([exe, "--nx", "160", "--ny", "96", "--steps", "12", "--dtau", "1e-3", "--muscl",
                        "--visc_substeps", "2", "--headless", "--dump", dump], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "Headless (stride=5):" in r.stdout and "Steps: 12" in r.stdout, r.stdout + r.stderr
    got, step, t = read_dump(dump)
    s = Burgers(Params(nx=160, ny=96, dtau=1e-3, muscl=1, visc_substeps=2)).init()
    s.step(12)
    u, v = s.download()
    assert step == 12 and np.array_equal(got[0], u) and np.array_equal(got[1], v)
    assert t == np.float32(s.clock()[0])
    s.close()
    # the reference's validation mode prints the relative L2 error against the exact solution
    r = subprocess.run([exe, "--nx", "256", "--colehopf", "--nu", "0.5", "--dtau", "5e-3", "--t0", "1e-3", "--ck", "2",
                        "--steps", "2200", "--stride", "100"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    last = [l for l in r.stdout.splitlines() if "relL2=" in l][-1]
    assert float(last.split("relL2=")[1]) < 2.5e-3
