"""GPU side of the `.4spl` export: tau_hyp3d_export_frame and the th3cs host program.

First hardware run in round 2 (profiles/r2_first_hw_run.md): all tests passed; the frame indices are integer
work and bit-identical to the reference's host loop (also in the CPU emulator, tests/test_hostemu_cpu.py)."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from fluid_sims_b200 import splat4
from fluid_sims_b200.hypersonic3d import Hypersonic3D, Params

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _require_healthy_cuda_context():
    """The reference drivers exit() on a CUDA error (the reference's CUDA_CHECK policy).  If an earlier failure of
    never-run code left a sticky error in this process, skip instead of letting them end the whole test run."""
    from fluid_sims_b200.gray_scott import GrayScott, Params as GsParams
    try:
        g = GrayScott(GsParams(nx=32, ny=32)).init()
        g.step(1)
        g.download()
        g.close()
    except Exception as e:      # noqa: BLE001
        pytest.skip(f"CUDA context unusable after an earlier failure: {e}")


def test_export_frame_equals_the_reference_host_loop():
    for n in (64, 40):                       # 40: dx = 1/40 is not a power of two (mode 8 != mode 0 in the last bit)
        s = Hypersonic3D(Params.default(n, n, n)).init()
        s.step(120)
        sch = s.vis(8)
        idx, mm = s.export_frame()
        want, want_mm = oracle.splat4_frame_indices(sch)
        assert np.array_equal(idx, want) and mm == want_mm
        assert len(np.unique(idx)) > 8     # 120 steps: the shock layer is forming (14 levels in the emulator)
        ref = oracle.hyp3d_vis(oracle.hyp3d_params(n, n, n), s.download()[0], s.download()[1], 8)
        assert np.abs(sch - ref).max() <= 2e-4 * np.abs(ref).max()     # the bound tests/test_hyp3d_gpu.py holds on hardware
        s.close()


@pytest.mark.skipif(not oracle.has_ref("ref_hyp3d"), reason="oracle/_ref not built")
def test_schlieren_mode_8_vs_reference_k_vis_on_the_exporters_grid():
    """64^3: dx = 1/64, so th3cs.cu's k_schlieren_export equals k_vis mode 0 of the solver it was cloned from"""
    s = Hypersonic3D(Params.default(64, 64, 64)).init()
    s.step(100)
    planes, _ = s.download()
    _require_healthy_cuda_context()
    ref = oracle.ref_hyp3d_vis(oracle.hyp3d_params(64, 64, 64), planes, 0)
    assert np.abs(s.vis(8) - ref).max() <= 2e-5 * np.abs(ref).max()
    s.close()


def test_th3cs_cli_writes_a_file_the_reference_viewer_parses(tmp_path):
    exe = os.path.join(ROOT, "fluid_sims_b200", "cli", "th3cs")
    subprocess.run(["make", "-C", os.path.dirname(exe)], check=True, capture_output=True)
    out = str(tmp_path / "v.4spl")
    r = subprocess.run([exe, "--n", "32", "--frames", "5", "--out", out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Frame 5/5 processed" in r.stdout and "Export Complete!" in r.stdout
    assert splat4.info(out) == dict(width=32, height=32, depth=32, frames=5, pSize=256, flags=4)
    d = splat4.parse(open(out, "rb").read())
    # the same run through the API
    s = Hypersonic3D(Params.default(32, 32, 32)).init()
    for f in range(5):
        s.step(4)
        idx, _ = s.export_frame()
        assert np.array_equal(d["indices"][f], idx), f
    s.close()
