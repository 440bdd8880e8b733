"""The C-ABI library loads without a GPU and exports every symbol include/tau_b200.h declares; the
product has no CPU fallback (creating a handle without a device fails loudly)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tau_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tau_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from fluid_sims_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(_lib.lib, s)]
    assert not missing, f"declared in tau_b200.h but not exported: {missing}"
    assert _lib.lib.tau_abi_version() == 1


def test_product_does_not_link_or_import_the_oracle():
    import subprocess
    from fluid_sims_b200 import LIB_PATH
    out = subprocess.run(["ldd", LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libref" not in out
    pkg = os.path.join(ROOT, "fluid_sims_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "liboracle" not in src, f


def test_no_cpu_fallback():
    from fluid_sims_b200 import TauError, device_count
    if device_count() > 0:
        pytest.skip("a GPU is present")
    from fluid_sims_b200.gray_scott import GrayScott, Params
    from fluid_sims_b200.hypersonic2d import Hypersonic2D, SimConfig
    with pytest.raises(TauError) as e:
        GrayScott(Params(nx=64, ny=64))
    assert e.value.code == -19
    with pytest.raises(TauError) as e:
        Hypersonic2D(SimConfig.default(64, 64))
    assert e.value.code == -19


def test_host_side_helpers_without_gpu():
    import numpy as np

    import oracle
    from fluid_sims_b200 import TauError
    from fluid_sims_b200.gray_scott import init_pattern
    from fluid_sims_b200.hypersonic2d import SimConfig
    u, v = init_pattern(96, 64, 1337)
    eu, ev = oracle.gs_init_pattern(96, 64, 1337)
    assert np.array_equal(u, eu) and np.array_equal(v, ev)
    c = SimConfig.default(8192, 1024)           # default_config() tau_hypersonic_cuda.cu:1394-1409
    assert (c.gamma, c.cfl, c.inflow_mach, c.geom_x0) == (1.1, 0.25, 25.0, 125.0)
    assert c.geom_cy == 512.0 and abs(c.geom_Rb - 1024 / 12) < 1e-12 and c.steps_per_frame == 2
    c.validate()
    for bad, pat in ((dict(gamma=1.0), "gamma"), (dict(cfl=0.0), "cfl"), (dict(visc_nu=-1.0), "visc-nu"),
                     (dict(inflow_mach=0.0), "mach"), (dict(geom_Rb=1.0), "geom-rb")):
        with pytest.raises(TauError, match=pat):
            SimConfig.default(8192, 1024, **bad).validate()


def test_cli_hosts_build_and_fail_loudly_without_gpu():
    """The C host programs link against the C-ABI only; without a GPU they exit(1) with the
    library's message (the reference's CK() policy), with one they run a tiny headless case."""
    import subprocess
    from fluid_sims_b200 import device_count
    cli = os.path.join(ROOT, "fluid_sims_b200", "cli")
    subprocess.run(["make", "-C", cli], check=True, capture_output=True)
    cases = {"tgs": ["--nx", "64", "--ny", "32", "--steps", "3", "--headless"],
             "tau_2d_hypersonic_cuda": ["--nx", "128", "--ny", "64", "--frames", "2"],
             "tau3d": ["--n", "16", "--frames", "1"],
             "tau_sph": ["--n", "2048", "--frames", "2"],
             "tau_burgers": ["--nx", "96", "--ny", "64", "--steps", "4", "--headless", "--dtau", "1e-3"],
             "tau_sw": ["--nx", "96", "--ny", "64", "--steps", "4", "--headless", "--dtau", "1e-3"],
             "th3cs": ["--n", "16", "--frames", "1", "--out", os.devnull],
             "tau_hypersonic": ["--nx", "64", "--ny", "48", "--frames", "3", "--speed-mode"]}
    for exe, args in cases.items():
        r = subprocess.run([os.path.join(cli, exe)] + args, capture_output=True, text=True, timeout=120)
        if device_count() > 0:
            assert r.returncode == 0, (exe, r.stderr)
            assert "updates/s" in r.stdout
        else:
            assert r.returncode == 1 and "no CUDA device" in r.stderr, (exe, r.stderr)
    # flag validation happens before any device work: reference message, exit code 1
    r = subprocess.run([os.path.join(cli, "tau_2d_hypersonic_cuda"), "--gamma", "0.5"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "Invalid --gamma" in r.stderr
    r = subprocess.run([os.path.join(cli, "tau_2d_hypersonic_cuda"), "--bogus"], capture_output=True, text=True)
    assert r.returncode == 1 and "Unknown or incomplete argument" in r.stderr


def test_hypc_init_sim_is_host_code_and_equals_the_oracle():
    """init_sim (tau_hypersonic.c:450-475) needs no device: the product's host routine against the oracle's,
    incl. the golden mask / initial planes of the compiled reference at 256 x 256"""
    import numpy as np
    import oracle
    from fluid_sims_b200.hypersonic_c import init_sim
    for W, H in ((256, 256), (300, 300), (37, 91)):
        planes, mask = init_sim(W, H)
        op, om = oracle.hypcpu_init(W, H)
        assert np.array_equal(mask.ravel(), om) and all(np.array_equal(a.ravel(), b) for a, b in zip(planes, op))
    g = np.load(os.path.join(ROOT, "tests", "golden", "hypcpu_ref_256x256.npz"))
    planes, mask = init_sim(256, 256)
    assert np.array_equal(mask.ravel(), g["mask"]) and np.array_equal(planes[3].ravel(), g["E0"])


def test_snapshot_file_format_without_gpu(tmp_path):
    """The regression-baseline text format (tau_hypersonic_cuda_tests.cu:84-125) and the verification
    tolerances (:527-557) are host code: usable on a GPU-less box."""
    from fluid_sims_b200.hypersonic2d import Snapshot
    s = Snapshot(24, 1000, 1234.5678901234567, -3.25, 1e-17, 9.75e8, 0.5, 0.25, 24.99, 1.5e9, -2.5e7, 3.0e12)
    p = str(tmp_path / "baseline.txt")
    s.write(p)
    txt = open(p).read()
    assert txt.startswith("steps 24\nfluid_cells 1000\nsum_rho 1234.5678901234567\n")
    assert Snapshot.read(p).as_tuple() == s.as_tuple()
    t = Snapshot.read(p)
    t.sum_rho *= 1 + 4e-8          # inside 5e-8 relative
    assert s.failures_against(t) == []
    t.sum_rho *= 1 + 1e-7          # outside
    t.min_p += 2e-9                # outside the absolute 1e-9
    assert s.failures_against(t) == ["FAIL: sum_rho matches baseline", "FAIL: min_p matches baseline"]
    (tmp_path / "bad.txt").write_text("steps 3\nfluid_cells\n")
    from fluid_sims_b200 import TauError
    import pytest
    with pytest.raises(TauError, match="of 12 fields"):
        Snapshot.read(str(tmp_path / "bad.txt"))


def test_hyp2d_row_schedule_properties():
    """tau_hyp2d_plan_layers (pure host code): the layers tile [0, h_local) exactly once; the guided
    schedule is tall-to-short within [min_rows, max_rows] (the last layer may be the remainder); uniform
    schedules have one height; a layer never exceeds the descriptor's 12-bit row field."""
    import ctypes as C
    import numpy as np
    from fluid_sims_b200._lib import lib
    f = lib.tau_hyp2d_plan_layers
    f.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_int)] * 2 + [C.c_int]
    f.restype = C.c_int
    rng = np.random.default_rng(0)
    cases = [(4096, 137, 2960, 0, 2, 4, 48), (512, 137, 2960, 0, 2, 4, 48), (7, 2, 2960, 0, 2, 4, 48),
             (1, 1, 4, 0, 1, 1, 1), (4096, 137, 2960, 24, 2, 4, 48), (5000, 3, 16, 5000, 1, 4, 48),
             (4096, 1, 1, 0, 1, 4, 100000)]
    for _ in range(200):
        cases.append((int(rng.integers(1, 9000)), int(rng.integers(1, 300)), int(rng.integers(1, 6000)),
                      int(rng.choice([0, 0, 4, 7, 64])), int(rng.integers(1, 5)), int(rng.integers(1, 9)),
                      int(rng.integers(9, 80))))
    for (hl, ns, rw, seg, k, mn, mx) in cases:
        ly, lh = (C.c_int * hl)(), (C.c_int * hl)()
        n = f(hl, ns, rw, seg, k, mn, mx, ly, lh, hl)
        assert n >= 1, (hl, ns, rw, seg, k, mn, mx)
        y = 0
        for i in range(n):
            assert ly[i] == y and 1 <= lh[i] <= 4095
            y += lh[i]
        assert y == hl
        hs = [lh[i] for i in range(n)]
        if seg:
            assert all(h == min(seg, 4095) for h in hs[:-1]) and hs[-1] <= min(seg, 4095)
        else:
            body = hs[:-1]
            assert all(a >= b for a, b in zip(body, body[1:]))            # tall to short
            assert all(mn <= h <= min(mx, 4095) for h in body)
    ly, lh = (C.c_int * 4)(), (C.c_int * 4)()
    assert f(100, 1, 1, 4, 1, 4, 48, ly, lh, 4) < 0                        # capacity exceeded: loud
    assert f(10, 1, 1, 0, 0, 4, 48, ly, lh, 4) < 0                         # taper_k < 1: loud
