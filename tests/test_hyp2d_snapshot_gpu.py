"""GPU tests of the on-disk artefacts around the 2-D hypersonic path: the reference's 12-scalar
regression snapshot (tau_hypersonic_cuda_tests.cu:20-36, :84-176, :527-557) and checkpoint/resume."""
import numpy as np
import pytest

import oracle
from fluid_sims_b200.hypersonic2d import Hypersonic2D, SimConfig, Snapshot

pytestmark = pytest.mark.gpu


def test_snapshot_equals_oracle_and_file_round_trip(tmp_path):
    W, H, steps = 256, 128, 24  # 24 = the reference's default --steps (:57)
    s = Hypersonic2D(SimConfig.default(W, H), dtype="f64").init()
    s.step(steps)
    snap = s.snapshot()
    planes, mask = s.download()
    want = oracle.hyp2d_snapshot(oracle.hyp2d_cfg(W, H), steps, planes, mask)
    # same state, same (sequential) summation order as compute_snapshot :143-176 -> identical bits
    assert snap.as_tuple() == tuple([int(want[0]), int(want[1])] + [float(x) for x in want[2:]])
    assert snap.fluid_cells > 0 and snap.min_rho >= 1e-25 and snap.min_p >= 1e-25  # :517-521
    path = str(tmp_path / "tau_hypersonic_cuda_baseline.txt")
    snap.write(path)
    lines = open(path).read().splitlines()
    assert [l.split()[0] for l in lines] == ["steps", "fluid_cells", "sum_rho", "sum_mx", "sum_my", "sum_E",
                                             "min_rho", "min_p", "max_mach", "checksum_rho", "checksum_mx",
                                             "checksum_E"]
    back = Snapshot.read(path)
    assert back.as_tuple() == snap.as_tuple()  # %.17g round-trips doubles
    assert snap.failures_against(back) == []
    # the fp32 handle verifies against the fp64 baseline at the reference's tolerances after 24 steps?
    # No: 5e-8 relative is below float rounding — the failures must be reported, not hidden.
    s32 = Hypersonic2D(SimConfig.default(W, H), dtype="f32").init()
    s32.step(steps)
    f = s32.snapshot().failures_against(back)
    assert all(x.startswith("FAIL: ") for x in f)
    # one more step changes the step count -> the reference's first check fails
    s.step(1)
    assert "FAIL: steps match baseline" in s.snapshot().failures_against(back)
    s.close()
    s32.close()


@pytest.mark.skipif(not oracle.has_ref("ref_hyp2d_1024x512"), reason="oracle/_ref not built")
def test_snapshot_verifies_against_reference_kernels_run():
    """A baseline computed from the reference kernels' fields verifies under the reference's own
    tolerances against the product's run of the same case."""
    W, H, steps = 1024, 512, 24
    ref, rmask, _, _, _ = oracle.ref_hyp2d_run(W, H, oracle.hyp2d_cfg(W, H).as11(), steps)
    want = oracle.hyp2d_snapshot(oracle.hyp2d_cfg(W, H), steps, ref, rmask)
    exp = Snapshot(int(want[0]), int(want[1]), *[float(x) for x in want[2:]])
    s = Hypersonic2D(SimConfig.default(W, H), dtype="f64").init()
    s.step(steps)
    assert s.snapshot().failures_against(exp) == []
    s.close()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_checkpoint_resume_is_bit_identical(tmp_path, dtype):
    W, H = 200, 120
    a = Hypersonic2D(SimConfig.default(W, H, geom_x0=60.0), dtype=dtype).init()
    a.step(17)
    path = str(tmp_path / "run.ckpt")
    a.checkpoint_save(path)
    info = Hypersonic2D.checkpoint_info(path)
    assert (info["W"], info["H"], info["dtype"], info["steps"]) == (W, H, dtype, 17)
    assert info["sim_t"] == a.clock()[0]
    a.step(23)
    want, _ = a.download()
    b = Hypersonic2D(SimConfig.default(W, H, geom_x0=60.0), dtype=dtype).checkpoint_load(path)
    assert b.steps_done == 17 and b.clock()[0] == info["sim_t"]
    b.step(23)
    got, _ = b.download()
    for x, y in zip(got, want):
        assert np.array_equal(x, y)
    assert b.clock() == a.clock()
    # a handle of another shape refuses the file, loudly
    from fluid_sims_b200 import TauError
    c = Hypersonic2D(SimConfig.default(W, H + 8), dtype=dtype)
    with pytest.raises(TauError, match="holds a"):
        c.checkpoint_load(path)
    for h in (a, b, c):
        h.close()
