"""Multi-GPU (>= 2 devices on the box): slab-decomposed 2-D hypersonic (chain + all-reduce max) and
Gray-Scott (ring) over NCCL reproduce the single-GPU run bit-for-bit.  Skipped on 1-GPU boxes; the
host-side exchange logic itself is covered on CPU by tests/test_slab_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_runs_match_single_gpu(tmp_path):
    from fluid_sims_b200 import device_count
    n = device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    if os.environ.get("TAU_TEST_WORLD"):          # e.g. 8 on a full box
        world = min(n, int(os.environ["TAU_TEST_WORLD"]))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "_mgpu_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert (tmp_path / "result.txt").read_text() == "OK"


def _dump_planes(path):
    import numpy as np
    raw = open(path, "rb").read()
    assert raw[:8] == b"TAUDUMP1"
    npl, es, d0, d1, d2 = np.frombuffer(raw, np.int32, 5, 8)
    dt = np.float32 if es == 4 else np.float64
    return np.frombuffer(raw, dt, npl * d0 * d1 * d2, 44).reshape(npl, d1, d0), float(np.frombuffer(raw, np.float64, 1, 36)[0])


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_group_handle_one_process_matches_single_gpu(dtype):
    """tau_hyp2d_group_* (multi-GPU behind the C boundary, one process, no torchrun / NCCL): bit-identical to the
    single-GPU handle — from k_init, and from an uploaded developed state — with as many devices as the box has
    (1 included: the group of one is the plain handle)."""
    import numpy as np
    from fluid_sims_b200 import device_count
    from fluid_sims_b200.hypersonic2d import Hypersonic2D, Hypersonic2DGroup, SimConfig
    n = min(device_count(), 4)
    W, H, steps = 1024, 512, 60
    cfg = SimConfig.default(W, H)
    one = Hypersonic2D(cfg, dtype=dtype).init()
    one.step(steps)
    ref, rmask = one.download()
    for ngpus in sorted({1, n}):
        g = Hypersonic2DGroup(cfg, ngpus, dtype=dtype).init()
        assert len(g.slabs()) == ngpus and sum(h for _, h in g.slabs()) == H
        g.step(steps // 2)
        g.step(steps - steps // 2)
        out, mask = g.download()
        assert np.array_equal(mask, rmask) and all(np.array_equal(a, b) for a, b in zip(out, ref))
        assert g.clock() == one.clock()
        assert g.launch_count >= steps * ngpus
        g.upload(ref, rmask)
        g.step(25)
        out, _ = g.download()
        px, mm = g.render(5)
        two = Hypersonic2D(cfg, dtype=dtype).upload(ref, rmask)
        two.step(25)
        exp, _ = two.download()
        assert all(np.array_equal(a, b) for a, b in zip(out, exp))
        epx, emm = two.render(5)
        assert mm == emm and np.array_equal(px.reshape(H, W, 4), np.asarray(epx).reshape(H, W, 4))
        g.close()
        two.close()
    one.close()


def test_cli_gpus_flag_dump_identical_to_one_gpu(tmp_path):
    """`tau_2d_hypersonic_cuda --gpus N` (C host, no Python in the loop) writes the same dump as `--gpus 1`"""
    import numpy as np
    from fluid_sims_b200 import device_count
    n = device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    exe = os.path.join(ROOT, "fluid_sims_b200", "cli", "tau_2d_hypersonic_cuda")
    dumps = []
    for g in (1, 2, min(n, 8)):
        d = tmp_path / f"g{g}.dump"
        r = subprocess.run([exe, "--nx", "1024", "--ny", "512", "--frames", "40", "--dtype", "f32", "--gpus", str(g),
                            "--dump", str(d)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert "Mcell-updates/s" in r.stdout
        dumps.append(_dump_planes(d))
    for planes, t in dumps[1:]:
        assert t == dumps[0][1] and np.array_equal(planes, dumps[0][0])


def test_tau3d_gpus_flag_dump_identical_to_one_gpu(tmp_path):
    """`tau3d --gpus N` (tau_hyp3d_group_*: one process, z-slab ring, ghost planes by cudaMemcpyPeerAsync, host max) writes
    the same dump as `--gpus 1`, byte for byte — 96^3, 30 steps from k_init"""
    from fluid_sims_b200 import device_count
    n = device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    exe = os.path.join(ROOT, "fluid_sims_b200", "cli", "tau3d")
    dumps = []
    for g in (1, 2, min(n, 8)):
        d = tmp_path / f"t3_{g}.dump"
        r = subprocess.run([exe, "--n", "96", "--frames", "15", "--gpus", str(g), "--dump", str(d)], capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0 and "Mcell-updates/s" in r.stdout, r.stderr
        dumps.append(open(d, "rb").read())
    assert all(x == dumps[0] for x in dumps[1:]) and len(dumps[0]) > 6 * 4 * 96 ** 3
