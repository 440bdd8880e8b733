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
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "_mgpu_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert (tmp_path / "result.txt").read_text() == "OK"
