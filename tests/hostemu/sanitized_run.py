"""TEST INFRASTRUCTURE ONLY — run by tests/test_hostemu_cpu.py in a subprocess: the emulated 2-D solver built
with -fsanitize=alignment,bounds (a misaligned float2/uint2 access faults on a GPU; UBSan aborts here)."""
import os
import sys

os.environ["TAU_HC_SANITIZE"] = "1"
os.environ["TAU_HC_SMS"] = "3"
os.environ["TAU_HC_CTAS_PER_SM"] = "2"
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(__file__))
import numpy as np  # noqa: E402
import hyp2d_emu  # noqa: E402

W, H = 308, 64
a, *_ = hyp2d_emu.run(W, H, 6, "f32", geom_x0=W / 3.0)
b, *_ = hyp2d_emu.run(W, H, 6, "f32", pair=True, geom_x0=W / 3.0)
assert hyp2d_emu.run.last_work_items[2] > 0
assert max(float(np.abs(x - y).max()) for x, y in zip(a, b)) < 1e-3
c, *_ = hyp2d_emu.run(W, H, 6, "f32", pair=2, geom_x0=W / 3.0)     # the fused kernel
assert all(np.array_equal(x, y) for x, y in zip(b, c))
hyp2d_emu.run_slabs(200, 60, 5, "f32", 2, pair=2, geom_x0=60.0)
hyp2d_emu.run(203, 57, 5, "f64", geom_x0=60.0)                 # generic (non-TMA) loader
hyp2d_emu.run_slabs(200, 60, 5, "f32", 2, pair=True, geom_x0=60.0)   # peer pushes of both kernels
print("sanitized run clean")
