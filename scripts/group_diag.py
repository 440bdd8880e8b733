"""tau_hyp2d_group (one process, n devices) against the single-GPU handle: where do they part?"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np

from fluid_sims_b200 import device_count
from fluid_sims_b200.hypersonic2d import Hypersonic2D, Hypersonic2DGroup, SimConfig

n = int(sys.argv[1]) if len(sys.argv) > 1 else min(device_count(), 2)
for dtype in ("f64", "f32"):
    for W, H in ((1024, 512), (4096, 4096)):
        cfg = SimConfig.default(W, H)
        one = Hypersonic2D(cfg, dtype=dtype).init()
        two = Hypersonic2D(cfg, dtype=dtype).init()
        g = Hypersonic2DGroup(cfg, n, dtype=dtype).init()
        done = 0
        for upto in (1, 2, 3, 5, 10, 20, 40, 60, 120, 300):
            one.step(upto - done)
            two.step(upto - done)
            g.step(upto - done)
            done = upto
            a, _ = one.download()
            b, _ = g.download()
            c, _ = two.download()
            d = [float(np.abs(x.astype(np.float64) - y).max()) for x, y in zip(a, b)]
            d1 = [float(np.abs(x.astype(np.float64) - y).max()) for x, y in zip(a, c)]
            rows = sorted(set(np.nonzero((a[0] != b[0]).any(axis=1))[0].tolist()))
            print(f"{dtype} {W}x{H} n={n} chunk={os.environ.get('TAU_HYP2D_GROUP_CHUNK', '16')} after {upto:4d} steps: group-vs-one max diff {max(d):.3e} "
                  f"(one-vs-one {max(d1):.1e}) clocks {g.clock()[0] == one.clock()[0]} rows differing {len(rows)} "
                  f"{rows[:4]}..{rows[-2:] if rows else ''}", flush=True)
        g.close(); one.close(); two.close()
