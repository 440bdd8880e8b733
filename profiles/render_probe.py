#!/usr/bin/env python
"""Drives the render pass at the BASELINE grid (4096^2 fp32) once per view mode; run under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:render`
to get per-kernel durations and DRAM traffic (profiles/r1_render_launches.txt)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fluid_sims_b200.hypersonic2d import Hypersonic2D, SimConfig  # noqa: E402

W = H = 4096
s = Hypersonic2D(SimConfig.default(W, H), dtype=sys.argv[1] if len(sys.argv) > 1 else "f32").init()
s.step(300)
s.sync()
for mode in range(7):
    t0 = time.perf_counter()
    rgba, mm = s.render(mode)
    print(f"mode {mode}: range [{mm[0]:.6g}, {mm[1]:.6g}]  {1e3 * (time.perf_counter() - t0):.2f} ms incl. the "
          f"64 MiB D2H copy of the frame")
