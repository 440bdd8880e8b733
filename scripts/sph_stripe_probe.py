"""per-kernel time of one sub-step: the single-GPU handle against a world = 1 stripe handle (same work + the stripe
bookkeeping: classify / absorb / id-ordering / cap-sized sort); run under `ncu --metrics gpu__time_duration.sum`"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fluid_sims_b200 import sph as S

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
which = sys.argv[2] if len(sys.argv) > 2 else "both"
P = S.Params(N=N)
pos0, vel0 = S.reset_particles(P)
if which in ("both", "single"):
    a = S.SPH(P).upload(pos0, vel0)
    a.step(4)
    a.sync()
if which in ("both", "stripe"):
    b = S.SPHStripes(P, 0, 1).upload(pos0, vel0)
    b.step(4)
    b.sync()
    print(b.status())
