"""Root-cause aid for the SPH parity test: how far apart are the product, the
reference's own (racy) kernels and the sorted-order CPU oracle, and how far is the reference from
itself run to run.  Prints, per configuration, the fraction of particles whose position differs by
more than 2e-5 and the largest difference.  Checker-side script (imports oracle/)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np

import oracle
from fluid_sims_b200.sph import SPH, Params, reset_particles


def product(P, pos0, vel0, frames):
    s = SPH(P).upload(pos0, vel0)
    s.step(frames)
    out = s.download()
    s.close()
    return out


def frac(a, b, tol=2e-5):
    d = np.abs(a - b).max(axis=1)
    return float((d > tol).mean()), float(d.max())


def main():
    cases = [(20000, 25, dict(viscSub=3)), (20000, 25, dict(viscSub=3, rain=0)), (20000, 25, dict(viscSub=1)),
             (20000, 8, dict(viscSub=3)), (65536, 30, {}), (65536, 30, dict(rain=0))]
    for N, frames, kw in cases:
        P = Params(N=N, **kw)
        op = oracle.sph_params(N, **kw)
        pos0, vel0 = reset_particles(P)
        prods = [product(P, pos0, vel0, frames) for _ in range(2)]
        refs = [oracle.ref_sph_run(op, pos0, vel0, frames) for _ in range(4)]
        print(f"N={N} frames={frames} {kw}")
        print("  product vs product      :", frac(prods[0][0], prods[1][0]))
        for i, r in enumerate(refs):
            print(f"  product vs reference[{i}] :", frac(prods[0][0], r[0]))
        for i in range(1, len(refs)):
            print(f"  reference[0] vs ref[{i}]   :", frac(refs[0][0], refs[i][0]))
        if N <= 20000:
            t0 = time.time()
            o = oracle.sph_run(op, pos0, vel0, frames)
            print(f"  product vs cpu oracle   :", frac(prods[0][0], o[0]), f"({time.time() - t0:.1f}s)")
            print(f"  reference[0] vs oracle  :", frac(refs[0][0], o[0]))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
