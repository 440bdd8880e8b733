#!/usr/bin/env python
"""Aggregate an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv
--log-file X.csv <cmd>`) per kernel: launches, total / mean duration, share of all GPU time.
Usage: python profiles/launch_summary.py gpurun_out/launches.csv [> profiles/<name>.txt]"""
import collections
import csv
import re
import sys


def main():
    rows = []
    with open(sys.argv[1], newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, im, iv, iu = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    for r in rd:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
        name = re.sub(r"\(.*", "", r[ik]).replace("<unnamed>::", "").replace("void ", "")
        rows.append((name, v * scale))
    agg = collections.OrderedDict()
    for n, us in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'mean us':>10s} {'share':>7s}")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:60]:60s} {c:8d} {us:12.1f} {us / c:10.2f} {100 * us / tot:6.1f}%")
    print(f"{'all':60s} {len(rows):8d} {tot:12.1f}")


if __name__ == "__main__":
    main()
