#!/usr/bin/env python
"""Instruction mix from an ncu source page (`ncu -i X.ncu-rep --page source --csv --print-source sass`):
executed warp-instructions by opcode, normalised per marched row, plus stall-sample totals.
Usage: python profiles/sass_mix.py src.csv ROWS   (ROWS = warp-rows of output, e.g. 4096*137)"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    nrow = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    hdr = rows[1]
    isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    ops, stalls, samples = collections.Counter(), collections.Counter(), collections.Counter()
    tot = 0
    for r in rows[2:]:
        if len(r) <= iex:
            continue
        t = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip())
        op = t.split()[0].split(".")[0] if t else "?"
        n = int(r[iex] or 0)
        ops[op] += n
        tot += n
        samples[op] += int(r[ismp] or 0)
        for i in stall_cols:
            stalls[hdr[i]] += int(r[i] or 0)
    print(f"executed warp-instructions: {tot}  per row: {tot / nrow:.1f}")
    for op, n in ops.most_common(32):
        print(f"  {op:10s} {n / nrow:8.1f}   samples {samples[op]}")
    st = sum(stalls.values())
    print("stall samples:")
    for k, v in stalls.most_common(12):
        print(f"  {k:28s} {100 * v / st:5.1f}%")


if __name__ == "__main__":
    main()
