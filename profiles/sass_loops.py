#!/usr/bin/env python
"""Largest loops (backward branches) of one kernel in a `cuobjdump -sass` listing, with their static opcode
mix.  Usage: cuobjdump -sass lib.so | python profiles/sass_loops.py <substring of the mangled kernel name>"""
import collections
import re
import sys

want = sys.argv[1]
txt = sys.stdin.read()
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if want not in name:
        continue
    lines = []
    for l in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            lines.append((int(m.group(1), 16), m.group(2).strip()))
    print(name[:100], "-", len(lines), "instructions")
    index = {a: i for i, (a, _) in enumerate(lines)}
    loops = []
    for i, (a, ins) in enumerate(lines):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", ins)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in index:
            loops.append((index[int(m.group(1), 16)], i))
    loops.sort(key=lambda x: x[1] - x[0], reverse=True)
    for s, e in loops[:int(sys.argv[2]) if len(sys.argv) > 2 else 4]:
        ops = collections.Counter()
        for _, ins in lines[s:e + 1]:
            op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0]
            ops["IMAD.MOV" if op.startswith("IMAD.MOV") else op.split(".")[0]] += 1
        print(f"  loop [{s}, {e}] {e - s + 1} instr: " + " ".join(f"{k}:{v}" for k, v in ops.most_common(22)))
