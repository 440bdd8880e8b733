#!/usr/bin/env python
"""Static opcode histogram per kernel of a `cuobjdump -sass` listing (stdin).  For straight-line
kernels (no loops) static = executed.  Usage: cuobjdump -sass X.cubin | python profiles/sass_count.py"""
import collections
import re
import sys

txt = sys.stdin.read()
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0]
    ops = collections.Counter()
    for l in f.split("\n"):
        m = re.match(r"\s*/\*[0-9a-f]{4,5}\*/\s+(.*?);", l)
        if m:
            ops[re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip()).split()[0].split(".")[0]] += 1
    tot = sum(ops.values())
    print(f"{name[:60]}: {tot} instructions")
    print("   " + " ".join(f"{k}:{v}" for k, v in ops.most_common(24)))
