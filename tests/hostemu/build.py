"""TEST INFRASTRUCTURE ONLY — builds build/hostemu/lib<name>_hostemu.so from fluid_sims_b200/csrc/<name>.cu
for tests/test_hostemu_cpu.py (see tests/hostemu/hostemu.h for what this is and is not)."""
import os
import re
import subprocess

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
OUT = os.path.join(ROOT, "build", "hostemu")


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def cuda_to_host(src: str) -> str:
    src = src.replace('#include "common.cuh"', '#include "hostemu.h"')
    out, pos = "", 0
    for m in re.finditer(r"(\w+)\s*<<<(.+?)>>>\s*\(", src, flags=re.S):
        if m.start() < pos:
            continue
        cfg = _split_top(m.group(2))
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        args = src[m.end():i - 1]
        j = src.index(";", i)
        out += src[pos:m.start()]
        out += f"tau_hc::launch(dim3({cfg[0]}), dim3({cfg[1]}), [&] {{ {m.group(1)}({args}); }});"
        pos = j + 1
    return out + src[pos:]


def build(name: str) -> str:
    os.makedirs(OUT, exist_ok=True)
    cu = os.path.join(ROOT, "fluid_sims_b200", "csrc", f"{name}.cu")
    cpp = os.path.join(OUT, f"{name}_host.cpp")
    so = os.path.join(OUT, f"lib{name}_hostemu.so")
    deps = [cu, os.path.join(ROOT, "tests", "hostemu", "hostemu.h"), __file__,
            os.path.join(ROOT, "include", "tau_b200.h")]
    if os.path.exists(so) and all(os.path.getmtime(so) > os.path.getmtime(d) for d in deps):
        return so
    text = cuda_to_host(open(cu).read())
    assert "<<<" not in text
    # the .cu includes the public header relative to csrc/
    text = text.replace('#include "../../include/tau_b200.h"', f'#include "{os.path.join(ROOT, "include", "tau_b200.h")}"')
    open(cpp, "w").write(text)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-mfma", "-ffp-contract=off", "-Wall", "-Wl,-Bsymbolic",
                    "-Wno-unused-function", "-Wno-unused-variable", "-Wno-unknown-pragmas",
                    "-I", os.path.join(ROOT, "tests", "hostemu"), cpp, "-o", so, "-lm"], check=True)
    return so
