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
    for m in re.finditer(r"(\w+(?:<[^<>;()]*>)?)\s*<<<((?:(?!>>>|<<<).)+?)>>>\s*\(", src, flags=re.S):
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


_ASM_RULES = (
    # (ptx prefix, C++ template over outputs o[] and inputs i[])
    ("rcp.approx.ftz.f32", "{o0} = 1.0f / ({i0});"),
    ("sqrt.approx.ftz.f32", "{o0} = sqrtf({i0});"),
    ("mov.u32 %0, %%tid.x", "{o0} = threadIdx.x;"),
    ("griddepcontrol.", ";"),
    ("mov.u64 %0, %globaltimer", "{o0} = ++tau_hc::fake_timer;"),
    ("ld.acquire.sys.global.u64", "{o0} = *(volatile unsigned long long *)({i0});"),
    ("st.release.sys.global.u64", "*(volatile unsigned long long *)({i0}) = ({i1});"),
)


def _operands(section):
    # '"=f"(r), "l"(a)' -> ['r', 'a'] (balanced parentheses)
    ops, i = [], 0
    while True:
        m = re.compile(r'"[^"]*"\s*\(').search(section, i)
        if not m:
            return ops
        j, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(section[j], 0)
            j += 1
        ops.append(section[m.end():j - 1].strip())
        i = j


def rewrite_asm(src: str) -> str:
    """every inline-PTX statement becomes its host equivalent; an unknown one is an error"""
    out, pos = "", 0
    for m in re.finditer(r"\basm\s*(?:volatile\s*)?\(", src):
        if m.start() < pos:
            continue
        i, depth, instr = m.end(), 1, False
        while depth:
            ch = src[i]
            if ch == '"' and src[i - 1] != "\\":
                instr = not instr
            elif not instr:
                depth += {"(": 1, ")": -1}.get(ch, 0)
            i += 1
        body = src[m.end():i - 1]
        end = src.index(";", i) + 1
        # split the body at top-level ':' outside strings
        parts, curp, instr, depth = [], "", False, 0
        for k, ch in enumerate(body):
            if ch == '"' and body[k - 1] != "\\":
                instr = not instr
            if not instr:
                depth += {"(": 1, ")": -1}.get(ch, 0)
            if ch == ":" and not instr and depth == 0:
                parts.append(curp)
                curp = ""
            else:
                curp += ch
        parts.append(curp)
        ptx = "".join(re.findall(r'"((?:[^"\\]|\\.)*)"', parts[0])).strip()
        outs = _operands(parts[1]) if len(parts) > 1 else []
        ins = _operands(parts[2]) if len(parts) > 2 else []
        for prefix, tmpl in _ASM_RULES:
            if ptx.startswith(prefix):
                kw = {f"o{n}": v for n, v in enumerate(outs)}
                kw.update({f"i{n}": v for n, v in enumerate(ins)})
                rep = tmpl.format(**kw)
                break
        else:
            raise RuntimeError(f"hostemu: no host equivalent for inline PTX {ptx!r}")
        out += src[pos:m.start()] + rep
        pos = end
    return out + src[pos:]


def preprocess(path: str) -> str:
    text = open(path).read()
    # local .cuh pieces are inlined (and rewritten the same way)
    def inline(m):
        return preprocess(os.path.join(os.path.dirname(path), m.group(1))).replace("#pragma once", "")
    text = re.sub(r'#include "(\w+\.(?:cuh|inc))"', lambda m: m.group(0) if m.group(1) == "common.cuh" else inline(m), text)
    text = re.sub(r"extern\s+__shared__\s+(.*?)(\w+)\[\];", r"__shared__ \1\2[TAU_HC_SMEM_BYTES];", text)
    text = re.sub(r"__shared__\s+alignas\((\w+)\)", r"__shared__ __attribute__((aligned(\1)))", text)  # g++: no alignas after static
    return cuda_to_host(rewrite_asm(text))


def build(name: str, defines=(), tag: str = "", contract: str = "off", sanitize: bool = False) -> str:
    """-> path of build/hostemu/lib<name><tag>_hostemu.so (rebuilt when a source is newer)"""
    os.makedirs(OUT, exist_ok=True)
    cu = os.path.join(ROOT, "fluid_sims_b200", "csrc", f"{name}.cu")
    cpp = os.path.join(OUT, f"{name}{tag}_host.cpp")
    so = os.path.join(OUT, f"lib{name}{tag}_hostemu.so")
    csrc_dir = os.path.join(ROOT, "fluid_sims_b200", "csrc")
    deps = [cu, os.path.join(ROOT, "tests", "hostemu", "hostemu.h"), __file__,
            *[os.path.join(csrc_dir, f) for f in os.listdir(csrc_dir) if f.endswith((".cuh", ".inc"))],
            os.path.join(ROOT, "include", "tau_b200.h")]
    if os.path.exists(so) and all(os.path.getmtime(so) > os.path.getmtime(d) for d in deps):
        return so
    text = preprocess(cu)
    assert not re.search(r"<<<\s*[^>\s]", text) and not re.search(r"\basm\s*(volatile\s*)?\(", text)
    # the .cu includes the public header relative to csrc/
    text = text.replace('#include "../../include/tau_b200.h"', f'#include "{os.path.join(ROOT, "include", "tau_b200.h")}"')
    open(cpp, "w").write(text)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-mfma", f"-ffp-contract={contract}", "-Wall", "-Wl,-Bsymbolic",
                    "-Wno-unused-function", "-Wno-unused-variable", "-Wno-unknown-pragmas", "-Wno-return-type",  # gcc 13 misreads `if constexpr … else return` in lambdas
                    *[f"-D{d}" for d in defines],
                    *(["-fsanitize=alignment,bounds", "-fno-sanitize-recover=all", "-g"] if sanitize else []),
                    "-I", os.path.join(ROOT, "tests", "hostemu"), cpp, "-o", so, "-lm"], check=True)
    return so


def build_all() -> str:
    """every product translation unit for the emulator in ONE library with the product's C-ABI:
    TAU_B200_LIB=build/hostemu/libtau_b200_hostemu.so runs Python code written for the GPU library on the CPU
    (a pre-flight for tests/ -m gpu at sizes the emulator can afford; never shipped, never a default)."""
    os.makedirs(OUT, exist_ok=True)
    names = ["burgers", "gray_scott", "hypersonic2d", "hypersonic3d", "hypersonic_c", "shallow_water", "snapshot", "splat4", "sph"]
    so = os.path.join(OUT, "libtau_b200_hostemu.so")
    csrc = os.path.join(ROOT, "fluid_sims_b200", "csrc")
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".inc"))] + [
        os.path.join(ROOT, "tests", "hostemu", "hostemu.h"), __file__, os.path.join(ROOT, "include", "tau_b200.h")]
    if os.path.exists(so) and all(os.path.getmtime(so) > os.path.getmtime(d) for d in deps):
        return so
    objs = []
    for name in names:
        cu = os.path.join(ROOT, "fluid_sims_b200", "csrc", f"{name}.cu")
        cpp = os.path.join(OUT, f"{name}_all_host.cpp")
        obj = os.path.join(OUT, f"{name}_all_host.o")
        text = preprocess(cu).replace('#include "../../include/tau_b200.h"',
                                      f'#include "{os.path.join(ROOT, "include", "tau_b200.h")}"')
        open(cpp, "w").write(text)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-c", "-mfma", "-ffp-contract=off", "-Wno-return-type",
                        "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "tests", "hostemu"), cpp, "-o", obj], check=True)
        objs.append(obj)
    subprocess.run(["g++", "-shared", "-Wl,-Bsymbolic", "-o", so] + objs + ["-lm"], check=True)
    return so


if __name__ == "__main__":
    # python hostemu_build.py rewrite IN.cu OUT.cpp — the CUDA -> host source rewriting alone (oracle/Makefile uses it to
    # run a reference translation unit on the CPU emulator)
    import sys
    if len(sys.argv) == 4 and sys.argv[1] == "rewrite":
        open(sys.argv[3], "w").write(preprocess(sys.argv[2]))
    else:
        sys.exit("usage: hostemu_build.py rewrite IN.cu OUT.cpp")
