"""TEST INFRASTRUCTURE ONLY.  Writes a float-typed scratch copy of the reference's tau_hypersonic_cuda.cu (hard part 2:
"a sed-generated double->float build of the reference is a useful secondary oracle to separate precision from algorithm"):
every `double` becomes `float`, every floating literal outside string literals gets an `f` suffix, W / H are rewritten.
usage: gen_f32_src.py IN.cu OUT.cu W H   (oracle/Makefile deletes OUT after compiling it; nothing is committed)"""
import re
import sys

src = open(sys.argv[1]).read()
W, H = sys.argv[3], sys.argv[4]
src = re.sub(r"^#define W 8192$", f"#define W {W}", src, flags=re.M)
src = re.sub(r"^#define H 1024$", f"#define H {H}", src, flags=re.M)
src = re.sub(r"\bdouble\b", "float", src)
lit = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")
out = []
for line in src.split("\n"):
    parts = line.split('"')
    for i in range(0, len(parts), 2):
        parts[i] = lit.sub(r"\1f", parts[i])
    out.append('"'.join(parts))
open(sys.argv[2], "w").write("\n".join(out))
