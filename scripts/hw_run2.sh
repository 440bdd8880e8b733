#!/usr/bin/env bash
# round 2, second hardware run: the restructured SPH parity tests, the whole GPU suite, the new bench line
set -u
OUT=gpurun_out/hw_run2
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 4 "$OUT/$name.log" | cut -c1-1500 | sed 's/^/   /' >> "$OUT/summary.txt"; }
run sph 600 python -m pytest tests/test_sph_gpu.py -m gpu -q -s
run gpu_suite 900 python -m pytest tests -m gpu -q
run bench 600 python bench.py
run bench_fused 600 env TAU_HYP2D_PAIR=2 python bench.py --no-other --no-extras --no-cpu
run bench_ref 600 python bench.py --impl reference --steps 5 --warmup 1
cat "$OUT/summary.txt"
