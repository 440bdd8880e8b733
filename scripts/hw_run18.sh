#!/usr/bin/env bash
# round 2, 1 GPU: whole GPU suite, smoke, bench (driver defaults) and its launch list after this session's SPH / 3-D changes
set -u
OUT=gpurun_out/hw_run18
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 4 "$OUT/$name.log" | cut -c1-3000 | sed 's/^/   /' >> "$OUT/summary.txt"; }
NCU="ncu --clock-control none"
run gpu_suite 1500 python -m pytest tests -m gpu -q
run smoke 300 python __graft_entry__.py smoke
run bench 400 python bench.py
run bench_ref 400 python bench.py --impl reference --steps 3 --warmup 1
run bench_launches 400 $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $OUT/bench_launches.csv python bench.py --steps 20 --warmup 3 --develop 100 --e2e-frames 4 --cpu-steps 1 --cpu-full-steps 0
cat "$OUT/summary.txt"
