#!/usr/bin/env bash
# round 2, 1 GPU: schedule constants at 2048 / 1024 rows (the N = 2 / N = 4 slab heights)
set -u
OUT=gpurun_out/hw_run28
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 1 "$OUT/$name.log" | cut -c1-300 | sed 's/^/   /' >> "$OUT/summary.txt"; }
B="python bench.py --no-e2e --no-cpu --no-extras --no-other --steps 200 --warmup 20"
for h in 2048 1024; do for k in 1 2; do for mx in 48 96; do
  run h${h}_k${k}_max${mx} 100 env TAU_HYP2D_TAPER_K=$k TAU_HYP2D_MAX_ROWS=$mx TAU_HYP2D_MIN_ROWS=6 $B --grid-h $h
done; done; done
run h2048_k2_max48_min4 100 env TAU_HYP2D_TAPER_K=2 TAU_HYP2D_MAX_ROWS=48 TAU_HYP2D_MIN_ROWS=4 $B --grid-h 2048
cat "$OUT/summary.txt"
