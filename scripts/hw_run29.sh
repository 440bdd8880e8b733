#!/usr/bin/env bash
# round 2, 1 GPU: K = 1 with the first-layer cap at half the fair share, at 1024 / 512 rows
set -u
OUT=gpurun_out/hw_run29
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 1 "$OUT/$name.log" | cut -c1-300 | sed 's/^/   /' >> "$OUT/summary.txt"; }
B="python bench.py --no-e2e --no-cpu --no-extras --no-other --steps 200 --warmup 20"
run h1024_k1_max24 100 env TAU_HYP2D_TAPER_K=1 TAU_HYP2D_MAX_ROWS=24 TAU_HYP2D_MIN_ROWS=6 $B --grid-h 1024
run h1024_k1_max32 100 env TAU_HYP2D_TAPER_K=1 TAU_HYP2D_MAX_ROWS=32 TAU_HYP2D_MIN_ROWS=6 $B --grid-h 1024
run h512_k1_max12 100 env TAU_HYP2D_TAPER_K=1 TAU_HYP2D_MAX_ROWS=12 TAU_HYP2D_MIN_ROWS=6 $B --grid-h 512
run h512_k1_max16 100 env TAU_HYP2D_TAPER_K=1 TAU_HYP2D_MAX_ROWS=16 TAU_HYP2D_MIN_ROWS=6 $B --grid-h 512
run h512_k1_max24 100 env TAU_HYP2D_TAPER_K=1 TAU_HYP2D_MAX_ROWS=24 TAU_HYP2D_MIN_ROWS=6 $B --grid-h 512
run h2048_k1_max64 100 env TAU_HYP2D_TAPER_K=1 TAU_HYP2D_MAX_ROWS=64 TAU_HYP2D_MIN_ROWS=6 $B --grid-h 2048
cat "$OUT/summary.txt"
