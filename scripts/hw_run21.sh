#!/usr/bin/env bash
# round 2, 1 GPU: start stagger of the resident warps (lockstep first items) at 512 / 4096 rows, static and guided schedules
set -u
OUT=gpurun_out/hw_run21
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 1 "$OUT/$name.log" | cut -c1-300 | sed 's/^/   /' >> "$OUT/summary.txt"; }
B="python bench.py --no-e2e --no-cpu --no-extras --no-other --steps 200 --warmup 20"
for st in 0 25 50 100 200 400; do
  run h512_auto_st$st 100 env TAU_HYP2D_STAGGER_NS=$st $B --grid-h 512
  run h4096_auto_st$st 100 env TAU_HYP2D_STAGGER_NS=$st $B
done
for st in 0 100 400; do run h512_seg26_st$st 100 env TAU_HYP2D_STAGGER_NS=$st TAU_HYP2D_SEG_ROWS=26 $B --grid-h 512; done
cat "$OUT/summary.txt"
