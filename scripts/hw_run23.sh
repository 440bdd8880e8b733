#!/usr/bin/env bash
# round 2, 1 GPU: schedule constants after the edge-strip fix (the edge items no longer dominate the tail)
set -u
OUT=gpurun_out/hw_run23
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 1 "$OUT/$name.log" | cut -c1-300 | sed 's/^/   /' >> "$OUT/summary.txt"; }
B="python bench.py --no-e2e --no-cpu --no-extras --no-other --steps 200 --warmup 20"
for k in 1 2; do for mx in 64 96 128 192; do for mn in 4 8; do
  run h4096_k${k}_max${mx}_min${mn} 100 env TAU_HYP2D_TAPER_K=$k TAU_HYP2D_MAX_ROWS=$mx TAU_HYP2D_MIN_ROWS=$mn $B
done; done; done
for k in 1 2 3; do for mn in 4 6 8 10; do
  run h512_k${k}_min${mn} 100 env TAU_HYP2D_TAPER_K=$k TAU_HYP2D_MIN_ROWS=$mn $B --grid-h 512
done; done
for s in 16 24 26; do run h512_seg$s 100 env TAU_HYP2D_SEG_ROWS=$s $B --grid-h 512; done
run h4096_seg190 100 env TAU_HYP2D_SEG_ROWS=190 $B
cat "$OUT/summary.txt"
