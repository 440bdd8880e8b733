#!/usr/bin/env bash
# round 2, 1 GPU: how much of the short-slab step time is warm-up vs imbalance — uniform segment heights against the guided
# schedule at 512 and 4096 rows; radix digit width of the SPH sort
set -u
OUT=gpurun_out/hw_run20
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 1 "$OUT/$name.log" | cut -c1-400 | sed 's/^/   /' >> "$OUT/summary.txt"; }
B="python bench.py --no-e2e --no-cpu --no-extras --no-other --steps 200 --warmup 20"
run h512_auto 100 $B --grid-h 512
for s in 12 16 21 24 26 32; do run h512_seg$s 100 env TAU_HYP2D_SEG_ROWS=$s $B --grid-h 512; done
run h512_k1 100 env TAU_HYP2D_TAPER_K=1 $B --grid-h 512
run h512_k3 100 env TAU_HYP2D_TAPER_K=3 $B --grid-h 512
run h4096_auto 100 $B
for s in 64 96 128 190; do run h4096_seg$s 100 env TAU_HYP2D_SEG_ROWS=$s $B; done
for b in 7 8 10; do run sph_bits$b 200 env TAU_SPH_SORT_BITS=$b python bench_all.py sph --steps-sph 30; done
cat "$OUT/summary.txt"
