#!/usr/bin/env bash
# round 2, 1 GPU: edge strips without the two full-slot fold sweeps (left: five columns, right: clamped reads) and edge items first;
# parity tests of the 2-D solver, bench at 4096 / 512 rows (state_crc must not move), schedule constants re-checked
set -u
OUT=gpurun_out/hw_run22
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 1 "$OUT/$name.log" | cut -c1-300 | sed 's/^/   /' >> "$OUT/summary.txt"; }
B="python bench.py --no-e2e --no-cpu --no-extras --no-other --steps 200 --warmup 20"
run h2_tests 900 python -m pytest tests/test_hyp2d_gpu.py tests/test_hyp2d_snapshot_gpu.py tests/test_hyp2d_render_gpu.py -m gpu -q
run h4096 100 $B
run h512 100 $B --grid-h 512
run h1024 100 $B --grid-h 1024
run h2048 100 $B --grid-h 2048
for m in 3 6; do run h512_min$m 100 env TAU_HYP2D_MIN_ROWS=$m $B --grid-h 512; run h4096_min$m 100 env TAU_HYP2D_MIN_ROWS=$m $B; done
run h512_k3 100 env TAU_HYP2D_TAPER_K=3 $B --grid-h 512
run h4096_k3 100 env TAU_HYP2D_TAPER_K=3 $B
run h4096_max64 100 env TAU_HYP2D_MAX_ROWS=64 $B
run h4096_max32 100 env TAU_HYP2D_MAX_ROWS=32 $B
run sph 200 python bench_all.py sph --steps-sph 30
cat "$OUT/summary.txt"
