#!/usr/bin/env bash
# round 2, 1 GPU, last run of the round: two small experiments (shortest segment of the 2-D schedule; 8 key columns per SPH cell),
# then the whole suite and the bench on the shipping build
set -u
OUT=gpurun_out/hw_run32
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 2 "$OUT/$name.log" | cut -c1-400 | sed 's/^/   /' >> "$OUT/summary.txt"; }
B="python bench.py --no-e2e --no-cpu --no-extras --no-other --steps 200 --warmup 20"
for m in 4 8; do run h512_min$m 60 env TAU_HYP2D_MIN_ROWS=$m $B --grid-h 512; run h4096_min$m 60 env TAU_HYP2D_MIN_ROWS=$m $B; done
run sph_subx8 120 env TAU_B200_LIB=scripts/variants/libtau_subx8.so python bench_all.py sph --steps-sph 30
run sph_subx8_tests 200 env TAU_B200_LIB=scripts/variants/libtau_subx8.so python -m pytest tests/test_sph_gpu.py -m gpu -q
run gpu_suite 1200 python -m pytest tests -m gpu -q
run bench 400 python bench.py
cat "$OUT/summary.txt"
