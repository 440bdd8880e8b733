#!/usr/bin/env bash
# round 2, 1 GPU: 3-D staging with the loads of half a tile column in flight; 8x8x8 tile variant (2 CTAs x 16 warps per SM)
set -u
OUT=gpurun_out/hw_run17
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 3 "$OUT/$name.log" | cut -c1-700 | sed 's/^/   /' >> "$OUT/summary.txt"; }
NCU="ncu --clock-control none"
run hyp3d_tests 900 python -m pytest tests/test_hyp3d_gpu.py -m gpu -q
run hyp3d_bench 300 python bench_all.py hyp3d
run hyp3d_bench_tz8 300 env TAU_B200_LIB=scripts/variants/libtau_tz8.so python bench_all.py hyp3d
run hyp3d_tests_tz8 900 env TAU_B200_LIB=scripts/variants/libtau_tz8.so python -m pytest tests/test_hyp3d_gpu.py -m gpu -q
run hyp3d_ncu 600 $NCU --set full --import-source on -k regex:hyp3d_step -s 40 -c 1 -o $OUT/hyp3d_step_r2d python bench_all.py hyp3d --steps3 5 --warm3 45
run hyp3d_ncu_tz8 600 env TAU_B200_LIB=scripts/variants/libtau_tz8.so $NCU --set full --import-source on -k regex:hyp3d_step -s 40 -c 1 -o $OUT/hyp3d_step_r2d_tz8 python bench_all.py hyp3d --steps3 5 --warm3 45
cat "$OUT/summary.txt"
