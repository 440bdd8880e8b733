#!/usr/bin/env bash
# round 2, 1 GPU: SPH parity with the frame-by-frame gates; 3-D step with the decoded-primitive side buffer (parity, A/B speed, ncu)
set -u
OUT=gpurun_out/hw_run15
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 4 "$OUT/$name.log" | cut -c1-2000 | sed 's/^/   /' >> "$OUT/summary.txt"; }
NCU="ncu --clock-control none"
run sph_tests 600 python -m pytest tests/test_sph_gpu.py -m gpu -q -s
run hyp3d_tests 900 python -m pytest tests/test_hyp3d_gpu.py -m gpu -q -s
run hyp3d_bench 300 python bench_all.py hyp3d
run hyp3d_bench_noprims 300 env TAU_HYP3D_PRIMS=0 python bench_all.py hyp3d
run hyp3d_ncu 600 $NCU --set full --import-source on -k regex:hyp3d_step -s 40 -c 1 -o $OUT/hyp3d_step_r2b python bench_all.py hyp3d --steps3 5 --warm3 45
cat "$OUT/summary.txt"
