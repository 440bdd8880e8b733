#!/usr/bin/env bash
# round 2, 1 GPU: 3-D staging loop as a column march over an AoS primitive side buffer (parity, speed, ncu);
# occupancy variants of the 2-D step kernel (CTA shape x register cap) at 4096 and 512 rows
set -u
OUT=gpurun_out/hw_run16
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 3 "$OUT/$name.log" | cut -c1-700 | sed 's/^/   /' >> "$OUT/summary.txt"; }
NCU="ncu --clock-control none"
run hyp3d_tests 900 python -m pytest tests/test_hyp3d_gpu.py -m gpu -q
run hyp3d_bench 300 python bench_all.py hyp3d
run hyp3d_ncu 600 $NCU --set full --import-source on -k regex:hyp3d_step -s 40 -c 1 -o $OUT/hyp3d_step_r2c python bench_all.py hyp3d --steps3 5 --warm3 45
B="python bench.py --no-e2e --no-cpu --no-extras --no-other --steps 200 --warmup 20"
run h2_default 200 $B
run h2_default_512 200 $B --grid-h 512
for v in w3c7 w1r88 w2r80; do
  run h2_$v 200 env TAU_B200_LIB=scripts/variants/libtau_$v.so $B
  run h2_${v}_512 200 env TAU_B200_LIB=scripts/variants/libtau_$v.so $B --grid-h 512
done
cat "$OUT/summary.txt"
