#!/usr/bin/env bash
# round 2, 1 GPU: profiles (launch lists + ncu --set full of the shipping kernels), SPH stripe overhead per kernel
set -u
OUT=gpurun_out/hw_run7
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 4 "$OUT/$name.log" | cut -c1-1500 | sed 's/^/   /' >> "$OUT/summary.txt"; }
NCU="ncu --clock-control none"
run bench 400 python bench.py
run bench_launches 400 $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $OUT/bench_launches.csv python bench.py --steps 20 --warmup 3 --develop 100 --e2e-frames 4 --cpu-steps 1 --cpu-full-steps 0
run hyp2d_ncu 400 $NCU --set full --import-source on -k regex:hyp2d_step -s 120 -c 1 -o $OUT/hyp2d_step_r2 python bench.py --steps 4 --warmup 3 --develop 150 --no-e2e --no-cpu --no-other --no-extras
run sph_launches 300 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $OUT/sph_launches.csv python scripts/sph_stripe_probe.py
run sph_ncu 600 $NCU --set full --import-source on -k regex:"sph_density|sph_forces_integrate" -s 4 -c 2 -o $OUT/sph_r2 python scripts/sph_stripe_probe.py 2097152 single
run hyp3d_ncu 600 $NCU --set full --import-source on -k regex:hyp3d_step -s 40 -c 1 -o $OUT/hyp3d_step_r2 python bench_all.py hyp3d --steps3 5 --warm3 45
cat "$OUT/summary.txt"
