#!/usr/bin/env bash
# round 2, 1 GPU: SPH with SUBX key columns + one 16-byte record per candidate in the force sweep (parity, speed, launch list, ncu)
set -u
OUT=gpurun_out/hw_run14
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 4 "$OUT/$name.log" | cut -c1-2000 | sed 's/^/   /' >> "$OUT/summary.txt"; }
NCU="ncu --clock-control none"
run sph_tests 600 python -m pytest tests/test_sph_gpu.py -m gpu -q -s
run sph_bench 300 python bench_all.py sph
run sph_launches 300 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/sph_launches.csv python scripts/sph_stripe_probe.py 2097152 single
run sph_ncu 600 $NCU --set full --import-source on -k regex:"sph_forces_integrate|sph_density" -s 4 -c 2 -o $OUT/sph_r2c python scripts/sph_stripe_probe.py 2097152 single
cat "$OUT/summary.txt"
