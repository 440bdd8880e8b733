#!/usr/bin/env bash
# round 2, 1 GPU: ncu --set full of the shipping 2-D step kernel (edge-strip fix, K = 1 schedule) at 4096 and 512 rows
set -u
OUT=gpurun_out/hw_run25
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 2 "$OUT/$name.log" | cut -c1-300 | sed 's/^/   /' >> "$OUT/summary.txt"; }
NCU="ncu --clock-control none"
run hyp2d_ncu 400 $NCU --set full --import-source on -k regex:hyp2d_step -s 120 -c 1 -o $OUT/hyp2d_step_r2d python bench.py --steps 4 --warmup 3 --develop 150 --no-e2e --no-cpu --no-other --no-extras
run hyp2d_ncu_512 400 $NCU --set full --import-source on -k regex:hyp2d_step -s 120 -c 1 -o $OUT/hyp2d_step_r2d_512 python bench.py --steps 4 --warmup 3 --develop 150 --no-e2e --no-cpu --no-other --no-extras --grid-h 512
run sph_ncu 600 $NCU --set full --import-source on -k regex:"sph_forces_integrate|sph_density" -s 4 -c 2 -o $OUT/sph_r2d python scripts/sph_stripe_probe.py 2097152 single
cat "$OUT/summary.txt"
