#!/usr/bin/env bash
# round 2, 1 GPU: the whole suite with the new tests (fp32 yardstick, config 1), 3-D kernel with the new face order
set -u
OUT=gpurun_out/hw_run8
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 6 "$OUT/$name.log" | cut -c1-1500 | sed 's/^/   /' >> "$OUT/summary.txt"; }
NCU="ncu --clock-control none"
run gpu_suite 1200 python -m pytest tests -m gpu -q -s
run bench3d 300 python bench_all.py hyp3d
run hyp3d_ncu 600 $NCU --set full --import-source on -k regex:hyp3d_step -s 40 -c 1 -o $OUT/hyp3d_step_r2b python bench_all.py hyp3d --steps3 5 --warm3 45
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
cat "$OUT/summary.txt"
