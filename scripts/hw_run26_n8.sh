#!/usr/bin/env bash
# round 2, N = 8: the driver's bench command after this session's changes (edge strips, schedule, 3-D side buffer, SPH key columns)
set -u
OUT=gpurun_out/hw_run26_n8
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 2 "$OUT/$name.log" | cut -c1-3000 | sed 's/^/   /' >> "$OUT/summary.txt"; }
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run bench_n8 300 $TR8 --master-port 29613 bench.py --gpus 8 --steps 20 --warmup 3 --trace-after 120 --total-timeout 200
cat "$OUT/summary.txt"
