#!/usr/bin/env bash
# round 2, N = 2: bench with the measured slab balance (plumbing check before the N = 8 run)
set -u
OUT=gpurun_out/hw_run27_n2
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 2 "$OUT/$name.log" | cut -c1-3000 | sed 's/^/   /' >> "$OUT/summary.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run bench_n2 300 $TR --master-port 29601 bench.py --gpus 2 --steps 20 --warmup 3 --trace-after 120 --total-timeout 200 --no-other
cat "$OUT/summary.txt"
