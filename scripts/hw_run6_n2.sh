#!/usr/bin/env bash
# round 2, N = 2 (fourth call): row-flag item table (bit-identity for any decomposition), SPH stripes on hardware,
# whole GPU suite on a 2-GPU box, crc at N = 1 vs N = 2
set -u
OUT=gpurun_out/hw_run6_n2
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 8 "$OUT/$name.log" | cut -c1-2500 | sed 's/^/   /' >> "$OUT/summary.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run group_diag 300 python scripts/group_diag.py 2
run gpu_suite 900 python -m pytest tests -m gpu -q -s
run bench_n1 300 python bench.py --steps 100 --warmup 5 --no-other --no-extras --no-cpu --no-e2e
run bench_n2 300 $TR --master-port 29571 bench.py --gpus 2 --steps 100 --warmup 5 --trace-after 100 --total-timeout 150
run sph_n2 300 $TR --master-port 29572 bench_all.py sph
cat "$OUT/summary.txt"
