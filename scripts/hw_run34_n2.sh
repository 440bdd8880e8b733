#!/usr/bin/env bash
# round 2, N = 2: bench.py's partition agreement step (plumbing check, short)
set -u
OUT=gpurun_out/hw_run34_n2
mkdir -p "$OUT"
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 20 --warmup 3 --no-other --no-e2e --no-cpu --develop 300 > $OUT/bench_n2.log 2>&1
echo "exit $?"; tail -n 1 $OUT/bench_n2.log | cut -c1-600
