#!/usr/bin/env bash
# round 2, N = 8 (second call): cost-weighted partition against equal slabs, N = 4, N = 1 crc at the driver's step counts
set -u
OUT=gpurun_out/hw_run10_n8
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 3 "$OUT/$name.log" | cut -c1-3000 | sed 's/^/   /' >> "$OUT/summary.txt"; }
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
X="--no-other --no-e2e --trace-after 100"
run n8_weighted 200 $TR8 --master-port 29591 bench.py --gpus 8 --steps 20 --warmup 3 $X
run n8_equal 200 $TR8 --master-port 29592 bench.py --gpus 8 --steps 20 --warmup 3 --equal-slabs $X
run n4_weighted 200 env CUDA_VISIBLE_DEVICES=0,1,2,3 $TR4 --master-port 29593 bench.py --gpus 4 --steps 20 --warmup 3 $X
run n4_equal 200 env CUDA_VISIBLE_DEVICES=0,1,2,3 $TR4 --master-port 29594 bench.py --gpus 4 --steps 20 --warmup 3 --equal-slabs $X
run n1 200 python bench.py --steps 20 --warmup 3 --no-other --no-e2e --no-cpu --no-extras
cat "$OUT/summary.txt"
