#!/usr/bin/env bash
# round 2, N = 2: 3-D sensitivity-yardstick parity test, other_configs with state_crc at N = 1 and N = 2 (512^3, SPH), NUMA binding of the e2e buffers
set -u
OUT=gpurun_out/hw_run11_n2
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 12 "$OUT/$name.log" | cut -c1-3000 | sed 's/^/   /' >> "$OUT/summary.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run hyp3d_tests 600 python -m pytest tests/test_hyp3d_gpu.py -m gpu -q -s -k "developed_flow"
run bench_n1 400 env CUDA_VISIBLE_DEVICES=0 python bench.py --steps 20 --warmup 3 --no-cpu
run bench_n2 400 $TR --master-port 29601 bench.py --gpus 2 --steps 20 --warmup 3 --trace-after 120 --total-timeout 200
cat "$OUT/summary.txt"
