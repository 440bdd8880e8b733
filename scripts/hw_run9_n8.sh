#!/usr/bin/env bash
# round 2, N = 8: multi-GPU parity at world 4 and 8, the group handle / --gpus on 8 devices, bench at N = 8 and N = 4
set -u
OUT=gpurun_out/hw_run9_n8
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 8 "$OUT/$name.log" | cut -c1-3000 | sed 's/^/   /' >> "$OUT/summary.txt"; }
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
run mgpu_tests 400 python -m pytest tests/test_multi_gpu.py -m gpu -q -s
run mgpu_world8 300 env TAU_TEST_WORLD=8 python -m pytest tests/test_multi_gpu.py -m gpu -q -s -k slab_runs
run bench_n8 300 $TR8 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 3 --trace-after 100 --total-timeout 150
run bench_n8_long 200 $TR8 --master-port 29582 bench.py --gpus 8 --steps 500 --warmup 20 --no-other --no-e2e --trace-after 100
run bench_n4 300 env CUDA_VISIBLE_DEVICES=0,1,2,3 $TR4 --master-port 29583 bench.py --gpus 4 --steps 20 --warmup 3 --trace-after 100 --total-timeout 150
run cli_gpus8 200 ./fluid_sims_b200/cli/tau_2d_hypersonic_cuda --nx 4096 --ny 4096 --dtype f32 --frames 1000 --gpus 8
cat "$OUT/summary.txt"
