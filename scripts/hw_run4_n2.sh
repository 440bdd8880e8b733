#!/usr/bin/env bash
# round 2, N = 2 (second call): the one-process group handle + `--gpus`, the new e2e default, crc at N = 1 vs 2,
# where rank 0's extra 20 us come from (same slab on either device; sampler off)
set -u
OUT=gpurun_out/hw_run4_n2
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 6 "$OUT/$name.log" | cut -c1-2500 | sed 's/^/   /' >> "$OUT/summary.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run group_tests 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "group or cli"
run bench_n1 600 python bench.py --steps 100 --warmup 5 --no-other --no-extras --no-cpu --no-e2e
run bench_n2 600 $TR --master-port 29551 bench.py --gpus 2 --steps 100 --warmup 5 --no-other
run half_dev0 300 env CUDA_VISIBLE_DEVICES=0 python bench.py --grid-h 2048 --steps 100 --warmup 5 --no-other --no-extras --no-cpu --no-e2e
run half_dev1 300 env CUDA_VISIBLE_DEVICES=1 python bench.py --grid-h 2048 --steps 100 --warmup 5 --no-other --no-extras --no-cpu --no-e2e
run cli_gpus2 300 ./fluid_sims_b200/cli/tau_2d_hypersonic_cuda --nx 4096 --ny 4096 --dtype f32 --frames 500 --gpus 2
run cli_gpus1 300 ./fluid_sims_b200/cli/tau_2d_hypersonic_cuda --nx 4096 --ny 4096 --dtype f32 --frames 500 --gpus 1
cat "$OUT/summary.txt"
