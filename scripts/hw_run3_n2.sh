#!/usr/bin/env bash
# round 2, N = 2: multi-GPU parity (2-D chain, GS ring, 3-D ring, SPH shards), the new bench line at N = 2,
# the device-side frame hand-over (e2e), the fused kernel in slab mode
set -u
OUT=gpurun_out/hw_run3_n2
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 6 "$OUT/$name.log" | cut -c1-3000 | sed 's/^/   /' >> "$OUT/summary.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run mgpu_tests 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -s
run hypc 300 python -m pytest tests/test_hypc_gpu.py tests/test_cabi.py -m gpu -q
run bench_n2 600 $TR --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 5
run bench_n2_async 600 $TR --master-port 29542 bench.py --gpus 2 --steps 100 --warmup 5 --e2e-peers-async --no-other
run bench_n2_async1 600 $TR --master-port 29543 bench.py --gpus 2 --steps 100 --warmup 5 --e2e-peers-async --e2e-lanes 1 --no-other
run bench_n2_fused 600 env TAU_HYP2D_PAIR=2 $TR --master-port 29544 bench.py --gpus 2 --steps 100 --warmup 5 --no-e2e --no-other
run bench_n2_20 600 $TR --master-port 29545 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-other
cat "$OUT/summary.txt"
