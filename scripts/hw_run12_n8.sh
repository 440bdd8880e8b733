#!/usr/bin/env bash
# round 2, N = 8: where a stripe sub-step's time goes (host enqueue vs device, with / without the exchanges), final bench line
set -u
OUT=gpurun_out/hw_run12_n8
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 3 "$OUT/$name.log" | cut -c1-3000 | sed 's/^/   /' >> "$OUT/summary.txt"; }
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run sph_n8 200 $TR8 --master-port 29611 bench_all.py sph
run sph_n8_noxchg 200 env TAU_SPH_STRIPE_NOXCHG=1 $TR8 --master-port 29612 bench_all.py sph
run mgpu_world8 300 env TAU_TEST_WORLD=8 python -m pytest tests/test_multi_gpu.py -m gpu -q -s -k slab_runs
run bench_n8 300 $TR8 --master-port 29613 bench.py --gpus 8 --steps 20 --warmup 3 --trace-after 120 --total-timeout 200
cat "$OUT/summary.txt"
