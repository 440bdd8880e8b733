#!/usr/bin/env bash
# round 2, N = 2 (third call): where the group handle parts from the single-GPU run (f64 test failure), where the default
# N = 2 bench hung (faulthandler after 60 s, tight timeouts), mirrored layer schedule
set -u
OUT=gpurun_out/hw_run5_n2
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 6 "$OUT/$name.log" | cut -c1-2500 | sed 's/^/   /' >> "$OUT/summary.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run group_diag 300 python scripts/group_diag.py 2
run group_diag_chunk1 300 env TAU_HYP2D_GROUP_CHUNK=1 python scripts/group_diag.py 2
run bench_n2 200 $TR --master-port 29561 bench.py --gpus 2 --steps 100 --warmup 5 --trace-after 60 --total-timeout 90
run bench_n2_noother 200 $TR --master-port 29562 bench.py --gpus 2 --steps 100 --warmup 5 --no-other --trace-after 60
cat "$OUT/summary.txt"
