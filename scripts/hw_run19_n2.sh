#!/usr/bin/env bash
# round 2, N = 2: multi-GPU tests (2-D slabs, 3-D ring with the primitive side buffer, GS ring, SPH stripes with SUBX key columns),
# bench at N = 2 with other_configs (state_crc against the N = 1 lines)
set -u
OUT=gpurun_out/hw_run19_n2
mkdir -p "$OUT"
run() { local name=$1 t=$2; shift 2; echo "== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $?" | tee -a "$OUT/summary.txt"; tail -n 6 "$OUT/$name.log" | cut -c1-3000 | sed 's/^/   /' >> "$OUT/summary.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run mgpu_tests 600 python -m pytest tests/test_multi_gpu.py tests/test_cli_gpu.py -m gpu -q -s
run bench_n2 400 $TR --master-port 29601 bench.py --gpus 2 --steps 20 --warmup 3 --trace-after 120 --total-timeout 200
run sph_n2 300 $TR --master-port 29602 bench_all.py sph
run hyp3d_n2 300 $TR --master-port 29603 bench_all.py hyp3d --n3 512 --steps3 12 --warm3 20
cat "$OUT/summary.txt"
