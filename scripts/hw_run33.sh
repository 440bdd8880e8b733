#!/usr/bin/env bash
# round 2, 1 GPU: ncu launch list of the bench command on the shipping build (headline part only)
set -u
OUT=gpurun_out/hw_run33
mkdir -p "$OUT"
timeout 150 ncu --clock-control none --metrics gpu__time_duration.sum -c 3000 --csv --log-file $OUT/bench_launches.csv python bench.py --steps 20 --warmup 3 --develop 100 --e2e-frames 4 --cpu-steps 1 --cpu-full-steps 0 --no-other --no-extras > $OUT/bench_launches.log 2>&1
echo "exit $?"; tail -n 1 $OUT/bench_launches.log | cut -c1-300
