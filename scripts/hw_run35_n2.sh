#!/usr/bin/env bash
# round 2, N = 2, last seconds of the budget: tau3d --gpus 2 against --gpus 1 on hardware (no Python)
cd fluid_sims_b200/cli
mkdir -p ../../gpurun_out/hw_run35_n2
./tau3d --n 96 --frames 10 --gpus 1 --dump /tmp/a.bin > ../../gpurun_out/hw_run35_n2/g1.log 2>&1; echo "g1 exit $?"
./tau3d --n 96 --frames 10 --gpus 2 --dump /tmp/b.bin > ../../gpurun_out/hw_run35_n2/g2.log 2>&1; echo "g2 exit $?"
cmp /tmp/a.bin /tmp/b.bin && echo "DUMPS IDENTICAL" | tee ../../gpurun_out/hw_run35_n2/cmp.log
tail -n 2 ../../gpurun_out/hw_run35_n2/g1.log ../../gpurun_out/hw_run35_n2/g2.log
