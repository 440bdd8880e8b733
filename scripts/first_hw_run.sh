#!/usr/bin/env bash
# First hardware run of everything that was written after the round-1 GPU budget was spent (NEXT.md).
# One gpurun call, each part under its own timeout so that a hang costs minutes, not the box:
#   /usr/local/graft/bin/gpurun --timeout 2700 -- 'bash scripts/first_hw_run.sh'
# Everything lands in gpurun_out/first_hw_run/.
set -u
OUT=gpurun_out/first_hw_run
mkdir -p "$OUT"
run() {  # name, timeout seconds, command...
  local name=$1 t=$2; shift 2
  echo "== $name" | tee -a "$OUT/summary.txt"
  timeout "$t" "$@" > "$OUT/$name.log" 2>&1
  echo "   exit $?" | tee -a "$OUT/summary.txt"
  tail -n 3 "$OUT/$name.log" | sed 's/^/   /' >> "$OUT/summary.txt"
}
# 0. the validated suite first: a regression there outranks everything below
run gpu_suite 900 python -m pytest tests -m gpu -q
run sph_diag 600 python scripts/sph_diag.py
# 1. pair kernel (hypersonic2d_pair.cuh): parity, then speed at the headline size and at a 512-row slab
run pair_parity 300 env TAU_TEST_PAIR=1 python -m pytest tests/test_hyp2d_gpu.py -m gpu -k pair -q -s
run pair_suite 600 env TAU_HYP2D_PAIR=1 python -m pytest tests/test_hyp2d_gpu.py -m gpu -q
run bench_prod 300 python bench.py --no-e2e --no-cpu
run bench_pair 300 env TAU_HYP2D_PAIR=1 python bench.py --no-e2e --no-cpu
run bench_prod_512 300 python bench.py --no-e2e --no-cpu --grid-h 512
run bench_pair_512 300 env TAU_HYP2D_PAIR=1 python bench.py --no-e2e --no-cpu --grid-h 512
# the fused kernel (TAU_HYP2D_PAIR=2: pair and production items in one kernel per step)
run fused_parity 300 env TAU_TEST_PAIR=1 TAU_HYP2D_PAIR=2 python -m pytest tests/test_hyp2d_gpu.py -m gpu -q
run bench_fused 300 env TAU_HYP2D_PAIR=2 python bench.py --no-e2e --no-cpu
run bench_fused_512 300 env TAU_HYP2D_PAIR=2 python bench.py --no-e2e --no-cpu --grid-h 512
for rr in 0 4 16; do
  run bench_pair_512_rest$rr 200 env TAU_HYP2D_PAIR=1 TAU_HYP2D_REST_ROWS=$rr python bench.py --no-e2e --no-cpu --grid-h 512
done
# 4 CTAs/SM build of the pair / fused kernels (128 registers, some spill) in a separate library
run ctas4_build 400 make -C fluid_sims_b200/csrc BUILD=../../build/csrc_ctas4 OUT=../libtau_b200_ctas4.so EXTRA_hypersonic2d=-DHP_MIN_CTAS=4
run bench_pair_ctas4 300 env TAU_B200_LIB=fluid_sims_b200/libtau_b200_ctas4.so TAU_HYP2D_PAIR=1 python bench.py --no-e2e --no-cpu
run bench_fused_ctas4 300 env TAU_B200_LIB=fluid_sims_b200/libtau_b200_ctas4.so TAU_HYP2D_PAIR=2 python bench.py --no-e2e --no-cpu
# 1b. ncu of the pair kernel (one launch of the developed flow; ~40 replays) and the launch list of a pair-mode step
run pair_ncu 600 env TAU_HYP2D_PAIR=1 ncu --set full --clock-control none --import-source on -k regex:hyp2d_step_pair \
  -s 300 -c 1 -o "$OUT/pair_prof" python bench.py --steps 4 --warmup 3 --develop 300 --no-e2e --no-cpu
run pair_launches 300 env TAU_HYP2D_PAIR=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 40 --csv \
  --log-file "$OUT/pair_launches.csv" python bench.py --steps 10 --warmup 3 --develop 300 --no-e2e --no-cpu
# 2. shallow water: parity (prints measured errors: replace the estimated bounds), golden, speed
run sw_parity 300 env TAU_TEST_SW=1 python -m pytest tests/test_sw_gpu.py -m gpu -q -s
run sw_golden 120 python tests/golden/make_golden_gpu.py sw
run sw_bench 120 python bench_all.py sw
# 3. .4spl export
run splat4 300 env TAU_TEST_4SPL=1 python -m pytest tests/test_splat4_gpu.py tests/test_hyp3d_gpu.py -m gpu -q -s
# 4. packed WENO5 build of the 3-D solver (separate library, the default one stays in place)
run weno_build 300 make -C fluid_sims_b200/csrc BUILD=../../build/csrc_weno OUT=../libtau_b200_weno.so EXTRA_hypersonic3d=-DT3_PACKED_WENO
run weno_parity 300 env TAU_B200_LIB=fluid_sims_b200/libtau_b200_weno.so python -m pytest tests/test_hyp3d_gpu.py -m gpu -q
run weno_bench 200 env TAU_B200_LIB=fluid_sims_b200/libtau_b200_weno.so python bench_all.py hyp3d
run default_bench3d 200 python bench_all.py hyp3d
cp tests/golden/sw_ref.npz "$OUT/" 2>/dev/null
cat "$OUT/summary.txt"
